#!/bin/bash
# memcheck + racecheck of one small frame (smoke) and of a converter frame; outputs under gpurun_out/TAG_*
TAG=${1:-san}
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|smoke ok" gpurun_out/${TAG}_memcheck.log | head -5
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "converter and twisted and float32 or hiz" > gpurun_out/${TAG}_memcheck2.log 2>&1; echo "memcheck2 rc=$?"
grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/${TAG}_memcheck2.log | head -5
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|hazard|smoke ok" gpurun_out/${TAG}_racecheck.log | head -5
