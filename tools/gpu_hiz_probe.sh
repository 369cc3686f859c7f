#!/bin/bash
# One-GPU probe of the raster's Hi-Z refresh: the GPU parity tests, then the frame with different cut lists.
# usage: tools/gpu_hiz_probe.sh TAG   (outputs under gpurun_out/TAG_*)
TAG=${1:-hiz}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
for cuts in "" "8" "4,16" "6,20,40" "2,6,14,30"; do
  CONTOUR3D_HIZ_CUTS="$cuts" timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench_cuts_${cuts//,/_}.json 2>> gpurun_out/${TAG}_bench.err
  python - "$cuts" gpurun_out/${TAG}_bench_cuts_${cuts//,/_}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("cuts=%-12s ms/step %.3f raster %.3f shade %.3f culled %d" % (sys.argv[1], d["ms_per_step"], d["stages_ms"]["raster"], d["stages_ms"]["shade"], d["culled_cubes"]))
except Exception as e:
    print("cuts=%s failed: %s" % (sys.argv[1], e))
PY
done
