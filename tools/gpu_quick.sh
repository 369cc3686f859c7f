#!/bin/bash
# quick 1-GPU visit: parity tests, one bench line (stages), per-kernel ncu times of kernels matching REGEX
# usage: tools/gpu_quick.sh TAG REGEX SKIP COUNT
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('ms/step %.3f' % d['ms_per_step'], {k: round(v,3) for k,v in d['stages_ms'].items()}, 'spills', d['shade_spills'], 'roofline %.3f' % d['roofline']['frac'])
"
if [ -n "$2" ]; then bash tools/gpu_kernel_times.sh $TAG "$2" ${3:-0} ${4:-40}; fi
