#!/bin/bash
# sweep time as a function of the share of planes handed out as small tail tiles and of their depth
for cfg in "20 4" "10 8" "20 8" "30 8" "15 2"; do
  set -- $cfg
  CONTOUR3D_SWEEP_TAIL=$1 CONTOUR3D_SWEEP_TAILDIV=$2 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('tail $1%% div $2: sweep %.4f ms  frac %.3f  ms/step %.3f' % (d['roofline']['ms_per_launch'], d['roofline']['frac'], d['ms_per_step']))"
done
