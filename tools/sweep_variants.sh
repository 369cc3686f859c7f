#!/bin/bash
# sweep-only timings under the experiment switches of c3d_sweep.cu
for promo in 0 2 3; do for warps in 12 16; do
  echo "== L2PROMO=$promo WARPS=$warps"
  CONTOUR3D_SWEEP_L2PROMO=$promo CONTOUR3D_SWEEP_WARPS=$warps python tools/sweep_probe.py 1024 2>&1 | grep -E "float32.*launch=3|float64.*launch=3|launch=1"
done; done
