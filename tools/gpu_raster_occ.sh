#!/bin/bash
for occ in 4 5 6 8; do
  CONTOUR3D_RASTER_OCC=$occ python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('occ $occ: raster %.3f ms  ms/step %.3f' % (d['stages_ms']['raster'], d['ms_per_step']))"
done
oracle/_ref/demo_case /dev/shm/demo.bin && MINIMPI_NP=1 contour3d_b200/c3d_case /dev/shm/demo.bin /dev/shm/demo.ppm 1 1 10 3 && md5sum /dev/shm/demo.ppm
CONTOUR3D_VERBOSE=1 MINIMPI_NP=1 contour3d_b200/c3d_case /dev/shm/demo.bin /dev/shm/demo.ppm 1 1 2 1 2>&1 | tail -3
