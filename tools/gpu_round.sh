#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list + one full capture of the frame's kernels.
# usage: tools/gpu_round.sh TAG   (outputs under gpurun_out/TAG_*)
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name 'regex:k_sweep_tiles|k_raster|k_shade_select|k_shade_setup|k_shade_normals|k_shade_blend|k_expand|k_bucket_count' \
    --launch-skip 16 --launch-count 8 -o gpurun_out/${TAG}_full -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
# config 5's size on one GPU (2048^3 fp32 = 34 GB, 8K image): a size-robustness data point, not the bench line
timeout 600 python bench.py --size 2048 --image 7680x4320 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench_2048.json 2> gpurun_out/${TAG}_bench_2048.err
ls -la gpurun_out
