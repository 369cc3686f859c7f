#!/bin/bash
# The measurement visits of round 2, as they were run through gpurun (one section per call; see profiles/README.md).
# Usage on the GPU box: bash tools/gpu_round2.sh one|two|four|eight TAG
set -u
what=${1:-one}; tag=${2:-r2}; out=gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" bench.py --gpus "$1" "${@:3}"; }
case "$what" in
  one)   # one B200: tests, every configuration, the reference arm, the ncu launch list and one full capture of a frame
    timeout 2300 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1
    python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json
    python bench.py --config 2 --steps 20 --warmup 5 > $out/${tag}_bench_config2.json
    python bench.py --config 4 --steps 20 --warmup 5 --e2e-max-gb 10 > $out/${tag}_bench_config4.json
    python bench.py --config 5 --steps 10 --warmup 3 > $out/${tag}_bench_config5_1gpu.json
    python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_reference_arm.json
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
    ncu --set full --clock-control none --import-source on --launch-skip 34 --launch-count 17 -o $out/${tag}_frame python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
    ;;
  two)   run 2 29641 --steps 20 --warmup 5 > $out/${tag}_scale_n2.json ;;
  four)  run 4 29651 --steps 20 --warmup 5 > $out/${tag}_scale_n4.json ;;
  eight) # never under ncu: a profiler cannot step through kernels that wait for another process
    run 8 29661 --steps 20 --warmup 5 > $out/${tag}_scale_n8.json
    run 8 29662 --config 5 --steps 20 --warmup 5 --no-e2e > $out/${tag}_config5_n8.json
    run 8 29663 --config 4 --steps 20 --warmup 5 --no-e2e > $out/${tag}_config4_n8.json
    ;;
esac
