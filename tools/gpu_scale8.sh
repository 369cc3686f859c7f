#!/bin/bash
# 8-GPU visit: the 8-rank parity case, then the bench at N=8 (config 3) and the 2048^3 / 8K variant (config 5's size).
TAG=${1:-s8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -k "1x4x2 or 1x1x4" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
run() {  # N extra-args tag
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29617 \
      bench.py --gpus $1 --steps 8 --warmup 3 --no-cpu-baseline $2 > gpurun_out/${TAG}_$3.json 2>> gpurun_out/${TAG}_bench.err
  python - gpurun_out/${TAG}_$3.json $3 <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    s = d["stages_ms"]
    print("%-10s ms/step %.3f value %.1f | e2e %s" % (sys.argv[2], d["ms_per_step"], d["value"], d["e2e"] and round(d["e2e"]["ms_per_step"], 1)))
    for r, pr in enumerate(d.get("stages_ms_per_rank") or []):
        print("   rank %d: sweep %.3f expand %.3f raster %.3f reduce %.3f shade %.3f total %.3f cubes %d" % (r, pr["sweep"], pr["expand"], pr["raster"], pr["reduce"], pr["shade"], pr["total"], pr["active_cubes"]))
except Exception as e:
    print(sys.argv[2], "failed:", e)
PY
}
run 8 "" n8
run 4 "--no-e2e" n4
run 8 "--size 2048 --image 7680x4320 --no-e2e" n8_2048
tail -3 gpurun_out/${TAG}_bench.err
