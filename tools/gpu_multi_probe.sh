#!/bin/bash
# N-GPU probe: multi-rank parity test, then the frame at N ranks with different Hi-Z cut lists.
# usage: tools/gpu_multi_probe.sh TAG N "cuts1;cuts2;..."   (outputs under gpurun_out/TAG_*)
TAG=${1:-multi}
N=${2:-2}
CUTS=${3:-"-;6,20,40"}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
IFS=';' read -ra LIST <<< "$CUTS"
for cuts in "${LIST[@]}"; do
  val="$cuts"; [ "$cuts" = "-" ] && val=""
  out=gpurun_out/${TAG}_n${N}_cuts_${cuts//,/_}.json
  CONTOUR3D_HIZ_CUTS="$val" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $N --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > $out 2>> gpurun_out/${TAG}_bench.err
  python - "$cuts" $out <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]).read().strip().splitlines() if l.startswith("{")][-1])
    s = d["stages_ms"]
    print("cuts=%-12s ms/step %.3f | sweep %.3f expand %.3f raster %.3f reduce %.3f shade %.3f | value %.1f" % (sys.argv[1], d["ms_per_step"], s["sweep"], s["expand"], s["raster"], s["reduce"], s["shade"], d["value"]))
except Exception as e:
    print("cuts=%s failed: %s" % (sys.argv[1], e))
PY
done
tail -5 gpurun_out/${TAG}_bench.err
