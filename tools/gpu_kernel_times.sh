#!/bin/bash
# Per-launch times (ncu, serialised, cold cache) of the frame's kernels matching a regex, on the bench workload.
# usage: tools/gpu_kernel_times.sh TAG REGEX [SKIP] [COUNT]
TAG=${1:-kt}; RE=${2:-k_}; SKIP=${3:-0}; COUNT=${4:-40}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none \
    --kernel-name "regex:$RE" --launch-skip $SKIP --launch-count $COUNT --csv --log-file gpurun_out/${TAG}_times.csv \
    python bench.py --steps 1 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_times.log 2>&1
python - gpurun_out/${TAG}_times.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
out = {}
for r in rows[1:]:
    out.setdefault((int(r[ii]), r[ki].split("(")[0][:40]), {})[r[mi]] = float(r[vi].replace(",", ""))
for (i, k), m in sorted(out.items()):
    print("%3d %-32s %9.1f us  %8.1f Minst  rd %7.1f MB  wr %7.1f MB  fp64 %4.1f%%  issue %4.1f%%  warps %4.1f%%" % (i, k[:32], m.get("gpu__time_duration.sum", 0) / 1e3,
          m.get("smsp__inst_executed.sum", 0) / 1e6, m.get("dram__bytes_read.sum", 0) / 1e6, m.get("dram__bytes_write.sum", 0) / 1e6,
          m.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 0), m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0),
          m.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0)))
PY
