#!/bin/bash
# tcgen05 CReFF engine: per-kernel durations, then the per-role wait trace from a -DARSEG_TTRACE build (built on the box)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "creff_tc" 2>&1 | tail -2
timeout 300 python tools/prof_creff.py --engine tc --frames 11 --iters 5 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:creff_tc --csv --log-file gpurun_out/tc_kernels.csv python tools/prof_creff.py --engine tc --frames 11 --iters 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(l for l in open("gpurun_out/tc_kernels.csv") if l.startswith('"'))]
h = rows[0]
for r in rows[1:]:
    d = dict(zip(h, r))
    print(d["Kernel Name"][:60], d["Metric Value"], d["Metric Unit"])
PY
if [ "$1" = "trace" ]; then
ARSEG_NVCC_EXTRA=-DARSEG_TTRACE python -m arseg_b200.build > /dev/null 2>&1
timeout 300 python tools/tc_trace.py 11 2>&1 | tee gpurun_out/tc_trace.log
fi
