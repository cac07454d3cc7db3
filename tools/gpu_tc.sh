#!/bin/bash
# tcgen05 CReFF engine iteration: tests, timing vs the march engine, optional ncu capture ($1 = ncu)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "creff_tc" 2>&1 | tail -3 | tee gpurun_out/t_tc.log
timeout 300 python tools/prof_creff.py --engine tc --frames 11 --iters 5 2>&1 | tail -1 | tee gpurun_out/tc_time.log
timeout 300 python tools/prof_creff.py --engine mma --frames 11 --iters 5 2>&1 | tail -1 | tee -a gpurun_out/tc_time.log
if [ "$1" = "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:creff_tc_kernel -c 1 -f -o gpurun_out/creff_tc_full python tools/prof_creff.py --engine tc --frames 11 --iters 1 2>&1 | tail -2
fi
# per-kernel durations (pre-passes + engine), cold-cache serialised
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:creff_tc --csv --log-file gpurun_out/tc_kernels.csv python tools/prof_creff.py --engine tc --frames 11 --iters 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(l for l in open("gpurun_out/tc_kernels.csv") if l.startswith('"'))]
h = rows[0]
for r in rows[1:]:
    d = dict(zip(h, r))
    print(d["Kernel Name"][:60], d["Metric Value"], d["Metric Unit"])
PY
