#!/bin/bash
# tcgen05 CReFF engine iteration: tests, timing vs the march engine, optional ncu capture ($1 = ncu)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "creff_tc" 2>&1 | tail -3 | tee gpurun_out/t_tc.log
timeout 300 python tools/prof_creff.py --engine tc --frames 11 --iters 5 2>&1 | tail -1 | tee gpurun_out/tc_time.log
timeout 300 python tools/prof_creff.py --engine mma --frames 11 --iters 5 2>&1 | tail -1 | tee -a gpurun_out/tc_time.log
if [ "$1" = "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:creff_tc_kernel -c 1 -f -o gpurun_out/creff_tc_full python tools/prof_creff.py --engine tc --frames 11 --iters 1 2>&1 | tail -2
fi
