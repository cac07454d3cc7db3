#!/bin/bash
# CReFF engine iteration: creff tests, A/B timing (march vs tile), optional ncu.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "creff" 2>&1 | tail -15 | tee gpurun_out/t_creff.log
timeout 300 python tools/prof_creff.py --frames 11 --iters 5 2>&1 | tail -2 | tee gpurun_out/creff_time.log
ARSEG_CREFF_TILE=1 timeout 300 python tools/prof_creff.py --frames 11 --iters 5 2>&1 | tail -1 | tee -a gpurun_out/creff_time.log
for sr in 720 240 144 96; do ARSEG_CREFF_SEG_ROWS=$sr timeout 300 python tools/prof_creff.py --frames 11 --iters 4 2>&1 | tail -1 | sed "s/^/seg_rows=$sr /" | tee -a gpurun_out/creff_time.log; done
if [ "$1" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:creff_march -c 1 -f -o gpurun_out/creff_march_full python tools/prof_creff.py --frames 1 --iters 1 2>&1 | tail -2
fi
