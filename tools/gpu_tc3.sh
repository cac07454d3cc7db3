#!/bin/bash
# tcgen05 CReFF engine variants: tests + per-kernel durations for two warp splits (built on the box), trace of the last one
mkdir -p gpurun_out
for split in 3 2; do
ARSEG_NVCC_EXTRA="-DARSEG_TC_KVSPLIT=$split" python -m arseg_b200.build > /dev/null 2>&1
echo "== KVSPLIT=$split"
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "creff_tc" 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:creff_tc_kernel --csv --log-file gpurun_out/tc_kernels.csv python tools/prof_creff.py --engine tc --frames 11 --iters 2 > /dev/null 2>&1
grep creff_tc_kernel gpurun_out/tc_kernels.csv | awk -F'","' '{print $NF}' | tr '\n' ' '; echo
done
if [ "$1" = "trace" ]; then
ARSEG_NVCC_EXTRA="-DARSEG_TTRACE -DARSEG_TC_KVSPLIT=${2:-3}" python -m arseg_b200.build > /dev/null 2>&1
timeout 300 python tools/tc_trace.py 11 2>&1 | tee gpurun_out/tc_trace.log | head -22
fi
