#!/bin/bash
# timing experiment on a -DARSEG_TTRACE build: per-role clock trace with the softmax (16) / epilogue (32) / depthwise (1) / gather (4) work removed
ARSEG_NVCC_EXTRA="-DARSEG_TTRACE" python -m arseg_b200.build > /dev/null 2>&1
for d in 0 16 32 48 5; do
echo "=== ARSEG_CREFF_DBG=$d"
ARSEG_CREFF_DBG=$d timeout 300 python tools/tc_trace.py 11 2>&1 | head -21 | cut -c1-110
done
