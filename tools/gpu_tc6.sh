#!/bin/bash
# per-role clock trace of the current engine (-DARSEG_TTRACE build on the box)
ARSEG_NVCC_EXTRA="-DARSEG_TTRACE" python -m arseg_b200.build > /dev/null 2>&1
timeout 300 python tools/tc_trace.py 11 2>&1 | tee gpurun_out/tc_trace.log | cut -c1-150
