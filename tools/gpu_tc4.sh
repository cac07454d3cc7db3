#!/bin/bash
# tcgen05 CReFF engine build variants (role splits): tc tests + engine kernel duration (ncu, serialised) per variant
mkdir -p gpurun_out
run() {
ARSEG_NVCC_EXTRA="$1" python -m arseg_b200.build > /dev/null 2>&1
T=$(timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "creff_tc" 2>&1 | tail -1)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:creff_tc_kernel --csv --log-file gpurun_out/tc_kernels.csv python tools/prof_creff.py --engine tc --frames 11 --iters 2 > /dev/null 2>&1
echo "variant [$1]: $(grep creff_tc_kernel gpurun_out/tc_kernels.csv | awk -F'","' '{print $NF}' | tr '\n' ' ') tests: $T"
}
for v in "$@"; do run "$v"; done
