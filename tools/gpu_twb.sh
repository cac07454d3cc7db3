#!/bin/bash
# sweep of the MV-warp pre-pass band height (rows of all N frames processed together), fp32 and f16 keyframe feature
for hr in "--hr32" ""; do for twb in 4 8 32 96 256 2048; do
ARSEG_TC_TWB=$twb timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:creff_tc_warp --csv --log-file gpurun_out/twb.csv python tools/prof_creff.py --engine tc $hr --frames 11 --iters 2 > /dev/null 2>&1
echo "hr=$hr TWB=$twb: $(grep creff_tc_warp gpurun_out/twb.csv | tail -1 | awk -F'","' '{print $NF}')"
done; done
