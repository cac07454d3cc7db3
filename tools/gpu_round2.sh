#!/bin/bash
# Full GPU pass at HEAD (round 2): parity suite, smoke, default bench (f16 + tf32 alt, CPU baseline, extras), A/B against the march
# engine, reference arm, ncu launch list, ncu --set full of the CReFF kernels and one conv layer from bench.py, config-5 sweep,
# configs 2-4 table.  tools/make_profiles.py <tag> turns gpurun_out/ into profiles/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
B="--no-cpu-baseline --alt-precision none --no-extras"
echo "=== gpu tests"
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/t_all.log
echo "=== smoke"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -8 | tee gpurun_out/t_smoke.log
echo "=== bench default"
timeout 900 python bench.py --profile > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 1500 gpurun_out/bench_default.json
echo "=== bench, march engine (ARSEG_CREFF_TC=0)"
ARSEG_CREFF_TC=0 timeout 600 python bench.py $B --profile > gpurun_out/bench_march.json 2> gpurun_out/bench_march.err; tail -c 600 gpurun_out/bench_march.json
echo "=== bench, pre-pass not hoisted (ARSEG_PLAN_OVERLAP=0)"
ARSEG_PLAN_OVERLAP=0 timeout 600 python bench.py $B > gpurun_out/bench_serialplan.json 2> gpurun_out/bench_serialplan.err; tail -c 600 gpurun_out/bench_serialplan.json
echo "=== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 1200 gpurun_out/bench_reference.json
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 $B > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/launches.csv
echo "=== ncu full: creff_tc_kernel, creff_tc_warp_kernel (f16 plan), creff_march (tf32 plan), conv up_1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:creff_tc_kernel -c 1 -f -o gpurun_out/creff_tc_f16 python bench.py --steps 1 --warmup 1 $B 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:creff_tc_warp_kernel -c 1 -f -o gpurun_out/creff_tc_warp_f16 python bench.py --steps 1 --warmup 1 $B 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:creff_march -c 1 -f -o gpurun_out/creff_march_tf32 python bench.py --precision tf32 --steps 1 --warmup 1 $B 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_halo_kernel -s 14 -c 1 -f -o gpurun_out/conv_halo_f16 python bench.py --steps 1 --warmup 1 $B 2>&1 | tail -1
echo "=== sweep"
timeout 900 python tools/sweep_creff.py --iters 3 2>&1 | tail -32
echo "=== configs"
timeout 600 python tools/bench_configs.py 2>&1 | tail -4
