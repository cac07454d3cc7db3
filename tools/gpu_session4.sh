#!/bin/bash
# Full GPU pass at HEAD: parity suite, smoke, default bench (f16 + tf32 alt, CPU baseline), reference arm,
# ncu launch list, ncu --set full of the CReFF kernel from bench.py, config-5 sweep, configs 2-4 table.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "=== gpu tests"
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/t_all.log
echo "=== smoke"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -8 | tee gpurun_out/t_smoke.log
echo "=== bench default"
timeout 900 python bench.py --profile > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 4000 gpurun_out/bench_default.json; head -40 gpurun_out/bench_default.err
echo "=== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 1500 gpurun_out/bench_reference.json
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --alt-precision none > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/launches.csv
echo "=== ncu full creff (f16 plan, from bench.py)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:creff_march -c 1 -f -o gpurun_out/creff_march_f16 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --alt-precision none 2>&1 | tail -2
echo "=== ncu full conv up_1 (f16 plan, from bench.py)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_halo_kernel.*256" -s 14 -c 1 -f -o gpurun_out/conv_halo_f16 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --alt-precision none 2>&1 | tail -2
echo "=== sweep"
timeout 600 python tools/sweep_creff.py --iters 3 2>&1 | tail -32
echo "=== configs"
timeout 600 python tools/bench_configs.py 2>&1 | tail -4
