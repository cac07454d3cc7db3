#!/bin/bash
# Baseline at HEAD: full GPU parity suite, smoke, bench (with CPU baseline), ncu launch list, ncu full capture of CReFF.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "=== gpu tests"
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/t_all.log
echo "=== smoke"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -8 | tee gpurun_out/t_smoke.log
echo "=== bench tf32"
timeout 900 python bench.py --steps 20 --warmup 3 --profile > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; tail -c 3000 gpurun_out/bench_tf32.json; head -50 gpurun_out/bench_tf32.err
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/launches.csv
echo "=== ncu full creff"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:creff_mma -c 1 -f -o gpurun_out/creff_full python tools/prof_creff.py --frames 1 --iters 1 2>&1 | tail -3
timeout 300 python tools/prof_creff.py --frames 11 --iters 5 2>&1 | tail -2
