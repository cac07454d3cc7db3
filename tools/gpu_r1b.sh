#!/bin/bash
# CReFF MMA engine bring-up: new tests, creff regression tests, smoke, short bench with per-kernel profile.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "creff" 2>&1 | tail -30 | tee gpurun_out/t_creff.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -8 | tee gpurun_out/t_smoke.log
timeout 600 python bench.py --precision tf32 --steps 5 --warmup 3 --no-cpu-baseline --profile > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; tail -c 2500 gpurun_out/bench_tf32.json; head -50 gpurun_out/bench_tf32.err
