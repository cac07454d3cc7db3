#!/bin/bash
# First GPU bring-up: parity tests (exact kernels first, tcgen05 in separate processes), smoke, first bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
echo "=== exact (fp32) tests" 
timeout 1200 python -m pytest tests -m gpu -q -k "not tcgen05 and not tf32 and not bf16" 2>&1 | tail -40 | tee gpurun_out/t1_exact.log
echo "=== tcgen05 conv tests"
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "tcgen05" 2>&1 | tail -40 | tee gpurun_out/t2_tc.log
echo "=== tf32/bf16 model tests"
timeout 900 python -m pytest tests -m gpu -q -k "(tf32 or bf16) and not tcgen05" 2>&1 | tail -40 | tee gpurun_out/t3_prec.log
echo "=== smoke"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -8 | tee gpurun_out/t4_smoke.log
echo "=== bench fp32"
timeout 600 python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline --profile > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; tail -c 1500 gpurun_out/bench_fp32.json; head -30 gpurun_out/bench_fp32.err
echo "=== bench tf32"
timeout 600 python bench.py --precision tf32 --steps 5 --warmup 3 --no-cpu-baseline --profile > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; tail -c 1500 gpurun_out/bench_tf32.json; head -30 gpurun_out/bench_tf32.err
