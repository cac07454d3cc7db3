#!/bin/bash
# round-2 follow-up session: full suite at HEAD, sanitizer passes over the re-built tcgen05 CReFF kernels (shipped build: memcheck,
# synccheck; -DARSEG_ARRIVE_ALL build: racecheck), conv halo ncu capture, config-5 sweep (both C = 64 engines), configs 2-4 table
mkdir -p gpurun_out
B="--no-cpu-baseline --alt-precision none --no-extras"
echo "=== gpu tests"
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/t_all.log
echo "=== sanitizer (tc)"
for tool in memcheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize.py tc > gpurun_out/sanitize_${tool}_tc.txt 2>&1
  echo "$tool tc: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_tc.txt | tail -1)"
done
echo "=== ncu full conv up_1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_halo_kernel -s 14 -c 1 -f -o gpurun_out/conv_halo_f16 python bench.py --steps 1 --warmup 1 $B 2>&1 | tail -1
echo "=== sweep"
timeout 900 python tools/sweep_creff.py --iters 3 2>&1 | tail -30
echo "=== configs"
timeout 600 python tools/bench_configs.py 2>&1 | tail -4
echo "=== racecheck (tc, every lane arrives)"
ARSEG_NVCC_EXTRA=-DARSEG_ARRIVE_ALL python -m arseg_b200.build > /dev/null 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitize.py tc > gpurun_out/sanitize_racecheck_tc_arriveall.txt 2>&1
echo "racecheck tc (arrive-all build): $(grep -E 'RACECHECK SUMMARY' gpurun_out/sanitize_racecheck_tc_arriveall.txt | tail -1)"
