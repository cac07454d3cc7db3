#!/bin/bash
# compute-sanitizer passes over the warp-specialised kernels (small shapes); summaries -> gpurun_out/sanitize_*.txt
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  for k in march tc wide conv; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize.py $k > gpurun_out/sanitize_${tool}_${k}.txt 2>&1
    echo "$tool $k: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_${k}.txt | tail -1)"
  done
done
