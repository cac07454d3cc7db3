#!/bin/bash
# conv engine iteration: conv + model parity tests, bench with per-kernel table.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "conv or models or dropin or engine or hr_keyframe or full_size" 2>&1 | tail -6 | tee gpurun_out/t_conv.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_tf32.json').read().strip().splitlines()[-1])
print("fps", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "conv", d["roofline_conv"]["achieved"], d["roofline_conv"]["ms_per_step"], "creff ms", d["roofline_creff"]["ms_per_launch"])
PY
head -32 gpurun_out/bench_tf32.err
