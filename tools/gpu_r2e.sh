#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -m gpu -x -q -k "tc or hoisted or march_engine" 2>&1 | tail -3
for mode in 1 0; do
ARSEG_PLAN_OVERLAP=$mode timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline --alt-precision none --profile > gpurun_out/bench_r2e_$mode.json 2> gpurun_out/bench_r2e_$mode.err
python - $mode <<'PY'
import json, sys
m = sys.argv[1]
d = json.loads([l for l in open("gpurun_out/bench_r2e_%s.json" % m).read().splitlines() if l.startswith("{")][-1])
print("overlap=%s value %.1f fps, %.3f ms/step, e2e %.1f" % (m, d["value"], d["ms_per_step"], d["e2e"]["value"]))
PY
grep "creff" gpurun_out/bench_r2e_$mode.err | head -3
done
