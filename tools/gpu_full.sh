#!/bin/bash
# Full GPU parity suite + benches in the given precisions (default: f16 tf32).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/t_all.log
for prec in ${@:-f16 tf32}; do
timeout 600 python bench.py --precision $prec --steps 10 --warmup 3 --no-cpu-baseline --profile > gpurun_out/bench_$prec.json 2> gpurun_out/bench_$prec.err; python - $prec <<'PY'
import json,sys
prec=sys.argv[1]
d=json.loads(open('gpurun_out/bench_%s.json'%prec).read().strip().splitlines()[-1])
print(prec, "fps", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "conv TF/s", d["roofline_conv"]["achieved"], "conv ms", d["roofline_conv"]["ms_per_step"], "creff ms", d["roofline_creff"]["ms_per_launch"])
PY
head -24 gpurun_out/bench_$prec.err
done
