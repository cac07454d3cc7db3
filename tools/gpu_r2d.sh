#!/bin/bash
# round-2 session d: full GPU suite, smoke, default bench (tcgen05 CReFF default, pre-pass hoisted), A/B against the march engine
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; tail -3 gpurun_out/bench_r2d.err
ARSEG_CREFF_TC=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline --alt-precision none > gpurun_out/bench_r2d_march.json 2> gpurun_out/bench_r2d_march.err
ARSEG_PLAN_OVERLAP=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline --alt-precision none > gpurun_out/bench_r2d_serial.json 2> gpurun_out/bench_r2d_serial.err
python - <<'PY'
import json
for f in ("bench_r2d", "bench_r2d_march", "bench_r2d_serial"):
    try:
        d = json.loads([l for l in open("gpurun_out/%s.json" % f).read().splitlines() if l.startswith("{")][-1])
        print(f, "value %.1f fps, %.3f ms/step, e2e %.1f, launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches")))
        print("   roofline", json.dumps(d.get("roofline"))[:400])
        for k in ("parity", "gop", "dropin_api"):
            if k in d: print("  ", k, json.dumps(d[k])[:300])
    except Exception as e:
        print(f, "failed", e)
PY
