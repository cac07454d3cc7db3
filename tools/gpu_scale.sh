#!/bin/bash
# N-GPU pass (run under `gpurun --gpus N`): correctness of frame-level sharding, then bench.py in both shard modes.
N=${1:-2}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $T tools/dist_check.py 2>&1 | grep dist_check
timeout 300 $T bench.py --gpus $N --steps 10 --warmup 3 --alt-precision none > gpurun_out/bench_n${N}_gop.json 2> gpurun_out/bench_n${N}_gop.err
timeout 300 $T bench.py --gpus $N --steps 10 --warmup 3 --shard frame --alt-precision none > gpurun_out/bench_n${N}_frame.json 2> gpurun_out/bench_n${N}_frame.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for m in ("gop", "frame"):
    try:
        d = json.loads(open("gpurun_out/bench_n%s_%s.json" % (n, m)).read().strip().splitlines()[-1])
        print("N=%s shard=%s: %.1f fps resident, %.1f fps e2e, %.3f ms/step, scaling %s" % (n, m, d["value"], d["e2e"]["value"], d["ms_per_step"], d["scaling"]))
    except Exception as e:
        print("N=%s shard=%s: failed (%s)" % (n, m, e))
PY
