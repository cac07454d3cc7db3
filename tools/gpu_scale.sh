#!/bin/bash
# N-GPU pass (run under `gpurun --gpus N`): correctness of frame-level sharding (bit-identical to 1 GPU), then bench.py
# (weak GOP-sharded value + the frame-sharded `strong` block in one run); $2 = workload (default camvid-psp18), $3 = nocheck skips the sharding check
N=${1:-2}
WL=${2:-camvid-psp18}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
[ "$3" = nocheck ] || timeout 300 $T tools/dist_check.py 2>&1 | grep dist_check | tee gpurun_out/dist_check_n${N}.log
timeout 600 $T bench.py --gpus $N --steps 10 --warmup 3 --alt-precision none --workload $WL > gpurun_out/bench_n${N}_${WL}.json 2> gpurun_out/bench_n${N}_${WL}.err
python - $N $WL <<'PY'
import json, sys
n, wl = sys.argv[1], sys.argv[2]
try:
    d = json.loads([l for l in open("gpurun_out/bench_n%s_%s.json" % (n, wl)).read().splitlines() if l.startswith("{")][-1])
    print("N=%s %s weak: %.1f fps resident, %.1f fps e2e, %.3f ms/step" % (n, wl, d["value"], d["e2e"]["value"], d["ms_per_step"]))
    s = d.get("strong")
    if s:
        print("N=%s %s strong: %.1f fps, %.3f ms/GOP, broadcast alone %.3f ms, frames/rank %s, %.2fx of one GPU (ceiling %.2fx, %.0f %% of it)" %
              (n, wl, s["value"], s["ms_per_gop"], s["ms_broadcast_alone"], s["frames_per_rank"], s["speedup_vs_one_gpu"], s["ceiling_speedup"],
               100 * s["frac_of_ceiling"]))
except Exception as e:
    print("N=%s: failed (%s)" % (n, e))
PY
