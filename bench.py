#!/usr/bin/env python
"""bench.py -- non-keyframe frames/sec @720x960 GOP-12 (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1      # the reference's CPU path (oracle port) on host cores

A "step" is one pass of the hot path over one GOP's 11 non-keyframes (CamVid 720x960, PSPNet-18, AR-0.5x,
k=7; BASELINE.json configs[1]) on every rank: frame down-scale -> LR-branch PSPNet-18 -> MV warp + CReFF +
classifier + argmax, keyframe feature resident on the device.  N>1: GOPs are sharded over ranks (each rank
owns whole GOPs, no data-path collective; weak scaling); `--shard frame` instead deals the 11 frames of ONE
GOP over the ranks after an NCCL broadcast of the keyframe feature (strong scaling, north_star's latency mode).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ARCH, H, W, SCALE, K_WIN, GOP = "camvid-psp18", 720, 960, 0.5, 7, 12
N_FRAMES = GOP - 1
METRIC = "non-keyframe frames/sec @720x960 GOP12 (CamVid PSPNet-18 AR-0.5x)"
UNIT = "frames/s"
# algorithmic bytes of the warp+CReFF+classifier kernel per frame (fp32 API-preserving minimum)
#   read HR p 64*720*960*4 + read LR p 64*360*480*4 + read MV int16 720*960*2*2 + write fused p + write logits 12*720*960*4
CREFF_BYTES_FULL = 64 * 720 * 960 * 4 * 2 + 64 * 360 * 480 * 4 + 720 * 960 * 4 + 12 * 720 * 960 * 4


def creff_bytes(lr_elem_bytes: int, write_p: bool, write_logits: bool) -> int:
    """Bytes the kernel must move per frame in the configuration the engine actually runs."""
    b = 64 * 720 * 960 * 4 + 64 * 360 * 480 * lr_elem_bytes + 720 * 960 * 4 + 720 * 960  # hr, lr, mv, argmax u8
    if write_p:
        b += 64 * 720 * 960 * 4
    if write_logits:
        b += 12 * 720 * 960 * 4
    return b


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows if len(r) > 3 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def make_inputs(rank: int):
    import numpy as np
    import torch
    from arseg_b200 import synth
    frames = torch.cat([synth.synth_frame(1, H, W, 1000 * rank + i) for i in range(N_FRAMES)])
    mvs = torch.from_numpy(np.stack([synth.synth_mv_int16(H, W, 2000 * rank + d, distance=d) for d in range(1, GOP)]))
    ref_p = synth.synth_feature(1, 64, H, W, 3000 + rank) * 0.5
    return frames, mvs, ref_p


def cpu_reference_run(steps: int, warmup: int, budget_s: float = 200.0):
    """The reference's own CPU path for this workload: oracle port (torch-CPU ops + plain-C localAttention) on all
    host cores, one 720x960 non-keyframe per step (a bounded sample of the 11-frame GOP step)."""
    import torch
    from arseg_b200 import models, synth
    from oracle import arseg_oracle as O
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.synth_state_dict(models.models_fuse[ARCH]().state_dict(), 4)
    frames, mvs, ref_p = make_inputs(0)
    times, stages = [], {}
    t_first = None
    done = 0
    for i in range(warmup + steps):
        tm = {}
        t0 = time.perf_counter()
        O.nonkey_step(ARCH, sd, frames[i % N_FRAMES:i % N_FRAMES + 1], ref_p, synth.mv_to_flow(mvs[i % N_FRAMES].numpy()),
                      SCALE, K_WIN, use_c=True, timings=tm)
        dt = time.perf_counter() - t0
        if t_first is None:
            t_first = dt
        if i >= warmup:
            times.append(dt)
            stages = tm
            done += 1
            if sum(times) + t_first * warmup > budget_s:
                break
    best = min(times)
    return {"fps": 1.0 / best, "s_per_frame": best, "steps_run": done, "cores": cores, "stages": stages}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("ARSEG_PRECISION", "tf32"), choices=["fp32", "tf32", "f16", "bf16"])
    ap.add_argument("--shard", default="gop", choices=["gop", "frame"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="also print the per-kernel time table to stderr")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(args.steps, max(args.warmup, 1))
        line = {"impl": "reference", "metric": METRIC, "value": r["fps"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": r["steps_run"], "warmup": max(args.warmup, 1), "ms_per_step": r["s_per_frame"] * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "CamVid 720x960 GOP-12 PSPNet-18 AR-0.5x k=7, non-keyframe path", "frames_per_step": 1},
                "cpu_baseline": {"value": r["fps"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                 "sample": "1 non-keyframe (720x960) per step, best of %d; stages %s" %
                                           (r["steps_run"], {k: round(v, 3) for k, v in r["stages"].items()})},
                "e2e": {"value": r["fps"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from arseg_b200 import evaluation as ev
    from arseg_b200 import models, synth

    torch.set_grad_enabled(False)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sd = synth.synth_state_dict(models.models_fuse[ARCH]().state_dict(), 4)
    frames, mvs, ref_p = make_inputs(rank)
    if args.shard == "frame" and world > 1:
        from arseg_b200 import dist as adist
        my = adist.frames_of_rank(N_FRAMES, world, rank)
        n_local = len(my)
        frames, mvs = frames[my], mvs[my]
    else:
        n_local = N_FRAMES
    eng = ev.NonKeyEngine(ARCH, sd, n_local, H, W, SCALE, args.precision, K_WIN, device=dev, want_logits=True)
    eng.set_inputs(frames.to(dev), mvs.to(dev), ref_p.to(dev))
    pin_f, pin_m = frames.pin_memory(), mvs.pin_memory()
    pin_o = torch.empty((n_local, H, W), dtype=torch.uint8).pin_memory()
    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def bcast_ref():
        if args.shard == "frame" and world > 1:
            dist.broadcast(eng.ref_p, src=0)     # keyframe feature, once per GOP (ncclBroadcast over NVLink)

    def timed(fn, steps):
        """K steps, L2 flushed between them; returns the sum of per-step device times (ms), max over ranks."""
        evs = []
        barrier()
        for _ in range(steps):
            l2_flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def dev_step():
        bcast_ref()
        eng.step()

    def e2e_step():
        bcast_ref()
        eng.step_host(pin_f, pin_m, pin_o)

    for _ in range(max(args.warmup, 3)):
        dev_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(dev_step, args.steps)
    if rank == 0:
        sampler.stop_flag = True
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    total_frames = (N_FRAMES if args.shard == "frame" else N_FRAMES * world) * args.steps
    value = total_frames / (ms / 1e3)
    e2e_value = total_frames / (ms_e2e / 1e3)

    # per-kernel device times (CUDA events on the launching stream) for the roofline objects
    prof = eng.plan.profile(iters=3, warmup=1) if rank == 0 else []
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm_peak, tc_burst, tc_sust, peak_src = peaks()
    t_all = sum(t for _, t in prof)
    creff_ms = sum(t for n, t in prof if n.startswith("creff"))
    tc_ms = sum(t for n, t in prof if "[tf32" in n or "[bf16" in n or "[f16" in n)
    simt_ms = sum(t for n, t in prof if "[simt" in n)
    lr_bytes = 2 if args.precision in ("bf16", "f16") else 4
    cbytes = creff_bytes(lr_bytes, write_p=False, write_logits=True) * n_local
    creff_gbs = cbytes / (creff_ms * 1e-3) / 1e9 if creff_ms > 0 else 0.0
    creff_gbs_full = CREFF_BYTES_FULL * n_local / (creff_ms * 1e-3) / 1e9 if creff_ms > 0 else 0.0
    conv_tflops = eng.plan.conv_flops / ((tc_ms + simt_ms) * 1e-3) / 1e12 if (tc_ms + simt_ms) > 0 else 0.0
    roof_creff = {"kernel": "creff_fused (warp+CReFF+classifier+argmax)", "bound": "hbm", "achieved": round(creff_gbs, 1),
                  "peak": hbm_peak, "unit": "GB/s", "frac": round(creff_gbs / hbm_peak, 4), "traffic": None,
                  "share_of_step": round(creff_ms / t_all, 3) if t_all else None, "ms_per_launch": round(creff_ms, 4),
                  "algorithmic_bytes_per_launch": cbytes, "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                  "achieved_if_fp32_p_and_logits_were_written": round(creff_gbs_full, 1)}
    roof_conv = {"kernel": "implicit-GEMM convs (tcgen05 + SIMT stride-2 layers), all layers", "bound": "tensor",
                 "achieved": round(conv_tflops, 2), "peak": tc_sust, "unit": "TFLOP/s", "frac": round(conv_tflops / tc_sust, 4),
                 "traffic": None, "share_of_step": round((tc_ms + simt_ms) / t_all, 3) if t_all else None,
                 "ms_per_step": round(tc_ms + simt_ms, 4), "algorithmic_flops_per_step": eng.plan.conv_flops,
                 "peak_source": peak_src + " (bf16_tflops_sustained; kernels timed inside a long step)"}
    dominant = roof_conv if (tc_ms + simt_ms) >= creff_ms else roof_creff
    if args.profile:
        for n, t in sorted(prof, key=lambda x: -x[1]):
            sys.stderr.write("%9.4f ms  %s\n" % (t, n))
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(2, 1, budget_s=40.0)
        cpu_base = {"value": round(r["fps"], 4), "unit": UNIT, "cores": r["cores"], "kind": "port",
                    "sample": "1 of the 11 non-keyframes (720x960) per run, best of %d after 1 warm-up; stages(s) %s" %
                              (r["steps_run"], {k: round(v, 3) for k, v in r["stages"].items()})}
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
        "scaling": "strong" if args.shard == "frame" else "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "tf32": "tf32 (fp32 storage, fp32 accumulate)", "f16": "f16 (fp16 storage: 11-bit significand as TF32, fp32 accumulate)", "bf16": "bf16 (fp32 accumulate)"}[args.precision],
        "data": "synthetic (seeded randn frames, block-constant int16 quarter-pel MV fields, name-keyed random weights)",
        "config": {"workload": "CamVid 720x960 GOP-12 PSPNet-18 AR-0.5x, k=7: 11 non-keyframes per step per rank (BASELINE configs[1])",
                   "frames_per_step_per_rank": n_local, "shard": args.shard, "precision": args.precision,
                   "l2": "256 MiB buffer written between timed steps (L2 flush); per-step inputs 122 MB",
                   "outputs": "log-prob maps [11,12,720,960] fp32 + argmax class maps uint8; fused p is not materialised "
                              "(evaluation.py:193 discards it)"},
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(pin_f.numel() * 4 + pin_m.numel() * 2),
                "d2h_bytes_per_step": int(pin_o.numel()), "ms_per_step": round(ms_e2e / args.steps, 4)},
        "gpu_launches": eng.launches_per_step * args.steps * 2,
        "launches_per_step": eng.launches_per_step,
        "clocks": sampler.summary(),
        "roofline": dominant, "roofline_creff": roof_creff, "roofline_conv": roof_conv,
        "cpu_baseline": cpu_base,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
