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
    """SM clock / throttle reasons sampled DURING the timed regions (NVML in-process, every 5 ms; nvidia-smi fallback)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.sm, self.bits, self.sm_max, self.stop_flag, self.how = index, [], 0, None, False, "nvml"

    def _visible_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip() != ""]
            if self.index < len(ids) and ids[self.index].strip().isdigit():
                return int(ids[self.index])
        return self.index

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.bits |= int(reasons(h))
                time.sleep(0.005)
            return
        except Exception:
            self.how = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = [0x8, 0x40, 0x20, 0x4]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self._visible_index()), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                r = [c.strip() for c in out.split(",")]
                self.sm.append(float(r[0]))
                self.sm_max = float(r[1])
                for i, b in enumerate(names):
                    if r[2 + i].lower().startswith("active"):
                        self.bits |= b
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unsampled"], "samples": 0}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.sm_max,
                "reasons": [n for b, n in self.REASONS.items() if self.bits & b], "samples": len(sm), "how": self.how}


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
    ap.add_argument("--precision", default=os.environ.get("ARSEG_PRECISION", "f16"), choices=["fp32", "tf32", "f16", "bf16"])
    ap.add_argument("--alt-precision", default="tf32", choices=["none", "fp32", "tf32", "f16", "bf16"],
                    help="second precision mode measured on the same inputs and reported under 'alt_precision'")
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
    # two pinned input sets (consecutive GOPs come from different host buffers) and two pinned result buffers
    pins = [(frames.pin_memory(), mvs.pin_memory(), torch.empty((n_local, H, W), dtype=torch.uint8).pin_memory()),
            (frames.flip(0).contiguous().pin_memory(), mvs.flip(0).contiguous().pin_memory(),
             torch.empty((n_local, H, W), dtype=torch.uint8).pin_memory())]
    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    total_frames = (N_FRAMES if args.shard == "frame" else N_FRAMES * world) * args.steps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def measure(precision):
        """Device-resident throughput, end-to-end throughput (pipelined and serial) and the per-kernel table."""
        frame_mode = args.shard == "frame" and world > 1
        eng = ev.NonKeyEngine(ARCH, sd, n_local, H, W, SCALE, precision, K_WIN, device=dev, want_logits=True, split_keyframe=frame_mode)
        eng.set_inputs(frames.to(dev), mvs.to(dev), ref_p.to(dev))
        bc_stream = torch.cuda.Stream(dev) if frame_mode else None

        def bcast_ref():
            if frame_mode:
                dist.broadcast(eng.ref_p, src=0)     # keyframe feature, once per GOP (ncclBroadcast over NVLink)

        def dev_step():
            if not frame_mode:
                eng.step()
                return
            # frame-level sharding: the broadcast of the keyframe feature (177 MB, once per GOP) runs on a side stream and
            # overlaps phase 1, which needs only the frames; the CReFF launches wait for it
            main = torch.cuda.current_stream()
            bc_stream.wait_stream(main)              # the previous GOP's CReFF has finished reading the feature
            with torch.cuda.stream(bc_stream):
                dist.broadcast(eng.ref_p, src=0)
            eng.step_phase1()
            main.wait_stream(bc_stream)
            eng.step_phase2()

        def timed_steps(fn, steps):
            """K steps, L2 flushed between them (outside the per-step event brackets); sum of step times, max over ranks."""
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
            return max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))

        def timed_stream(steps):
            """K end-to-end steps through HostPipeline: one event pair around the whole run (copies of step i+1 overlap
            the compute of step i, so per-step brackets would not mean anything); the L2 flush stays inside."""
            pipe = eng.host_pipeline()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                l2_flush.zero_()
                bcast_ref()
                pipe.submit(*pins[i & 1])
            pipe.drain()
            e1.record()
            barrier()
            return max_over_ranks(e0.elapsed_time(e1))

        for _ in range(max(args.warmup, 3)):
            dev_step()
        ms = timed_steps(dev_step, args.steps)
        timed_stream(max(args.warmup, 3))
        ms_e2e = timed_stream(args.steps)
        for _ in range(2):
            eng.step_host(*pins[0])
        ms_serial = timed_steps(lambda: (bcast_ref(), eng.step_host(*pins[0])), args.steps)
        prof = eng.plan.profile(iters=3, warmup=1) if rank == 0 else []
        barrier()
        return {"eng": eng, "ms": ms, "ms_e2e": ms_e2e, "ms_serial": ms_serial, "prof": prof,
                "value": total_frames / (ms / 1e3), "e2e": total_frames / (ms_e2e / 1e3), "e2e_serial": total_frames / (ms_serial / 1e3)}

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    main_r = measure(args.precision)
    if rank == 0:
        sampler.stop_flag = True
    alt_r = None
    if args.alt_precision not in ("none", args.precision):
        eng0 = main_r.pop("eng")
        main_r["launches_per_step"], main_r["conv_flops"] = eng0.launches_per_step, eng0.plan.conv_flops
        del eng0
        torch.cuda.empty_cache()
        alt_r = measure(args.alt_precision)
        alt_r.pop("eng")
    else:
        eng0 = main_r.pop("eng")
        main_r["launches_per_step"], main_r["conv_flops"] = eng0.launches_per_step, eng0.plan.conv_flops
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, tc_burst, tc_sust, peak_src = peaks()
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")      # dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full)
    if os.path.exists(tp):
        traffic = json.load(open(tp))

    def rooflines(r, precision):
        prof = r["prof"]
        t_all = sum(t for _, t in prof)
        creff_ms = sum(t for n, t in prof if n.startswith("creff"))
        tc_ms = sum(t for n, t in prof if "[tf32" in n or "[bf16" in n or "[f16" in n)
        simt_ms = sum(t for n, t in prof if "[simt" in n)
        lr_bytes = 2 if precision in ("bf16", "f16") else 4
        cbytes = creff_bytes(lr_bytes, write_p=False, write_logits=True) * n_local
        creff_gbs = cbytes / (creff_ms * 1e-3) / 1e9 if creff_ms > 0 else 0.0
        creff_gbs_full = CREFF_BYTES_FULL * n_local / (creff_ms * 1e-3) / 1e9 if creff_ms > 0 else 0.0
        conv_tflops = r["conv_flops"] / ((tc_ms + simt_ms) * 1e-3) / 1e12 if (tc_ms + simt_ms) > 0 else 0.0
        tr = traffic.get("creff_" + precision)
        roof_creff = {"kernel": "creff_march_kernel (MV warp + CReFF + classifier + log-softmax + argmax, one launch per step)",
                      "bound": "hbm", "achieved": round(creff_gbs, 1),
                      "peak": hbm_peak, "unit": "GB/s", "frac": round(creff_gbs / hbm_peak, 4),
                      "traffic": int(tr * n_local / N_FRAMES) if tr else None,
                      "share_of_step": round(creff_ms / t_all, 3) if t_all else None, "ms_per_launch": round(creff_ms, 4),
                      "algorithmic_bytes_per_launch": cbytes, "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                      "achieved_survey_8d_bytes": round(creff_gbs_full, 1), "frac_survey_8d_bytes": round(creff_gbs_full / hbm_peak, 4),
                      "note": "achieved = bytes the launch must move (hr fp32 + lr + int16 MV in; fp32 log-probs + u8 class map out) / time; "
                              "achieved_survey_8d_bytes also counts the 176.9 MB/frame fused-p write of SURVEY 8(d), which "
                              "evaluation.py:193 discards and the engine does not materialise"}
        roof_conv = {"kernel": "conv_tc kernels (tcgen05 implicit GEMM, all conv layers of phase 1)", "bound": "tensor",
                     "achieved": round(conv_tflops, 2), "peak": tc_sust, "unit": "TFLOP/s", "frac": round(conv_tflops / tc_sust, 4),
                     "traffic": None, "share_of_step": round((tc_ms + simt_ms) / t_all, 3) if t_all else None,
                     "ms_per_step": round(tc_ms + simt_ms, 4), "algorithmic_flops_per_step": r["conv_flops"],
                     "peak_source": peak_src + " (bf16_tflops_sustained; kernels timed inside a long step)"}
        return roof_creff, roof_conv, (roof_conv if (tc_ms + simt_ms) >= creff_ms else roof_creff)

    roof_creff, roof_conv, dominant = rooflines(main_r, args.precision)
    if args.profile:
        for n, t in sorted(main_r["prof"], key=lambda x: -x[1]):
            sys.stderr.write("%9.4f ms  %s\n" % (t, n))
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(2, 1, budget_s=40.0)
        cpu_base = {"value": round(r["fps"], 4), "unit": UNIT, "cores": r["cores"], "kind": "port",
                    "sample": "1 of the 11 non-keyframes (720x960) per run, best of %d after 1 warm-up; stages(s) %s" %
                              (r["steps_run"], {k: round(v, 3) for k, v in r["stages"].items()})}
    dnames = {"fp32": "f32", "tf32": "tf32 (fp32 storage, fp32 accumulate)",
              "f16": "f16 (fp16 activation/weight storage: 11-bit significand like TF32; fp32 accumulate, fp32 CReFF inputs/softmax/outputs)",
              "bf16": "bf16 (fp32 accumulate)"}
    h2d = int(pins[0][0].numel() * 4 + pins[0][1].numel() * 2)
    line = {
        "metric": METRIC, "value": round(main_r["value"], 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(main_r["ms"] / args.steps, 4), "higher_is_better": True,
        "scaling": "strong" if args.shard == "frame" else "weak", "vs_baseline": None,
        "dtype": dnames[args.precision],
        "data": "synthetic (seeded randn frames, block-constant int16 quarter-pel MV fields, name-keyed random weights)",
        "config": {"workload": "CamVid 720x960 GOP-12 PSPNet-18 AR-0.5x, k=7: 11 non-keyframes per step per rank (BASELINE configs[1])",
                   "frames_per_step_per_rank": n_local, "shard": args.shard, "precision": args.precision,
                   "l2": "256 MiB buffer written between timed steps (L2 flush); per-step inputs 122 MB",
                   "outputs": "log-prob maps [11,12,720,960] fp32 + argmax class maps uint8; fused p is not materialised "
                              "(evaluation.py:193 discards it)"},
        "e2e": {"value": round(main_r["e2e"], 2), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": int(pins[0][2].numel()), "ms_per_step": round(main_r["ms_e2e"] / args.steps, 4),
                "how": "NonKeyEngine.host_pipeline(): every step copies its frames + int16 MV fields from pinned host memory and "
                       "its class maps back; the copies of step i+1 overlap the compute of step i (one event pair around all K steps, "
                       "L2 flush inside)",
                "serial_value": round(main_r["e2e_serial"], 2), "serial_ms_per_step": round(main_r["ms_serial"] / args.steps, 4)},
        "gpu_launches": main_r["launches_per_step"] * args.steps,
        "launches_per_step": main_r["launches_per_step"],
        "clocks": sampler.summary(),
        "roofline": dominant, "roofline_creff": roof_creff, "roofline_conv": roof_conv,
        "cpu_baseline": cpu_base,
    }
    if alt_r is not None:
        a_creff, a_conv, _ = rooflines(dict(alt_r, conv_flops=main_r["conv_flops"]), args.alt_precision)
        line["alt_precision"] = {"precision": args.alt_precision, "dtype": dnames[args.alt_precision], "value": round(alt_r["value"], 2),
                                 "ms_per_step": round(alt_r["ms"] / args.steps, 4), "e2e": round(alt_r["e2e"], 2),
                                 "creff_ms": a_creff["ms_per_launch"], "conv_tflops": a_conv["achieved"], "conv_frac": a_conv["frac"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
