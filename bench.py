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

SCALE, K_WIN, GOP = 0.5, 7, 12
N_FRAMES = GOP - 1
UNIT = "frames/s"
# --workload: BASELINE.json configs[1] (default, the configuration the metric is quoted on), configs[2], configs[3]
WORKLOADS = {
    # name: (arch, H, W, C of p, stride of p, n_classes, label)
    "camvid-psp18": ("camvid-psp18", 720, 960, 64, 1, 12, "CamVid 720x960 GOP-12 PSPNet-18 AR-0.5x"),
    "camvid-bise18": ("camvid-bise18", 720, 960, 256, 8, 12, "CamVid 720x960 GOP-12 BiSeNet-18 AR-0.5x"),
    "cityscapes-psp18": ("cityscapes-psp18", 1024, 2048, 512, 8, 19, "Cityscapes 1024x2048 GOP-12 PSPNet-18 AR-0.5x"),
}
ARCH, H, W, C_P, STRIDE_P, N_CLS, WL_LABEL = WORKLOADS["camvid-psp18"]


def set_workload(name: str) -> None:
    global ARCH, H, W, C_P, STRIDE_P, N_CLS, WL_LABEL
    ARCH, H, W, C_P, STRIDE_P, N_CLS, WL_LABEL = WORKLOADS[name]


def metric_name() -> str:
    dataset = WL_LABEL.split(" ")[0]                              # "CamVid" / "Cityscapes"
    return "non-keyframe frames/sec @%dx%d GOP12 (%s %s)" % (H, W, dataset, WL_LABEL.split(" GOP-12 ")[1])


def creff_bytes_survey_8d(lr_numel_per_frame: int, logits_numel_per_frame: int) -> int:
    """algorithmic bytes of the warp + CReFF + classifier step PER FRAME, the fp32 API-preserving minimum:
    read HR p + read LR p + read MV (int16, frame resolution) + write fused p + write logits (all fp32).
    CamVid-PSP: 176.9 + 44.2 + 2.8 + 176.9 + 33.2 = 434.0 MB."""
    hf, wf = H // STRIDE_P, W // STRIDE_P
    return C_P * hf * wf * 4 * 2 + lr_numel_per_frame * 4 + H * W * 4 + logits_numel_per_frame * 4


def creff_bytes_moved(n_frames: int, hr_elem: int, lr_numel_per_frame: int, lr_elem: int, logits_numel_per_frame: int,
                      write_p: bool) -> int:
    """Bytes ONE launch has to move in the configuration the engine runs: the keyframe feature ONCE (the frames of a GOP share
    it and it stays in L2 across them), and per frame the LR feature, the int16 MV field, the fp32 logits and the uint8 class map."""
    hf, wf = H // STRIDE_P, W // STRIDE_P
    per_frame = lr_numel_per_frame * lr_elem + H * W * 4 + logits_numel_per_frame * 4 + H * W
    if write_p:
        per_frame += C_P * hf * wf * 4
    return C_P * hf * wf * hr_elem + n_frames * per_frame


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
    """One GOP of synthetic input: 11 decoded frames (uint8 HWC, what a video decoder hands over), their int16 quarter-pel MV
    fields at keyframe distance 1..11 (dataset/camvid.py:624-626 wire format), the keyframe itself and its feature."""
    import numpy as np
    import torch
    from arseg_b200 import synth
    g = torch.Generator().manual_seed(1000 * rank + 17)
    frames_u8 = torch.randint(0, 256, (N_FRAMES, H, W, 3), generator=g, dtype=torch.uint8)
    key_u8 = torch.randint(0, 256, (1, H, W, 3), generator=g, dtype=torch.uint8)
    mvs = torch.from_numpy(np.stack([synth.synth_mv_int16(H, W, 2000 * rank + d, distance=d) for d in range(1, GOP)]))
    ref_p = synth.synth_feature(1, C_P, H // STRIDE_P, W // STRIDE_P, 3000 + rank) * 0.5
    return frames_u8, key_u8, mvs, ref_p


def normalise(frames_u8):
    """transforms.ToTensor + Normalize with the dataset constants (dataset/camvid.py:182-185): what the reference's DataLoader yields."""
    from arseg_b200 import ops
    from oracle import arseg_oracle as O
    return O.ingest_u8(frames_u8, ops.CAMVID_MEAN, ops.CAMVID_STD, (H, W))


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
    frames_u8, _, mvs, ref_p = make_inputs(0)
    frames = normalise(frames_u8)
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
    ap.add_argument("--workload", default="camvid-psp18", choices=list(WORKLOADS))
    ap.add_argument("--precision", default=os.environ.get("ARSEG_PRECISION", "f16"), choices=["fp32", "tf32", "f16", "bf16"])
    ap.add_argument("--alt-precision", default="tf32", choices=["none", "fp32", "tf32", "f16", "bf16"],
                    help="second precision mode measured on the same inputs and reported under 'alt_precision'")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the frame-sharded (one GOP over all ranks) measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the parity / whole-GOP / drop-in API measurements")
    ap.add_argument("--profile", action="store_true", help="also print the per-kernel time table to stderr")
    args = ap.parse_args()
    set_workload(args.workload)
    METRIC = metric_name()
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
                "config": {"workload": WL_LABEL + " k=7, non-keyframe path", "frames_per_step": 1},
                "cpu_baseline": {"value": r["fps"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                 "sample": "1 non-keyframe (%dx%d) per step, best of %d; stages %s" %
                                           (H, W, r["steps_run"], {k: round(v, 3) for k, v in r["stages"].items()})},
                "e2e": {"value": r["fps"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # stdout carries exactly ONE line (the JSON): whatever libraries print while initialising (e.g. NCCL's version banner)
    # goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from arseg_b200 import dist as adist
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
    frames_u8, key_u8, mvs, ref_p = make_inputs(rank)
    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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

    def measure(precision, frame_mode):
        """Device-resident throughput, end-to-end throughput (pipelined and serial) and the per-kernel table.
        frame_mode False: every rank owns whole GOPs (weak scaling, no data-path collective).  True: ONE GOP, its 11 frames
        dealt over the ranks after an NCCL broadcast of the keyframe feature (north_star's split; strong scaling)."""
        my = adist.frames_of_rank(N_FRAMES, world, rank) if frame_mode else list(range(N_FRAMES))
        n_local = len(my)
        fr, mv = frames_u8[my].contiguous(), mvs[my].contiguous()
        total_frames = (N_FRAMES if frame_mode else N_FRAMES * world) * args.steps
        # two pinned input sets (consecutive GOPs come from different host buffers) and two pinned result buffers
        pins = [(fr.pin_memory(), mv.pin_memory(), torch.empty((n_local, H, W), dtype=torch.uint8).pin_memory()),
                (fr.flip(0).contiguous().pin_memory(), mv.flip(0).contiguous().pin_memory(),
                 torch.empty((n_local, H, W), dtype=torch.uint8).pin_memory())]
        # the keyframe feature is resident in the engine's internal layout (fp32 NHWC, what the keyframe engine produces): the
        # tensor-core CReFF engines read it in place; the exact fp32 plan keeps the API layout (NCHW)
        ref_buf = None if precision == "fp32" else torch.empty((1, H // STRIDE_P, W // STRIDE_P, C_P), dtype=ev.internal_ref_dtype(ARCH, precision, K_WIN), device=dev)
        eng = ev.NonKeyEngine(ARCH, sd, n_local, H, W, SCALE, precision, K_WIN, device=dev, want_logits=True,
                              split_keyframe=frame_mode, uint8_frames=True, ref_nhwc=ref_buf)
        eng.set_inputs(fr.to(dev), mv.to(dev), ref_p.to(dev))
        ref_dev = eng.ref_nhwc if eng.ref_nhwc is not None else eng.ref_p
        bc_stream = torch.cuda.Stream(dev) if frame_mode else None

        def bcast_ref():
            if frame_mode:
                dist.broadcast(ref_dev, src=0)       # keyframe feature, once per GOP (ncclBroadcast over NVLink)

        def dev_step():
            if not frame_mode:
                eng.step()
                return
            # the broadcast of the keyframe feature runs on a side stream and overlaps phase 1, which needs only the frames;
            # the CReFF launches wait for it
            main = torch.cuda.current_stream()
            bc_stream.wait_stream(main)              # the previous GOP's CReFF has finished reading the feature
            with torch.cuda.stream(bc_stream):
                dist.broadcast(ref_dev, src=0)
            eng.step_phase1()
            main.wait_stream(bc_stream)
            eng.step_phase2()

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
        ms_bcast = None
        if frame_mode:
            for _ in range(3):
                bcast_ref()
            ms_bcast = timed_steps(bcast_ref, args.steps) / args.steps
        prof = list(zip(eng.plan.names, [t for _, t in eng.plan.profile(iters=3, warmup=1)], eng.plan.step_flops)) if rank == 0 else []
        barrier()
        return {"eng": eng, "ms": ms, "ms_e2e": ms_e2e, "ms_serial": ms_serial, "ms_bcast": ms_bcast, "prof": prof, "n_local": n_local,
                "frames": my, "value": total_frames / (ms / 1e3), "e2e": total_frames / (ms_e2e / 1e3),
                "e2e_serial": total_frames / (ms_serial / 1e3),
                "h2d": int(pins[0][0].numel() + pins[0][1].numel() * 2), "d2h": int(pins[0][2].numel())}

    def engine_facts(r):
        eng = r.pop("eng")
        lr = eng.lr_p
        r["launches_per_step"], r["conv_flops"] = eng.launches_per_step, eng.plan.conv_flops
        r["lr_numel"], r["lr_elem"] = lr.numel() // max(1, r["n_local"]), lr.element_size()
        r["logits_numel"] = eng.logits.numel() // max(1, r["n_local"]) if eng.logits is not None else 0
        r["creff_names"] = [n for n in eng.plan.names if n.startswith("creff")]
        r["tc"] = any(n.endswith("_tc") for n in r["creff_names"])
        r["hr_elem"] = eng.ref_nhwc.element_size() if eng.ref_nhwc is not None else 4
        # tcgen05 engine: the MV-warped keyframe rows go through a workspace (written by the pre-pass, read back by the attention kernel)
        r["ws_bytes"] = int(eng.N * (H + 17) * ((W + 15) // 16 * 16 + 8) * 128) if r["tc"] else 0      # k = 7: 6 + 11 border rows, 8 border columns (creff_tc.cu t_hp / t_wp)
        return eng

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    main_r = measure(args.precision, False)
    eng0 = engine_facts(main_r)

    # ---------------- extras on rank 0 of a single-GPU run: parity vs the oracle, whole-GOP rate, the literal drop-in API ----------
    extras = {}
    if world == 1 and not args.no_extras:
        extras = run_extras(args, eng0, sd, frames_u8, key_u8, mvs, ref_p, dev, l2_flush, timed_steps)
    del eng0
    torch.cuda.empty_cache()

    strong_r = None
    if world > 1 and not args.no_strong:
        strong_r = measure(args.precision, True)
        engine_facts(strong_r)
        torch.cuda.empty_cache()
    if rank == 0:
        sampler.stop_flag = True
    alt_r = None
    if args.alt_precision not in ("none", args.precision):
        alt_r = measure(args.alt_precision, False)
        engine_facts(alt_r)
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
        t_all = sum(t for _, t, _ in prof)
        creff_ms = sum(t for n, t, _ in prof if n.startswith("creff"))
        # tensor-core / SIMT implicit-GEMM conv launches only: their own FLOPs over their own time
        conv = [(n, t, f) for n, t, f in prof if "[tf32" in n or "[bf16" in n or "[f16" in n or "[simt" in n]
        conv_ms, conv_flops = sum(t for _, t, _ in conv), sum(f for _, _, f in conv)
        n = r["n_local"]
        cbytes = creff_bytes_moved(n, r["hr_elem"], r["lr_numel"], r["lr_elem"], r["logits_numel"], write_p=False)
        full = creff_bytes_survey_8d(r["lr_numel"], r["logits_numel"]) * n
        creff_gbs = cbytes / (creff_ms * 1e-3) / 1e9 if creff_ms > 0 else 0.0
        creff_gbs_full = full / (creff_ms * 1e-3) / 1e9 if creff_ms > 0 else 0.0
        conv_tflops = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        tr = traffic.get("creff_f16" if r["tc"] else "creff_tf32") if ARCH == "camvid-psp18" else None      # tcgen05 kernels / march kernel
        roof_creff = {"kernel": "%s (MV warp + CReFF + classifier + log-softmax + argmax; %d launch(es) per step)" %
                                ("creff_tc_warp_kernel + creff_tc_kernel [tcgen05]" if r["tc"] else "creff_march_kernel [mma.sync]" if C_P == 64 else "creff_wide kernels [mma.sync]",
                                 len(r["creff_names"])),
                      "bound": "hbm", "achieved": round(creff_gbs, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(creff_gbs / hbm_peak, 4),
                      "traffic": int(tr * n / N_FRAMES) if tr else None,
                      "traffic_source": "static: profiles/ncu_traffic.json (ncu --set full capture of this command, scaled to this launch's frame count)" if tr else None,
                      "share_of_step": round(creff_ms / t_all, 3) if t_all else None, "ms_per_launch": round(creff_ms, 4),
                      "algorithmic_bytes_per_launch": cbytes,
                      "workspace_round_trip_bytes": 2 * r["ws_bytes"],
                      "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                      "achieved_survey_8d_bytes": round(creff_gbs_full, 1), "frac_survey_8d_bytes": round(creff_gbs_full / hbm_peak, 4),
                      "note": "achieved = bytes ONE launch must move / its CUDA-event time: the keyframe feature once (shared by the "
                              "frames of the GOP) + per frame the LR feature (%d B/elem), the int16 MV field, fp32 logits and the u8 class "
                              "map.  achieved_survey_8d_bytes = SURVEY 8(d)'s 434.0 MB/frame figure (counts the keyframe feature per frame "
                              "and the fused-p write that evaluation.py:193 discards and the engine does not materialise).  "
                              "workspace_round_trip_bytes: the tcgen05 engine's pre-pass writes the MV-warped keyframe rows (fp16) to a "
                              "workspace and the attention kernel reads them back -- real DRAM traffic (it is in `traffic`) that the "
                              "algorithm does not need, so it earns no credit in `achieved`" % r["lr_elem"]}
        roof_conv = {"kernel": "conv_tc / conv_simt kernels (implicit-GEMM conv launches of phase 1; stem, pyramid 1x1 and linear layers excluded "
                               "from both FLOPs and time)", "bound": "tensor",
                     "achieved": round(conv_tflops, 2), "peak": tc_burst, "unit": "TFLOP/s", "frac": round(conv_tflops / tc_burst, 4),
                     "frac_of_sustained": round(conv_tflops / tc_sust, 4),
                     "traffic": None, "share_of_step": round(conv_ms / t_all, 3) if t_all else None,
                     "ms_per_step": round(conv_ms, 4), "algorithmic_flops_per_step": conv_flops,
                     "peak_source": peak_src + " (MEASURED_PEAKS.json bf16_tflops, burst: the kernels are timed one by one between CUDA events)"}
        return roof_creff, roof_conv, (roof_conv if conv_ms >= creff_ms else roof_creff)

    roof_creff, roof_conv, dominant = rooflines(main_r, args.precision)
    if args.profile:
        for n, t, _ in sorted(main_r["prof"], key=lambda x: -x[1]):
            sys.stderr.write("%9.4f ms  %s\n" % (t, n))
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(2, 1, budget_s=40.0)
        cpu_base = {"value": round(r["fps"], 4), "unit": UNIT, "cores": r["cores"], "kind": "port",
                    "sample": "1 of the 11 non-keyframes (%dx%d) per run, best of %d after 1 warm-up; stages(s) %s" %
                              (H, W, r["steps_run"], {k: round(v, 3) for k, v in r["stages"].items()})}
    dnames = {"fp32": "f32", "tf32": "tf32 (fp32 storage, fp32 accumulate)",
              "f16": "f16 (fp16 activation / weight / feature storage: 11-bit significand like TF32; fp32 accumulate, softmax and outputs)",
              "bf16": "bf16 (fp32 accumulate)"}
    line = {
        "metric": METRIC, "value": round(main_r["value"], 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(main_r["ms"] / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": dnames[args.precision],
        "data": "synthetic (seeded uint8 frames, block-constant int16 quarter-pel MV fields, name-keyed random weights)",
        "config": {"workload": WL_LABEL + ", k=7: 11 non-keyframes per step per rank" + (" (BASELINE configs[1])" if ARCH == "camvid-psp18" else ""),
                   "frames_per_step_per_rank": main_r["n_local"], "shard": "gop (every rank owns whole GOPs; no data-path collective)",
                   "precision": args.precision,
                   "l2": "256 MiB buffer written between timed steps (L2 flush)",
                   "inputs": "uint8 HWC frames (ToTensor + Normalize fused into the LR down-scale kernel), int16 MV fields, keyframe feature "
                             "resident in the engine's internal layout (NHWC, %s, as the keyframe engine of this plan writes it)" % str(ev.internal_ref_dtype(ARCH, args.precision, K_WIN)).replace("torch.", ""),
                   "outputs": "log-prob maps fp32 + argmax class maps uint8; fused p is not materialised (evaluation.py:193 discards it)"},
        "e2e": {"value": round(main_r["e2e"], 2), "unit": UNIT, "h2d_bytes_per_step": main_r["h2d"],
                "d2h_bytes_per_step": main_r["d2h"], "ms_per_step": round(main_r["ms_e2e"] / args.steps, 4),
                "how": "NonKeyEngine.host_pipeline(): every step copies its uint8 frames + int16 MV fields from pinned host memory and "
                       "its class maps back; the copies of step i+1 overlap the compute of step i (one event pair around all K steps, "
                       "L2 flush inside)",
                "serial_value": round(main_r["e2e_serial"], 2), "serial_ms_per_step": round(main_r["ms_serial"] / args.steps, 4)},
        "gpu_launches": main_r["launches_per_step"] * args.steps,
        "launches_per_step": main_r["launches_per_step"],
        "clocks": sampler.summary(),
        "roofline": dominant, "roofline_creff": roof_creff, "roofline_conv": roof_conv,
        "cpu_baseline": cpu_base,
    }
    line.update(extras)
    if strong_r is not None:
        s_creff, s_conv, _ = rooflines(strong_r, args.precision)
        ceiling = N_FRAMES / float(-(-N_FRAMES // world))
        per_gpu = main_r["value"] / world
        line["strong"] = {
            "what": "ONE GOP's 11 non-keyframes dealt over the %d ranks (north_star's split): rank 0's keyframe feature (%.1f MB fp32) is "
                    "broadcast once per GOP with ncclBroadcast on a side stream that overlaps phase 1; only the CReFF launches wait for it"
                    % (world, C_P * (H // STRIDE_P) * (W // STRIDE_P) * 4 / 1e6),
            "value": round(strong_r["value"], 2), "unit": UNIT, "ms_per_gop": round(strong_r["ms"] / args.steps, 4),
            "ms_broadcast_alone": round(strong_r["ms_bcast"], 4) if strong_r["ms_bcast"] is not None else None,
            "frames_per_rank": [len(adist.frames_of_rank(N_FRAMES, world, rk)) for rk in range(world)],
            "ceiling_speedup": round(ceiling, 3), "speedup_vs_one_gpu": round(strong_r["value"] / per_gpu, 3),
            "frac_of_ceiling": round(strong_r["value"] / (per_gpu * ceiling), 3),
            "e2e": round(strong_r["e2e"], 2), "creff_ms": s_creff["ms_per_launch"], "conv_tflops": s_conv["achieved"],
            "scaling": "strong"}
    if alt_r is not None:
        a_creff, a_conv, _ = rooflines(alt_r, args.alt_precision)
        line["alt_precision"] = {"precision": args.alt_precision, "dtype": dnames[args.alt_precision], "value": round(alt_r["value"], 2),
                                 "ms_per_step": round(alt_r["ms"] / args.steps, 4), "e2e": round(alt_r["e2e"], 2),
                                 "creff_ms": a_creff["ms_per_launch"], "conv_tflops": a_conv["achieved"], "conv_frac": a_conv["frac"]}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


def run_extras(args, eng, sd, frames_u8, key_u8, mvs, ref_p, dev, l2_flush, timed_steps):
    """Single-GPU extras, all outside the timed regions of `value` / `e2e`:
    parity   -- frame 0 of the benchmarked step against the oracle's evaluation.py:176-204 (CPU, a few seconds);
    gop      -- whole-GOP rate: HR keyframe forward (KeyFrameEngine writing the feature where the CReFF kernel reads it) + the 11
                non-keyframes (SURVEY 8(d) config 2 "also report whole-GOP fps");
    dropin_api -- the literal per-frame API sequence of evaluation.py:177-204 (warpFeature, forward_phase1, forward_phase2) through
                the drop-in modules, batch 1, from pinned host buffers."""
    import torch
    from arseg_b200 import evaluation as ev
    from arseg_b200 import models, ops, synth
    from oracle import arseg_oracle as O
    out = {}
    # ---- parity
    eng.set_inputs(frames_u8.to(dev), mvs.to(dev), ref_p.to(dev))
    preds = eng.step()[0].cpu()
    logits = eng.logits[0:1].cpu() if eng.logits is not None else None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    o_preds, o_logits, _, _ = O.nonkey_step(ARCH, sd, normalise(frames_u8[0:1]), ref_p, synth.mv_to_flow(mvs[0].numpy()), SCALE, K_WIN, use_c=True)
    par = {"precision": args.precision, "frame": 0, "argmax_mismatch": round(float((preds.long() != o_preds[0]).float().mean()), 6),
           "oracle_s": round(time.perf_counter() - t0, 2), "reference": "oracle nonkey_step (evaluation.py:176-204 restated, fp32 CPU)"}
    if logits is not None and tuple(logits.shape) == tuple(o_logits.shape):
        par["logits_rel_err"] = round(float((logits - o_logits).abs().max() / o_logits.abs().max()), 6)
    out["parity"] = par
    # ---- whole GOP
    sd_hr = synth.synth_state_dict(models.models[ARCH]().state_dict(), 4)
    kf = ev.KeyFrameEngine(ARCH, sd_hr, H, W, args.precision, device=dev, uint8_frames=True, api_layout=False)
    kf.img.copy_(key_u8.to(dev))
    if kf.p_nhwc is not None and args.precision != "fp32":
        # the keyframe engine's feature buffer IS the non-keyframe engine's keyframe-feature input (internal fp32 NHWC layout)
        nk = ev.NonKeyEngine(ARCH, sd, N_FRAMES, H, W, SCALE, args.precision, K_WIN, device=dev, uint8_frames=True, ref_nhwc=kf.p_nhwc)
    else:
        kf = ev.KeyFrameEngine(ARCH, sd_hr, H, W, args.precision, device=dev, uint8_frames=True)
        kf.img.copy_(key_u8.to(dev))
        nk = ev.NonKeyEngine(ARCH, sd, N_FRAMES, H, W, SCALE, args.precision, K_WIN, device=dev, uint8_frames=True)
        kf_out = kf.p
    nk.set_inputs(frames_u8.to(dev), mvs.to(dev), None)

    def gop_step():
        kf.step()
        if nk.ref_nhwc is None:
            nk.ref_p.copy_(kf_out, non_blocking=True)
        nk.step()

    for _ in range(3):
        gop_step()
    ms_gop = timed_steps(gop_step, args.steps) / args.steps
    ms_kf = timed_steps(lambda: kf.step(), args.steps) / args.steps
    out["gop"] = {"value": round(GOP / (ms_gop * 1e-3), 2), "unit": "frames/s (all 12 frames of a GOP: HR keyframe forward + 11 non-keyframes)",
                  "ms_per_gop": round(ms_gop, 4), "ms_keyframe": round(ms_kf, 4), "keyframe_launches": kf.launches_per_step,
                  "launches_per_gop": kf.launches_per_step + nk.launches_per_step,
                  "keyframe_conv_tflops": round(kf.conv_flops / (ms_kf * 1e-3) / 1e12, 1)}
    del kf, nk
    torch.cuda.empty_cache()
    # ---- literal drop-in API, one frame at a time
    net = models.models_fuse[ARCH]()
    net.load_state_dict(sd)
    net.precision = args.precision
    net = net.to(dev).eval()
    frames = normalise(frames_u8).pin_memory()
    flows = torch.stack([synth.mv_to_flow(mvs[i].numpy())[0] for i in range(N_FRAMES)]).pin_memory()     # f64 [11,H,W,2] as the dataset yields it
    ref_d = ref_p.to(dev)
    pred_pin = torch.empty((1, H, W), dtype=torch.uint8).pin_memory()

    def api_frame(i):
        img = frames[i:i + 1].to(dev, non_blocking=True)
        fl = flows[i:i + 1].to(dev, non_blocking=True)
        p, _, _, _ = ev.nonkey_step(net, img, ref_d, fl, SCALE)
        pred_pin.copy_(p, non_blocking=True)

    for static in (False, True):
        net.static_outputs = static
        for i in range(3):
            api_frame(i)
        ms = timed_steps(lambda: [api_frame(i) for i in range(N_FRAMES)], max(2, args.steps // 4)) / max(2, args.steps // 4)
        key = "dropin_api_static_outputs" if static else "dropin_api"
        out[key] = {"value": round(N_FRAMES / (ms * 1e-3), 2), "unit": UNIT, "ms_per_frame": round(ms / N_FRAMES, 4),
                    "how": "ev.nonkey_step = warpFeature + F.interpolate-equivalent + net.forward_phase1 + net.forward_phase2 + resize/argmax, "
                           "batch 1, fp32 frame + f64 flow copied from pinned host memory every frame"
                           + ("; module outputs are the plans' static buffers" if static else "; module outputs cloned (reference semantics)")}
    plans = [pl for _, (pl, _) in net._plans.values()]
    out["dropin_api"]["launches_per_frame"] = sum(pl.n_launches for pl in plans) + 3
    return out


if __name__ == "__main__":
    main()
