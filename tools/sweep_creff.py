"""BASELINE config 5: CReFF local-window sweep (k = 3/5/7/9) x LR-scale sweep (0.3 .. 0.9) at 720x960, C = 64.

For every point the fused MV-warp + CReFF + classifier + argmax kernel runs on the 11 non-keyframes of one GOP
(shared keyframe feature); reports ms/frame and achieved GB/s = algorithmic bytes / time, against the measured HBM
peak (MEASURED_PEAKS.json).  Two byte counts: `moved` = what the launch has to move (hr fp32 + lr fp32 +
int16 MV in, fp32 log-probs + u8 class map out), `8d` = SURVEY's API-preserving figure that also writes the fused p.

    python tools/sweep_creff.py [--frames 11] [--iters 5] [--out gpurun_out/creff_sweep.md]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from arseg_b200 import _lib as L  # noqa: E402
from arseg_b200 import ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=11)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "creff_sweep.md"))
    a = ap.parse_args()
    torch.set_grad_enabled(False)
    dev = "cuda:0"
    C, H, W, ncls, N = 64, 720, 960, 12, a.frames
    peak = 6548.8
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("hbm_gbs", peak)
    hr = ops.nchw_to_nhwc(synth.synth_feature(1, C, H, W, 1).to(dev) * 0.5)
    hr16 = hr.half()          # the f16 plan hands the tcgen05 engine an fp16 keyframe feature (evaluation.internal_ref_dtype)
    mv = torch.from_numpy(np.stack([synth.synth_mv_int16(H, W, 10 + i, distance=1 + i % 11) for i in range(N)])).to(dev)
    g = torch.Generator().manual_seed(5)
    ws = []
    for _ in range(3):
        ws += [(torch.randn(C * 9, generator=g) * 0.3).to(dev), (torch.randn(C, generator=g) * 0.1).to(dev)]
    wcls, bcls = (torch.randn(ncls, C, generator=g) * 0.2).to(dev), (torch.randn(ncls, generator=g) * 0.1).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for scale in (0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9):
        h, w = int(H * scale), int(W * scale)
        lr = ops.nchw_to_nhwc(synth.synth_feature(N, C, h, w, 2).to(dev) * 0.5)
        lr16 = ops.nchw_to_nhwc(synth.synth_feature(N, C, h, w, 2).to(dev) * 0.5, torch.float16)
        for k in (3, 5, 7, 9):
            best = {}
            for eng_name, eng, lr_in in (("march", L.CREFF_MMA_F16, lr), ("tc", L.CREFF_TCGEN05, lr16)):
                if eng_name == "tc" and k > 7:
                    best[eng_name] = float("nan")
                    continue
                ts = []
                for it in range(a.iters + 1):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    ops.creff_fused(hr16 if eng_name == "tc" else hr, lr_in, *ws, k, flow=mv, wcls=wcls, bcls=bcls, log_softmax=True, lr_layout=L.NHWC, hr_layout=L.NHWC,
                                    engine=eng, want_p=False, want_logits=True, want_argmax=True, hr_shared=True)
                    e1.record()
                    torch.cuda.synchronize()
                    if it:
                        ts.append(e0.elapsed_time(e1))
                best[eng_name] = min(ts) / N
            ms = best["march"]
            moved = C * H * W * 4 + C * h * w * 4 + H * W * 4 + ncls * H * W * 4 + H * W
            full = 2 * C * H * W * 4 + C * h * w * 4 + H * W * 4 + ncls * H * W * 4
            rows.append((scale, "%dx%d" % (h, w), k, ms, moved / ms / 1e6, moved / ms / 1e6 / peak, full / ms / 1e6, full / ms / 1e6 / peak, best["tc"]))
        del lr, lr16
    out = ["| LR scale | LR p | k | ms/frame march (mma.sync, fp32 LR p) | GB/s (moved) | frac of %.0f GB/s | GB/s (8d bytes) | frac | ms/frame tcgen05 (pre-pass + engine, fp16 keyframe feature and LR p) |" % peak,
           "|---|---|---|---|---|---|---|---|---|"]
    for r in rows:
        out.append("| %.1f | %s | %d | %.4f | %.0f | %.3f | %.0f | %.3f | %.4f |" % r)
    txt = "\n".join(out)
    print(txt)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    open(a.out, "w").write("CReFF sweep, %d frames per launch, best of %d, L2 flushed between launches (CUDA events)\n\n" % (N, a.iters) + txt + "\n")


if __name__ == "__main__":
    main()
