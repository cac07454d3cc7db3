"""Runs the fused warp+CReFF kernel alone at the CamVid 720x960 shape (for ncu / quick timing).

    python tools/prof_creff.py [--frames N] [--k 7] [--engine tc|mma|exact] [--iters I] [--want-p]
"""
import argparse
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
    ap.add_argument("--frames", type=int, default=2)
    ap.add_argument("--k", type=int, default=7)
    ap.add_argument("--scale", type=float, default=0.5)
    ap.add_argument("--engine", default="mma")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--want-p", action="store_true")
    ap.add_argument("--H", type=int, default=720)
    ap.add_argument("--W", type=int, default=960)
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--hr32", action="store_true", help="tc engine: fp32 NHWC keyframe feature (engine asked for by name)")
    a = ap.parse_args()
    torch.set_grad_enabled(False)
    dev = "cuda:0"
    C, H, W, ncls = 64, a.H, a.W, 12
    h, w = int(H * a.scale), int(W * a.scale)
    hr = synth.synth_feature(1, C, H, W, 1).to(dev) * 0.5
    lr = synth.synth_feature(a.frames, C, h, w, 2).to(dev) * 0.5
    mv = torch.from_numpy(np.stack([synth.synth_mv_int16(H, W, 10 + i, distance=1 + i % 11) for i in range(a.frames)])).to(dev)
    g = torch.Generator().manual_seed(5)
    ws = []
    for _ in range(3):
        ws += [(torch.randn(C * 9, generator=g) * 0.3).to(dev), (torch.randn(C, generator=g) * 0.1).to(dev)]
    wcls, bcls = (torch.randn(ncls, C, generator=g) * 0.2).to(dev), (torch.randn(ncls, generator=g) * 0.1).to(dev)
    lr_nhwc = ops.nchw_to_nhwc(lr, torch.float16 if a.engine == "tc" else torch.float32)
    if a.engine == "tc":      # tcgen05 engine: f16 keyframe feature + f16 LR feature
        hr_in, kw = ops.nchw_to_nhwc(hr, torch.float32 if a.hr32 else torch.float16), dict(hr_layout=L.NHWC, engine=L.CREFF_TCGEN05)
    elif a.engine == "mma":
        hr_in, kw = ops.nchw_to_nhwc(hr), dict(hr_layout=L.NHWC, engine=L.CREFF_MMA_F16)
    else:
        hr_in, kw = hr, dict(hr_layout=L.NCHW, engine=L.CREFF_EXACT_F32)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    times = []
    for it in range(a.iters):
        if not a.no_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.creff_fused(hr_in, lr_nhwc, *ws, a.k, flow=mv, wcls=wcls, bcls=bcls, log_softmax=True, lr_layout=L.NHWC,
                        want_p=a.want_p, want_logits=True, want_argmax=True, hr_shared=True, **kw)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    print("creff[%s] k=%d scale=%.1f frames=%d want_p=%d: %s ms (incl. output alloc), best %.4f ms/frame" %
          (a.engine, a.k, a.scale, a.frames, a.want_p, ["%.3f" % t for t in times], min(times) / a.frames))


if __name__ == "__main__":
    main()
