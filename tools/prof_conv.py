"""Runs one tcgen05 conv layer alone (for ncu / quick timing / env-var A-B tests).

    python tools/prof_conv.py --cin 64 --cout 64 --h 360 --w 480 [--n 11] [--k 3] [--dil 1] [--iters 5] [--dtype tf32|bf16]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from arseg_b200 import _lib as L  # noqa: E402
from arseg_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=11)
    ap.add_argument("--cin", type=int, default=64)
    ap.add_argument("--cout", type=int, default=64)
    ap.add_argument("--h", type=int, default=360)
    ap.add_argument("--w", type=int, default=480)
    ap.add_argument("--k", type=int, default=3)
    ap.add_argument("--dil", type=int, default=1)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--dtype", default="tf32")
    ap.add_argument("--res", action="store_true")
    a = ap.parse_args()
    torch.set_grad_enabled(False)
    dev = "cuda:0"
    dt = torch.float32 if a.dtype == "tf32" else torch.bfloat16
    eng = L.CONV_TC_TF32 if a.dtype == "tf32" else L.CONV_TC_BF16
    g = torch.Generator().manual_seed(0)
    x = torch.randn(a.n, a.h, a.w, a.cin, generator=g).to(dev, dt)
    w = (torch.randn(a.cout, a.k, a.k, a.cin, generator=g) * (1.0 / (a.cin * a.k * a.k) ** 0.5)).to(dev, dt)
    scale, shift = (torch.rand(a.cout, generator=g) + 0.5).to(dev), (torch.randn(a.cout, generator=g) * 0.1).to(dev)
    res = torch.randn(a.n, a.h, a.w, a.cout, generator=g).to(dev, dt) if a.res else None
    out = torch.empty(a.n, a.h, a.w, a.cout, dtype=dt, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pad = a.dil * (a.k // 2)
    times = []
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv2d_nhwc(x, w, scale, shift, res, 1, pad, a.dil, L.ACT_PRELU, 0.25, eng, out=out)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    flops = 2.0 * a.n * a.h * a.w * a.cin * a.cout * a.k * a.k
    best = min(times)
    print("conv %dx%d %d->%d @%dx%d x%d dil %d %s: best %.4f ms  %.1f TFLOP/s  (%s)" %
          (a.k, a.k, a.cin, a.cout, a.h, a.w, a.n, a.dil, a.dtype, best, flops / best / 1e9, " ".join("%.3f" % t for t in times)))


if __name__ == "__main__":
    main()
