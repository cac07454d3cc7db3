"""Small invocations of the warp-specialised kernels for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize.py [march|wide|tc|conv|all]

Shapes are tiny (the tools slow kernels down by 100-1000x) but cover >= 2 march steps / tiles, a ragged strip and a segment
boundary, the halo conv kernel with a residual epilogue, and both classifier widths."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from arseg_b200 import _lib as L  # noqa: E402
from arseg_b200 import ops, synth  # noqa: E402

torch.set_grad_enabled(False)
DEV = "cuda:0"


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


def creff_case(C, H, W, h, w, k, hr_dtype, lr_dtype, ncls):
    hr, lr = rnd(1, C, H, W, seed=1) * 0.5, rnd(2, C, h, w, seed=2) * 0.4
    ws = []
    for i in range(3):
        ws += [(rnd(C * 9, seed=10 + i) * 0.3).to(DEV), (rnd(C, seed=20 + i) * 0.1).to(DEV)]
    wcls, bcls = (rnd(ncls, C, seed=5) * 0.2).to(DEV), (rnd(ncls, seed=6) * 0.1).to(DEV)
    mv = torch.from_numpy(np.stack([synth.synth_mv_int16(H, W, 30 + i, distance=3 + i) for i in range(2)])).to(DEV)
    out = ops.creff_fused(ops.nchw_to_nhwc(hr.to(DEV), hr_dtype), ops.nchw_to_nhwc(lr.to(DEV), lr_dtype), *ws, k, flow=mv, wcls=wcls, bcls=bcls,
                          log_softmax=True, want_argmax=True, hr_shared=True, lr_layout=L.NHWC, hr_layout=L.NHWC, engine=L.CREFF_MMA_F16)
    torch.cuda.synchronize()
    return float(out[0].abs().mean())


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    os.environ["ARSEG_CREFF_SEG_ROWS"] = "16"
    if what in ("march", "all"):
        print("creff_march (C=64, mma.sync, mbarrier hand-offs):", creff_case(64, 28, 40, 14, 20, 7, torch.float32, torch.float32, 12))
    if what in ("tc", "all"):
        print("creff_tc (C=64, tcgen05/TMEM):", creff_case(64, 28, 40, 14, 20, 7, torch.float16, torch.float16, 19))
    if what in ("wide", "all"):
        print("creff_wide (C=128, cp.async double buffers):", creff_case(128, 12, 24, 6, 12, 5, torch.float32, torch.float32, 12))
    if what in ("conv", "all"):
        x = (rnd(1, 24, 20, 64, seed=3) * 0.5).half().to(DEV)
        w = (rnd(64, 3, 3, 64, seed=4) * 0.05).half().to(DEV)
        res = (rnd(1, 24, 20, 64, seed=7) * 0.5).half().to(DEV)
        sc, sh = torch.ones(64, device=DEV), torch.zeros(64, device=DEV)
        y = ops.conv2d_nhwc(x, w, sc, sh, residual=res, pad=1, act=L.ACT_RELU, engine=L.CONV_TC_F16)          # halo kernel
        y2 = ops.conv2d_nhwc(x, (rnd(128, 1, 1, 64, seed=8) * 0.1).half().to(DEV), engine=L.CONV_TC_F16)        # tap-box kernel (1x1)
        torch.cuda.synchronize()
        print("conv_tc (halo 3x3 + residual, 1x1):", float(y.float().abs().mean()), float(y2.float().abs().mean()))


if __name__ == "__main__":
    main()
