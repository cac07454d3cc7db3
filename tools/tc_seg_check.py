"""Does the tcgen05 CReFF result depend on the row segmentation?  Prints max |diff| and the fraction of differing pixels."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from arseg_b200 import _lib as L, ops, synth
from test_gpu_ops import creff_sd, creff_args, rnd, _h, _tc_run, DEV

for k in (3, 5, 7):
    C, ncls, H, W, h, w = 64, 12, 42, 52, 21, 26
    sd = creff_sd(C)
    hr, lr = _h(rnd(1, C, H, W, seed=153) * 0.6), _h(rnd(2, C, h, w, seed=154) * 0.4)
    wcls, bcls = rnd(ncls, C, seed=155) * 0.2, rnd(ncls, seed=156) * 0.1
    mvs = torch.from_numpy(np.stack([synth.synth_mv_int16(H, W, 160 + i, distance=4 + 5 * i) for i in range(2)])).to(DEV)
    outs = {}
    for seg in (4096, 8, 12, 16, 24):
        os.environ["ARSEG_CREFF_SEG_ROWS"] = str(seg)
        outs[seg] = _tc_run(hr, lr, sd, k, flow=mvs, wcls=wcls.to(DEV), bcls=bcls.to(DEV), log_softmax=True, want_argmax=True, hr_shared=True)
    for seg in (8, 12, 16, 24):
        d = (outs[4096][0] - outs[seg][0]).abs()
        rows = (d.amax(dim=(0, 1, 3)) > 0).nonzero().flatten().tolist()
        print("k=%d seg=%d: max|dp| %.3g, pixels differing %.4f, rows %s, argmax diff %d" % (k, seg, d.max().item(), (d.amax(1) > 0).float().mean().item(), rows[:24], (outs[4096][2] != outs[seg][2]).sum().item()))
