import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from arseg_b200 import ops, _lib as L, synth
from tools.tc_debug import rnd, sd_of, args_of, DEV
C, ncls, H, W, h, w, k = 64, 12, 42, 52, 21, 26, 7
sd = sd_of(C)
hr, lr = (rnd(1, C, H, W, seed=153) * 0.6).half().float(), (rnd(2, C, h, w, seed=154) * 0.4).half().float()
wcls, bcls = rnd(ncls, C, seed=155) * 0.2, rnd(ncls, seed=156) * 0.1
mvs = torch.from_numpy(np.stack([synth.synth_mv_int16(H, W, 160 + i, distance=4 + 5 * i) for i in range(2)])).to(DEV)
def run():
    return ops.creff_fused(ops.nchw_to_nhwc(hr.to(DEV), torch.float16), ops.nchw_to_nhwc(lr.to(DEV), torch.float16), *args_of(sd), k, flow=mvs,
                           wcls=wcls.to(DEV), bcls=bcls.to(DEV), log_softmax=True, want_argmax=True, hr_shared=True,
                           lr_layout=L.NHWC, hr_layout=L.NHWC, engine=L.CREFF_MMA_F16)
for seg in ("4096", "4096", "8", "8", "24"):
    os.environ["ARSEG_CREFF_SEG_ROWS"] = seg
    r = run()
    if seg == "4096" and "base" not in globals():
        base = r
        continue
    d = (r[0] - base[0]).abs()
    rows = d.amax((0, 1, 3)).cpu()
    print("seg_rows=%s: equal p=%s logits=%s; max|dp|=%.3e; rows with differences: %s" % (seg, torch.equal(r[0], base[0]), torch.equal(r[1], base[1]), float(d.max()),
          [i for i, v in enumerate(rows.tolist()) if v > 0]))
os.environ["ARSEG_CREFF_SEG_ROWS"] = "8"
r = run()
d = (r[0] - base[0]).abs()[0]          # frame 0: [C, H, W]
idx = (d > 0).nonzero()
print("frame 0: %d differing elements; channels %s" % (len(idx), sorted(set(idx[:, 0].tolist()))[:20]))
print("cols:", sorted(set(idx[:, 2].tolist())))
print("(row, col) pairs:", sorted(set((int(a), int(b)) for a, b in zip(idx[:, 1].tolist(), idx[:, 2].tolist())))[:40])
dl = (r[1] - base[1]).abs()[0]
print("logits differing elements:", int((dl > 0).sum()), "max", float(dl.max()))
