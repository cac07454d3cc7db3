import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from arseg_b200 import ops, _lib as L, synth
from tools.tc_debug import rnd, sd_of, args_of, DEV
C, ncls, H, W, h, w = 64, 12, 42, 52, 21, 26
sd = sd_of(C)
def case(name, k=7, flow=True, lr_zero=False, hr_const=False, seg="8", dbg="0"):
    os.environ["ARSEG_CREFF_DBG"] = dbg
    hr = (rnd(1, C, H, W, seed=153) * 0.6).half().float()
    lr = (rnd(2, C, h, w, seed=154) * 0.4).half().float()
    if lr_zero: lr = lr * 0
    if hr_const: hr = hr * 0 + 0.25
    mvs = torch.from_numpy(np.stack([synth.synth_mv_int16(H, W, 160 + i, distance=4 + 5 * i) for i in range(2)])).to(DEV) if flow else None
    def run():
        return ops.creff_fused(ops.nchw_to_nhwc(hr.to(DEV), torch.float16), ops.nchw_to_nhwc(lr.to(DEV), torch.float16), *args_of(sd), k, flow=mvs,
                               want_logits=False, hr_shared=True, lr_layout=L.NHWC, hr_layout=L.NHWC, engine=L.CREFF_MMA_F16)[0]
    os.environ["ARSEG_CREFF_SEG_ROWS"] = "4096"; a = run()
    os.environ["ARSEG_CREFF_SEG_ROWS"] = seg; b = run()
    d = (a - b).abs()
    px = (d.amax(1) > 0).nonzero()
    print("%-28s differing pixels %d (max %.2e) %s" % (name, len(px), float(d.max()), [tuple(v[1:]) for v in px.tolist()[:8]]))
case("base k=7")
case("no flow", flow=False)
case("lr zero", lr_zero=True)
case("hr const", hr_const=True)
case("k=3", k=3)
case("k=5", k=5)
case("seg 16", seg="16")
case("seg 40", seg="40")
case("no overlap (dbg 8)", dbg="8")
