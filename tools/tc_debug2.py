import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from arseg_b200 import ops, _lib as L, synth
from oracle import arseg_oracle as O
from tools.tc_debug import rnd, rel, sd_of, args_of, DEV
C = 64; sd = sd_of(C); k = 7
H, W, h, w = 16, 16, 8, 8
hr, lr = (rnd(1, C, H, W, seed=51) * 0.6).half().float(), (rnd(1, C, h, w, seed=52) * 0.4).half().float()
ref = O.creff(sd, "fuse_attention.", hr, lr, k)
lr_up = F.interpolate(lr, size=(H, W), mode="bilinear", align_corners=True)
out_p, _, _ = ops.creff_fused(ops.nchw_to_nhwc(hr.to(DEV), torch.float16), ops.nchw_to_nhwc(lr.to(DEV), torch.float16), *args_of(sd), k,
                              want_logits=False, lr_layout=L.NHWC, hr_layout=L.NHWC, engine=L.CREFF_MMA_F16)
o = out_p.cpu()
att_ref = ref - lr_up
att = o - lr_up
print("rows: |att - att_ref| max per row:", " ".join("%.3f" % v for v in (att - att_ref).abs().amax((0, 1, 3)).tolist()))
print("      |att_ref| max per row      :", " ".join("%.3f" % v for v in att_ref.abs().amax((0, 1, 3)).tolist()))
print("      |att| max per row          :", " ".join("%.3f" % v for v in att.abs().amax((0, 1, 3)).tolist()))
print("nan:", torch.isnan(o).sum().item())
# does tile 0 look like attention restricted to some key rows?  compare with attention computed on hr with rows zeroed
for y in range(8):
    print("row %d: err per col" % y, " ".join("%.2f" % v for v in (att - att_ref)[0, :, y, :].abs().amax(0).tolist()))
