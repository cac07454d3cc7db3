"""First-light / debugging run of the tcgen05 CReFF engine (csrc/creff_tc.cu) against the oracle on small cases."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from arseg_b200 import ops, _lib as L, synth
from oracle import arseg_oracle as O

DEV = torch.device("cuda:0")
def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / max(1e-6, float(b.abs().max())))
def sd_of(C, seed=50):
    spec = {"fuse_attention.%s.%s" % (n, l): torch.empty((C, 1, 3, 3) if l == "weight" else (C,))
            for n in ("lr_query_conv", "hr_key_conv", "hr_value_conv") for l in ("weight", "bias")}
    return synth.synth_state_dict(spec, seed)
def args_of(sd):
    return [sd["fuse_attention.%s.%s" % (n, l)].reshape(-1).contiguous().to(DEV)
            for n in ("lr_query_conv", "hr_key_conv", "hr_value_conv") for l in ("weight", "bias")]

C = 64
sd = sd_of(C)
for (k, H, W, h, w) in [(7, 32, 48, 16, 24), (7, 21, 37, 11, 19), (5, 32, 48, 16, 24), (3, 40, 16, 20, 8)]:
    hr, lr = (rnd(1, C, H, W, seed=51) * 0.6).half().float(), (rnd(1, C, h, w, seed=52) * 0.4).half().float()
    ncls = 12
    wcls, bcls = rnd(ncls, C, seed=55) * 0.2, rnd(ncls, seed=56) * 0.1
    ref = O.creff(sd, "fuse_attention.", hr, lr, k)
    logits = F.log_softmax(F.conv2d(ref, wcls.view(ncls, C, 1, 1), bcls), dim=1)
    out_p, out_l, out_a = ops.creff_fused(ops.nchw_to_nhwc(hr.to(DEV), torch.float16), ops.nchw_to_nhwc(lr.to(DEV), torch.float16), *args_of(sd), k,
                                          wcls=wcls.to(DEV), bcls=bcls.to(DEV), log_softmax=True, want_argmax=True,
                                          lr_layout=L.NHWC, hr_layout=L.NHWC, engine=L.CREFF_MMA_F16)
    torch.cuda.synchronize()
    d = (out_p.cpu() - ref).abs()
    print("k=%d %dx%d: rel_err p=%.3e logits=%.3e argmax mismatch=%.4f  worst at %s" % (
        k, H, W, rel(out_p, ref), rel(out_l, logits), (out_a.cpu().long() != logits.argmax(1)).float().mean().item(),
        str(np.unravel_index(int(d.argmax()), d.shape))), flush=True)
    if rel(out_p, ref) > 3e-3:
        e = d[0].amax(0)          # per-pixel max error map
        rows = (e > 3e-3 * ref.abs().max()).float().mean(1)
        cols = (e > 3e-3 * ref.abs().max()).float().mean(0)
        print("  bad fraction per row:", " ".join("%.1f" % v for v in rows.tolist()))
        print("  bad fraction per col:", " ".join("%.1f" % v for v in cols.tolist()))
