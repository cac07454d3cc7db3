"""Per-warp wait-time table of the tcgen05 CReFF engine (needs a -DARSEG_TTRACE build:
ARSEG_NVCC_EXTRA=-DARSEG_TTRACE python -m arseg_b200.build).  One CTA records, per warp, the cycles spent waiting on every
hand-off barrier and the cycles of the whole role."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arseg_b200 import _lib as L, ops, synth

TAGS = {1: "gfull/ready", 2: "ddone(G)", 3: "lrfree", 4: "ofull(D)", 5: "sfull(D)", 6: "ddone(M)", 7: "pfull", 8: "afull", 9: "sfull(S)", 10: "ofull(E)", 11: "lfull", 12: "ofree"}
ROLE = ["K0", "K1", "K2", "V0", "V1", "V2", "Qa", "Qb", "G0", "G1", "G2", "M", "E0", "E1", "E2", "E3", "S0", "S1", "S2", "S3"]

def main():
    torch.set_grad_enabled(False)
    dev = "cuda:0"
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 11
    Cc, H, W, ncls, k = 64, 720, 960, 12, 7
    hr = ops.nchw_to_nhwc(synth.synth_feature(1, Cc, H, W, 1).to(dev) * 0.5, torch.float16)
    lr = ops.nchw_to_nhwc(synth.synth_feature(frames, Cc, H // 2, W // 2, 2).to(dev) * 0.5, torch.float16)
    mv = torch.from_numpy(np.stack([synth.synth_mv_int16(H, W, 10 + i, distance=1 + i % 11) for i in range(frames)])).to(dev)
    g = torch.Generator().manual_seed(5)
    ws = []
    for _ in range(3):
        ws += [(torch.randn(Cc * 9, generator=g) * 0.3).to(dev), (torch.randn(Cc, generator=g) * 0.1).to(dev)]
    wcls, bcls = (torch.randn(ncls, Cc, generator=g) * 0.2).to(dev), (torch.randn(ncls, generator=g) * 0.1).to(dev)
    lib = L.load()
    buf = (C.c_longlong * 512)()
    for it in range(2):
        lib.arseg_debug_creff_tc_trace(buf, 1)
        ops.creff_fused(hr, lr, *ws, k, flow=mv, wcls=wcls, bcls=bcls, log_softmax=True, lr_layout=L.NHWC, want_p=False,
                        want_logits=True, want_argmax=True, hr_shared=True, hr_layout=L.NHWC, engine=L.CREFF_MMA_F16)
        torch.cuda.synchronize()
    lib.arseg_debug_creff_tc_trace(buf, 0)
    t = np.array(list(buf), dtype=np.int64).reshape(32, 16)
    S = max(1, int(t[16, 14]))
    print("tiles in the traced CTA: %d; cycles per tile by warp (total | waits by barrier)" % S)
    for w in range(len(ROLE)):
        tot = t[w, 15] / S
        waits = {TAGS[j]: t[w, j] / S for j in TAGS if t[w, j]}
        busy = tot - sum(waits.values())
        print("%-3s total %7.0f  busy %7.0f  " % (ROLE[w], tot, busy) + "  ".join("%s %.0f" % kv for kv in waits.items()))

    ev = (C.c_longlong * (3 * 64 * 12))()
    lib.arseg_debug_creff_tc_events(ev)
    e = np.array(list(ev), dtype=np.int64).reshape(3, 64, 12)
    mn = {0: "ready", 1: "QK issued", 2: "pfull", 3: "PV issued", 4: "afull", 5: "CLS issued"}
    sn = {1: "sfull", 2: "pass1", 3: "pass2", 4: "P arrive"}
    en = {0: "ready", 5: "ofull", 6: "A arrive", 7: "lfull", 8: "done"}
    for tile in (10, 11, 12):
        t0 = e[1, tile, 1]
        print("tile %d (cycles after S saw sfull): M " % tile + "  ".join("%s %d" % (mn[j], e[0, tile, j] - t0) for j in mn if e[0, tile, j]))
        print("     S " + "  ".join("%s %d" % (sn[j], e[1, tile, j] - t0) for j in sn))
        print("     E " + "  ".join("%s %d" % (en[j], e[2, tile, j] - t0) for j in en))
        print("     next tile: S sfull %d" % (e[1, tile + 1, 1] - t0))


if __name__ == "__main__":
    main()
