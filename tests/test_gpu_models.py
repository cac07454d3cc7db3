"""GPU parity tests, module and pipeline level.

* drop-in modules (forward_phase1 / forward_phase2 / warpFeature) against the golden vectors recorded from
  the UNMODIFIED reference (tests/golden/*.npz) and against the oracle;
* the fused NonKeyEngine against the literal per-frame API sequence;
* full-size (720x960) size-independent properties.

Tolerances: 'fp32' precision = exact-arithmetic kernels: 1e-4 of the tensor's max magnitude and identical
argmax maps (up to ties: <= 0.05 % pixels).  'tf32' (tcgen05 kind::tf32, the arithmetic cuDNN applies to the
reference on Ampere-or-newer GPUs): 1e-2 max / 2e-3 rms, argmax agreement >= 99 %.  'f16' (fp16 storage = TF32's
11-bit significand, fp32 accumulate): the same bounds as 'tf32'.  'bf16': 6e-2 max /
2e-2 rms, argmax agreement >= 95 % on these random-weight nets (near-tie logits).
"""
import numpy as np
import pytest
import torch

from arseg_b200 import _lib as L
from arseg_b200 import evaluation as ev
from arseg_b200 import models, ops, synth
from oracle import arseg_oracle as O
from tests.util import CASES, case_setup, load_golden, rel_err, rms_err

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = "cuda:0"

TOL = {"fp32": (1e-4, 5e-5, 5e-4), "tf32": (1e-2, 2e-3, 1e-2), "f16": (1e-2, 2e-3, 1e-2), "bf16": (6e-2, 2e-2, 5e-2)}  # max, rms, argmax mismatch


def _net(arch, sd, precision):
    net = models.models_fuse[arch]()
    net.load_state_dict(sd)
    net.precision = precision
    return net.to(DEV).eval()


@pytest.mark.parametrize("precision", ["fp32", "tf32", "f16", "bf16"])
@pytest.mark.parametrize("name", CASES)
def test_dropin_modules_match_reference_golden(name, precision):
    g = load_golden(name)
    arch, _, sd, imgs, ref_p, mv, flow, scale = case_setup(g)
    net = _net(arch, sd, precision)
    tmax, trms, targ = TOL[precision]
    preds, logits, fused, lr_p = ev.nonkey_step(net, imgs.to(DEV), ref_p.to(DEV), flow.to(DEV), scale)
    assert rel_err(lr_p[:, ::4], torch.from_numpy(g["lr_p"])) < tmax
    assert rms_err(lr_p[:, ::4], torch.from_numpy(g["lr_p"])) < trms
    assert rel_err(fused[:, ::8], torch.from_numpy(g["fused"])) < tmax
    assert rel_err(logits, torch.from_numpy(g["logits"])) < tmax
    mism = float((preds.cpu().numpy() != g["preds"]).mean())
    assert mism <= targ, mism


@pytest.mark.parametrize("name", CASES)
def test_warp_feature_matches_reference_golden(name):
    g = load_golden(name)
    arch, _, sd, imgs, ref_p, mv, flow, scale = case_setup(g)
    fl = ev.resize_flow(flow.to(DEV), ref_p.shape[-2], ref_p.shape[-1])
    warped = ev.warpFeature(ref_p.to(DEV), fl)
    assert rel_err(warped[:, ::8], torch.from_numpy(g["warped"])) < 2e-5


@pytest.mark.parametrize("name", ["camvid_psp18_s05", "camvid_bise18_s05"])
def test_phase1_api_tuple(name):
    """forward_phase1 returns the reference's tuple (aux outputs included) -- checked against the oracle."""
    g = load_golden(name)
    arch, _, sd, imgs, ref_p, mv, flow, scale = case_setup(g)
    net = _net(arch, sd, "fp32")
    x = torch.nn.functional.interpolate(imgs, synth.lr_size(imgs.shape[2], imgs.shape[3], scale), mode="bilinear", align_corners=True)
    got = net.forward_phase1(x.to(DEV))
    ref = O._PHASES[arch][0](sd, x)
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert tuple(a.shape) == tuple(b.shape)
        assert rel_err(a, b) < 1e-4


@pytest.mark.parametrize("arch,H,W", [("camvid-psp18", 64, 96), ("camvid-bise18", 96, 128), ("cityscapes-psp18", 64, 128)])
def test_hr_keyframe_forward(arch, H, W):
    """HR branch on the same kernels: last output = p, first = logits (evaluation.py:119,173-174)."""
    net = models.models[arch]()
    sd = synth.synth_state_dict(net.state_dict(), 4)
    net.load_state_dict(sd)
    net.precision = "fp32"
    net = net.to(DEV).eval()
    x = synth.synth_frame(1, H, W, 2)
    got = net(x.to(DEV))
    ref = {"camvid-psp18": O.pspnet_hr, "camvid-bise18": O.bisenet_hr, "cityscapes-psp18": O.semseg_hr}[arch](sd, x)
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        assert tuple(a.shape) == tuple(b.shape)
        assert rel_err(a, b) < 1e-4


@pytest.mark.parametrize("precision", ["fp32", "tf32", "f16"])
@pytest.mark.parametrize("name", CASES)
def test_engine_matches_literal_api(name, precision):
    """NonKeyEngine (batched, int16 MVs, fused warp+CReFF+classifier+argmax, CUDA graph) == per-frame API path."""
    g = load_golden(name)
    arch, _, sd, imgs, ref_p, mv, flow, scale = case_setup(g)
    H, W = imgs.shape[-2:]
    N = 3
    frames = torch.cat([imgs] + [synth.synth_frame(1, H, W, 100 + i) for i in range(N - 1)])
    mvs = np.stack([mv] + [synth.synth_mv_int16(H, W, 200 + i, distance=2 + 4 * i) for i in range(N - 1)])
    eng = ev.NonKeyEngine(arch, sd, N, H, W, scale, precision, want_logits=True, want_p=True, device=DEV)
    eng.set_inputs(frames.to(DEV), torch.from_numpy(mvs).to(DEV), ref_p.to(DEV))
    preds = eng.step().cpu()
    torch.cuda.synchronize()
    net = _net(arch, sd, precision)
    tmax, trms, targ = TOL[precision]
    for i in range(N):
        fl = torch.from_numpy(mvs[i].astype(np.float64) / 4.0).unsqueeze(0)
        p_i, logits_i, fused_i, _ = ev.nonkey_step(net, frames[i:i + 1].to(DEV), ref_p.to(DEV), fl.to(DEV), scale)
        assert rel_err(eng.fused_p[i:i + 1], fused_i) < max(tmax, 1e-4)
        assert float((preds[i] != p_i[0].cpu()).float().mean()) <= targ
    if precision == "fp32":   # frame 0 is the golden case
        assert float((preds[0].numpy() != g["preds"][0]).mean()) <= 5e-4
    # replaying the captured graph is deterministic
    again = eng.step().cpu()
    assert torch.equal(again, preds)


def test_full_size_properties_camvid_psp():
    """720x960 (BASELINE config 2 shapes), size-independent properties of the fused kernel:
    (1) warp inside the kernel == warpFeature kernel followed by the pre-warped kernel;
    (2) frames of a batch are independent (batched == one at a time, bit-exact);
    (3) argmax output == argmax of the returned log-probs; log-probs exponentiate to 1."""
    C, H, W, h, w, k, ncls = 64, 720, 960, 360, 480, 7, 12
    net = models.models_fuse["camvid-psp18"]()
    sd = synth.synth_state_dict(net.state_dict(), 4)
    args = [sd["fuse_attention.%s.%s" % (n, l)].reshape(-1).contiguous().to(DEV)
            for n in ("lr_query_conv", "hr_key_conv", "hr_value_conv") for l in ("weight", "bias")]
    wcls, bcls = sd["final_conv.weight"].reshape(ncls, C).to(DEV), sd["final_conv.bias"].to(DEV)
    hr = (synth.synth_feature(1, C, H, W, 1) * 0.5).to(DEV)
    lr = (synth.synth_feature(2, C, h, w, 2) * 0.4).to(DEV)
    mvs = torch.from_numpy(np.stack([synth.synth_mv_int16(H, W, 3, 11), synth.synth_mv_int16(H, W, 4, 5)])).to(DEV)
    p, l, a = ops.creff_fused(hr, lr, *args, k, flow=mvs, wcls=wcls, bcls=bcls, log_softmax=True, want_argmax=True, hr_shared=True)
    for i in range(2):
        flow = (mvs[i:i + 1].double() / 4.0)
        warped = ops.warp_feature(hr, flow)
        p1, l1, a1 = ops.creff_fused(warped, lr[i:i + 1].contiguous(), *args, k, wcls=wcls, bcls=bcls, log_softmax=True, want_argmax=True)
        assert rel_err(p1, p[i:i + 1]) < 1e-6
        assert rel_err(l1, l[i:i + 1]) < 1e-6
        pi, li, ai = ops.creff_fused(hr, lr[i:i + 1].contiguous(), *args, k, flow=mvs[i:i + 1].contiguous(), wcls=wcls, bcls=bcls,
                                     log_softmax=True, want_argmax=True, hr_shared=True)
        assert torch.equal(pi, p[i:i + 1]) and torch.equal(li, l[i:i + 1]) and torch.equal(ai, a[i:i + 1])
    assert torch.equal(a.long(), l.argmax(1))
    assert (l.exp().sum(1) - 1).abs().max() < 1e-4


def test_host_pipeline_matches_resident_steps():
    """HostPipeline (H2D of step i+1 overlapping the compute of step i, D2H on a third stream) returns, for every
    step, exactly the class maps of the same inputs run one at a time from device-resident buffers."""
    g = load_golden("camvid_psp18_s05")
    arch, _, sd, imgs, ref_p, mv, flow, scale = case_setup(g)
    H, W = imgs.shape[-2:]
    N, STEPS = 2, 5
    eng = ev.NonKeyEngine(arch, sd, N, H, W, scale, "f16", device=DEV)
    batches = []
    for s in range(STEPS):
        fr = torch.cat([synth.synth_frame(1, H, W, 300 + 10 * s + i) for i in range(N)]).pin_memory()
        mvs = torch.from_numpy(np.stack([synth.synth_mv_int16(H, W, 400 + 10 * s + i, distance=1 + 2 * s + i) for i in range(N)])).pin_memory()
        batches.append((fr, mvs, torch.zeros((N, H, W), dtype=torch.uint8).pin_memory()))
    want = []
    for fr, mvs, _ in batches:
        eng.set_inputs(fr.to(DEV), mvs.to(DEV), ref_p.to(DEV))
        want.append(eng.step().cpu().clone())
    torch.cuda.synchronize()
    pipe = eng.host_pipeline()
    for fr, mvs, out in batches:
        pipe.submit(fr, mvs, out)
    pipe.drain()
    torch.cuda.synchronize()
    for s in range(STEPS):
        assert torch.equal(batches[s][2], want[s]), s
    assert not torch.equal(want[0], want[1])


def test_split_keyframe_engine_matches_single_graph():
    """split_keyframe=True (two graphs: phase 1 | keyframe-feature consumers, so a broadcast of the feature can overlap
    phase 1) gives exactly the class maps of the single-graph engine, also when the feature arrives between the halves."""
    g = load_golden("camvid_psp18_s05")
    arch, _, sd, imgs, ref_p, mv, flow, scale = case_setup(g)
    H, W = imgs.shape[-2:]
    one = ev.NonKeyEngine(arch, sd, 1, H, W, scale, "f16", device=DEV)
    two = ev.NonKeyEngine(arch, sd, 1, H, W, scale, "f16", device=DEV, split_keyframe=True)
    mvd = torch.from_numpy(mv).unsqueeze(0).to(DEV)
    one.set_inputs(imgs.to(DEV), mvd, ref_p.to(DEV))
    want = one.step().clone()
    two.set_inputs(imgs.to(DEV), mvd, torch.zeros_like(ref_p).to(DEV))
    two.step_phase1()
    two.ref_p.copy_(ref_p.to(DEV))          # the keyframe feature lands after phase 1 was enqueued
    got = two.step_phase2()
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    assert torch.equal(two.step(), want)


# ---------------------------------------------------------------- full-shape parity (BASELINE configs[0..3] shapes)
# One frame at the benchmarked resolutions through the captured engine against the oracle's literal
# evaluation.py:176-204 step (CPU, a few seconds per architecture, cached across the precisions): tile-edge, ring-wrap
# and 32-bit index behaviour at 720x960 / 1024x2048 that the 64x96 goldens cannot reach.
_FULL = {"camvid-psp18": (720, 960, 1, 64), "camvid-bise18": (720, 960, 8, 256), "cityscapes-psp18": (1024, 2048, 8, 512)}
_FULL_CACHE = {}


def _full_case(arch):
    if arch not in _FULL_CACHE:
        H, W, stride, C = _FULL[arch]
        net = models.models_fuse[arch]()
        sd = synth.synth_state_dict(net.state_dict(), 4)
        imgs = synth.synth_frame(1, H, W, 2)
        ref_p = synth.synth_feature(1, C, H // stride, W // stride, 1)
        mv = synth.synth_mv_int16(H, W, 3, distance=7)
        preds, logits, fused, lr_p = O.nonkey_step(arch, sd, imgs, ref_p, synth.mv_to_flow(mv), 0.5)
        _FULL_CACHE[arch] = (sd, imgs, ref_p, mv, preds, logits, fused, lr_p)
    return _FULL_CACHE[arch]


@pytest.mark.parametrize("precision", ["fp32", "tf32", "f16"])
@pytest.mark.parametrize("arch", list(_FULL))
def test_full_shape_engine_matches_oracle(arch, precision):
    H, W, stride, C = _FULL[arch]
    sd, imgs, ref_p, mv, preds, logits, fused, lr_p = _full_case(arch)
    eng = ev.NonKeyEngine(arch, sd, 1, H, W, 0.5, precision, want_logits=True, want_p=True, device=DEV)
    eng.set_inputs(imgs.to(DEV), torch.from_numpy(mv).unsqueeze(0).to(DEV), ref_p.to(DEV))
    got = eng.step().cpu()
    torch.cuda.synchronize()
    tmax, trms, targ = TOL[precision]
    assert tuple(got.shape) == (1, H, W)
    assert rel_err(eng.fused_p, fused) < max(tmax, 1e-4), rel_err(eng.fused_p, fused)
    assert rms_err(eng.fused_p, fused) < max(trms, 5e-5)
    mism = float((got.long() != preds).float().mean())
    assert mism <= targ, mism
    if arch == "camvid-psp18":          # logits at frame resolution come straight out of the fused kernel
        assert rel_err(eng.logits, logits) < max(tmax, 1e-4)
    del eng
    torch.cuda.empty_cache()


@pytest.mark.parametrize("ref_layout", ["nchw", "nhwc"])
def test_full_shape_tcgen05_creff_matches_oracle(ref_layout):
    """BASELINE configs[1] shapes through the tcgen05 CReFF engine (the f16 plan's default at C = 64): the MV-warp pre-pass is a
    hoisted side-stream branch of the captured graph; the keyframe feature is read as fp32, NCHW (API layout, transposed inside
    the plan) or NHWC in place."""
    arch = "camvid-psp18"
    H, W, stride, C = _FULL[arch]
    sd, imgs, ref_p, mv, preds, logits, fused, lr_p = _full_case(arch)
    ref_nhwc = torch.empty((1, H, W, C), device=DEV) if ref_layout == "nhwc" else None
    eng = ev.NonKeyEngine(arch, sd, 1, H, W, 0.5, "f16", want_logits=True, want_p=True, device=DEV, ref_nhwc=ref_nhwc)
    assert any(nm.endswith("_tc") for nm in eng.plan.names) and any(nm.endswith("_tc_prewarp") for nm in eng.plan.names), eng.plan.names
    assert eng.plan.hoisted, "the pre-pass must be a hoisted step"
    eng.set_inputs(imgs.to(DEV), torch.from_numpy(mv).unsqueeze(0).to(DEV), ref_p.to(DEV))
    got = eng.step().cpu()
    torch.cuda.synchronize()
    tmax, trms, targ = TOL["f16"]
    assert rel_err(eng.fused_p, fused) < tmax and rms_err(eng.fused_p, fused) < trms
    assert rel_err(eng.logits, logits) < tmax
    assert float((got.long() != preds).float().mean()) <= targ


def test_hoisted_prepass_equals_serial_plan(monkeypatch):
    """The side-stream fork / join of the pre-pass is a scheduling change only: the same engine with ARSEG_PLAN_OVERLAP=0 (every
    step in order on one stream) and the eager (uncaptured) plan give bit-identical class maps and logits, run after run."""
    g = load_golden("camvid_psp18_s05")
    arch, _, sd, imgs, ref_p, mv, flow, scale = case_setup(g)
    H, W = imgs.shape[-2:]
    mvd = torch.from_numpy(mv).unsqueeze(0).to(DEV)

    def run(graph):
        eng = ev.NonKeyEngine(arch, sd, 1, H, W, scale, "f16", want_logits=True, device=DEV, graph=graph)
        assert any(nm.endswith("_tc") for nm in eng.plan.names)
        eng.set_inputs(imgs.to(DEV), mvd, ref_p.to(DEV))
        outs = []
        for _ in range(3):
            outs.append((eng.step().clone(), eng.logits.clone()))
        torch.cuda.synchronize()
        for o in outs[1:]:
            assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1])
        return outs[0]
    a = run(True)
    b = run(False)
    monkeypatch.setenv("ARSEG_PLAN_OVERLAP", "0")
    c = run(True)
    for x in (b, c):
        assert torch.equal(a[0], x[0]) and torch.equal(a[1], x[1])


def test_f16_plan_keeps_the_march_engine_selectable(monkeypatch):
    """ARSEG_CREFF_TC=0: the f16 plan runs the mma.sync march engine (fp32 LR feature p) as in round 1."""
    monkeypatch.setenv("ARSEG_CREFF_TC", "0")
    arch = "camvid-psp18"
    H, W, stride, C = _FULL[arch]
    sd, imgs, ref_p, mv, preds, logits, fused, lr_p = _full_case(arch)
    eng = ev.NonKeyEngine(arch, sd, 1, H, W, 0.5, "f16", want_logits=True, want_p=True, device=DEV)
    assert any(nm.endswith("_mma") for nm in eng.plan.names) and not eng.plan.hoisted, eng.plan.names
    eng.set_inputs(imgs.to(DEV), torch.from_numpy(mv).unsqueeze(0).to(DEV), ref_p.to(DEV))
    got = eng.step().cpu()
    torch.cuda.synchronize()
    tmax, trms, targ = TOL["f16"]
    assert rel_err(eng.fused_p, fused) < tmax and rel_err(eng.logits, logits) < tmax
    assert float((got.long() != preds).float().mean()) <= targ


def test_dropin_models_run_as_dataparallel_replicas():
    """evaluation.py:41,54 wrap every net in nn.DataParallel and :173 calls the keyframe net through it: with more than one
    visible GPU the forward runs on replicas whose parameters are plain attributes (no state_dict entries)."""
    g = load_golden("camvid_psp18_s05")
    arch, _, sd, imgs, ref_p, mv, flow, scale = case_setup(g)
    net = _net(arch, sd, "fp32")
    x = torch.nn.functional.interpolate(imgs, synth.lr_size(imgs.shape[2], imgs.shape[3], scale), mode="bilinear", align_corners=True).to(DEV)
    want = net.forward_phase1(x)
    replica = torch.nn.parallel.replicate(net, [0, 0])[1]
    assert len(replica.state_dict()) < len(net.state_dict())        # the situation the fix is for
    got = replica.forward_phase1(x)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    out_r, fused_r = replica.forward_phase2(got[-1], ref_p.to(DEV))
    out_n, fused_n = net.forward_phase2(want[-1], ref_p.to(DEV))
    assert torch.equal(out_r, out_n) and torch.equal(fused_r, fused_n)
    hr = models.models[arch]()
    hr.load_state_dict(synth.synth_state_dict(hr.state_dict(), 4))
    hr.precision = "fp32"
    hr = hr.to(DEV).eval()
    yr = torch.nn.parallel.replicate(hr, [0, 0])[1](imgs.to(DEV))
    yn = hr(imgs.to(DEV))
    for a, b in zip(yr, yn):
        assert torch.equal(a, b)


# ---------------------------------------------------------------- drop-in `localAttention` module, as the reference imports it
def test_import_localAttention_through_the_reference_wrappers():
    """`dropin/` on sys.path makes `import localAttention` resolve to this repo; the reference's autograd wrappers
    similarFunction / weightingFunction (model/attention.py:13-53, restated here) then run forward AND backward on the sm_100a
    kernels.  Checked against the oracle ops under torch autograd."""
    import importlib
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.modules.pop("localAttention", None)
    sys.path.insert(0, os.path.join(root, "dropin"))
    try:
        la = importlib.import_module("localAttention")
    finally:
        sys.path.pop(0)
    assert os.path.dirname(la.__file__).startswith(os.path.join(root, "dropin"))

    class similarFunction(torch.autograd.Function):          # model/attention.py:13-30
        @staticmethod
        def forward(ctx, x_ori, x_loc, kH, kW):
            ctx.save_for_backward(x_ori, x_loc)
            ctx.kHW = (kH, kW)
            return la.similar_forward(x_ori, x_loc, kH, kW)

        @staticmethod
        def backward(ctx, grad_outputs):
            x_ori, x_loc = ctx.saved_tensors
            kH, kW = ctx.kHW
            return (la.similar_backward(x_loc, grad_outputs, kH, kW, True), la.similar_backward(x_ori, grad_outputs, kH, kW, False),
                    None, None)

    class weightingFunction(torch.autograd.Function):        # model/attention.py:33-50
        @staticmethod
        def forward(ctx, x_ori, x_weight, kH, kW):
            ctx.save_for_backward(x_ori, x_weight)
            ctx.kHW = (kH, kW)
            return la.weighting_forward(x_ori, x_weight, kH, kW)

        @staticmethod
        def backward(ctx, grad_outputs):
            x_ori, x_weight = ctx.saved_tensors
            kH, kW = ctx.kHW
            return la.weighting_backward_ori(x_weight, grad_outputs, kH, kW), la.weighting_backward_weight(x_ori, grad_outputs, kH, kW), None, None

    g = torch.Generator().manual_seed(31)
    q, kk, v = (torch.randn(2, 16, 11, 13, generator=g) for _ in range(3))
    kH = kW = 5
    with torch.enable_grad():
        # MyAttention.forward's core (model/attention.py:199-207) on the GPU through the wrappers
        qg, kg, vg = (t.clone().to(DEV).requires_grad_(True) for t in (q, kk, v))
        a = torch.softmax(similarFunction.apply(qg, kg, kH, kW), dim=3)
        o = weightingFunction.apply(vg, a, kH, kW)
        (o * o).sum().backward()
        # the same with the oracle ops (differentiable torch restatements) on the CPU
        qc, kc, vc = (t.clone().requires_grad_(True) for t in (q, kk, v))
        ac = torch.softmax(O.similar_forward(qc, kc, kH, kW), dim=3)
        oc = O.weighting_forward(vc, ac, kH, kW)
        (oc * oc).sum().backward()
    assert rel_err(o, oc) < 1e-5
    for got, want in ((qg.grad, qc.grad), (kg.grad, kc.grad), (vg.grad, vc.grad)):
        assert rel_err(got, want) < 1e-4


# ---------------------------------------------------------------- HR keyframe branch in the engine, uint8 ingest
@pytest.mark.parametrize("precision", ["fp32", "tf32", "f16"])
@pytest.mark.parametrize("arch,H,W", [("camvid-psp18", 64, 96), ("camvid-bise18", 96, 128), ("cityscapes-psp18", 64, 128)])
def test_keyframe_engine_matches_oracle(arch, H, W, precision):
    """KeyFrameEngine (captured HR forward up to p) == highres_net(ref_imgs)[-1] of evaluation.py:173-174 (oracle)."""
    net = models.models[arch]()
    sd = synth.synth_state_dict(net.state_dict(), 4)
    x = synth.synth_frame(1, H, W, 2)
    ref = {"camvid-psp18": O.pspnet_hr, "camvid-bise18": O.bisenet_hr, "cityscapes-psp18": O.semseg_hr}[arch](sd, x)[-1]
    eng = ev.KeyFrameEngine(arch, sd, H, W, precision, device=DEV)
    eng.img.copy_(x.to(DEV))
    p = eng.step()
    tmax, trms, _ = TOL[precision]
    assert tuple(p.shape) == tuple(ref.shape)
    assert rel_err(p, ref) < max(tmax, 1e-4) and rms_err(p, ref) < max(trms, 5e-5)


def test_keyframe_engine_full_shape_feeds_the_nonkey_engine():
    """720x960 PSPNet-18 keyframe in the f16 plan against the oracle, written straight into a NonKeyEngine's keyframe-feature
    buffer (whole-GOP path: HR forward + 11 non-keyframes)."""
    arch, H, W = "camvid-psp18", 720, 960
    net = models.models[arch]()
    sd = synth.synth_state_dict(net.state_dict(), 4)
    x = synth.synth_frame(1, H, W, 2)
    ref = O.pspnet_hr(sd, x)[-1]
    nk = ev.NonKeyEngine(arch, synth.synth_state_dict(models.models_fuse[arch]().state_dict(), 4), 1, H, W, 0.5, "f16", device=DEV)
    kf = ev.KeyFrameEngine(arch, sd, H, W, "f16", device=DEV, out=nk.ref_p)
    kf.img.copy_(x.to(DEV))
    p = kf.step()
    assert p.data_ptr() == nk.ref_p.data_ptr()
    tmax, trms, _ = TOL["f16"]
    assert rel_err(p, ref) < tmax and rms_err(p, ref) < trms
    nk.step()
    torch.cuda.synchronize()


@pytest.mark.parametrize("precision", ["f16", "tf32"])
def test_nonkey_engine_reads_the_keyframe_feature_in_nhwc(precision):
    """ref_nhwc: the keyframe feature in the internal NHWC layout (fp16 where the tcgen05 CReFF engine reads it, fp32 otherwise)
    is read in place -- same class maps and logits as the API-layout (NCHW) engine, one launch less."""
    g = load_golden("camvid_psp18_s05")
    arch, _, sd, imgs, ref_p, mv, flow, scale = case_setup(g)
    H, W = imgs.shape[-2:]
    mvd = torch.from_numpy(mv).unsqueeze(0).to(DEV)
    a = ev.NonKeyEngine(arch, sd, 1, H, W, scale, precision, device=DEV, want_logits=True)
    dt = ev.internal_ref_dtype(arch, precision)
    assert dt == (torch.float16 if precision == "f16" else torch.float32)
    buf = torch.empty((1, H, W, ref_p.shape[1]), dtype=dt, device=DEV)
    b = ev.NonKeyEngine(arch, sd, 1, H, W, scale, precision, device=DEV, want_logits=True, ref_nhwc=buf)
    a.set_inputs(imgs.to(DEV), mvd, ref_p.to(DEV))
    b.set_inputs(imgs.to(DEV), mvd, ref_p.to(DEV))
    assert b.launches_per_step == a.launches_per_step - 1
    pa, pb = a.step().clone(), b.step().clone()
    assert torch.equal(pa, pb) and torch.equal(a.logits, b.logits)
    if precision == "f16":      # an fp32 internal feature is accepted as well (the pre-pass then gathers fp32 taps)
        buf32 = torch.empty((1, H, W, ref_p.shape[1]), dtype=torch.float32, device=DEV)
        c = ev.NonKeyEngine(arch, sd, 1, H, W, scale, precision, device=DEV, want_logits=True, ref_nhwc=buf32)
        c.set_inputs(imgs.to(DEV), mvd, ref_p.to(DEV))
        pc = c.step().clone()
        assert float((pc != pa).float().mean()) < 2e-3 and rel_err(c.logits, a.logits.cpu()) < 3e-3


def test_uint8_frame_ingest_engine_matches_float_engine():
    """uint8 HWC frames + on-device ToTensor / Normalize (dataset/camvid.py:182-185) == the fp32 engine fed with the frames
    normalised on the host the way the dataset does."""
    g = load_golden("camvid_psp18_s05")
    arch, _, sd, imgs, ref_p, mv, flow, scale = case_setup(g)
    H, W = imgs.shape[-2:]
    gen = torch.Generator().manual_seed(12)
    u8 = torch.randint(0, 256, (2, H, W, 3), generator=gen, dtype=torch.uint8)
    f32 = O.ingest_u8(u8, ops.CAMVID_MEAN, ops.CAMVID_STD, (H, W))        # ToTensor + Normalize at full resolution
    mvs = torch.from_numpy(np.stack([mv, synth.synth_mv_int16(H, W, 77, distance=3)]))
    a = ev.NonKeyEngine(arch, sd, 2, H, W, scale, "fp32", device=DEV, want_logits=True)
    b = ev.NonKeyEngine(arch, sd, 2, H, W, scale, "fp32", device=DEV, want_logits=True, uint8_frames=True)
    a.set_inputs(f32.to(DEV), mvs.to(DEV), ref_p.to(DEV))
    b.set_inputs(u8.to(DEV), mvs.to(DEV), ref_p.to(DEV))
    pa, pb = a.step().clone(), b.step().clone()
    assert rel_err(b.logits, a.logits) < 1e-5
    assert float((pa != pb).float().mean()) < 1e-3
