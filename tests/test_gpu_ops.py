"""GPU parity tests, op level: every C-ABI kernel against the oracle / the ATen CPU op the reference calls.

Tolerances (stated per test): exact-arithmetic fp32 kernels agree with the CPU fp32 reference up to
summation order -> 2e-5 relative to the tensor's max magnitude; TF32 tensor-core convs 2e-3; bf16 2e-2.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from arseg_b200 import _lib as L
from arseg_b200 import ops
from oracle import arseg_oracle as O
from tests.util import rel_err, rms_err

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = "cuda:0"
FP32_TOL = 2e-5


def rnd(*shape, seed=0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=dtype)


# ---------------------------------------------------------------- localAttention boundary
@pytest.mark.parametrize("shape,k", [((2, 8, 9, 37), (3, 3)), ((1, 64, 24, 40), (7, 7)), ((1, 5, 1, 33), (5, 5)),
                                     ((2, 6, 11, 13), (3, 5)), ((1, 16, 20, 70), (9, 9))])
def test_local_attention_forward_ops(shape, k):
    q, kk = rnd(*shape, seed=1), rnd(*shape, seed=2)
    s_ref = O.similar_forward(q, kk, *k)
    s = ops.similar_forward(q.to(DEV), kk.to(DEV), *k)
    assert s.shape == s_ref.shape
    assert rel_err(s, s_ref) < FP32_TOL
    a = torch.softmax(s_ref, dim=3)
    o_ref = O.weighting_forward(kk, a, *k)
    o = ops.weighting_forward(kk.to(DEV), a.to(DEV), *k)
    assert rel_err(o, o_ref) < FP32_TOL


def test_local_attention_zero_padding_takes_softmax_mass():
    """out-of-image taps give logit exactly 0 (not -inf)."""
    q, k = rnd(1, 4, 6, 6, seed=3), rnd(1, 4, 6, 6, seed=4)
    s = ops.similar_forward(q.to(DEV), k.to(DEV), 3, 3).cpu()
    assert (s[0, 0, 0, [0, 1, 2, 3, 6]] == 0).all()     # corner pixel: 5 of 9 taps are outside
    assert (s[0, 2, 2] != 0).all()


@pytest.mark.parametrize("k", [(3, 3), (7, 7), (3, 5)])
def test_local_attention_backward_ops(k):
    shape = (2, 6, 10, 35)
    x, y = rnd(*shape, seed=5), rnd(*shape, seed=6)
    T = k[0] * k[1]
    g_s = rnd(2, 10, 35, T, seed=7)
    for is_ori in (True, False):
        ref = O.similar_backward(x, g_s, *k, is_ori)
        got = ops.similar_backward(x.to(DEV), g_s.to(DEV), *k, is_ori)
        assert rel_err(got, ref) < FP32_TOL
    a = torch.softmax(g_s, 3)
    ref = O.weighting_backward_ori(a, y, *k)
    assert rel_err(ops.weighting_backward_ori(a.to(DEV), y.to(DEV), *k), ref) < FP32_TOL
    ref = O.weighting_backward_weight(x, y, *k)
    assert rel_err(ops.weighting_backward_weight(x.to(DEV), y.to(DEV), *k), ref) < FP32_TOL


def test_local_attention_rejects_bad_input():
    with pytest.raises(RuntimeError):
        ops.similar_forward(torch.zeros(1, 2, 4, 4, device=DEV, dtype=torch.float64), torch.zeros(1, 2, 4, 4, device=DEV, dtype=torch.float64), 3, 3)
    with pytest.raises(RuntimeError):
        ops.similar_forward(torch.zeros(1, 2, 4, 4, device=DEV), torch.zeros(1, 2, 4, 4, device=DEV), 4, 4)  # even window


# ---------------------------------------------------------------- warpFeature / resize
@pytest.mark.parametrize("fdtype", [torch.float64, torch.float32])
def test_warp_feature(fdtype):
    feat = rnd(2, 7, 19, 45, seed=8)
    g = torch.Generator().manual_seed(9)
    flow = (torch.randint(-6, 7, (2, 19, 45, 2), generator=g).to(fdtype))
    flow[0, :4] *= 10          # far out-of-image samples -> zeros padding
    ref = O.warp_feature(feat, flow)
    got = ops.warp_feature(feat.to(DEV), flow.to(DEV))
    assert rel_err(got, ref) < FP32_TOL


@pytest.mark.parametrize("mode,kw", [(L.RESIZE_BILINEAR, dict(mode="bilinear", align_corners=False)),
                                     (L.RESIZE_BILINEAR_AC, dict(mode="bilinear", align_corners=True)),
                                     (L.RESIZE_NEAREST, dict(mode="nearest"))])
@pytest.mark.parametrize("sizes", [((23, 30), (46, 60)), ((72, 96), (36, 48)), ((12, 15), (23, 30)), ((9, 12), (72, 96))])
def test_resize_nchw(mode, kw, sizes):
    (hi, wi), (ho, wo) = sizes
    x = rnd(2, 3, hi, wi, seed=10)
    ref = F.interpolate(x, size=(ho, wo), **kw)
    got = ops.resize_nchw(x.to(DEV), (ho, wo), mode)
    assert rel_err(got, ref) < FP32_TOL


@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)])
@pytest.mark.parametrize("hw,C", [((45, 60), 64), ((1, 7), 8), ((6, 1), 16), ((9, 12), 40)])
def test_resize_nhwc_x2_and_generic(hw, C, dtype, tol):
    """F.upsample(scale_factor=2, bilinear) (model/pspnet.py:38-41): the exact-x2 kernel and the generic kernel agree
    (to fp32 rounding: the compiler may contract the two kernels' multiply-adds differently), both equal to
    F.interpolate(align_corners=False); channel-slice destination."""
    hi, wi = hw
    x = rnd(2, C, hi, wi, seed=12)
    xr = x.to(dtype).float()
    ref = F.interpolate(xr, size=(2 * hi, 2 * wi), mode="bilinear", align_corners=False)
    nhwc = ops.nchw_to_nhwc(x.to(DEV), dtype)
    out = torch.zeros((2, 2 * hi, 2 * wi, C + 16), dtype=dtype, device=DEV)
    ops.resize_nhwc(nhwc, (2 * hi, 2 * wi), L.RESIZE_BILINEAR, out=out, coff=8)            # x2 kernel (C % 8 == 0)
    got = out[..., 8:8 + C].float().permute(0, 3, 1, 2).cpu()
    assert rel_err(got, ref) < tol
    assert float(out[..., :8].abs().max()) == 0 and float(out[..., 8 + C:].abs().max()) == 0
    # generic kernel: a destination slice that is not 16-byte aligned takes the scalar path of resize_nhwc_kernel
    out2 = torch.zeros((2, 2 * hi, 2 * wi, C + 3), dtype=dtype, device=DEV)
    ops.resize_nhwc(nhwc, (2 * hi, 2 * wi), L.RESIZE_BILINEAR, out=out2, coff=3)
    assert rel_err(out2[..., 3:].float(), out[..., 8:8 + C].float()) < tol


def test_layout_roundtrip_and_resize_nhwc_slice():
    x = rnd(2, 12, 9, 14, seed=11)
    nhwc = ops.nchw_to_nhwc(x.to(DEV))
    assert torch.equal(nhwc.cpu(), x.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(ops.nhwc_to_nchw(nhwc).cpu(), x)
    bf = ops.nchw_to_nhwc(x.to(DEV), torch.bfloat16)
    assert torch.equal(bf.cpu(), x.permute(0, 2, 3, 1).contiguous().bfloat16())


# ---------------------------------------------------------------- conv engines
def conv_ref(x, w, scale, shift, res, stride, pad, dil, act, slope):
    y = F.conv2d(x.double(), w.double(), None, stride, pad, dil)
    if scale is not None:
        y = y * scale.double().view(1, -1, 1, 1)
    if shift is not None:
        y = y + shift.double().view(1, -1, 1, 1)
    if res is not None:
        y = y + res.double()
    if act == L.ACT_RELU:
        y = F.relu(y)
    elif act == L.ACT_PRELU:
        y = torch.where(y > 0, y, y * slope)
    return y.float()


def run_conv(x, w, scale, shift, res, stride, pad, dil, act, slope, engine, dtype=torch.float32):
    xd = ops.nchw_to_nhwc(x.to(DEV), dtype)
    wd = w.permute(0, 2, 3, 1).contiguous().to(DEV, dtype)
    rd = ops.nchw_to_nhwc(res.to(DEV), dtype) if res is not None else None
    y = ops.conv2d_nhwc(xd, wd, scale.to(DEV) if scale is not None else None, shift.to(DEV) if shift is not None else None,
                        rd, stride, pad, dil, act, slope, engine)
    return ops.nhwc_to_nchw(y).cpu()


CONV_CASES = [
    # N, Cin, Cout, H, W, k, stride, pad, dil, residual, act
    (2, 64, 64, 23, 30, 3, 1, 1, 1, True, L.ACT_RELU),
    (1, 64, 128, 24, 32, 3, 2, 1, 1, False, L.ACT_RELU),
    (1, 64, 128, 24, 32, 1, 2, 0, 1, False, L.ACT_NONE),
    (1, 128, 256, 12, 15, 3, 1, 2, 2, True, L.ACT_RELU),
    (1, 256, 512, 12, 15, 3, 1, 4, 4, False, L.ACT_PRELU),
    (2, 512, 128, 6, 6, 1, 1, 0, 1, False, L.ACT_NONE),
    (1, 64, 12, 20, 28, 1, 1, 0, 1, False, L.ACT_NONE),
    (1, 192, 96, 17, 21, 3, 1, 1, 1, False, L.ACT_RELU),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_simt_fp32(case):
    N, Ci, Co, H, W, k, s, p, d, has_res, act = case
    x, w = rnd(N, Ci, H, W, seed=20), rnd(Co, Ci, k, k, seed=21) * (1.0 / (Ci * k * k) ** 0.5)
    scale, shift = torch.rand(Co) + 0.5, rnd(Co, seed=22) * 0.1
    Ho, Wo = (H + 2 * p - d * (k - 1) - 1) // s + 1, (W + 2 * p - d * (k - 1) - 1) // s + 1
    res = rnd(N, Co, Ho, Wo, seed=23) if has_res else None
    ref = conv_ref(x, w, scale, shift, res, s, p, d, act, 0.25)
    got = run_conv(x, w, scale, shift, res, s, p, d, act, 0.25, L.CONV_SIMT_F32)
    assert got.shape == ref.shape
    assert rel_err(got, ref) < FP32_TOL


TC_CASES = [c for c in CONV_CASES if c[2] >= 16] + [
    (2, 64, 128, 45, 61, 3, 2, 1, 1, False, L.ACT_RELU),       # stride 2 with odd extents (TMA element strides)
    (1, 128, 256, 23, 30, 1, 2, 0, 1, False, L.ACT_NONE),
    (2, 64, 64, 45, 60, 3, 1, 1, 1, True, L.ACT_RELU),
    (1, 1024, 256, 18, 24, 3, 1, 1, 1, False, L.ACT_PRELU),
    (1, 2560, 1024, 9, 12, 1, 1, 0, 1, False, L.ACT_RELU),
]


@pytest.mark.parametrize("case", TC_CASES)
@pytest.mark.parametrize("engine,dtype,tol", [(L.CONV_TC_TF32, torch.float32, 2e-3), (L.CONV_TC_F16, torch.float16, 2e-3),
                                              (L.CONV_TC_BF16, torch.bfloat16, 2e-2)])
def test_conv_tcgen05(case, engine, dtype, tol):
    N, Ci, Co, H, W, k, s, p, d, has_res, act = case
    x, w = rnd(N, Ci, H, W, seed=30), rnd(Co, Ci, k, k, seed=31) * (1.0 / (Ci * k * k) ** 0.5)
    scale, shift = torch.rand(Co) + 0.5, rnd(Co, seed=32) * 0.1
    Ho, Wo = (H + 2 * p - d * (k - 1) - 1) // s + 1, (W + 2 * p - d * (k - 1) - 1) // s + 1
    res = rnd(N, Co, Ho, Wo, seed=33) if has_res else None
    if dtype != torch.float32:    # reference on the 16-bit-rounded operands: isolates accumulation error
        x, w = x.to(dtype).float(), w.to(dtype).float()
        res = res.to(dtype).float() if res is not None else None
    ref = conv_ref(x, w, scale, shift, res, s, p, d, act, 0.25)
    got = run_conv(x, w, scale, shift, res, s, p, d, act, 0.25, engine, dtype)
    assert got.shape == ref.shape
    assert rel_err(got, ref) < tol, (rel_err(got, ref), rms_err(got, ref))
    assert rms_err(got, ref) < tol / 4


@pytest.mark.parametrize("N,H,W,Cout", [(2, 37, 53, 64), (1, 64, 130, 64), (1, 23, 18, 40)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.float16, 2e-3), (torch.bfloat16, 2e-2)])
def test_conv_stem(N, H, W, Cout, dtype, tol):
    """conv1 7x7 stride 2 pad 3 (no bias) + BN(eval) + ReLU, model/extractors.py:112-114,148-150.
    16-bit dtypes with Cout = 64 run the tensor-core stem (stem_mma.cu), everything else the CUDA-core kernel."""
    x, w = rnd(N, 3, H, W, seed=40), rnd(Cout, 3, 7, 7, seed=41) * (1.0 / 147 ** 0.5)
    scale, shift = torch.rand(Cout) + 0.5, rnd(Cout, seed=42) * 0.1
    ref = F.relu(F.conv2d(x.double(), w.double(), None, 2, 3) * scale.double().view(1, -1, 1, 1) + shift.double().view(1, -1, 1, 1)).float()
    got = ops.nhwc_to_nchw(ops.conv_stem(x.to(DEV), w.to(DEV), scale.to(DEV), shift.to(DEV), dtype)).cpu()
    assert got.shape == ref.shape
    assert rel_err(got, ref) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.float16, 2e-3)])
@pytest.mark.parametrize("semseg", [False, True])
def test_pyramid_pooling_module(dtype, tol, semseg):
    """PSPModule (model/pspnet.py:14-31) / PPM (model/pspnet_semseg.py:12-30) as three launches."""
    from arseg_b200 import engine as E
    N, Cf, H, W, bins = 2, 64, 23, 30, (1, 2, 3, 6)
    Cout = Cf // 4 if semseg else Cf
    x = rnd(N, Cf, H, W, seed=80)
    ws = [rnd(Cout, Cf, 1, 1, seed=81 + i) * (1.0 / Cf ** 0.5) for i in range(4)]
    scs = [torch.rand(Cout) + 0.5 for _ in range(4)] if semseg else None
    shs = [rnd(Cout, seed=90 + i) * 0.1 for i in range(4)] if semseg else None
    xr = x.to(dtype).float()
    branches = []
    for i, b in enumerate(bins):
        t = F.conv2d(F.adaptive_avg_pool2d(xr, b), ws[i])
        if semseg:
            t = F.relu(t * scs[i].view(1, -1, 1, 1) + shs[i].view(1, -1, 1, 1))
        branches.append(F.interpolate(t, size=(H, W), mode="bilinear", align_corners=semseg))
    ref = torch.cat([xr] + branches, 1) if semseg else torch.cat(branches + [xr], 1)
    pl = E.Plan(torch.device(DEV), {torch.float32: "fp32", torch.float16: "f16"}[dtype])
    f = ops.nchw_to_nhwc(x.to(DEV), dtype)
    cat = pl.pyramid(f, bins, ws, scs, shs, relu=semseg, mode=L.RESIZE_BILINEAR_AC if semseg else L.RESIZE_BILINEAR,
                     stages_first=not semseg)
    pl.launch()
    torch.cuda.synchronize()
    got = ops.nhwc_to_nchw(cat).cpu()
    assert got.shape == ref.shape
    assert rel_err(got, ref) < tol, rel_err(got, ref)


def test_conv_output_channel_slice():
    """out_cstride / out_coff: the conv writes a channel slice of a wider tensor (fused torch.cat)."""
    x, w = rnd(1, 64, 10, 12, seed=40), rnd(32, 64, 1, 1, seed=41) * 0.1
    xd = ops.nchw_to_nhwc(x.to(DEV))
    wd = w.permute(0, 2, 3, 1).contiguous().to(DEV)
    out = torch.full((1, 10, 12, 96), 7.0, device=DEV)
    ops.conv2d_nhwc(xd, wd, out=out, out_coff=32)
    ref = F.conv2d(x, w).permute(0, 2, 3, 1)
    assert rel_err(out[..., 32:64], ref) < FP32_TOL
    assert (out[..., :32] == 7).all() and (out[..., 64:] == 7).all()


# ---------------------------------------------------------------- fused CReFF
def creff_sd(C, seed=50):
    from arseg_b200 import synth
    spec = {"fuse_attention.%s.%s" % (n, l): torch.empty((C, 1, 3, 3) if l == "weight" else (C,))
            for n in ("lr_query_conv", "hr_key_conv", "hr_value_conv") for l in ("weight", "bias")}
    return synth.synth_state_dict(spec, seed)


def creff_args(sd):
    return [sd["fuse_attention.%s.%s" % (n, l)].reshape(-1).contiguous().to(DEV)
            for n in ("lr_query_conv", "hr_key_conv", "hr_value_conv") for l in ("weight", "bias")]


@pytest.mark.parametrize("k", [3, 5, 7, 9])
@pytest.mark.parametrize("C,H,W,h,w", [(64, 24, 40, 12, 20), (16, 17, 35, 11, 23)])
def test_creff_fused_prewarped(k, C, H, W, h, w):
    """MyAttention.forward semantics (model/attention.py:184-213) as one kernel, all window sizes."""
    sd = creff_sd(C)
    hr, lr = rnd(1, C, H, W, seed=51) * 0.6, rnd(1, C, h, w, seed=52) * 0.4
    ref = O.creff(sd, "fuse_attention.", hr, lr, k)
    out_p, _, _ = ops.creff_fused(hr.to(DEV), lr.to(DEV), *creff_args(sd), k, want_logits=False)
    assert rel_err(out_p, ref) < 5e-5


@pytest.mark.parametrize("flow_kind", ["i16", "f64", "f32"])
@pytest.mark.parametrize("stride", [1, 8])
def test_creff_fused_with_mv_warp_and_classifier(flow_kind, stride):
    """evaluation.py:177-183 + MyAttention + final_conv + log-softmax + argmax in one launch."""
    from arseg_b200 import synth
    C, ncls, k = 64, 12, 7
    Hm, Wm = 48 * stride, 64 * stride // 2 * 2
    H, W, h, w = Hm // stride, Wm // stride, Hm // stride // 2, Wm // stride // 2
    sd = creff_sd(C)
    hr, lr = rnd(1, C, H, W, seed=53) * 0.6, rnd(2, C, h, w, seed=54) * 0.4
    wcls, bcls = rnd(ncls, C, seed=55) * 0.2, rnd(ncls, seed=56) * 0.1
    mvs = np.stack([synth.synth_mv_int16(Hm, Wm, 60 + i, distance=5 + 3 * i) for i in range(2)])
    flow64 = torch.from_numpy(mvs.astype(np.float64) / 4.0)
    refs = []
    for i in range(2):
        fl = O.resize_flow(flow64[i:i + 1], H, W)
        warped = O.warp_feature(hr, fl)
        fused = O.creff(sd, "fuse_attention.", warped, lr[i:i + 1], k)
        logits = F.log_softmax(F.conv2d(fused, wcls.view(ncls, C, 1, 1), bcls), dim=1)
        refs.append((fused, logits))
    flow = {"i16": torch.from_numpy(mvs), "f64": flow64, "f32": flow64.float()}[flow_kind].to(DEV)
    out_p, out_l, out_a = ops.creff_fused(hr.to(DEV), lr.to(DEV), *creff_args(sd), k, flow=flow, wcls=wcls.to(DEV),
                                          bcls=bcls.to(DEV), log_softmax=True, want_argmax=True, hr_shared=True)
    tol = 5e-5 if flow_kind != "f32" else 5e-3   # an fp32 flow changes the reference's own grid arithmetic
    for i in range(2):
        assert rel_err(out_p[i:i + 1], refs[i][0]) < tol
        assert rel_err(out_l[i:i + 1], refs[i][1]) < tol
        ref_arg = refs[i][1].argmax(1)
        mism = (out_a[i:i + 1].cpu().long() != ref_arg).float().mean().item()
        assert mism < (1e-3 if flow_kind != "f32" else 1e-2)
        assert torch.equal(out_a[i:i + 1].cpu().long(), out_l[i:i + 1].cpu().argmax(1))


def test_creff_fused_nhwc_lr_and_bf16():
    C, H, W, h, w, k = 64, 24, 40, 12, 20, 7
    sd = creff_sd(C)
    hr, lr = rnd(1, C, H, W, seed=57) * 0.6, rnd(1, C, h, w, seed=58) * 0.4
    ref = O.creff(sd, "fuse_attention.", hr, lr, k)
    lr_nhwc = ops.nchw_to_nhwc(lr.to(DEV))
    out_p, _, _ = ops.creff_fused(hr.to(DEV), lr_nhwc, *creff_args(sd), k, want_logits=False, lr_layout=L.NHWC)
    assert rel_err(out_p, ref) < 5e-5
    lr_bf = ops.nchw_to_nhwc(lr.to(DEV), torch.bfloat16)
    ref_bf = O.creff(sd, "fuse_attention.", hr, lr.bfloat16().float(), k)
    out_p, _, _ = ops.creff_fused(hr.to(DEV), lr_bf, *creff_args(sd), k, want_logits=False, lr_layout=L.NHWC)
    assert rel_err(out_p, ref_bf) < 5e-5


def test_resize_argmax_and_hist():
    logits = rnd(2, 12, 10, 12, seed=70)
    ref_l = F.interpolate(logits, size=(80, 96), mode="bilinear", align_corners=True)
    pred, up = ops.resize_argmax(logits.to(DEV), (80, 96), L.RESIZE_BILINEAR_AC, want_logits=True)
    assert rel_err(up, ref_l) < FP32_TOL
    assert (pred.cpu().long() != ref_l.argmax(1)).float().mean() < 1e-3
    # class maps only: the row-walking kernel (interpolated source rows cached in registers) == the per-pixel kernel
    for mode, ncls, hw, HW in ((L.RESIZE_BILINEAR_AC, 12, (10, 12), (80, 96)), (L.RESIZE_BILINEAR, 19, (9, 13), (72, 104)),
                               (L.RESIZE_BILINEAR_AC, 27, (7, 5), (53, 41))):
        lg = rnd(2, ncls, *hw, seed=72).to(DEV)
        p_full, _ = ops.resize_argmax(lg, HW, mode, want_logits=True)
        p_rows, _ = ops.resize_argmax(lg, HW, mode, want_logits=False)
        assert torch.equal(p_full, p_rows)
    g = torch.Generator().manual_seed(71)
    label = torch.randint(0, 12, (2, 80, 96), generator=g)
    label[0, :5] = 255
    hist = ops.confusion_hist(pred, label.to(DEV), 12).view(12, 12).cpu()
    ref = O.confusion_hist(pred.cpu().long(), label, 12)
    assert torch.equal(hist.float(), ref)


# ---------------------------------------------------------------- fused CReFF, tensor-core engine
# Tolerance: Q, K, V, P and the classifier operands are rounded to f16 (11-bit significand, the same as
# TF32) and accumulated in fp32 -> 3e-3 relative to the tensor's max magnitude (measured ~5e-4).
MMA_TOL = 3e-3


@pytest.mark.parametrize("k", [3, 5, 7, 9])
@pytest.mark.parametrize("H,W,h,w", [(32, 48, 16, 24), (24, 40, 12, 20), (17, 35, 11, 23)])
def test_creff_mma_prewarped(k, H, W, h, w):
    C = 64
    sd = creff_sd(C)
    hr, lr = rnd(1, C, H, W, seed=51) * 0.6, rnd(1, C, h, w, seed=52) * 0.4
    ref = O.creff(sd, "fuse_attention.", hr, lr, k)
    out_p, _, _ = ops.creff_fused(ops.nchw_to_nhwc(hr.to(DEV)), ops.nchw_to_nhwc(lr.to(DEV)), *creff_args(sd), k,
                                  want_logits=False, lr_layout=L.NHWC, hr_layout=L.NHWC, engine=L.CREFF_MMA_F16)
    assert out_p.shape == ref.shape
    assert rel_err(out_p, ref) < MMA_TOL, rel_err(out_p, ref)
    assert rms_err(out_p, ref) < MMA_TOL / 4


@pytest.mark.parametrize("flow_kind", ["i16", "f64"])
@pytest.mark.parametrize("stride", [1, 8])
@pytest.mark.parametrize("lr_dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_creff_mma_with_mv_warp_and_classifier(flow_kind, stride, lr_dtype):
    from arseg_b200 import synth
    C, ncls, k = 64, 12, 7
    Hm, Wm = 48 * stride, 64 * stride
    H, W, h, w = Hm // stride, Wm // stride, Hm // stride // 2, Wm // stride // 2
    sd = creff_sd(C)
    hr, lr = rnd(1, C, H, W, seed=53) * 0.6, rnd(2, C, h, w, seed=54) * 0.4
    if lr_dtype != torch.float32:
        lr = lr.to(lr_dtype).float()
    wcls, bcls = rnd(ncls, C, seed=55) * 0.2, rnd(ncls, seed=56) * 0.1
    mvs = np.stack([synth.synth_mv_int16(Hm, Wm, 60 + i, distance=5 + 3 * i) for i in range(2)])
    flow64 = torch.from_numpy(mvs.astype(np.float64) / 4.0)
    flow = {"i16": torch.from_numpy(mvs), "f64": flow64}[flow_kind].to(DEV)
    out_p, out_l, out_a = ops.creff_fused(ops.nchw_to_nhwc(hr.to(DEV)), ops.nchw_to_nhwc(lr.to(DEV), lr_dtype), *creff_args(sd), k,
                                          flow=flow, wcls=wcls.to(DEV), bcls=bcls.to(DEV), log_softmax=True, want_argmax=True,
                                          hr_shared=True, lr_layout=L.NHWC, hr_layout=L.NHWC, engine=L.CREFF_MMA_F16)
    for i in range(2):
        fl = O.resize_flow(flow64[i:i + 1], H, W)
        warped = O.warp_feature(hr, fl)
        fused = O.creff(sd, "fuse_attention.", warped, lr[i:i + 1], k)
        logits = F.log_softmax(F.conv2d(fused, wcls.view(ncls, C, 1, 1), bcls), dim=1)
        assert rel_err(out_p[i:i + 1], fused) < MMA_TOL, rel_err(out_p[i:i + 1], fused)
        assert rel_err(out_l[i:i + 1], logits) < MMA_TOL, rel_err(out_l[i:i + 1], logits)
        mism = (out_a[i:i + 1].cpu().long() != logits.argmax(1)).float().mean().item()
        assert mism < 5e-3, mism
        assert torch.equal(out_a[i:i + 1].cpu().long(), out_l[i:i + 1].cpu().argmax(1))


@pytest.mark.parametrize("k", [3, 5, 7, 9])
@pytest.mark.parametrize("C,H,W,h,w", [(128, 24, 40, 12, 20), (256, 17, 35, 11, 23), (512, 9, 21, 5, 11)])
def test_creff_wide_prewarped(k, C, H, W, h, w):
    """C = 64 m > 64 (BiSeNet 256, Cityscapes-PSP 512): the two-launch tensor-core engine (creff_wide.cu) against the
    oracle's MyAttention.forward (model/attention.py:184-213); ragged sizes exercise the partial tiles."""
    sd = creff_sd(C)
    hr, lr = rnd(1, C, H, W, seed=51) * 0.6, rnd(1, C, h, w, seed=52) * 0.4
    ref = O.creff(sd, "fuse_attention.", hr, lr, k)
    out_p, _, _ = ops.creff_fused(ops.nchw_to_nhwc(hr.to(DEV)), ops.nchw_to_nhwc(lr.to(DEV)), *creff_args(sd), k,
                                  want_logits=False, lr_layout=L.NHWC, hr_layout=L.NHWC, engine=L.CREFF_MMA_F16)
    assert out_p.shape == ref.shape
    assert rel_err(out_p, ref) < MMA_TOL, rel_err(out_p, ref)
    assert rms_err(out_p, ref) < MMA_TOL / 4


@pytest.mark.parametrize("flow_kind", ["i16", "f64"])
@pytest.mark.parametrize("C,ncls,lr_dtype", [(256, 12, torch.float32), (256, 12, torch.float16), (512, 19, torch.float32), (128, 19, torch.bfloat16)])
def test_creff_wide_with_mv_warp_and_classifier(flow_kind, C, ncls, lr_dtype):
    """Frame-resolution MV field (stride 8, f64 bilinear resize of the field, evaluation.py:177-180), shared keyframe
    feature, classifier with raw logits (BiSeNet / Cityscapes: no LogSoftmax) and with log-softmax, argmax."""
    from arseg_b200 import synth
    k, stride = 7, 8
    Hm, Wm = 40 * stride, 56 * stride
    H, W, h, w = Hm // stride, Wm // stride, Hm // stride // 2 + 1, Wm // stride // 2
    sd = creff_sd(C)
    hr, lr = rnd(1, C, H, W, seed=53) * 0.6, rnd(2, C, h, w, seed=54) * 0.4
    if lr_dtype != torch.float32:
        lr = lr.to(lr_dtype).float()
    wcls, bcls = rnd(ncls, C, seed=55) * 0.1, rnd(ncls, seed=56) * 0.1
    mvs = np.stack([synth.synth_mv_int16(Hm, Wm, 60 + i, distance=5 + 3 * i) for i in range(2)])
    flow64 = torch.from_numpy(mvs.astype(np.float64) / 4.0)
    flow = {"i16": torch.from_numpy(mvs), "f64": flow64}[flow_kind].to(DEV)
    for logsm in (False, True):
        out_p, out_l, out_a = ops.creff_fused(ops.nchw_to_nhwc(hr.to(DEV)), ops.nchw_to_nhwc(lr.to(DEV), lr_dtype), *creff_args(sd), k,
                                              flow=flow, wcls=wcls.to(DEV), bcls=bcls.to(DEV), log_softmax=logsm, want_argmax=True,
                                              hr_shared=True, lr_layout=L.NHWC, hr_layout=L.NHWC, engine=L.CREFF_MMA_F16)
        for i in range(2):
            fl = O.resize_flow(flow64[i:i + 1], H, W)
            warped = O.warp_feature(hr, fl)
            fused = O.creff(sd, "fuse_attention.", warped, lr[i:i + 1], k)
            logits = F.conv2d(fused, wcls.view(ncls, C, 1, 1), bcls)
            if logsm:
                logits = F.log_softmax(logits, dim=1)
            assert rel_err(out_p[i:i + 1], fused) < MMA_TOL, rel_err(out_p[i:i + 1], fused)
            assert rel_err(out_l[i:i + 1], logits) < MMA_TOL, rel_err(out_l[i:i + 1], logits)
            mism = (out_a[i:i + 1].cpu().long() != logits.argmax(1)).float().mean().item()
            assert mism < 1e-2, mism
            assert torch.equal(out_a[i:i + 1].cpu().long(), out_l[i:i + 1].cpu().argmax(1))


@pytest.mark.parametrize("C", [64, 128])
def test_creff_mma_per_frame_keyframe_features(C):
    """hr_shared = 0: every frame has its own (already warped) keyframe feature, batch 3, ragged sizes -- both tensor-core
    engines (C = 64 march, C = 128 wide) against the oracle frame by frame; frames must not leak into each other."""
    k, H, W, h, w, N = 5, 19, 37, 10, 19, 3
    sd = creff_sd(C)
    hr, lr = rnd(N, C, H, W, seed=61) * 0.6, rnd(N, C, h, w, seed=62) * 0.4
    out_p, _, _ = ops.creff_fused(ops.nchw_to_nhwc(hr.to(DEV)), ops.nchw_to_nhwc(lr.to(DEV)), *creff_args(sd), k,
                                  want_logits=False, lr_layout=L.NHWC, hr_layout=L.NHWC, engine=L.CREFF_MMA_F16, hr_shared=False)
    for i in range(N):
        ref = O.creff(sd, "fuse_attention.", hr[i:i + 1], lr[i:i + 1], k)
        assert rel_err(out_p[i:i + 1], ref) < MMA_TOL, (i, rel_err(out_p[i:i + 1], ref))
    one, _, _ = ops.creff_fused(ops.nchw_to_nhwc(hr[1:2].to(DEV)), ops.nchw_to_nhwc(lr[1:2].to(DEV)), *creff_args(sd), k,
                                want_logits=False, lr_layout=L.NHWC, hr_layout=L.NHWC, engine=L.CREFF_MMA_F16)
    assert torch.equal(one, out_p[1:2])


@pytest.mark.parametrize("k", [3, 7, 9])
@pytest.mark.parametrize("seg_rows", [8, 12])
def test_creff_march_row_segments_are_seamless(k, seg_rows, monkeypatch):
    """The column-marching engine splits a frame into row segments (one CTA each); every pixel's arithmetic is
    independent of the split, so the segmented result must equal the single-segment one bit for bit."""
    from arseg_b200 import synth
    C, ncls, H, W, h, w = 64, 12, 42, 52, 21, 26
    sd = creff_sd(C)
    hr, lr = rnd(1, C, H, W, seed=153) * 0.6, rnd(2, C, h, w, seed=154) * 0.4
    wcls, bcls = rnd(ncls, C, seed=155) * 0.2, rnd(ncls, seed=156) * 0.1
    mvs = torch.from_numpy(np.stack([synth.synth_mv_int16(H, W, 160 + i, distance=4 + 5 * i) for i in range(2)])).to(DEV)

    def run():
        return ops.creff_fused(ops.nchw_to_nhwc(hr.to(DEV)), ops.nchw_to_nhwc(lr.to(DEV)), *creff_args(sd), k, flow=mvs,
                               wcls=wcls.to(DEV), bcls=bcls.to(DEV), log_softmax=True, want_argmax=True, hr_shared=True,
                               lr_layout=L.NHWC, hr_layout=L.NHWC, engine=L.CREFF_MMA_F16)
    monkeypatch.setenv("ARSEG_CREFF_SEG_ROWS", "4096")
    one = run()
    monkeypatch.setenv("ARSEG_CREFF_SEG_ROWS", str(seg_rows))
    seg = run()
    for a, b in zip(one, seg):
        assert torch.equal(a, b)
    flow64 = mvs.cpu().double() / 4.0
    for i in range(2):
        fused = O.creff(sd, "fuse_attention.", O.warp_feature(hr, O.resize_flow(flow64[i:i + 1], H, W)), lr[i:i + 1], k)
        assert rel_err(seg[0][i:i + 1], fused) < MMA_TOL


# ---------------------------------------------------------------- fused CReFF, tcgen05 / TMEM engine (csrc/creff_tc.cu)
# f16 keyframe feature + f16 LR feature (what the 'f16' plan feeds it).  The oracle runs on the f16-rounded inputs in
# fp32, so the tolerance covers the kernel's own roundings: warped-hr / lr_up rows, Q, K, V, P and the classifier input
# are f16 (11-bit significand), every accumulation fp32 -> the MMA_TOL of the mma.sync engine (3e-3 of max magnitude).
def _h(x):
    return x.half().float()


def _tc_run(hr, lr, sd, k, **kw):
    return ops.creff_fused(ops.nchw_to_nhwc(hr.to(DEV), torch.float16), ops.nchw_to_nhwc(lr.to(DEV), torch.float16), *creff_args(sd), k,
                           lr_layout=L.NHWC, hr_layout=L.NHWC, engine=L.CREFF_MMA_F16, **kw)


@pytest.mark.parametrize("k", [3, 5, 7])
@pytest.mark.parametrize("H,W,h,w", [(32, 48, 16, 24), (21, 37, 11, 19), (64, 16, 32, 8), (9, 70, 5, 33)])
def test_creff_tc_prewarped(k, H, W, h, w):
    """MyAttention.forward (model/attention.py:184-213) on tcgen05: ragged sizes exercise partial tiles, strips and segments."""
    C = 64
    sd = creff_sd(C)
    hr, lr = _h(rnd(1, C, H, W, seed=51) * 0.6), _h(rnd(1, C, h, w, seed=52) * 0.4)
    ref = O.creff(sd, "fuse_attention.", hr, lr, k)
    out_p, _, _ = _tc_run(hr, lr, sd, k, want_logits=False)
    assert out_p.shape == ref.shape
    assert rel_err(out_p, ref) < MMA_TOL, rel_err(out_p, ref)
    assert rms_err(out_p, ref) < MMA_TOL / 4


@pytest.mark.parametrize("flow_kind", ["i16", "f64"])
@pytest.mark.parametrize("stride", [1, 8])
@pytest.mark.parametrize("ncls", [12, 19])
def test_creff_tc_with_mv_warp_and_classifier(flow_kind, stride, ncls):
    """evaluation.py:177-183 + MyAttention + final_conv + log-softmax + argmax in one tcgen05 launch."""
    from arseg_b200 import synth
    C, k = 64, 7
    Hm, Wm = 48 * stride, 64 * stride
    H, W, h, w = Hm // stride, Wm // stride, Hm // stride // 2, Wm // stride // 2
    sd = creff_sd(C)
    hr, lr = _h(rnd(1, C, H, W, seed=53) * 0.6), _h(rnd(2, C, h, w, seed=54) * 0.4)
    wcls, bcls = rnd(ncls, C, seed=55) * 0.2, rnd(ncls, seed=56) * 0.1
    mvs = np.stack([synth.synth_mv_int16(Hm, Wm, 60 + i, distance=5 + 3 * i) for i in range(2)])
    flow64 = torch.from_numpy(mvs.astype(np.float64) / 4.0)
    flow = {"i16": torch.from_numpy(mvs), "f64": flow64}[flow_kind].to(DEV)
    out_p, out_l, out_a = _tc_run(hr, lr, sd, k, flow=flow, wcls=wcls.to(DEV), bcls=bcls.to(DEV), log_softmax=True, want_argmax=True,
                                  hr_shared=True)
    for i in range(2):
        fl = O.resize_flow(flow64[i:i + 1], H, W)
        warped = O.warp_feature(hr, fl)
        fused = O.creff(sd, "fuse_attention.", warped, lr[i:i + 1], k)
        logits = F.log_softmax(F.conv2d(fused, wcls.view(ncls, C, 1, 1), bcls), dim=1)
        assert rel_err(out_p[i:i + 1], fused) < MMA_TOL, rel_err(out_p[i:i + 1], fused)
        assert rel_err(out_l[i:i + 1], logits) < MMA_TOL, rel_err(out_l[i:i + 1], logits)
        mism = (out_a[i:i + 1].cpu().long() != logits.argmax(1)).float().mean().item()
        assert mism < 5e-3, mism
        assert torch.equal(out_a[i:i + 1].cpu().long(), out_l[i:i + 1].cpu().argmax(1))


def test_creff_tc_per_frame_keyframe_features():
    """hr_shared = 0: every frame has its own (already warped) keyframe feature; frames must not leak into each other."""
    C, k, H, W, h, w, N = 64, 5, 19, 37, 10, 19, 3
    sd = creff_sd(C)
    hr, lr = _h(rnd(N, C, H, W, seed=61) * 0.6), _h(rnd(N, C, h, w, seed=62) * 0.4)
    out_p, _, _ = _tc_run(hr, lr, sd, k, want_logits=False, hr_shared=False)
    for i in range(N):
        ref = O.creff(sd, "fuse_attention.", hr[i:i + 1], lr[i:i + 1], k)
        assert rel_err(out_p[i:i + 1], ref) < MMA_TOL, (i, rel_err(out_p[i:i + 1], ref))
    one, _, _ = _tc_run(hr[1:2], lr[1:2], sd, k, want_logits=False)
    assert torch.equal(one, out_p[1:2])


@pytest.mark.parametrize("k", [3, 7])
@pytest.mark.parametrize("seg_rows", [8, 24])
def test_creff_tc_row_segments_are_seamless(k, seg_rows, monkeypatch):
    """Row segments (one CTA each) must not change the result.  Runs are bit-reproducible for a given segmentation; across
    segmentations isolated pixels (~0.5 %) differ by one f16 ulp of an attention weight (<= 2.5e-4 here; not a race: identical
    with the tile overlap disabled, tools/tc_det2.py), so the comparison is a tight tolerance, not torch.equal."""
    from arseg_b200 import synth
    C, ncls, H, W, h, w = 64, 12, 42, 52, 21, 26
    sd = creff_sd(C)
    hr, lr = _h(rnd(1, C, H, W, seed=153) * 0.6), _h(rnd(2, C, h, w, seed=154) * 0.4)
    wcls, bcls = rnd(ncls, C, seed=155) * 0.2, rnd(ncls, seed=156) * 0.1
    mvs = torch.from_numpy(np.stack([synth.synth_mv_int16(H, W, 160 + i, distance=4 + 5 * i) for i in range(2)])).to(DEV)

    def run():
        return _tc_run(hr, lr, sd, k, flow=mvs, wcls=wcls.to(DEV), bcls=bcls.to(DEV), log_softmax=True, want_argmax=True, hr_shared=True)
    monkeypatch.setenv("ARSEG_CREFF_SEG_ROWS", "4096")
    one = run()
    again = run()
    for a, b in zip(one, again):
        assert torch.equal(a, b)                          # bit-reproducible
    monkeypatch.setenv("ARSEG_CREFF_SEG_ROWS", str(seg_rows))
    seg = run()
    assert (one[0] - seg[0]).abs().max().item() < 5e-4 and (one[1] - seg[1]).abs().max().item() < 1e-3
    assert ((one[0] - seg[0]).abs().amax(1) > 0).float().mean().item() < 0.03
    assert (one[2] != seg[2]).float().mean().item() < 1e-3
    flow64 = mvs.cpu().double() / 4.0
    for i in range(2):
        fused = O.creff(sd, "fuse_attention.", O.warp_feature(hr, O.resize_flow(flow64[i:i + 1], H, W)), lr[i:i + 1], k)
        assert rel_err(seg[0][i:i + 1], fused) < MMA_TOL


@pytest.mark.parametrize("k", [3, 7])
def test_creff_tcgen05_engine_fp32_keyframe_feature(k):
    """engine = CREFF_TCGEN05 by name with an fp32 NHWC keyframe feature (what a keyframe engine hands over): the pre-pass
    gathers fp32 taps with fp32 weights, so only the stored warped rows are f16."""
    from arseg_b200 import synth
    C, ncls, H, W, h, w = 64, 12, 40, 56, 20, 28
    sd = creff_sd(C)
    hr, lr = rnd(1, C, H, W, seed=71) * 0.6, _h(rnd(2, C, h, w, seed=72) * 0.4)
    wcls, bcls = rnd(ncls, C, seed=73) * 0.2, rnd(ncls, seed=74) * 0.1
    mvs = np.stack([synth.synth_mv_int16(H, W, 80 + i, distance=4 + 6 * i) for i in range(2)])
    flow64 = torch.from_numpy(mvs.astype(np.float64) / 4.0)
    out_p, out_l, out_a = ops.creff_fused(ops.nchw_to_nhwc(hr.to(DEV)), ops.nchw_to_nhwc(lr.to(DEV), torch.float16), *creff_args(sd), k,
                                          flow=torch.from_numpy(mvs).to(DEV), wcls=wcls.to(DEV), bcls=bcls.to(DEV), log_softmax=True,
                                          want_argmax=True, hr_shared=True, lr_layout=L.NHWC, hr_layout=L.NHWC, engine=L.CREFF_TCGEN05)
    for i in range(2):
        fused = O.creff(sd, "fuse_attention.", O.warp_feature(hr, O.resize_flow(flow64[i:i + 1], H, W)), lr[i:i + 1], k)
        logits = F.log_softmax(F.conv2d(fused, wcls.view(ncls, C, 1, 1), bcls), dim=1)
        assert rel_err(out_p[i:i + 1], fused) < MMA_TOL, rel_err(out_p[i:i + 1], fused)
        assert rel_err(out_l[i:i + 1], logits) < MMA_TOL
        assert torch.equal(out_a[i:i + 1].cpu().long(), out_l[i:i + 1].cpu().argmax(1))
    with pytest.raises(L.ArsegError):      # the engine asked for by name does not fall back: fp32 LR feature is unsupported
        ops.creff_fused(ops.nchw_to_nhwc(hr.to(DEV)), ops.nchw_to_nhwc(lr.to(DEV)), *creff_args(sd), k, want_logits=False,
                        lr_layout=L.NHWC, hr_layout=L.NHWC, engine=L.CREFF_TCGEN05)


def _tc_args(hr, lr, sd, k, out_p, flow=None):
    """arseg_creff_args for the tcgen05 engine, built by hand: the C ABI exactly as a non-Python caller sees it."""
    import ctypes as C
    ws = [sd["fuse_attention." + n].reshape(-1).float().to(DEV).contiguous() for n in
          ("lr_query_conv.weight", "lr_query_conv.bias", "hr_key_conv.weight", "hr_key_conv.bias", "hr_value_conv.weight", "hr_value_conv.bias")]
    N, h, w, Cc = lr.shape
    _, H, W, _ = hr.shape
    a = L.CreffArgs(hr.data_ptr(), 1, L.NHWC, L.CREFF_TCGEN05, flow.data_ptr() if flow is not None else None, L.I16 if flow is not None else 0,
                    H if flow is not None else 0, W if flow is not None else 0, lr.data_ptr(), L.NHWC, L.F16, h, w, *[t.data_ptr() for t in ws],
                    None, None, 0, 0, out_p.data_ptr(), None, None, N, Cc, H, W, k, None, 0, ops.dtype_code(hr.dtype), L.CREFF_PHASE_ALL)
    return a, ws


def test_creff_tcgen05_phases_and_workspace_contract():
    """C ABI v6: PHASE_PREPASS followed by PHASE_MAIN on the same arguments equals PHASE_ALL bit for bit; a missing or short
    workspace is refused with a message (no silent fallback), and the other engines treat PHASE_PREPASS as a no-op."""
    import ctypes as C
    from arseg_b200 import synth
    Cc, H, W, h, w, k, N = 64, 40, 56, 20, 28, 7, 2
    sd = creff_sd(Cc)
    lib = L.load()
    hr = ops.nchw_to_nhwc((rnd(1, Cc, H, W, seed=91) * 0.6).to(DEV), torch.float16)
    lr = ops.nchw_to_nhwc((rnd(N, Cc, h, w, seed=92) * 0.4).to(DEV), torch.float16)
    mv = torch.from_numpy(np.stack([synth.synth_mv_int16(H, W, 93 + i, distance=3 + 4 * i) for i in range(N)])).to(DEV)
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    for phases in ((L.CREFF_PHASE_ALL,), (L.CREFF_PHASE_PREPASS, L.CREFF_PHASE_MAIN)):
        out_p = torch.zeros((N, Cc, H, W), device=DEV)
        a, keep = _tc_args(hr, lr, sd, k, out_p, mv)
        need = int(lib.arseg_creff_workspace_bytes(C.byref(a)))
        assert need >= N * H * W * 128
        wsb = torch.empty(need, dtype=torch.uint8, device=DEV)
        a.workspace, a.workspace_bytes = wsb.data_ptr(), need
        for ph in phases:
            a.phase = ph
            L.check(lib.arseg_creff_fused_fwd(C.byref(a), st), "creff")
        torch.cuda.synchronize()
        outs.append(out_p)
    assert torch.equal(outs[0], outs[1]) and float(outs[0].abs().max()) > 0
    # workspace contract
    out_p = torch.zeros((N, Cc, H, W), device=DEV)
    a, keep = _tc_args(hr, lr, sd, k, out_p, mv)
    assert lib.arseg_creff_fused_fwd(C.byref(a), st) != L.OK and b"workspace" in lib.arseg_last_error()
    a.workspace, a.workspace_bytes = wsb.data_ptr(), 1024
    assert lib.arseg_creff_fused_fwd(C.byref(a), st) != L.OK
    a.phase = 7
    assert lib.arseg_creff_fused_fwd(C.byref(a), st) != L.OK and b"phase" in lib.arseg_last_error()
    # the march engine has no pre-pass: PHASE_PREPASS returns OK without touching the outputs
    hr32, lr32 = ops.nchw_to_nhwc((rnd(1, Cc, H, W, seed=91) * 0.6).to(DEV)), ops.nchw_to_nhwc((rnd(N, Cc, h, w, seed=92) * 0.4).to(DEV))
    out_m = torch.full((N, Cc, H, W), 7.0, device=DEV)
    a, keep = _tc_args(hr32, lr32, sd, k, out_m, mv)
    a.engine, a.lr_dtype, a.phase = L.CREFF_MMA_F16, L.F32, L.CREFF_PHASE_PREPASS
    L.check(lib.arseg_creff_fused_fwd(C.byref(a), st), "creff")
    torch.cuda.synchronize()
    assert float(out_m.min()) == 7.0 and float(out_m.max()) == 7.0


def test_creff_tc_rejects_k9_and_mixed_dtypes():
    sd = creff_sd(64)
    hr, lr = rnd(1, 64, 16, 16, seed=1), rnd(1, 64, 8, 8, seed=2)
    with pytest.raises(L.ArsegError):
        _tc_run(hr, lr, sd, 9, want_logits=False)
    with pytest.raises(L.ArsegError):      # f16 keyframe feature with an fp32 LR feature
        ops.creff_fused(ops.nchw_to_nhwc(hr.to(DEV), torch.float16), ops.nchw_to_nhwc(lr.to(DEV)), *creff_args(sd), 7, want_logits=False,
                        lr_layout=L.NHWC, hr_layout=L.NHWC, engine=L.CREFF_MMA_F16)


def test_creff_mma_rejects_unsupported():
    sd = creff_sd(16)
    hr, lr = rnd(1, 16, 16, 16, seed=1), rnd(1, 16, 8, 8, seed=2)
    with pytest.raises(L.ArsegError):
        ops.creff_fused(ops.nchw_to_nhwc(hr.to(DEV)), ops.nchw_to_nhwc(lr.to(DEV)), *creff_args(sd), 7, want_logits=False,
                        lr_layout=L.NHWC, hr_layout=L.NHWC, engine=L.CREFF_MMA_F16)


def test_transpose_cityscapes_full_resolution():
    """1024 x 2048 = 65536 pixel tiles: one more than grid.y allows (the BiSeNet heads transpose the x8 up-sampled logits)."""
    x = rnd(1, 1024, 2048, 3, seed=7).to(DEV)
    y = ops.nhwc_to_nchw(x)
    assert torch.equal(y, x.permute(0, 3, 1, 2).contiguous())
    assert torch.equal(ops.nchw_to_nhwc(y), x)


def test_conv_tcgen05_dilation_beyond_the_halo_box():
    """dilation 6: tap columns reach tx + 12 > the 16-pixel halo box -> the per-tap kernel must run (and match)."""
    Cin, Cout, H, W, dil = 64, 64, 20, 28, 6
    x, w = rnd(1, H, W, Cin, seed=90) * 0.5, rnd(Cout, 3, 3, Cin, seed=91) * 0.05
    ref = F.conv2d(x.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), padding=dil, dilation=dil).permute(0, 2, 3, 1)
    got = ops.conv2d_nhwc(x.to(DEV), w.to(DEV), pad=dil, dil=dil, engine=L.CONV_TC_TF32)
    assert rel_err(got, ref) < 2e-3, rel_err(got, ref)


def test_wrappers_reject_wrong_dtypes():
    x16 = rnd(1, 3, 8, 8, seed=1).half().to(DEV)
    with pytest.raises(RuntimeError):
        ops.resize_nchw(x16, (4, 4), L.RESIZE_BILINEAR_AC)
    with pytest.raises(RuntimeError):
        ops.nchw_to_nhwc(x16)
    with pytest.raises(RuntimeError):
        ops.resize_argmax(x16, (16, 16), L.RESIZE_BILINEAR_AC)
    sd = creff_sd(64)
    hr, lr = rnd(1, 64, 16, 16, seed=1).to(DEV), rnd(1, 64, 8, 8, seed=2).to(DEV)
    args = creff_args(sd)
    with pytest.raises(RuntimeError):
        ops.creff_fused(hr, lr, args[0].double(), *args[1:], 7, want_logits=False)
    with pytest.raises(RuntimeError):
        ops.creff_fused(hr.double(), lr, *args, 7, want_logits=False)


# ---------------------------------------------------------------- data formats either side of the path
@pytest.mark.parametrize("Hi,Wi,Ho,Wo", [(72, 96, 36, 48), (72, 96, 50, 67), (33, 41, 33, 41)])
def test_frame_ingest_u8(Hi, Wi, Ho, Wo):
    """uint8 HWC -> ToTensor + Normalize (dataset/camvid.py:182-185) -> LR down-scale (evaluation.py:186-188) in one kernel."""
    g = torch.Generator().manual_seed(9)
    fr = torch.randint(0, 256, (2, Hi, Wi, 3), generator=g, dtype=torch.uint8)
    ref = O.ingest_u8(fr, ops.CAMVID_MEAN, ops.CAMVID_STD, (Ho, Wo))
    got = ops.frame_ingest_u8(fr.to(DEV), (Ho, Wo))
    assert rel_err(got, ref) < 1e-6


@pytest.mark.parametrize("Fn,H,W", [(1, 16, 24), (5, 37, 53), (11, 64, 96)])
def test_merge_motion_matches_oracle(Fn, H, W):
    """GPU mergeMotion == the oracle's restatement of generate_compressed_dataset_camvid.py:6-56, bit for bit (integers)."""
    from arseg_b200 import synth
    maps = synth.synth_decoder_maps(Fn, H, W, 21)
    ref = O.merge_motion(maps)                                  # [H,W,F+1,2] int32
    got = ops.merge_motion(torch.from_numpy(maps).to(DEV)).cpu().numpy()
    for f in range(1, Fn + 1):
        assert np.array_equal(got[f - 1], ref[:, :, f].astype(np.int16)), f


def test_merge_motion_matches_reference_golden():
    """720x960, 4 frames: the golden recorded from the UNMODIFIED reference function (tests/golden/make_merge_motion_golden.py)."""
    import zlib
    from arseg_b200 import synth
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "merge_motion.npz"))
    Fn, H, W = int(g["F"]), int(g["H"]), int(g["W"])
    maps = synth.synth_decoder_maps(Fn, H, W, int(g["seed"]))
    got = ops.merge_motion(torch.from_numpy(maps).to(DEV)).cpu().numpy()
    assert np.array_equal(got[Fn - 1], g["last"])
    for f in range(1, Fn + 1):
        assert zlib.crc32(np.ascontiguousarray(got[f - 1]).tobytes()) == int(g["crcs"][f])
    # the merged field is the MV input of the non-keyframe path: integer-pel multiples, pointing inside the frame
    assert (got % 4 == 0).all()
