"""Tensor-level wrappers over the C ABI (include/arseg.h).  PyTorch is plumbing only: it owns device
memory and the stream; every op below is one hand-written sm_100a kernel launch."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L

_TORCH2ARSEG = {torch.float32: L.F32, torch.bfloat16: L.BF16, torch.float64: L.F64, torch.int16: L.I16, torch.float16: L.F16}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk_cuda(name: str, *ts: torch.Tensor) -> None:
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("%s: expected a CUDA tensor (no CPU fallback exists)" % name)
        if not t.is_contiguous():
            raise RuntimeError("%s: expected a contiguous tensor" % name)


def _chk_f32(name: str, *ts: torch.Tensor) -> None:
    """The kernels behind these wrappers read fp32: any other dtype would be reinterpreted silently, so it raises (the
    localAttention wrappers do the same, as the upstream extension does)."""
    for t in ts:
        if t is not None and t.dtype != torch.float32:
            raise RuntimeError("%s: expected a float32 tensor, got %s" % (name, t.dtype))


def dtype_code(dt: torch.dtype) -> int:
    return _TORCH2ARSEG[dt]


# ------------------------------------------------------------------ localAttention boundary
def _la_check(name, a, b):
    _chk_cuda(name, a, b)
    if a.dtype != torch.float32 or b.dtype != torch.float32:
        raise RuntimeError("%s: only float32 tensors are supported" % name)
    if a.device != b.device:
        raise RuntimeError("%s: tensors on different devices" % name)


def similar_forward(x_ori, x_loc, kH, kW):
    _la_check("similar_forward", x_ori, x_loc)
    N, Cc, H, W = x_ori.shape
    if tuple(x_loc.shape) != (N, Cc, H, W):
        raise RuntimeError("similar_forward: shape mismatch")
    out = torch.empty((N, H, W, kH * kW), dtype=torch.float32, device=x_ori.device)
    with torch.cuda.device(x_ori.device):
        L.check(L.load().arseg_local_similar_fwd(_p(x_ori), _p(x_loc), _p(out), N, Cc, H, W, kH, kW, _stream()), "similar_forward")
    return out


def weighting_forward(x_ori, x_weight, kH, kW):
    _la_check("weighting_forward", x_ori, x_weight)
    N, Cc, H, W = x_ori.shape
    if tuple(x_weight.shape) != (N, H, W, kH * kW):
        raise RuntimeError("weighting_forward: shape mismatch")
    out = torch.empty_like(x_ori)
    with torch.cuda.device(x_ori.device):
        L.check(L.load().arseg_local_weighting_fwd(_p(x_ori), _p(x_weight), _p(out), N, Cc, H, W, kH, kW, _stream()), "weighting_forward")
    return out


def similar_backward(x, grad_out, kH, kW, is_ori):
    _la_check("similar_backward", x, grad_out)
    N, Cc, H, W = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        L.check(L.load().arseg_local_similar_bwd(_p(x), _p(grad_out), _p(out), N, Cc, H, W, kH, kW, int(bool(is_ori)), _stream()),
                "similar_backward")
    return out


def weighting_backward_ori(x_weight, grad_out, kH, kW):
    _la_check("weighting_backward_ori", x_weight, grad_out)
    N, Cc, H, W = grad_out.shape
    out = torch.empty_like(grad_out)
    with torch.cuda.device(grad_out.device):
        L.check(L.load().arseg_local_weighting_bwd_ori(_p(x_weight), _p(grad_out), _p(out), N, Cc, H, W, kH, kW, _stream()),
                "weighting_backward_ori")
    return out


def weighting_backward_weight(x_ori, grad_out, kH, kW):
    _la_check("weighting_backward_weight", x_ori, grad_out)
    N, Cc, H, W = x_ori.shape
    out = torch.empty((N, H, W, kH * kW), dtype=torch.float32, device=x_ori.device)
    with torch.cuda.device(x_ori.device):
        L.check(L.load().arseg_local_weighting_bwd_weight(_p(x_ori), _p(grad_out), _p(out), N, Cc, H, W, kH, kW, _stream()),
                "weighting_backward_weight")
    return out


# ------------------------------------------------------------------ evaluation.py helpers
def warp_feature(feature: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    """warpFeature(feature[B,C,H,W] fp32, flow[B,H,W,2] f32|f64) (evaluation.py:61-87)."""
    flow = flow.contiguous()
    feature = feature.contiguous()
    _chk_cuda("warpFeature", feature, flow)
    if feature.dtype != torch.float32 or flow.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("warpFeature: feature must be float32 and flow float32/float64")
    B, Cc, H, W = feature.shape
    if tuple(flow.shape) != (B, H, W, 2):
        raise RuntimeError("warpFeature: flow must be [B,H,W,2] at the feature resolution")
    out = torch.empty_like(feature)
    with torch.cuda.device(feature.device):
        L.check(L.load().arseg_warp_feature_nchw(_p(feature), _p(flow), dtype_code(flow.dtype), _p(out), B, Cc, H, W, _stream()),
                "warpFeature")
    return out


def resize_nchw(x: torch.Tensor, size, mode: int) -> torch.Tensor:
    x = x.contiguous()
    _chk_cuda("resize_nchw", x)
    _chk_f32("resize_nchw", x)
    N, Cc, H, W = x.shape
    out = torch.empty((N, Cc, size[0], size[1]), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.load().arseg_resize_nchw_f32(_p(x), _p(out), N * Cc, H, W, size[0], size[1], mode, _stream()), "resize_nchw")
    return out


def resize_nhwc(x: torch.Tensor, size, mode: int, out: Optional[torch.Tensor] = None, coff: int = 0) -> torch.Tensor:
    """Resize an NHWC tensor (fp32 / fp16 / bf16) into `out[..., coff:coff+C]` (a channel slice = fused torch.cat)."""
    _chk_cuda("resize_nhwc", x)
    N, Hi, Wi, C = x.shape
    Ho, Wo = int(size[0]), int(size[1])
    if out is None:
        out = torch.empty((N, Ho, Wo, C), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.load().arseg_resize_nhwc(_p(x), _p(out), dtype_code(x.dtype), N, Hi, Wi, C, Ho, Wo, out.shape[-1], coff, mode, _stream()),
                "resize_nhwc")
    return out


def resize_argmax(logits: torch.Tensor, size, mode: int, want_logits: bool = False):
    logits = logits.contiguous()
    _chk_cuda("resize_argmax", logits)
    _chk_f32("resize_argmax", logits)
    N, K, H, W = logits.shape
    pred = torch.empty((N, size[0], size[1]), dtype=torch.uint8, device=logits.device)
    up = torch.empty((N, K, size[0], size[1]), dtype=torch.float32, device=logits.device) if want_logits else None
    with torch.cuda.device(logits.device):
        L.check(L.load().arseg_resize_argmax_nchw(_p(logits), _p(up), _p(pred), N, K, H, W, size[0], size[1], mode, _stream()),
                "resize_argmax")
    return pred, up


def confusion_hist(pred: torch.Tensor, label: torch.Tensor, n_classes: int, ignore_label: int = 255,
                   hist: Optional[torch.Tensor] = None) -> torch.Tensor:
    _chk_cuda("confusion_hist", pred, label)
    if pred.dtype != torch.uint8 or label.dtype != torch.int64:
        raise RuntimeError("confusion_hist: pred uint8 / label int64 expected")
    if hist is None:
        hist = torch.zeros(n_classes * n_classes, dtype=torch.int64, device=pred.device)
    with torch.cuda.device(pred.device):
        L.check(L.load().arseg_confusion_hist(_p(pred), _p(label), _p(hist), pred.numel(), n_classes, ignore_label, _stream()),
                "confusion_hist")
    return hist


CAMVID_MEAN, CAMVID_STD = (0.39068785, 0.40521392, 0.41434407), (0.29652068, 0.30514979, 0.30080369)   # dataset/camvid.py:184


def frame_ingest_u8(frames: torch.Tensor, size, mean=CAMVID_MEAN, std=CAMVID_STD, mode: int = L.RESIZE_BILINEAR_AC,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """uint8 HWC frames [N,H,W,3] -> ToTensor + Normalize (dataset/camvid.py:182-185) -> bilinear resize to `size`
    (evaluation.py:186-188), fp32 NCHW, one kernel."""
    _chk_cuda("frame_ingest_u8", frames, out)
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise RuntimeError("frame_ingest_u8: expected uint8 [N,H,W,3] frames")
    N, Hi, Wi, _ = frames.shape
    Ho, Wo = int(size[0]), int(size[1])
    if out is None:
        out = torch.empty((N, 3, Ho, Wo), dtype=torch.float32, device=frames.device)
    m, s_ = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    with torch.cuda.device(frames.device):
        L.check(L.load().arseg_frame_ingest_u8(_p(frames), m, s_, _p(out), N, Hi, Wi, Ho, Wo, mode, _stream()), "frame_ingest_u8")
    return out


def merge_motion(maps: torch.Tensor) -> torch.Tensor:
    """mergeMotion (pre-process/generate_compressed_dataset_camvid.py:6-56) on the device: decoder maps int16 [F,H,W,3]
    (mvx, mvy quarter-pel, refIdx; frames 1..F of a GOP) -> merged quarter-pel MV fields int16 [F,H,W,2] pointing to the keyframe
    (plane F-1 = the `.bin` of the frame at keyframe distance F, dataset/camvid.py:624-626)."""
    _chk_cuda("merge_motion", maps)
    if maps.dtype != torch.int16 or maps.dim() != 4 or maps.shape[-1] != 3:
        raise RuntimeError("merge_motion: expected int16 [F,H,W,3] decoder maps")
    Fn, H, W, _ = maps.shape
    out = torch.empty((Fn, H, W, 2), dtype=torch.int16, device=maps.device)
    with torch.cuda.device(maps.device):
        lib = L.load()
        need = int(lib.arseg_merge_motion_workspace_bytes(Fn, H, W))
        ws = torch.empty(need, dtype=torch.uint8, device=maps.device)
        L.check(lib.arseg_merge_motion(_p(maps), _p(ws), need, _p(out), Fn, H, W, _stream()), "merge_motion")
    return out


def nchw_to_nhwc(x: torch.Tensor, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    x = x.contiguous()
    _chk_cuda("nchw_to_nhwc", x)
    _chk_f32("nchw_to_nhwc", x)
    N, Cc, H, W = x.shape
    out = torch.empty((N, H, W, Cc), dtype=dtype, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.load().arseg_nchw_to_nhwc(_p(x), _p(out), dtype_code(dtype), N, Cc, H, W, _stream()), "nchw_to_nhwc")
    return out


def nhwc_to_nchw(x: torch.Tensor) -> torch.Tensor:
    _chk_cuda("nhwc_to_nchw", x)
    N, H, W, Cc = x.shape
    out = torch.empty((N, Cc, H, W), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.load().arseg_nhwc_to_nchw(_p(x), dtype_code(x.dtype), _p(out), N, Cc, H, W, _stream()), "nhwc_to_nchw")
    return out


def conv2d_nhwc(x, w, scale=None, shift=None, residual=None, stride=1, pad=0, dil=1, act=L.ACT_NONE, slope=0.0,
                engine=L.CONV_SIMT_F32, out=None, out_coff=0):
    """x [N,H,W,Cin] NHWC, w [Cout,KH,KW,Cin]; returns NHWC [N,Ho,Wo,Cout] (or writes a channel slice of `out`)."""
    _chk_cuda("conv2d_nhwc", x, w, scale, shift, residual, out)
    _chk_f32("conv2d_nhwc (scale / shift)", scale, shift)
    if residual is not None and residual.dtype != x.dtype:
        raise RuntimeError("conv2d_nhwc: residual dtype %s != activation dtype %s" % (residual.dtype, x.dtype))
    N, Hi, Wi, Cin = x.shape
    Cout, KH, KW, Cin2 = w.shape
    assert Cin == Cin2 and x.dtype == w.dtype
    Ho = (Hi + 2 * pad - dil * (KH - 1) - 1) // stride + 1
    Wo = (Wi + 2 * pad - dil * (KW - 1) - 1) // stride + 1
    if out is None:
        out = torch.empty((N, Ho, Wo, Cout), dtype=x.dtype, device=x.device)
    d = L.ConvDesc(_p(x), _p(w), _p(scale), _p(shift), _p(residual), _p(out), dtype_code(x.dtype), N, Hi, Wi, Cin, Cout,
                   KH, KW, stride, pad, dil, out.shape[-1], out_coff, act, float(slope), engine,
                   1 if (out.dtype == torch.float32 and x.dtype != torch.float32) else 0)
    with torch.cuda.device(x.device):
        L.check(L.load().arseg_conv2d_nhwc(C.byref(d), _stream()), "conv2d_nhwc")
    return out


def conv_stem(x_nchw, w_oihw, scale, shift, dtype=torch.float32):
    """conv1 7x7 s2 p3 + folded BN + ReLU (model/extractors.py:112-114): NCHW fp32 [N,3,H,W] -> NHWC [N,Ho,Wo,Cout]."""
    _chk_cuda("conv_stem", x_nchw, w_oihw, scale, shift)
    _chk_f32("conv_stem", x_nchw, w_oihw, scale, shift)
    N, Ci, H, W = x_nchw.shape
    Cout = w_oihw.shape[0]
    assert Ci == 3 and tuple(w_oihw.shape[1:]) == (3, 7, 7)
    Ho, Wo = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    w = w_oihw.permute(0, 2, 3, 1).contiguous()
    out = torch.empty((N, Ho, Wo, Cout), dtype=dtype, device=x_nchw.device)
    with torch.cuda.device(x_nchw.device):
        L.check(L.load().arseg_conv_stem7x7s2(_p(x_nchw), _p(w), _p(scale), _p(shift), _p(out), dtype_code(dtype), N, H, W, Cout,
                                              _stream()), "conv_stem")
    return out


def creff_fused(hr, lr, wq, bq, wk, bk, wv, bv, k, flow=None, flow_hw=None, wcls=None, bcls=None, log_softmax=False,
                lr_layout=L.NCHW, want_p=True, want_logits=True, want_argmax=False, hr_shared=False, n_frames=None,
                engine=L.CREFF_EXACT_F32, hr_layout=L.NCHW):
    """Fused (MV warp +) CReFF (+ classifier).  hr fp32 NCHW [1|N,C,H,W] (or NHWC [1|N,H,W,C] with hr_layout=NHWC);
    lr NCHW fp32 [N,C,h,w] or NHWC [N,h,w,C].  engine: L.CREFF_EXACT_F32 (fp32 SIMT, NCHW hr) or L.CREFF_MMA_F16
    (mma.sync window attention, TF32-class error, C = 64 m, NHWC hr and lr) or L.CREFF_TCGEN05 (tcgen05 / TMEM engine: C = 64,
    k <= 7, NHWC fp32 or fp16 hr, NHWC fp16 lr; CREFF_MMA_F16 with an fp16 hr is routed to it as well)."""
    _chk_cuda("creff_fused", hr, lr, wq, bq, wk, bk, wv, bv, flow, wcls, bcls)
    _chk_f32("creff_fused (depthwise / classifier weights)", wq, bq, wk, bk, wv, bv, wcls, bcls)
    if hr.dtype not in (torch.float32, torch.float16) or (hr.dtype == torch.float16 and hr_layout != L.NHWC):
        raise RuntimeError("creff_fused: hr must be float32, or float16 NHWC (tcgen05 engine); got %s" % hr.dtype)
    if lr_layout == L.NCHW and lr.dtype != torch.float32:
        raise RuntimeError("creff_fused: an NCHW lr must be float32, got %s" % lr.dtype)
    if flow is not None and flow.dtype not in (torch.int16, torch.float32, torch.float64):
        raise RuntimeError("creff_fused: flow must be int16 (quarter-pel), float32 or float64, got %s" % flow.dtype)
    if hr_layout == L.NHWC:
        Nh, H, W, Cc = hr.shape
    else:
        Nh, Cc, H, W = hr.shape
    if lr_layout == L.NCHW:
        N, C2, h, w = lr.shape
    else:
        N, h, w, C2 = lr.shape
    assert C2 == Cc
    if n_frames is not None:
        assert n_frames == N
    dev = hr.device
    ncls = 0 if wcls is None else wcls.shape[0]
    out_p = torch.empty((N, Cc, H, W), dtype=torch.float32, device=dev) if want_p else None
    out_l = torch.empty((N, ncls, H, W), dtype=torch.float32, device=dev) if (want_logits and ncls) else None
    out_a = torch.empty((N, H, W), dtype=torch.uint8, device=dev) if (want_argmax and ncls) else None
    Hm, Wm = (flow.shape[1], flow.shape[2]) if flow is not None else (0, 0)
    a = L.CreffArgs(_p(hr), int(hr_shared), hr_layout, engine, _p(flow), dtype_code(flow.dtype) if flow is not None else 0, Hm, Wm,
                    _p(lr), lr_layout, dtype_code(lr.dtype), h, w, _p(wq), _p(bq), _p(wk), _p(bk), _p(wv), _p(bv),
                    _p(wcls), _p(bcls), ncls, int(log_softmax), _p(out_p), _p(out_l), _p(out_a), N, Cc, H, W, k, None, 0,
                    dtype_code(hr.dtype))
    with torch.cuda.device(dev):
        lib = L.load()
        need = int(lib.arseg_creff_workspace_bytes(C.byref(a)))
        ws = torch.empty(need, dtype=torch.uint8, device=dev) if need else None      # caller-owned scratch (C > 64 MMA engine)
        a.workspace, a.workspace_bytes = _p(ws), need
        L.check(lib.arseg_creff_fused_fwd(C.byref(a), _stream()), "creff_fused")
    return out_p, out_l, out_a
