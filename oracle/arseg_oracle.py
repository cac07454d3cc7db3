"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not a product path, not a fallback.

CPU restatement (torch-CPU functional ops + numpy / plain C for the two
`localAttention` operators) of the AR-Seg per-non-keyframe inference path.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module.

Every function cites the reference file:line it restates (paths relative to
/root/reference).  Parameters are read from a plain state_dict with the
reference's key names, so the same dict drives the reference
modules, this oracle and the B200 engine.

Pinning status:
  * everything except the two `localAttention` ops is pinned against the
    UNMODIFIED reference modules imported from /root/reference
    (tests/golden/make_golden.py; fixtures in tests/golden/*.npz);
  * `similar_forward` / `weighting_forward` live in an un-vendored pip
    dependency (zzd1992/Image-Local-Attention @ master, requirements.txt:7).
    `weighting_forward` is pinned against the reference's own in-repo
    restatement `f_weighting_cpu` (model/attention.py:75-85);
    `similar_forward` has no working in-repo restatement (`f_similar_cpu`,
    model/attention.py:55-73, is broken) -> **parity unpinned** for that op:
    it follows the published algorithm and the layout `f_weighting_cpu` fixes.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
_HERE = os.path.dirname(os.path.abspath(__file__))

# --------------------------------------------------------------------------
# localAttention operators (external dependency; call sites model/attention.py:18,38)
# --------------------------------------------------------------------------


def similar_forward(x_ori: torch.Tensor, x_loc: torch.Tensor, kH: int, kW: int) -> torch.Tensor:
    """S[n,y,x,i*kW+j] = sum_c Q[n,c,y,x] * K[n,c,y+i-kH//2,x+j-kW//2]; OOB taps give exactly 0.

    Call site model/attention.py:18 (from MyAttention.forward:199).  Output layout
    [N,H,W,kH*kW] as consumed by nn.Softmax(dim=3) (model/attention.py:168,203).
    Shift-and-accumulate over the taps (memory-lean form of the unfold definition).
    """
    N, C, H, W = x_ori.shape
    rh, rw = kH // 2, kW // 2
    kp = F.pad(x_loc, (rw, rw, rh, rh))
    out = x_ori.new_empty(N, H, W, kH * kW)
    for i in range(kH):
        for j in range(kW):
            out[..., i * kW + j] = (x_ori * kp[:, :, i:i + H, j:j + W]).sum(dim=1)
    return out


def weighting_forward(x_ori: torch.Tensor, x_weight: torch.Tensor, kH: int, kW: int) -> torch.Tensor:
    """O[n,c,y,x] = sum_ij V[n,c,y+i-r,x+j-r] * A[n,y,x,i*kW+j]  (zero padded).

    Call site model/attention.py:38 (from MyAttention.forward:207); tap order pinned by
    f_weighting_cpu (model/attention.py:75-85).
    """
    N, C, H, W = x_ori.shape
    rh, rw = kH // 2, kW // 2
    vp = F.pad(x_ori, (rw, rw, rh, rh))
    out = torch.zeros_like(x_ori)
    for i in range(kH):
        for j in range(kW):
            out += vp[:, :, i:i + H, j:j + W] * x_weight[..., i * kW + j].unsqueeze(1)
    return out


def similar_backward(x: torch.Tensor, grad_out: torch.Tensor, kH: int, kW: int, is_ori: bool) -> torch.Tensor:
    """Gradients of similar_forward (call sites model/attention.py:27-28).

    is_ori=True : x is x_loc, returns dL/dx_ori[n,c,y,x] = sum_ij g[n,y,x,ij] * x_loc[n,c,y+i-r,x+j-r]
    is_ori=False: x is x_ori, returns dL/dx_loc[n,c,y',x'] = sum_ij g[n,y'-i+r,x'-j+r,ij] * x_ori[n,c,y'-i+r,x'-j+r]
    """
    N, C, H, W = x.shape
    rh, rw = kH // 2, kW // 2
    out = torch.zeros_like(x)
    if is_ori:
        xp = F.pad(x, (rw, rw, rh, rh))
        for i in range(kH):
            for j in range(kW):
                out += xp[:, :, i:i + H, j:j + W] * grad_out[..., i * kW + j].unsqueeze(1)
    else:
        outp = F.pad(out, (rw, rw, rh, rh))
        for i in range(kH):
            for j in range(kW):
                outp[:, :, i:i + H, j:j + W] += x * grad_out[..., i * kW + j].unsqueeze(1)
        out = outp[:, :, rh:rh + H, rw:rw + W].contiguous()
    return out


def weighting_backward_ori(x_weight: torch.Tensor, grad_out: torch.Tensor, kH: int, kW: int) -> torch.Tensor:
    """dL/dV[n,c,y',x'] = sum_ij A[n,y'-i+r,x'-j+r,ij] * g[n,c,y'-i+r,x'-j+r]  (model/attention.py:47)."""
    N, C, H, W = grad_out.shape
    rh, rw = kH // 2, kW // 2
    outp = F.pad(torch.zeros_like(grad_out), (rw, rw, rh, rh))
    for i in range(kH):
        for j in range(kW):
            outp[:, :, i:i + H, j:j + W] += grad_out * x_weight[..., i * kW + j].unsqueeze(1)
    return outp[:, :, rh:rh + H, rw:rw + W].contiguous()


def weighting_backward_weight(x_ori: torch.Tensor, grad_out: torch.Tensor, kH: int, kW: int) -> torch.Tensor:
    """dL/dA[n,y,x,ij] = sum_c V[n,c,y+i-r,x+j-r] * g[n,c,y,x]  (model/attention.py:48) == similar_forward(g, V)."""
    return similar_forward(grad_out, x_ori, kH, kW)


# ---- plain-C twins (oracle/local_attention.c), used for speed in the CPU baseline ----------
_clib = None


def _load_c():
    global _clib
    if _clib is None:
        path = os.path.join(_HERE, "_build", "liboracle_local_attention.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle` or __graft_entry__.build())")
        lib = ctypes.CDLL(path)
        fp = ctypes.POINTER(ctypes.c_float)
        lib.oracle_similar_forward.argtypes = [fp, fp, fp] + [ctypes.c_int] * 8
        lib.oracle_weighting_forward.argtypes = [fp, fp, fp] + [ctypes.c_int] * 8
        lib.oracle_similar_forward.restype = None
        lib.oracle_weighting_forward.restype = None
        _clib = lib
    return _clib


def _fp(t: torch.Tensor):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))


def _split_threads(fn, total: int):
    """Run fn(lo, hi) over [0,total) on all host cores (ctypes releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    nt = max(1, min(os.cpu_count() or 1, total))
    step = (total + nt - 1) // nt
    with ThreadPoolExecutor(nt) as ex:
        list(ex.map(lambda i: fn(i * step, min(total, (i + 1) * step)), range(nt)))


def similar_forward_c(x_ori, x_loc, kH, kW):
    lib = _load_c()
    x_ori, x_loc = x_ori.contiguous().float(), x_loc.contiguous().float()
    N, C, H, W = x_ori.shape
    out = torch.empty(N, H, W, kH * kW)
    _split_threads(lambda lo, hi: lib.oracle_similar_forward(
        _fp(x_ori), _fp(x_loc), _fp(out), N, C, H, W, kH, kW, lo, hi), N * H)
    return out


def weighting_forward_c(x_ori, x_weight, kH, kW):
    lib = _load_c()
    x_ori, x_weight = x_ori.contiguous().float(), x_weight.contiguous().float()
    N, C, H, W = x_ori.shape
    out = torch.empty(N, C, H, W)
    _split_threads(lambda lo, hi: lib.oracle_weighting_forward(
        _fp(x_ori), _fp(x_weight), _fp(out), N, C, H, W, kH, kW, lo, hi), N * C)
    return out


# --------------------------------------------------------------------------
# evaluation.py pieces
# --------------------------------------------------------------------------


def resize_flow(flow: torch.Tensor, Hf: int, Wf: int) -> torch.Tensor:
    """evaluation.py:177-180.  flow f64 [B,H,W,2] -> [B,Hf,Wf,2]; BOTH components scaled by Hf/H."""
    f = flow.permute(0, 3, 1, 2)
    f = f * Hf / f.shape[-2]
    f = F.interpolate(f, [Hf, Wf], mode="bilinear", align_corners=True)
    return f.permute(0, 2, 3, 1)


def warp_feature(feature: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    """evaluation.py:61-87.  Grid normalised with the align_corners=True formula (80-81) in the flow's
    dtype (float64), cast to fp32 (83), sampled by F.grid_sample with its defaults = bilinear, zeros
    padding, align_corners=False (85)  ->  zero flow is NOT the identity."""
    B, C, H, W = feature.shape
    fl = flow.permute(0, 3, 1, 2)
    xx = torch.arange(0, W).view(1, 1, 1, W).expand(B, 1, H, W)
    yy = torch.arange(0, H).view(1, 1, H, 1).expand(B, 1, H, W)
    grid = torch.cat((xx, yy), 1).float()
    vgrid = grid + fl                                     # promotes to the flow dtype (f64)
    gx = 2.0 * vgrid[:, 0] / max(W - 1, 1) - 1.0
    gy = 2.0 * vgrid[:, 1] / max(H - 1, 1) - 1.0
    vg = torch.stack((gx, gy), dim=-1).float()
    return F.grid_sample(feature, vg, mode="bilinear", padding_mode="zeros", align_corners=False)


# --------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------


def _bn(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                        False, 0.0, 1e-5)


def _conv(sd: SD, p: str, x, stride=1, padding=0, dilation=1, groups=1):
    return F.conv2d(x, sd[p + "weight"], sd.get(p + "bias"), stride, padding, dilation, groups)


def creff(sd: SD, p: str, hr_feat: torch.Tensor, lr_feat: torch.Tensor, k: int = 7,
          use_c: bool = False) -> torch.Tensor:
    """MyAttention.forward, model/attention.py:184-213 (params 161-164)."""
    N, C, H, W = hr_feat.shape
    lr_up = F.interpolate(lr_feat, (H, W), mode="bilinear", align_corners=True)            # :191
    v = _conv(sd, p + "hr_value_conv.", hr_feat, padding=1, groups=C)                        # :194
    kk = _conv(sd, p + "hr_key_conv.", hr_feat, padding=1, groups=C)                         # :196
    q = _conv(sd, p + "lr_query_conv.", lr_up, padding=1, groups=C)                          # :197
    s = (similar_forward_c if use_c else similar_forward)(q, kk, k, k)                      # :199
    a = torch.softmax(s, dim=3)                                                              # :203
    o = (weighting_forward_c if use_c else weighting_forward)(v, a, k, k)                   # :207
    return lr_up + o                                                                         # :210


def _basic_block(sd: SD, p: str, x, stride: int, dil1: int, dil2: int, ds_stride: Optional[int]):
    """BasicBlock.forward model/extractors.py:48-66 (== model/bisenet.py:47-60 up to add order)."""
    out = F.relu(_bn(sd, p + "bn1.", _conv(sd, p + "conv1.", x, stride, dil1, dil1)))
    out = _bn(sd, p + "bn2.", _conv(sd, p + "conv2.", out, 1, dil2, dil2))
    res = x
    if (p + "downsample.0.weight") in sd:
        res = _bn(sd, p + "downsample.1.", _conv(sd, p + "downsample.0.", x, ds_stride))
    return F.relu(out + res)


def resnet18_os8(sd: SD, p: str, x, semseg: bool = False):
    """ResNet.forward model/extractors.py:146-158 built by resnet18 (:340) / _make_layer (:130-144).

    CamVid variant: layer3/4 stride 1, dilation 2/4 on block 1 ONLY (block 0 gets no dilation, :139).
    semseg variant (model/pspnet_semseg.py:145-154): every conv2 of layer3/4 dilated 2/4 (block 0
    included); conv1 keeps what _make_layer gave it (block 0: d=1, block 1: d=2/4).
    Key prefixes: CamVid 'feats.conv1/bn1/layerN', semseg 'layer0.0/layer0.1/layerN'.
    """
    if semseg:
        c1, b1, lp = p + "layer0.0.", p + "layer0.1.", p
    else:
        c1, b1, lp = p + "feats.conv1.", p + "feats.bn1.", p + "feats."
    x = F.relu(_bn(sd, b1, _conv(sd, c1, x, 2, 3)))
    x = F.max_pool2d(x, 3, 2, 1)
    x = _basic_block(sd, lp + "layer1.0.", x, 1, 1, 1, None)
    x = _basic_block(sd, lp + "layer1.1.", x, 1, 1, 1, None)
    x = _basic_block(sd, lp + "layer2.0.", x, 2, 1, 1, 2)
    x = _basic_block(sd, lp + "layer2.1.", x, 1, 1, 1, None)
    x = _basic_block(sd, lp + "layer3.0.", x, 1, 1, 2 if semseg else 1, 1)
    x3 = _basic_block(sd, lp + "layer3.1.", x, 1, 2, 2, None)
    x = _basic_block(sd, lp + "layer4.0.", x3, 1, 1, 4 if semseg else 1, 1)
    x4 = _basic_block(sd, lp + "layer4.1.", x, 1, 4, 4, None)
    return x4, x3


# --------------------------------------------------------------------------
# CamVid PSPNet-18 (model/pspnet.py)
# --------------------------------------------------------------------------


def _psp_module(sd: SD, p: str, feats, sizes=(1, 2, 3, 6)):
    """PSPModule.forward model/pspnet.py:27-31 (stages :22-25)."""
    h, w = feats.shape[2:]
    priors = []
    for i, s in enumerate(sizes):
        t = F.adaptive_avg_pool2d(feats, (s, s))
        t = F.conv2d(t, sd[p + "stages.%d.1.weight" % i])
        priors.append(F.interpolate(t, size=(h, w), mode="bilinear", align_corners=False))
    priors.append(feats)
    return F.relu(_conv(sd, p + "bottleneck.", torch.cat(priors, 1)))


def _psp_up(sd: SD, p: str, x):
    """PSPUpsample.forward model/pspnet.py:43-46: x2 bilinear (align_corners=False) -> 3x3+bias -> BN -> PReLU."""
    h, w = 2 * x.shape[2], 2 * x.shape[3]
    x = F.interpolate(x, size=(h, w), mode="bilinear", align_corners=False)
    x = _bn(sd, p + "conv.1.", _conv(sd, p + "conv.0.", x, 1, 1))
    return F.prelu(x, sd[p + "conv.2.weight"])


def _psp_trunk(sd: SD, p: str, x):
    f, class_f = resnet18_os8(sd, p, x)
    t = _psp_module(sd, p + "psp.", f)
    t = _psp_up(sd, p + "up_1.", t)
    t = _psp_up(sd, p + "up_2.", t)
    t = _psp_up(sd, p + "up_3.", t)
    aux = F.adaptive_max_pool2d(class_f, (1, 1)).view(-1, class_f.size(1))
    cls = F.linear(F.relu(F.linear(aux, sd[p + "classifier.0.weight"], sd[p + "classifier.0.bias"])),
                   sd[p + "classifier.2.weight"], sd[p + "classifier.2.bias"])
    return cls, t


def pspnet_phase1(sd: SD, x, p: str = ""):
    """PSPNetWithFuse.forward_phase1 model/pspnet.py:198-217 -> (cls [N,n_cls], p [N,64,h,w])."""
    return _psp_trunk(sd, p, x)


def pspnet_phase2(sd: SD, lr_p, ref_p, k: int = 7, p: str = "", use_c: bool = False):
    """PSPNetWithFuse.forward_phase2 model/pspnet.py:219-231 -> (log-probs, fused p)."""
    N, C, H, W = ref_p.shape
    fused = creff(sd, p + "fuse_attention.", ref_p, lr_p, k, use_c)
    out = _conv(sd, p + "final_conv.", fused)
    out = F.interpolate(out, (H, W), mode="bilinear", align_corners=True)
    return F.log_softmax(out, dim=1), fused


def pspnet_hr(sd: SD, x, p: str = ""):
    """PSPNet.forward model/pspnet.py:76-100 -> (log-probs, cls, p)."""
    N, C, H, W = x.shape
    cls, t = _psp_trunk(sd, p, x)
    out = _conv(sd, p + "final_conv.", t)
    out = F.interpolate(out, (H, W), mode="bilinear", align_corners=True)
    return F.log_softmax(out, dim=1), cls, t


# --------------------------------------------------------------------------
# Cityscapes PSPNet-18 "semseg" (model/pspnet_semseg.py)
# --------------------------------------------------------------------------


def _ppm(sd: SD, p: str, x, bins=(1, 2, 3, 6)):
    """PPM.forward model/pspnet_semseg.py:25-30: x first, then bins; branch = pool -> 1x1 -> BN -> ReLU -> up(ac=True)."""
    out = [x]
    for i, b in enumerate(bins):
        t = F.adaptive_avg_pool2d(x, b)
        t = F.relu(_bn(sd, p + "features.%d.2." % i, F.conv2d(t, sd[p + "features.%d.1.weight" % i])))
        out.append(F.interpolate(t, x.shape[2:], mode="bilinear", align_corners=True))
    return torch.cat(out, 1)


def semseg_phase1(sd: SD, x, p: str = ""):
    """pspnet_semseg.PSPNetWithFuse.forward_phase1 model/pspnet_semseg.py:223-235 -> (x_tmp, p)."""
    x4, x3 = resnet18_os8(sd, p, x, semseg=True)
    t = _ppm(sd, p + "ppm.", x4)
    t = F.relu(_bn(sd, p + "cls.1.", _conv(sd, p + "cls.0.", t, 1, 1)))
    return x3, t


def semseg_phase2(sd: SD, lr_p, ref_p, k: int = 7, p: str = "", use_c: bool = False):
    """pspnet_semseg.PSPNetWithFuse.forward_phase2 model/pspnet_semseg.py:237-250 (no upsampling here)."""
    fused = creff(sd, p + "fuse_attention.", ref_p, lr_p, k, use_c)
    return _conv(sd, p + "final_conv.", fused), fused


def semseg_hr(sd: SD, x, p: str = ""):
    """pspnet_semseg.PSPNetWithFuse.forward mode='normal' model/pspnet_semseg.py:184-219 -> (x, aux, p)."""
    n, c, h, w = x.shape
    x3, t = semseg_phase1(sd, x, p)
    out = _conv(sd, p + "cls.4.", t)
    out = F.interpolate(out, size=(h, w), mode="bilinear", align_corners=True)
    aux = F.relu(_bn(sd, p + "aux.1.", _conv(sd, p + "aux.0.", x3, 1, 1)))
    aux = _conv(sd, p + "aux.4.", aux)
    aux = F.interpolate(aux, size=(h, w), mode="bilinear", align_corners=True)
    return out, aux, t


# --------------------------------------------------------------------------
# BiSeNetV1-18 (model/bisenet.py)
# --------------------------------------------------------------------------


def _cbr(sd: SD, p: str, x, stride=1, padding=1):
    """ConvBNReLU.forward model/bisenet.py:176-180."""
    return F.relu(_bn(sd, p + "bn.", _conv(sd, p + "conv.", x, stride, padding)))


def _bise_resnet(sd: SD, p: str, x):
    """Resnet18.forward model/bisenet.py:84-94 (standard stride-32 ResNet-18)."""
    x = F.relu(_bn(sd, p + "bn1.", _conv(sd, p + "conv1.", x, 2, 3)))
    x = F.max_pool2d(x, 3, 2, 1)
    x = _basic_block(sd, p + "layer1.0.", x, 1, 1, 1, None)
    x = _basic_block(sd, p + "layer1.1.", x, 1, 1, 1, None)
    x = _basic_block(sd, p + "layer2.0.", x, 2, 1, 1, 2)
    f8 = _basic_block(sd, p + "layer2.1.", x, 1, 1, 1, None)
    x = _basic_block(sd, p + "layer3.0.", f8, 2, 1, 1, 2)
    f16 = _basic_block(sd, p + "layer3.1.", x, 1, 1, 1, None)
    x = _basic_block(sd, p + "layer4.0.", f16, 2, 1, 1, 2)
    f32 = _basic_block(sd, p + "layer4.1.", x, 1, 1, 1, None)
    return f8, f16, f32


def _arm(sd: SD, p: str, x):
    """AttentionRefinementModule.forward model/bisenet.py:252-260."""
    feat = _cbr(sd, p + "conv.", x)
    att = feat.mean(dim=(2, 3), keepdim=True)
    att = _bn(sd, p + "bn_atten.", F.conv2d(att, sd[p + "conv_atten.weight"])).sigmoid()
    return feat * att


def _context_path(sd: SD, p: str, x):
    """ContextPath.forward model/bisenet.py:289-306."""
    f8, f16, f32 = _bise_resnet(sd, p + "resnet.", x)
    avg = _cbr(sd, p + "conv_avg.", f32.mean(dim=(2, 3), keepdim=True), 1, 0)
    f32s = _arm(sd, p + "arm32.", f32) + avg
    f32u = F.interpolate(f32s, scale_factor=2.0, mode="nearest")
    f32u = F.interpolate(f32u, [f16.shape[-2], f16.shape[-1]], mode="bilinear", align_corners=True)
    f32u = _cbr(sd, p + "conv_head32.", f32u)
    f16s = _arm(sd, p + "arm16.", f16) + f32u
    f16u = F.interpolate(f16s, scale_factor=2.0, mode="nearest")
    f16u = _cbr(sd, p + "conv_head16.", f16u)
    return f16u, f32u


def _spatial_path(sd: SD, p: str, x):
    """SpatialPath.forward model/bisenet.py:335-340."""
    x = _cbr(sd, p + "conv1.", x, 2, 3)
    x = _cbr(sd, p + "conv2.", x, 2, 1)
    x = _cbr(sd, p + "conv3.", x, 2, 1)
    return _cbr(sd, p + "conv_out.", x, 1, 0)


def _ffm(sd: SD, p: str, fsp, fcp):
    """FeatureFusionModule.forward model/bisenet.py:387-399."""
    feat = _cbr(sd, p + "convblk.", torch.cat([fsp, fcp], dim=1), 1, 0)
    att = feat.mean(dim=(2, 3), keepdim=True)
    att = _bn(sd, p + "bn.", F.conv2d(att, sd[p + "conv.weight"])).sigmoid()
    return feat * att + feat


def _bise_out(sd: SD, p: str, x, up: int):
    """BiSeNetOutput.forward model/bisenet.py:219-223."""
    x = _conv(sd, p + "conv_out.", _cbr(sd, p + "conv.", x))
    return F.interpolate(x, scale_factor=float(up), mode="bilinear", align_corners=False)


def _bise_trunk(sd: SD, p: str, x):
    cp8, cp16 = _context_path(sd, p + "cp.", x)
    sp = _spatial_path(sd, p + "sp.", x)
    sp = F.interpolate(sp, [cp8.shape[-2], cp8.shape[-1]], mode="bilinear", align_corners=True)
    fuse = _ffm(sd, p + "ffm.", sp, cp8)
    mid = _cbr(sd, p + "conv_out.conv.", fuse)
    return cp8, cp16, mid


def bisenet_phase1(sd: SD, x, p: str = ""):
    """BiSeNetV1WithFuse.forward_phase1 model/bisenet.py:546-563, aux_mode='train' (ctor default :483)."""
    cp8, cp16, mid = _bise_trunk(sd, p, x)
    return _bise_out(sd, p + "conv_out16.", cp8, 8), _bise_out(sd, p + "conv_out32.", cp16, 16), mid


def bisenet_phase2(sd: SD, mid, ref_p, k: int = 7, p: str = "", use_c: bool = False):
    """BiSeNetV1WithFuse.forward_phase2 model/bisenet.py:565-575 -> (raw logits x8 upsampled, fused p)."""
    fused = creff(sd, p + "fuse_attention.", ref_p, mid, k, use_c)
    out = _conv(sd, p + "conv_out.conv_out.", fused)
    return F.interpolate(out, scale_factor=8.0, mode="bilinear", align_corners=False), fused


def bisenet_hr(sd: SD, x, p: str = ""):
    """BiSeNetV1.forward model/bisenet.py:438-461 aux_mode='train' -> (out, out16, out32, feat_fuse)."""
    cp8, cp16, mid = _bise_trunk(sd, p, x)
    out = _conv(sd, p + "conv_out.conv_out.", mid)
    out = F.interpolate(out, scale_factor=8.0, mode="bilinear", align_corners=False)
    return out, _bise_out(sd, p + "conv_out16.", cp8, 8), _bise_out(sd, p + "conv_out32.", cp16, 16), mid


# --------------------------------------------------------------------------
# The per-non-keyframe step: evaluation.py:173-204 (EvalAlterRes.__call__ body)
# --------------------------------------------------------------------------

_PHASES = {
    "camvid-psp18": (pspnet_phase1, pspnet_phase2),
    "camvid-bise18": (bisenet_phase1, bisenet_phase2),
    "cityscapes-psp18": (semseg_phase1, semseg_phase2),
    "cityscapes-bise18": (bisenet_phase1, bisenet_phase2),
}


def nonkey_step(arch: str, sd: SD, imgs: torch.Tensor, ref_p: torch.Tensor, flow: torch.Tensor,
                scale: float = 0.5, k: int = 7, label_size: Optional[Sequence[int]] = None,
                use_c: bool = False, timings: Optional[dict] = None):
    """evaluation.py:176-204 given the keyframe feature `ref_p` (= highres_net(ref)[-1], :173-174).

    Returns (preds int64 [N,H,W], logits [N,n_cls,H,W] after the final interpolation, fused p, lr p).
    """
    import time
    ph1, ph2 = _PHASES[arch]
    t0 = time.perf_counter()
    fl = resize_flow(flow, ref_p.shape[-2], ref_p.shape[-1])                                   # :177-180
    warped = warp_feature(ref_p, fl)                                                            # :183
    t1 = time.perf_counter()
    N, C, H, W = imgs.shape
    new_hw = [int(H * scale), int(W * scale)]                                                   # :186-187
    x = F.interpolate(imgs, new_hw, mode="bilinear", align_corners=True)                        # :188
    lr_p = ph1(sd, x)[-1]                                                                       # :190-191
    t2 = time.perf_counter()
    out, fused = ph2(sd, lr_p, warped, k, use_c=use_c)                                          # :193
    size = list(label_size) if label_size is not None else [H, W]
    logits = F.interpolate(out, size=size, mode="bilinear", align_corners=True)                 # :201-202
    preds = torch.argmax(torch.softmax(logits, dim=1), dim=1)                                   # :203-204
    t3 = time.perf_counter()
    if timings is not None:
        timings["flow_warp_s"] = t1 - t0
        timings["resize_phase1_s"] = t2 - t1
        timings["phase2_post_s"] = t3 - t2
    return preds, logits, fused, lr_p


def confusion_hist(preds: torch.Tensor, label: torch.Tensor, n_classes: int, ignore_label: int = 255):
    """evaluation.py:205-209."""
    keep = label != ignore_label
    return torch.bincount(label[keep] * n_classes + preds[keep], minlength=n_classes ** 2
                          ).view(n_classes, n_classes).float()


# ----------------------------------------------------------------------------------------------
# data formats either side of the path
# ----------------------------------------------------------------------------------------------
def ingest_u8(frames_u8: torch.Tensor, mean, std, size) -> torch.Tensor:
    """uint8 HWC frames [N,H,W,3] -> transforms.ToTensor + Normalize (dataset/camvid.py:182-185) -> the LR down-scale of
    evaluation.py:186-188 (bilinear, align_corners=True)."""
    x = frames_u8.permute(0, 3, 1, 2).to(torch.float32).div(255)                    # ToTensor
    m = torch.tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(1, 3, 1, 1)
    x = (x - m) / s                                                                  # Normalize
    return F.interpolate(x, list(size), mode="bilinear", align_corners=True)


def merge_motion(maps):
    """mergeMotion(workspace_dir, 0, F) (pre-process/generate_compressed_dataset_camvid.py:6-56) on in-memory decoder maps.

    maps: int16 [F,H,W,3] = (mvx, mvy quarter-pel, refIdx) of frames 1..F as dumped by the patched dec265 (`test_%03d.bin`).
    Returns int32 [H,W,F+1,2] exactly like the reference (plane 0 stays -1; planes 1..F = merged quarter-pel MVs)."""
    import numpy as np
    maps = np.asarray(maps)
    Fn, h, w, _ = maps.shape
    max_ref_num = 3                                                                  # :10
    dp = np.ones([h, w, Fn + 1, 3], dtype=np.int32) * -1                             # :12
    k1, j1 = np.meshgrid(range(w), range(h))                                         # :24
    for f1 in range(1, Fn + 1):
        flow = maps[f1 - 1].astype(np.int16).copy()
        intra = np.logical_or(flow[..., 2] < 0, flow[..., 2] >= max_ref_num)         # :20
        flow[intra] = 0                                                              # :21-22
        j2 = np.clip(j1 + np.round(flow[..., 1] / 4).astype(int), 0, h - 1)          # :25, :33 (np.round: half to even)
        k2 = np.clip(k1 + np.round(flow[..., 0] / 4).astype(int), 0, w - 1)          # :26, :34
        f2 = np.maximum(0, f1 - flow[..., 2].astype(int) - 1)                        # :27
        parent = dp[j2, k2, f2]                                                      # [h,w,3]
        father = np.stack([k2, j2, f2], axis=-1).astype(np.int32)
        # :37-48 (the `== 90` branch is dead: refIdx 90 was zeroed by the intra mask above)
        dp[:, :, f1] = np.where((parent[..., 2] != -1)[..., None], parent, father)
    dp[:, :, 1:, 0] = (dp[:, :, 1:, 0] - k1[..., None]) * 4                          # :53
    dp[:, :, 1:, 1] = (dp[:, :, 1:, 1] - j1[..., None]) * 4                          # :54
    return dp[:, :, :, :2]
