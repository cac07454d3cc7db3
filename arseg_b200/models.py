"""Drop-in model classes: same names, constructor arguments, state_dict keys and forward API as the
reference's (evaluation.py:24-36 registry), but every forward runs hand-written sm_100a kernels through
the C ABI.  The torch.nn layers below are PARAMETER HOLDERS only (they give `load_state_dict` the
reference's key names); they are never called.

API mirrored (reference file:line):
  PSPNetWithFuse            model/pspnet.py:103-231      forward / forward_phase1 / forward_phase2
  PSPNet                    model/pspnet.py:49-100       forward -> (logp, cls, p)
  BiSeNetV1WithFuse         model/bisenet.py:481-575
  BiSeNetV1                 model/bisenet.py:419-461     forward -> (out, out16, out32, feat_fuse)
  PSPNetWithFuse_Cityscapes model/pspnet_semseg.py:118-250
  MyAttention               model/attention.py:157-213
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch
from torch import nn

from . import _lib as L
from . import engine as E


def default_precision() -> str:
    return os.environ.get("ARSEG_PRECISION", "tf32")


# ----------------------------------------------------------------------------------------------
# parameter holders
# ----------------------------------------------------------------------------------------------
def _block(cin: int, cout: int, downsample: bool) -> nn.Module:
    m = nn.Module()
    m.conv1 = nn.Conv2d(cin, cout, 3, bias=False)
    m.bn1 = nn.BatchNorm2d(cout)
    m.conv2 = nn.Conv2d(cout, cout, 3, bias=False)
    m.bn2 = nn.BatchNorm2d(cout)
    if downsample:
        m.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, bias=False), nn.BatchNorm2d(cout))
    return m


def _resnet18_holder() -> nn.Module:
    m = nn.Module()
    m.conv1 = nn.Conv2d(3, 64, 7, bias=False)
    m.bn1 = nn.BatchNorm2d(64)
    chans = [64, 64, 128, 256, 512]
    for i in range(1, 5):
        cin, cout = chans[i - 1], chans[i]
        setattr(m, "layer%d" % i, nn.Sequential(_block(cin, cout, cin != cout), _block(cout, cout, False)))
    return m


def _cbr_holder(cin: int, cout: int, ks: int) -> nn.Module:
    m = nn.Module()
    m.conv = nn.Conv2d(cin, cout, ks, bias=False)
    m.bn = nn.BatchNorm2d(cout)
    return m


class MyAttention(nn.Module):
    """CReFF (model/attention.py:157-213).  forward(hr_feat, lr_feat) -> lr_up + attention(hr) as ONE fused kernel."""

    def __init__(self, feat_dim: int, kW: int, kH: int):
        super().__init__()
        if kW != kH:
            raise NotImplementedError("MyAttention: only square windows (the reference always passes kH == kW)")
        self.lr_query_conv = nn.Conv2d(feat_dim, feat_dim, 3, padding=1, groups=feat_dim)
        self.hr_key_conv = nn.Conv2d(feat_dim, feat_dim, 3, padding=1, groups=feat_dim)
        self.hr_value_conv = nn.Conv2d(feat_dim, feat_dim, 3, padding=1, groups=feat_dim)
        self.kW, self.kH = kW, kH
        for ly in (self.lr_query_conv, self.hr_key_conv, self.hr_value_conv):   # model/attention.py:178-182
            nn.init.kaiming_normal_(ly.weight, a=1)
            nn.init.constant_(ly.bias, 0)

    def forward(self, hr_feat: torch.Tensor, lr_feat: torch.Tensor) -> torch.Tensor:
        from . import ops
        sd = {k: getattr(getattr(self, k.split(".")[0]), k.split(".")[1]) for k in
              ("lr_query_conv.weight", "lr_query_conv.bias", "hr_key_conv.weight", "hr_key_conv.bias", "hr_value_conv.weight",
               "hr_value_conv.bias")}                      # attribute access works on DataParallel replicas too
        out_p, _, _ = ops.creff_fused(
            hr_feat.contiguous().float(), lr_feat.contiguous().float(),
            *[sd[k].reshape(-1).contiguous() for k in ("lr_query_conv.weight", "lr_query_conv.bias", "hr_key_conv.weight",
                                                        "hr_key_conv.bias", "hr_value_conv.weight", "hr_value_conv.bias")],
            self.kH, want_logits=False)
        return out_p


class _PlannedNet(nn.Module):
    """Shared machinery: plan cache keyed by (kind, input shape, device, precision, parameter version)."""

    precision: Optional[str] = None

    def _init_plans(self):
        self._plans: Dict[tuple, tuple] = {}

    def _version(self) -> int:
        # nn.DataParallel replicas (evaluation.py:41,54,173) get fresh broadcast copies on every forward: their tensors'
        # _version counters say nothing, so the weight generation (bumped by load_state_dict / .to()) is part of the key
        return getattr(self, "_wgen", 0) + sum(t._version for t in self._sd().values())

    def _sd(self) -> Dict[str, torch.Tensor]:
        """name -> tensor for every parameter and buffer.  Not state_dict(): a replica made by
        nn.parallel.replicate keeps its parameter copies as plain attributes (`_parameters` is empty), so the tensors are
        collected from `_parameters`, the instance dict and `_buffers` of every sub-module."""
        out: Dict[str, torch.Tensor] = {}
        for prefix, m in self.named_modules(remove_duplicate=False):     # aliases (final_conv = cls.4 / conv_out.conv_out) keep both names
            names = list(m._parameters.keys()) + [k for k, v in m.__dict__.items() if isinstance(v, torch.Tensor)] + list(m._buffers.keys())
            for k in names:
                v = m._parameters.get(k)
                if v is None:
                    v = m.__dict__.get(k)
                if v is None:
                    v = m._buffers.get(k)
                if isinstance(v, torch.Tensor):
                    out[(prefix + "." if prefix else "") + k] = v.detach()
        return out

    def load_state_dict(self, *a, **k):
        self._wgen = getattr(self, "_wgen", 0) + 1
        return super().load_state_dict(*a, **k)

    static_outputs = False     # True: forwards return the plan's own output buffers (valid until the next call) instead of clones

    def _out(self, t):
        if t is None or self.static_outputs:
            return t
        return t.clone()

    def _get_plan(self, kind: str, shapes: tuple, device: torch.device, build):
        prec = self.precision or default_precision()
        key = (kind, shapes, str(device), prec)
        ver = self._version()
        hit = self._plans.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        if not device.type == "cuda":
            raise RuntimeError("arseg_b200 models run on CUDA only (no CPU fallback); got device %s" % device)
        with torch.no_grad(), torch.cuda.device(device):
            pl = E.Plan(device, prec)
            io = build(pl, self._sd())
        self._plans[key] = (ver, (pl, io))
        return pl, io

    def _apply(self, fn, *a, **k):  # .cuda()/.to(): cached plans hold device buffers of the old placement
        self._plans = {}
        self._wgen = getattr(self, "_wgen", 0) + 1
        return super()._apply(fn, *a, **k)


def _run_phase2(net: _PlannedNet, p: torch.Tensor, ref_p: torch.Tensor, *, final_prefix: str, log_softmax: bool,
                up: Optional[Tuple[int, int]]):
    """forward_phase2 of the three fused nets: CReFF + final 1x1 conv (+ log-softmax) (+ bilinear x-up)."""
    p = p.contiguous().float()
    ref_p = ref_p.contiguous().float()
    N, Cc, H, W = ref_p.shape
    k = net.fuse_attention.kH

    def build(pl: E.Plan, sd):
        hr = pl.empty((N, Cc, H, W), torch.float32)
        lr = pl.empty(tuple(p.shape), torch.float32)
        lr_in, lr_layout = lr, L.NCHW
        if pl.precision != "fp32" and Cc % 64 == 0:
            # the tensor-core CReFF engines take NHWC operands: one transpose launch for the LR feature (the keyframe
            # feature's is inside Plan.creff), instead of falling back to the exact SIMT kernel (8x slower at 720x960)
            lr_in, lr_layout = pl.to_nhwc(lr, name="lr_nchw_to_nhwc"), L.NHWC
        out_p, out_l, _ = pl.creff(hr, lr_in, sd, "fuse_attention.", k, lr_layout=lr_layout,
                                   wcls=sd[final_prefix + "weight"], bcls=sd[final_prefix + "bias"],
                                   log_softmax=log_softmax)
        if up is not None:
            out_l = pl.resize_nchw(out_l, up[0], up[1], L.RESIZE_BILINEAR, name="out_upsample")
        return hr, lr, out_l, out_p

    pl, (hr, lr, out_l, out_p) = net._get_plan("phase2", (tuple(p.shape), tuple(ref_p.shape)), ref_p.device, build)
    hr.copy_(ref_p)
    lr.copy_(p)
    pl.run()
    if getattr(net, "static_outputs", False):
        # the plan's own output buffers: valid until the next forward_phase2 of this module on this shape (the contract
        # of a CUDA-graph static output); saves two device copies (210 MB at 720x960) per call
        return out_l, out_p
    return out_l.clone(), out_p.clone()


# ----------------------------------------------------------------------------------------------
# CamVid PSPNet-18
# ----------------------------------------------------------------------------------------------
class _PSPBase(_PlannedNet):
    def _holders(self, n_classes, sizes, psp_size, deep_features_size):
        self.feats = _resnet18_holder()
        self.psp = nn.Module()
        self.psp.stages = nn.ModuleList([nn.Sequential(nn.Identity(), nn.Conv2d(psp_size, psp_size, 1, bias=False))
                                         for _ in sizes])
        self.psp.bottleneck = nn.Conv2d(psp_size * (len(sizes) + 1), 1024, 1)
        for name, (ci, co) in (("up_1", (1024, 256)), ("up_2", (256, 64)), ("up_3", (64, 64))):
            m = nn.Module()
            m.conv = nn.Sequential(nn.Conv2d(ci, co, 3, padding=1), nn.BatchNorm2d(co), nn.PReLU())
            setattr(self, name, m)
        self.final_conv = nn.Conv2d(64, n_classes, 1)
        self.classifier = nn.Sequential(nn.Linear(deep_features_size, 256), nn.ReLU(), nn.Linear(256, n_classes))
        self.sizes = tuple(sizes)
        self.n_classes = n_classes

    def _trunk(self, x: torch.Tensor, head: bool):
        x = x.contiguous().float()
        N, _, H, W = x.shape

        def build(pl: E.Plan, sd):
            xin = pl.empty((N, 3, H, W), torch.float32)
            cls, t = E.build_psp_phase1(pl, sd, xin, "", self.sizes, aux=True)
            p_nchw = pl.to_nchw(t, name="p_to_nchw")
            logp = None
            if head:   # model/pspnet.py:95-98: final_conv -> interpolate(H,W) -> LogSoftmax
                lg = pl.conv(t, sd["final_conv.weight"], None, sd["final_conv.bias"], name="final_conv")
                lg = pl.to_nchw(lg)
                if lg.shape[-2:] != (H, W):
                    lg = pl.resize_nchw(lg, H, W, L.RESIZE_BILINEAR_AC)
                logp = pl.log_softmax_nchw(lg)
            return xin, cls, p_nchw, logp

        pl, (xin, cls, p_nchw, logp) = self._get_plan("hr" if head else "phase1", (N, H, W), x.device, build)
        xin.copy_(x)
        pl.run()
        c = self._out
        return c(logp), c(cls), c(p_nchw)


class PSPNet(_PSPBase):
    """HR keyframe net (model/pspnet.py:49-100)."""

    def __init__(self, input_channel=3, n_classes=18, sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024,
                 backend="resnet34", pretrained=True):
        super().__init__()
        if backend != "resnet18" or input_channel != 3:
            raise NotImplementedError("only the resnet18 / RGB configuration of the shipped checkpoints is built")
        self._init_plans()
        self._holders(n_classes, sizes, psp_size, deep_features_size)

    def forward(self, x):
        return self._trunk(x, head=True)


class PSPNetWithFuse(_PSPBase):
    """LR non-keyframe net with CReFF (model/pspnet.py:103-231)."""

    def __init__(self, input_channel=3, n_classes=18, sizes=(1, 2, 3, 6), psp_size=2048, deep_features_size=1024,
                 backend="resnet34", pretrained=True, attention_type="local", atten_k=7):
        super().__init__()
        if backend != "resnet18" or input_channel != 3 or attention_type != "local":
            raise NotImplementedError("only resnet18 / RGB / attention_type='local' (what evaluation.py builds)")
        self._init_plans()
        self._holders(n_classes, sizes, psp_size, deep_features_size)
        self.middle_dim = 64
        self.attention_type = attention_type
        self.fuse_attention = MyAttention(self.middle_dim, kH=atten_k, kW=atten_k)

    def forward_phase1(self, x):
        _, cls, p = self._trunk(x, head=False)
        return cls, p

    def forward_phase2(self, p, ref_p):
        return _run_phase2(self, p, ref_p, final_prefix="final_conv.", log_softmax=True, up=None)

    def forward(self, x, mode="normal", ref_p=None):
        if mode == "normal":
            return self._trunk(x, head=True)
        if mode == "merge":
            out_cls, out_p = self.forward_phase1(x)
            out, out_p = self.forward_phase2(out_p, ref_p)
            return out, out_cls, out_p
        raise ValueError(mode)


# ----------------------------------------------------------------------------------------------
# Cityscapes PSPNet-18 ("semseg" style)
# ----------------------------------------------------------------------------------------------
class PSPNetWithFuse_Cityscapes(_PlannedNet):
    """model/pspnet_semseg.py:118-250 (used for both the HR and the LR net on Cityscapes, evaluation.py:27,34)."""

    def __init__(self, layers=50, bins=(1, 2, 3, 6), dropout=0.1, classes=2, zoom_factor=8, feat_dim=2048, use_ppm=True,
                 criterion=None, pretrained=True, attention_type="local", atten_k=7):
        super().__init__()
        if layers != 18 or not use_ppm or attention_type != "local" or feat_dim != 512:
            raise NotImplementedError("only layers=18 / feat_dim=512 / use_ppm / 'local' (what evaluation.py builds)")
        self._init_plans()
        r = _resnet18_holder()
        self.layer0 = nn.Sequential(r.conv1, r.bn1, nn.ReLU(), nn.Identity())
        self.layer1, self.layer2, self.layer3, self.layer4 = r.layer1, r.layer2, r.layer3, r.layer4
        self.ppm = nn.Module()
        self.ppm.features = nn.ModuleList([
            nn.Sequential(nn.Identity(), nn.Conv2d(feat_dim, feat_dim // len(bins), 1, bias=False),
                          nn.BatchNorm2d(feat_dim // len(bins)), nn.ReLU()) for _ in bins])
        self.cls = nn.Sequential(nn.Conv2d(feat_dim * 2, 512, 3, padding=1, bias=False), nn.BatchNorm2d(512), nn.ReLU(),
                                 nn.Dropout2d(p=dropout), nn.Conv2d(512, classes, 1))
        self.final_conv = self.cls[-1]
        self.aux = nn.Sequential(nn.Conv2d(feat_dim // 2, 256, 3, padding=1, bias=False), nn.BatchNorm2d(256), nn.ReLU(),
                                 nn.Dropout2d(p=dropout), nn.Conv2d(256, classes, 1))
        self.bins = tuple(bins)
        self.zoom_factor = zoom_factor
        self.middle_dim = 512
        self.fuse_attention = MyAttention(self.middle_dim, kH=atten_k, kW=atten_k)

    def _trunk(self, x, full: bool):
        x = x.contiguous().float()
        N, _, H, W = x.shape

        def build(pl: E.Plan, sd):
            xin = pl.empty((N, 3, H, W), torch.float32)
            x3, t = E.build_semseg_phase1(pl, sd, xin, "", self.bins)
            out = aux = None
            if full:   # model/pspnet_semseg.py:198-213
                lg = pl.to_nchw(pl.conv(t, sd["cls.4.weight"], None, sd["cls.4.bias"], name="cls.4"))
                out = pl.resize_nchw(lg, H, W, L.RESIZE_BILINEAR_AC) if self.zoom_factor != 1 else lg
                sc, sh = E.fold_bn(sd, "aux.1.")
                a = pl.conv(x3, sd["aux.0.weight"], sc, sh, pad=1, act=L.ACT_RELU, name="aux.0")
                a = pl.to_nchw(pl.conv(a, sd["aux.4.weight"], None, sd["aux.4.bias"], name="aux.4"))
                aux = pl.resize_nchw(a, H, W, L.RESIZE_BILINEAR_AC) if self.zoom_factor != 1 else a
            return xin, pl.to_nchw(x3, name="x_tmp_to_nchw"), pl.to_nchw(t, name="p_to_nchw"), out, aux

        pl, (xin, x3, p, out, aux) = self._get_plan("hr" if full else "phase1", (N, H, W), x.device, build)
        xin.copy_(x)
        pl.run()
        c = self._out
        return c(x3), c(p), c(out), c(aux)

    def forward_phase1(self, x):
        x3, p, _, _ = self._trunk(x, full=False)
        return x3, p

    def forward_phase2(self, p, ref_p):
        return _run_phase2(self, p, ref_p, final_prefix="final_conv.", log_softmax=False, up=None)

    def forward(self, x, mode="normal", ref_p=None):
        if mode == "normal":
            _, p, out, aux = self._trunk(x, full=True)
            return out, aux, p
        if mode == "merge":
            n, c, h, w = x.shape
            x3, p = self.forward_phase1(x)
            out, p = self.forward_phase2(p, ref_p)
            # model/pspnet_semseg.py:212-214: aux head on x_tmp, resized to the INPUT size
            _, _, _, aux = self._trunk(x, full=True)
            return out, aux, p
        raise ValueError(mode)


# ----------------------------------------------------------------------------------------------
# BiSeNetV1-18
# ----------------------------------------------------------------------------------------------
def _bise_output_holder(cin, mid, ncls) -> nn.Module:
    m = nn.Module()
    m.conv = _cbr_holder(cin, mid, 3)
    m.conv_out = nn.Conv2d(mid, ncls, 1, bias=True)
    return m


class _BiSeBase(_PlannedNet):
    def _holders(self, n_classes, backend, aux_mode):
        if backend != "resnet18":
            raise NotImplementedError("only backend='resnet18' (what evaluation.py builds)")
        self.cp = nn.Module()
        self.cp.resnet = _resnet18_holder()
        for name, cin in (("arm16", 256), ("arm32", 512)):
            a = nn.Module()
            a.conv = _cbr_holder(cin, 128, 3)
            a.conv_atten = nn.Conv2d(128, 128, 1, bias=False)
            a.bn_atten = nn.BatchNorm2d(128)
            setattr(self.cp, name, a)
        self.cp.conv_head32 = _cbr_holder(128, 128, 3)
        self.cp.conv_head16 = _cbr_holder(128, 128, 3)
        self.cp.conv_avg = _cbr_holder(512, 128, 1)
        self.sp = nn.Module()
        self.sp.conv1 = _cbr_holder(3, 64, 7)
        self.sp.conv2 = _cbr_holder(64, 64, 3)
        self.sp.conv3 = _cbr_holder(64, 64, 3)
        self.sp.conv_out = _cbr_holder(64, 128, 1)
        self.ffm = nn.Module()
        self.ffm.convblk = _cbr_holder(256, 256, 1)
        self.ffm.conv = nn.Conv2d(256, 256, 1, bias=False)
        self.ffm.bn = nn.BatchNorm2d(256)
        self.conv_out = _bise_output_holder(256, 256, n_classes)
        self.feat_conv_out = self.conv_out.conv          # aliases, model/bisenet.py:490-491
        self.final_conv = self.conv_out.conv_out
        self.aux_mode = aux_mode
        if aux_mode == "train":
            self.conv_out16 = _bise_output_holder(128, 64, n_classes)
            self.conv_out32 = _bise_output_holder(128, 64, n_classes)
        self.n_classes = n_classes

    def _trunk(self, x, head: bool):
        x = x.contiguous().float()
        N, _, H, W = x.shape
        aux = self.aux_mode == "train"

        def build(pl: E.Plan, sd):
            xin = pl.empty((N, 3, H, W), torch.float32)
            o16, o32, mid = E.build_bisenet_phase1(pl, sd, xin, "", aux=aux)
            out = None
            if head:   # model/bisenet.py:446-448
                lg = pl.conv(mid, sd["conv_out.conv_out.weight"], None, sd["conv_out.conv_out.bias"], name="final_conv")
                _, h8, w8, _ = lg.shape
                out = pl.to_nchw(pl.resize(lg, h8 * 8, w8 * 8, L.RESIZE_BILINEAR, name="out_upsample"))
            return xin, o16, o32, pl.to_nchw(mid, name="mid_to_nchw"), out

        pl, (xin, o16, o32, mid, out) = self._get_plan("hr" if head else "phase1", (N, H, W), x.device, build)
        xin.copy_(x)
        pl.run()
        c = self._out
        return c(out), c(o16), c(o32), c(mid)


class BiSeNetV1(_BiSeBase):
    """HR keyframe net (model/bisenet.py:419-461)."""

    def __init__(self, n_classes, backend, aux_mode="train", *args, **kwargs):
        super().__init__()
        self._init_plans()
        self._holders(n_classes, backend, aux_mode)

    def forward(self, x):
        out, o16, o32, mid = self._trunk(x, head=True)
        if self.aux_mode == "train":
            return out, o16, o32, mid
        if self.aux_mode == "eval":
            return out,
        if self.aux_mode == "pred":
            return out.argmax(dim=1)
        raise NotImplementedError


class BiSeNetV1WithFuse(_BiSeBase):
    """LR non-keyframe net with CReFF (model/bisenet.py:481-575)."""

    def __init__(self, n_classes, backend, aux_mode="train", attention_type="local", atten_k=7, *args, **kwargs):
        super().__init__()
        if attention_type != "local":
            raise NotImplementedError("only attention_type='local'")
        self._init_plans()
        self._holders(n_classes, backend, aux_mode)
        self.middle_dim = 256
        self.fuse_attention = MyAttention(self.middle_dim, kH=atten_k, kW=atten_k)

    def forward_phase1(self, x):
        _, o16, o32, mid = self._trunk(x, head=False)
        if self.aux_mode == "train":
            return o16, o32, mid
        if self.aux_mode == "eval":
            return mid
        raise NotImplementedError

    def forward_phase2(self, middle_feat, ref_p):
        N, Cc, H, W = ref_p.shape
        return _run_phase2(self, middle_feat, ref_p, final_prefix="final_conv.", log_softmax=False, up=(H * 8, W * 8))

    def forward(self, x, mode="normal", ref_p=None):
        if mode == "normal":
            out, o16, o32, mid = self._trunk(x, head=True)
            if self.aux_mode == "train":
                return out, o16, o32, mid
            if self.aux_mode == "eval":
                return out,
            if self.aux_mode == "pred":
                return out.argmax(dim=1)
            raise NotImplementedError
        if mode == "merge":
            ph1 = self.forward_phase1(x)
            mid = ph1[-1] if self.aux_mode == "train" else ph1
            out, out_p = self.forward_phase2(mid, ref_p)
            if self.aux_mode == "train":
                return out, ph1[0], ph1[1], out_p
            return out,
        raise ValueError(mode)


# registry with the reference's keys (evaluation.py:24-36)
models = {
    "camvid-psp18": lambda: PSPNet(sizes=(1, 2, 3, 6), n_classes=12, psp_size=512, deep_features_size=256, backend="resnet18"),
    "camvid-bise18": lambda: BiSeNetV1(n_classes=12, backend="resnet18"),
    "cityscapes-psp18": lambda: PSPNetWithFuse_Cityscapes(bins=(1, 2, 3, 6), classes=19, feat_dim=512, layers=18),
    "cityscapes-bise18": lambda: BiSeNetV1(n_classes=19, backend="resnet18"),
}
models_fuse = {
    "camvid-psp18": lambda: PSPNetWithFuse(sizes=(1, 2, 3, 6), n_classes=12, psp_size=512, deep_features_size=256,
                                           backend="resnet18", atten_k=7),
    "camvid-bise18": lambda: BiSeNetV1WithFuse(n_classes=12, backend="resnet18"),
    "cityscapes-psp18": lambda: PSPNetWithFuse_Cityscapes(bins=(1, 2, 3, 6), classes=19, feat_dim=512, layers=18),
    "cityscapes-bise18": lambda: BiSeNetV1WithFuse(n_classes=19, backend="resnet18"),
}
