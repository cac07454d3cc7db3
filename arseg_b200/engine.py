"""Execution plans: the LR-branch networks of AR-Seg expressed as sequences of C-ABI kernel launches.

A `Plan` is built once per (architecture, batch, resolution, precision): it folds BatchNorm into
per-channel scale/shift, repacks weights OIHW -> [O][KH][KW][I], pre-allocates every activation buffer
and records the launches as pre-bound ctypes calls.  `Plan.run()` replays them on the current stream
(optionally through a CUDA graph), so the steady-state step has no Python-side tensor work.

Precision modes
  'fp32' : fp32 NHWC activations, SIMT implicit-GEMM convs (exact fp32 FMA; the parity gate)
  'tf32' : fp32 NHWC activations, tcgen05 kind::tf32 convs (what cuDNN does for the reference on Ampere+)
  'f16'  : fp16 NHWC activations and weights (11-bit significand = TF32's, saturating at +-65504), tcgen05 kind::f16
           convs at twice the TF32 rate, fp32 accumulate / epilogue
  'bf16' : bf16 NHWC activations, tcgen05 kind::f16 convs, fp32 accumulate / epilogue
Layers the tcgen05 engine does not take (stride 2, Cout < 16, tiny spatial extents) run on the SIMT engine
in the same dtype.

Reference structure restated here (file:line in /root/reference): model/extractors.py:30-66,108-158;
model/pspnet.py:14-46,103-131,198-231; model/bisenet.py:25-113,162-180,207-223,243-306,326-340,360-399,
481-575; model/pspnet_semseg.py:12-30,118-250.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from .ops import dtype_code

SD = Dict[str, torch.Tensor]

_PRECISIONS = {
    "fp32": (torch.float32, L.CONV_SIMT_F32),
    "tf32": (torch.float32, L.CONV_TC_TF32),
    "f16": (torch.float16, L.CONV_TC_F16),
    "bf16": (torch.bfloat16, L.CONV_TC_BF16),
}


def fold_bn(sd: SD, bn: str, conv_bias: Optional[torch.Tensor] = None, eps: float = 1e-5) -> Tuple[torch.Tensor, torch.Tensor]:
    """BatchNorm2d(eval) folded to y = x*scale + shift (conv bias absorbed)."""
    g, b = sd[bn + "weight"].double(), sd[bn + "bias"].double()
    m, v = sd[bn + "running_mean"].double(), sd[bn + "running_var"].double()
    scale = g / torch.sqrt(v + eps)
    shift = b - m * scale
    if conv_bias is not None:
        shift = shift + conv_bias.double() * scale
    return scale.float(), shift.float()


class Plan:
    def __init__(self, device: torch.device, precision: str = "fp32"):
        if precision not in _PRECISIONS:
            raise ValueError("precision must be one of %s" % list(_PRECISIONS))
        self.device = torch.device(device)
        self.precision = precision
        self.act_dtype, self.engine = _PRECISIONS[precision]
        self.lib = L.load()
        self.steps: List[Callable[[int], int]] = []
        self.names: List[str] = []
        self.keep: List[object] = []       # keeps device tensors / ctypes structs alive
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.parts: Optional[List[torch.cuda.CUDAGraph]] = None    # split capture: [phase 1, CReFF + post]
        self.split_at: Optional[int] = None
        self.conv_flops = 0                # 2*MACs of conv/linear layers as executed
        self.step_flops: List[int] = []    # the same, per launch (0 for launches that are not contractions)
        self.n_launches = 0
        # steps that depend only on the plan's INPUTS (not on earlier steps): launch() forks them onto a side stream at the start
        # of the plan and joins right before the step that follows them in the serial order, so a memory-bound pre-pass (the MV
        # warp of the keyframe feature) overlaps the tensor-core-bound LR branch.  launch_range() / profile() run them in place.
        self.hoisted: List[int] = []
        self.overlap = os.environ.get("ARSEG_PLAN_OVERLAP", "1") != "0"
        self._side: Optional[torch.cuda.Stream] = None

    # -- helpers -------------------------------------------------------------------------------
    def dev(self, t: torch.Tensor, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        t = t.detach().to(device=self.device, dtype=dtype or t.dtype).contiguous()
        if t.data_ptr() % 16:       # e.g. a DataParallel replica's parameter: a view into a coalesced broadcast buffer
            t = t.clone()           # (the kernels read weights with 16-byte vector loads)
        self.keep.append(t)
        return t

    def empty(self, shape: Sequence[int], dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        t = torch.empty(tuple(shape), dtype=dtype or self.act_dtype, device=self.device)
        self.keep.append(t)
        return t

    def _add(self, name: str, fn: Callable[[int], int], flops: int = 0, hoist: bool = False, kernels: int = 1) -> None:
        if hoist:
            self.hoisted.append(len(self.steps))
        self.steps.append(fn)
        self.names.append(name)
        self.step_flops.append(int(flops))
        self.conv_flops += int(flops)
        self.n_launches += kernels          # kernel launches (a step is one C-ABI call; the CReFF pre-pass launches two kernels)

    def conv_weight(self, w: torch.Tensor) -> torch.Tensor:
        """OIHW -> [O][KH][KW][I] in the activation dtype."""
        return self.dev(w.detach().permute(0, 2, 3, 1).contiguous(), self.act_dtype)

    # -- ops -----------------------------------------------------------------------------------
    def conv(self, x: torch.Tensor, w_oihw: torch.Tensor, scale=None, shift=None, *, stride=1, pad=0, dil=1,
             act=L.ACT_NONE, slope=0.0, residual=None, out=None, coff=0, name="conv", out_f32=False) -> torch.Tensor:
        N, Hi, Wi, Cin = x.shape
        Cout, _, KH, KW = w_oihw.shape
        Ho = (Hi + 2 * pad - dil * (KH - 1) - 1) // stride + 1
        Wo = (Wi + 2 * pad - dil * (KW - 1) - 1) // stride + 1
        w = self.conv_weight(w_oihw)
        sc = self.dev(scale, torch.float32) if scale is not None else None
        sh = self.dev(shift, torch.float32) if shift is not None else None
        engine = self.engine
        es = 4 if self.act_dtype == torch.float32 else 2
        tc_ok = (engine != L.CONV_SIMT_F32 and stride in (1, 2) and Cin % (128 // es) == 0 and Cout >= 16
                 and 2 * pad == dil * (KH - 1) and 2 * pad == dil * (KW - 1) and Ho * Wo >= 64)
        if not tc_ok:
            engine = L.CONV_SIMT_F32
        # out_f32: fp32 output from a 16-bit plan (tcgen05 engines only; other plans / engines ignore the request)
        out_f32 = bool(out_f32 and tc_ok and self.act_dtype != torch.float32)
        if out is None:
            out = self.empty((N, Ho, Wo, Cout), torch.float32 if out_f32 else None)
        d = L.ConvDesc(x.data_ptr(), w.data_ptr(), sc.data_ptr() if sc is not None else None,
                       sh.data_ptr() if sh is not None else None, residual.data_ptr() if residual is not None else None,
                       out.data_ptr(), dtype_code(self.act_dtype), N, Hi, Wi, Cin, Cout, KH, KW, stride, pad, dil,
                       out.shape[-1], coff, act, float(slope), engine, 1 if out_f32 else 0)
        self.keep.append(d)
        fn = self.lib.arseg_conv2d_nhwc
        self._add("%s[%s %dx%d %d->%d @%dx%d]" % (name, {1: "simt", 2: "tf32", 3: "bf16", 4: "f16"}[engine], KH, KW, Cin, Cout, Ho, Wo),
                  lambda s, d=d: fn(C.byref(d), s), flops=2 * N * Ho * Wo * Cout * Cin * KH * KW)
        return out

    def stem(self, x_nchw: torch.Tensor, w_oihw: torch.Tensor, scale, shift, name="stem") -> torch.Tensor:
        N, _, H, W = x_nchw.shape
        Cout = w_oihw.shape[0]
        Ho, Wo = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        w = self.dev(w_oihw.detach().permute(0, 2, 3, 1).contiguous(), torch.float32)
        sc, sh = self.dev(scale, torch.float32), self.dev(shift, torch.float32)
        out = self.empty((N, Ho, Wo, Cout))
        fn = self.lib.arseg_conv_stem7x7s2
        args = (x_nchw.data_ptr(), w.data_ptr(), sc.data_ptr(), sh.data_ptr(), out.data_ptr(), dtype_code(self.act_dtype),
                N, H, W, Cout)
        self._add(name, lambda s: fn(*args, s), flops=2 * N * Ho * Wo * Cout * 147)
        return out

    def maxpool(self, x: torch.Tensor) -> torch.Tensor:
        N, H, W, Cc = x.shape
        out = self.empty((N, (H - 1) // 2 + 1, (W - 1) // 2 + 1, Cc))
        fn = self.lib.arseg_maxpool3x3s2_nhwc
        args = (x.data_ptr(), out.data_ptr(), dtype_code(self.act_dtype), N, H, W, Cc)
        self._add("maxpool", lambda s: fn(*args, s))
        return out

    def resize(self, x: torch.Tensor, Ho: int, Wo: int, mode: int, out=None, coff=0, name="resize") -> torch.Tensor:
        N, Hi, Wi, Cc = x.shape
        if out is None:
            out = self.empty((N, Ho, Wo, Cc), x.dtype)
        fn = self.lib.arseg_resize_nhwc
        args = (x.data_ptr(), out.data_ptr(), dtype_code(x.dtype), N, Hi, Wi, Cc, Ho, Wo, out.shape[-1], coff, mode)
        self._add(name, lambda s: fn(*args, s))
        return out

    def resize_nchw(self, x: torch.Tensor, Ho: int, Wo: int, mode: int, name="resize_nchw") -> torch.Tensor:
        N, Cc, Hi, Wi = x.shape
        out = self.empty((N, Cc, Ho, Wo), torch.float32)
        fn = self.lib.arseg_resize_nchw_f32
        args = (x.data_ptr(), out.data_ptr(), N * Cc, Hi, Wi, Ho, Wo, mode)
        self._add(name, lambda s: fn(*args, s))
        return out

    def avgpool(self, x: torch.Tensor, Ho: int, Wo: int, fp32_out: bool = False) -> torch.Tensor:
        N, H, W, Cc = x.shape
        odt = torch.float32 if fp32_out else x.dtype
        out = self.empty((N, Ho, Wo, Cc), odt)
        fn = self.lib.arseg_adaptive_avgpool_nhwc
        args = (x.data_ptr(), out.data_ptr(), dtype_code(x.dtype), dtype_code(odt), N, H, W, Cc, Ho, Wo)
        self._add("avgpool%dx%d" % (Ho, Wo), lambda s: fn(*args, s))
        return out

    def pyramid(self, f: torch.Tensor, bins: Sequence[int], weights: Sequence[torch.Tensor], scales=None, shifts=None,
                relu: bool = False, mode: int = L.RESIZE_BILINEAR, stages_first: bool = True, name: str = "psp") -> torch.Tensor:
        """Pyramid pooling module as three launches: all levels of AdaptiveAvgPool2d, the per-level 1x1 convolutions
        (+ folded BN, + ReLU) and bilinear up-sampling + concat with `f` (PSPModule model/pspnet.py:14-31 puts the
        stages first, PPM model/pspnet_semseg.py:12-30 puts x first)."""
        N, H, W, Cf = f.shape
        nlev, Cout = len(bins), weights[0].shape[0]
        B = sum(b * b for b in bins)
        cbins = (C.c_int * nlev)(*bins)
        self.keep.append(cbins)
        pooled = self.empty((N, B, Cf), torch.float32)
        stage = self.empty((N, B, Cout), torch.float32)
        w = self.dev(torch.stack([x.detach().reshape(Cout, Cf) for x in weights]).contiguous(), torch.float32)
        sc = self.dev(torch.stack(list(scales)).contiguous(), torch.float32) if scales is not None else None
        sh = self.dev(torch.stack(list(shifts)).contiguous(), torch.float32) if shifts is not None else None
        out = self.empty((N, H, W, nlev * Cout + Cf))
        f1, f2, f3 = self.lib.arseg_pyramid_pool_nhwc, self.lib.arseg_pyramid_conv1x1, self.lib.arseg_pyramid_upsample_concat
        a1 = (f.data_ptr(), pooled.data_ptr(), dtype_code(f.dtype), N, H, W, Cf, cbins, nlev)
        a2 = (pooled.data_ptr(), w.data_ptr(), sc.data_ptr() if sc is not None else None, sh.data_ptr() if sh is not None else None,
              int(relu), stage.data_ptr(), N, Cf, Cout, cbins, nlev)
        s_off, f_off = (0, nlev * Cout) if stages_first else (Cf, 0)
        a3 = (stage.data_ptr(), f.data_ptr(), out.data_ptr(), dtype_code(f.dtype), N, H, W, Cout, Cf, s_off, f_off, mode, cbins, nlev)
        self._add(name + ".pool", lambda s: f1(*a1, s))
        self._add(name + ".conv1x1", lambda s: f2(*a2, s), flops=2 * N * B * Cout * Cf)
        self._add(name + ".upsample_concat", lambda s: f3(*a3, s))
        return out

    def gmaxpool(self, x: torch.Tensor) -> torch.Tensor:
        N, H, W, Cc = x.shape
        out = self.empty((N, Cc), torch.float32)
        fn = self.lib.arseg_global_maxpool_nhwc
        args = (x.data_ptr(), out.data_ptr(), dtype_code(x.dtype), N, H, W, Cc)
        self._add("gmaxpool", lambda s: fn(*args, s))
        return out

    def linear(self, x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], relu: bool, name="linear") -> torch.Tensor:
        N, K = x.shape[0], x.numel() // x.shape[0]
        M = w.shape[0]
        wd = self.dev(w.reshape(M, K), torch.float32)
        bd = self.dev(b, torch.float32) if b is not None else None
        out = self.empty((N, M), torch.float32)
        fn = self.lib.arseg_linear_f32
        args = (x.data_ptr(), wd.data_ptr(), bd.data_ptr() if bd is not None else None, out.data_ptr(), N, K, M, int(relu))
        self._add(name, lambda s: fn(*args, s), flops=2 * N * K * M)
        return out

    def gate(self, feat: torch.Tensor, gate: torch.Tensor, gs, gb, add_identity=False, add_chan=None, add_pix=None,
             name="gate") -> torch.Tensor:
        N, H, W, Cc = feat.shape
        gsd = self.dev(gs, torch.float32) if gs is not None else None
        gbd = self.dev(gb, torch.float32) if gb is not None else None
        out = self.empty((N, H, W, Cc), feat.dtype)
        fn = self.lib.arseg_gate_nhwc
        args = (feat.data_ptr(), gate.data_ptr(), gsd.data_ptr() if gsd is not None else None,
                gbd.data_ptr() if gbd is not None else None, int(add_identity),
                add_chan.data_ptr() if add_chan is not None else None, add_pix.data_ptr() if add_pix is not None else None,
                out.data_ptr(), dtype_code(feat.dtype), N, H, W, Cc)
        self._add(name, lambda s: fn(*args, s))
        return out

    def frame_ingest_u8(self, frames: torch.Tensor, Ho: int, Wo: int, mean, std, name="frame_ingest_u8") -> torch.Tensor:
        """uint8 HWC frames -> ToTensor + Normalize + bilinear (align_corners=True) resize, fp32 NCHW (csrc/ingest.cu)."""
        N, Hi, Wi, _ = frames.shape
        out = self.empty((N, 3, Ho, Wo), torch.float32)
        m, s_ = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
        self.keep += [m, s_]
        fn = self.lib.arseg_frame_ingest_u8
        args = (frames.data_ptr(), m, s_, out.data_ptr(), N, Hi, Wi, Ho, Wo, L.RESIZE_BILINEAR_AC)
        self._add(name, lambda s: fn(*args, s))
        return out

    def to_nhwc(self, x: torch.Tensor, dtype: Optional[torch.dtype] = None, name="nchw_to_nhwc") -> torch.Tensor:
        """fp32 NCHW -> NHWC of `dtype` (default: the plan's activation dtype)."""
        N, Cc, H, W = x.shape
        out = self.empty((N, H, W, Cc), dtype or self.act_dtype)
        fn = self.lib.arseg_nchw_to_nhwc
        args = (x.data_ptr(), out.data_ptr(), dtype_code(out.dtype), N, Cc, H, W)
        self._add(name, lambda s: fn(*args, s))
        return out

    def to_nchw(self, x: torch.Tensor, name="nhwc_to_nchw", out: Optional[torch.Tensor] = None) -> torch.Tensor:
        N, H, W, Cc = x.shape
        if out is None:
            out = self.empty((N, Cc, H, W), torch.float32)
        elif tuple(out.shape) != (N, Cc, H, W) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError("to_nchw: `out` must be a contiguous fp32 [%d,%d,%d,%d] tensor" % (N, Cc, H, W))
        fn = self.lib.arseg_nhwc_to_nchw
        args = (x.data_ptr(), dtype_code(x.dtype), out.data_ptr(), N, Cc, H, W)
        self._add(name, lambda s: fn(*args, s))
        return out

    def creff(self, hr: torch.Tensor, lr: torch.Tensor, sd: SD, prefix: str, k: int, *, flow: Optional[torch.Tensor] = None,
              hr_shared=False, lr_layout=L.NHWC, wcls=None, bcls=None, log_softmax=False, want_p=True, want_logits=True,
              want_argmax=False, name="creff_fused", engine=None, hr_layout=L.NCHW, hoist_prepass=False):
        """Fused MV-warp + CReFF (+ classifier).  hr NCHW fp32 [1|N,C,H,W]; lr [N,h,w,C] NHWC (or NCHW fp32).
        engine None: exact fp32 SIMT kernel in 'fp32' plans; tensor-core window attention (f16 operands, fp32
        accumulate) in 'tf32'/'f16'/'bf16' plans when C is a multiple of 64 (hr is first converted to NHWC by one transpose
        launch; C = 64: column-marching engine, or the tcgen05 engine when lr is f16; C > 64: the two-launch wide engine with a
        plan-owned workspace).  hoist_prepass: hr and flow are INPUTS of the plan (not produced by earlier steps), so the tcgen05
        engine's MV-warp pre-pass may run on the side stream from the start of the plan."""
        if hr_layout == L.NHWC:         # the keyframe feature already in the internal layout (fp32 [1|N,H,W,C]): no transpose launch
            _, H, W, Cc = hr.shape
        else:
            _, Cc, H, W = hr.shape
        if engine is None:
            mma_ok = (self.precision != "fp32" and Cc % 64 == 0 and Cc <= 1024 and lr_layout == L.NHWC and k in (3, 5, 7, 9)
                      and (wcls is None or wcls.shape[0] <= 32) and min(H, W) >= 2)
            engine = L.CREFF_MMA_F16 if mma_ok else L.CREFF_EXACT_F32
        hr_in_nhwc = hr_layout == L.NHWC
        # C = 64 with an f16 LR feature (the 'f16' plan): the tcgen05 / TMEM engine (csrc/creff_tc.cu, k <= 7), 3.15 ms per 11
        # CamVid frames against 3.34 ms for the mma.sync march engine, and 0.39 ms of that is the MV-warp pre-pass, which this
        # plan runs on a side stream under the LR branch.  ARSEG_CREFF_TC=0 keeps the march engine (A/B measurements).
        tc = (engine == L.CREFF_MMA_F16 and Cc == 64 and k <= 7 and lr_layout == L.NHWC and lr.dtype == torch.float16
              and os.environ.get("ARSEG_CREFF_TC", "1") != "0")
        if engine in (L.CREFF_MMA_F16, L.CREFF_TCGEN05):
            if hr_in_nhwc:
                if hr.dtype != torch.float32 and not (tc and hr.dtype == torch.float16):
                    raise ValueError("creff: an NHWC keyframe feature must be fp32 (or fp16 for the tcgen05 engine)")
            else:
                # the transpose writes what the engine wants to read: fp32, or fp16 for the tcgen05 engine (its pre-pass stores fp16
                # rows anyway, and an fp16 keyframe feature is half the L2 footprint of the gather: 0.36 instead of 0.70 ms)
                hr_nhwc = self.empty((hr.shape[0], H, W, Cc), torch.float16 if tc else torch.float32)
                fn_t = self.lib.arseg_nchw_to_nhwc
                targs = (hr.data_ptr(), hr_nhwc.data_ptr(), L.F16 if tc else L.F32, hr.shape[0], Cc, H, W)
                self._add("hr_nchw_to_nhwc", lambda s: fn_t(*targs, s), hoist=hoist_prepass and tc)
                hr, hr_layout = hr_nhwc, L.NHWC
            if tc:
                engine = L.CREFF_TCGEN05
            name = name + ("_tc" if tc else "_mma")
        elif hr_in_nhwc:
            raise ValueError("creff: the exact fp32 engine takes the keyframe feature as NCHW")
        if lr_layout == L.NHWC:
            N, h, w, _ = lr.shape
        else:
            N, _, h, w = lr.shape
        f32 = torch.float32
        ws = [self.dev(sd[prefix + n].reshape(-1), f32) for n in
              ("lr_query_conv.weight", "lr_query_conv.bias", "hr_key_conv.weight", "hr_key_conv.bias",
               "hr_value_conv.weight", "hr_value_conv.bias")]
        ncls = 0
        wc = bc = None
        if wcls is not None:
            ncls = wcls.shape[0]
            wc = self.dev(wcls.reshape(ncls, Cc), f32)
            bc = self.dev(bcls, f32) if bcls is not None else None
        out_p = self.empty((N, Cc, H, W), f32) if want_p else None
        out_l = self.empty((N, ncls, H, W), f32) if (want_logits and ncls) else None
        out_a = self.empty((N, H, W), torch.uint8) if (want_argmax and ncls) else None
        Hm, Wm = (flow.shape[1], flow.shape[2]) if flow is not None else (0, 0)
        a = L.CreffArgs(hr.data_ptr(), int(hr_shared), hr_layout, engine, flow.data_ptr() if flow is not None else None,
                        dtype_code(flow.dtype) if flow is not None else 0, Hm, Wm, lr.data_ptr(), lr_layout,
                        dtype_code(lr.dtype), h, w, *[t.data_ptr() for t in ws],
                        wc.data_ptr() if wc is not None else None, bc.data_ptr() if bc is not None else None, ncls,
                        int(log_softmax), out_p.data_ptr() if out_p is not None else None,
                        out_l.data_ptr() if out_l is not None else None, out_a.data_ptr() if out_a is not None else None,
                        N, Cc, H, W, k, None, 0, dtype_code(hr.dtype), L.CREFF_PHASE_ALL)
        need = int(self.lib.arseg_creff_workspace_bytes(C.byref(a)))
        if need:
            wsb = self.empty((need,), torch.uint8)
            a.workspace, a.workspace_bytes = wsb.data_ptr(), need
        self.keep.append(a)
        fn = self.lib.arseg_creff_fused_fwd
        if tc:
            # two calls on the same arguments: the workspace pre-pass (MV warp of the keyframe feature: reads hr + flow only, so it
            # may run from the start of the plan when those are plan inputs) and the attention kernel
            pre = L.CreffArgs()
            C.memmove(C.byref(pre), C.byref(a), C.sizeof(a))
            pre.phase, a.phase = L.CREFF_PHASE_PREPASS, L.CREFF_PHASE_MAIN
            self.keep.append(pre)
            self._add(name + "_prewarp", lambda s, a=pre: fn(C.byref(a), s), hoist=hoist_prepass, kernels=2)
        self._add(name, lambda s, a=a: fn(C.byref(a), s))
        return out_p, out_l, out_a

    def log_softmax_nchw(self, x: torch.Tensor) -> torch.Tensor:
        N, K, H, W = x.shape
        out = self.empty((N, K, H, W), torch.float32)
        fn = self.lib.arseg_log_softmax_nchw
        args = (x.data_ptr(), out.data_ptr(), N, K, H, W)
        self._add("log_softmax", lambda s: fn(*args, s))
        return out

    def resize_argmax(self, logits: torch.Tensor, Ho: int, Wo: int, mode: int, want_logits=False):
        N, K, H, W = logits.shape
        pred = self.empty((N, Ho, Wo), torch.uint8)
        up = self.empty((N, K, Ho, Wo), torch.float32) if want_logits else None
        fn = self.lib.arseg_resize_argmax_nchw
        args = (logits.data_ptr(), up.data_ptr() if up is not None else None, pred.data_ptr(), N, K, H, W, Ho, Wo, mode)
        self._add("resize_argmax", lambda s: fn(*args, s))
        return pred, up

    # -- execution -----------------------------------------------------------------------------
    def launch(self) -> None:
        """Enqueue every kernel of the plan on the current stream (hoisted steps: on a side stream forked at the start of the
        plan and joined before the step that consumes them; inside a stream capture this becomes a parallel graph branch)."""
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream()
            s = main.cuda_stream
            hoisted = self.hoisted if self.overlap else []
            join_at = {}
            if hoisted:
                if self._side is None:
                    self._side = torch.cuda.Stream(priority=0)          # lowest priority: fills what the main branch leaves free
                fork = torch.cuda.Event()
                fork.record(main)
                self._side.wait_event(fork)
                ss = self._side.cuda_stream
                for i in hoisted:
                    rc = self.steps[i](ss)
                    if rc != L.OK:
                        L.check(rc, self.names[i])
                    ev = torch.cuda.Event()
                    ev.record(self._side)
                    join_at[i + 1] = ev
            for i, (name, fn) in enumerate(zip(self.names, self.steps)):
                if i in join_at:
                    main.wait_event(join_at[i])
                if hoisted and i in hoisted:
                    continue
                rc = fn(s)
                if rc != L.OK:
                    L.check(rc, name)
            if len(self.steps) in join_at:
                main.wait_event(join_at[len(self.steps)])

    def launch_range(self, lo: int, hi: int) -> None:
        with torch.cuda.device(self.device):
            s = torch.cuda.current_stream().cuda_stream
            for name, fn in zip(self.names[lo:hi], self.steps[lo:hi]):
                rc = fn(s)
                if rc != L.OK:
                    L.check(rc, name)

    def _capture_stream(self) -> torch.cuda.Stream:
        # the main branch is captured on a high-priority stream: kernel nodes keep their stream's priority, so the block scheduler
        # serves the LR-branch kernels first and the hoisted (priority 0) pre-pass takes the SM capacity they leave
        if getattr(self, "_cap", None) is None:
            self._cap = torch.cuda.Stream(priority=-1)
        return self._cap

    def capture(self, split_at: Optional[int] = None) -> None:
        """Capture the launch sequence into a CUDA graph (replayed by run()).  split_at = i: two graphs, launches [0, i) and
        [i, n) -- run_part(0) / run_part(1) replay them separately (the caller may put a stream dependency, e.g. the NCCL
        broadcast of the keyframe feature, between phase 1 and the CReFF launches)."""
        if split_at is not None and 0 < split_at < len(self.steps):
            with torch.cuda.device(self.device):
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self.launch()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                parts = []
                for lo, hi in ((0, split_at), (split_at, len(self.steps))):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self._capture_stream()):
                        self.launch_range(lo, hi)
                    parts.append(g)
                self.parts, self.split_at = parts, split_at
            return
        with torch.cuda.device(self.device):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.launch()          # warm-up outside capture (function attributes, TMA entry point)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._capture_stream()):
                self.launch()
            self.graph = g

    def run(self) -> None:
        if self.parts is not None:
            self.parts[0].replay()
            self.parts[1].replay()
        elif self.graph is not None:
            self.graph.replay()
        else:
            self.launch()

    def run_part(self, i: int) -> None:
        """Replay one half of a split capture (0: up to the split, 1: from the split on)."""
        if self.parts is None:
            raise RuntimeError("plan was not captured with split_at")
        self.parts[i].replay()

    def profile(self, iters: int = 5, warmup: int = 2) -> List[Tuple[str, float]]:
        """Per-kernel device time (ms, mean over `iters`) with CUDA events on the launching stream."""
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream()
            s = stream.cuda_stream
            n = len(self.steps)
            tot = [0.0] * n
            for it in range(warmup + iters):
                evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
                evs[0].record(stream)
                for i, fn in enumerate(self.steps):
                    rc = fn(s)
                    if rc != L.OK:
                        L.check(rc, self.names[i])
                    evs[i + 1].record(stream)
                torch.cuda.synchronize()
                if it >= warmup:
                    for i in range(n):
                        tot[i] += evs[i].elapsed_time(evs[i + 1])
            return [(self.names[i], tot[i] / iters) for i in range(n)]


# ==============================================================================================
# network builders (functional over a reference-keyed state_dict)
# ==============================================================================================

def _cbr(pl: Plan, sd: SD, p: str, x, *, stride=1, pad=1, out=None, coff=0, name=None):
    """ConvBNReLU (model/bisenet.py:162-180): conv(no bias) + BN + ReLU."""
    sc, sh = fold_bn(sd, p + "bn.")
    return pl.conv(x, sd[p + "conv.weight"], sc, sh, stride=stride, pad=pad, act=L.ACT_RELU, out=out, coff=coff,
                   name=name or p.rstrip("."))


def _basic_block(pl: Plan, sd: SD, p: str, x, stride: int, dil1: int, dil2: int, ds_stride: int):
    """BasicBlock (model/extractors.py:36-66 / model/bisenet.py:31-60)."""
    s1, b1 = fold_bn(sd, p + "bn1.")
    t = pl.conv(x, sd[p + "conv1.weight"], s1, b1, stride=stride, pad=dil1, dil=dil1, act=L.ACT_RELU, name=p + "conv1")
    res = x
    if (p + "downsample.0.weight") in sd:
        sd_, bd_ = fold_bn(sd, p + "downsample.1.")
        res = pl.conv(x, sd[p + "downsample.0.weight"], sd_, bd_, stride=ds_stride, pad=0, name=p + "downsample")
    s2, b2 = fold_bn(sd, p + "bn2.")
    return pl.conv(t, sd[p + "conv2.weight"], s2, b2, pad=dil2, dil=dil2, act=L.ACT_RELU, residual=res, name=p + "conv2")


def _resnet_os8(pl: Plan, sd: SD, p: str, x_nchw, semseg: bool):
    """Dilated ResNet-18, output stride 8 (model/extractors.py:108-158; semseg rewrite model/pspnet_semseg.py:145-154)."""
    if semseg:
        c1, b1, lp = p + "layer0.0.", p + "layer0.1.", p
    else:
        c1, b1, lp = p + "feats.conv1.", p + "feats.bn1.", p + "feats."
    sc, sh = fold_bn(sd, b1)
    x = pl.stem(x_nchw, sd[c1 + "weight"], sc, sh)
    x = pl.maxpool(x)
    x = _basic_block(pl, sd, lp + "layer1.0.", x, 1, 1, 1, 1)
    x = _basic_block(pl, sd, lp + "layer1.1.", x, 1, 1, 1, 1)
    x = _basic_block(pl, sd, lp + "layer2.0.", x, 2, 1, 1, 2)
    x = _basic_block(pl, sd, lp + "layer2.1.", x, 1, 1, 1, 1)
    x = _basic_block(pl, sd, lp + "layer3.0.", x, 1, 1, 2 if semseg else 1, 1)
    x3 = _basic_block(pl, sd, lp + "layer3.1.", x, 1, 2, 2, 1)
    x = _basic_block(pl, sd, lp + "layer4.0.", x3, 1, 1, 4 if semseg else 1, 1)
    x4 = _basic_block(pl, sd, lp + "layer4.1.", x, 1, 4, 4, 1)
    return x4, x3


def build_psp_phase1(pl: Plan, sd: SD, x_nchw: torch.Tensor, p: str = "", sizes=(1, 2, 3, 6), aux: bool = True):
    """PSPNetWithFuse.forward_phase1 (model/pspnet.py:198-217) -> (cls fp32 [N,ncls] | None, p NHWC [N,h,w,64])."""
    f, x3 = _resnet_os8(pl, sd, p, x_nchw, semseg=False)
    N, h, w, Cf = f.shape
    # PSPModule stages + concat (model/pspnet.py:22-30)
    cat = pl.pyramid(f, sizes, [sd[p + "psp.stages.%d.1.weight" % i] for i in range(len(sizes))], mode=L.RESIZE_BILINEAR,
                     stages_first=True, name="psp")
    t = pl.conv(cat, sd[p + "psp.bottleneck.weight"], None, sd[p + "psp.bottleneck.bias"], act=L.ACT_RELU, name="psp.bottleneck")
    for u in ("up_1.", "up_2.", "up_3."):                           # PSPUpsample (model/pspnet.py:34-46)
        Nn, hh, ww, _ = t.shape
        t = pl.resize(t, 2 * hh, 2 * ww, L.RESIZE_BILINEAR, name=u + "upsample")
        sc, sh = fold_bn(sd, p + u + "conv.1.", sd[p + u + "conv.0.bias"])
        # up_3 produces the LR feature p (the residual of the CReFF kernel): kept fp32 in 16-bit plans -- the march engine gathers
        # it fastest as fp32 (3.33 ms per 11 frames, 4.39 ms from an f16 feature) -- except for the opt-in tcgen05 CReFF engine,
        # whose operands are f16
        tc_creff = pl.precision == "f16" and os.environ.get("ARSEG_CREFF_TC", "1") != "0"
        t = pl.conv(t, sd[p + u + "conv.0.weight"], sc, sh, pad=1, act=L.ACT_PRELU,
                    slope=float(sd[p + u + "conv.2.weight"].reshape(-1)[0]), name=u + "conv",
                    out_f32=(u == "up_3." and not tc_creff))
    cls = None
    if aux:                                                         # model/pspnet.py:215-217
        a = pl.gmaxpool(x3)
        a = pl.linear(a, sd[p + "classifier.0.weight"], sd[p + "classifier.0.bias"], True, "classifier.0")
        cls = pl.linear(a, sd[p + "classifier.2.weight"], sd[p + "classifier.2.bias"], False, "classifier.2")
    return cls, t


def build_semseg_phase1(pl: Plan, sd: SD, x_nchw: torch.Tensor, p: str = "", bins=(1, 2, 3, 6)):
    """pspnet_semseg.PSPNetWithFuse.forward_phase1 (model/pspnet_semseg.py:223-235) -> (x_tmp NHWC, p NHWC [N,h,w,512])."""
    x4, x3 = _resnet_os8(pl, sd, p, x_nchw, semseg=True)
    N, h, w, Cf = x4.shape
    red = Cf // len(bins)
    folded = [fold_bn(sd, p + "ppm.features.%d.2." % i) for i in range(len(bins))]
    cat = pl.pyramid(x4, bins, [sd[p + "ppm.features.%d.1.weight" % i] for i in range(len(bins))], [f_[0] for f_ in folded],
                     [f_[1] for f_ in folded], relu=True, mode=L.RESIZE_BILINEAR_AC, stages_first=False, name="ppm")   # x first (:27)
    sc, sh = fold_bn(sd, p + "cls.1.")
    t = pl.conv(cat, sd[p + "cls.0.weight"], sc, sh, pad=1, act=L.ACT_RELU, name="cls.0")
    return x3, t


def _arm(pl: Plan, sd: SD, p: str, x, add_chan=None, add_pix=None):
    """AttentionRefinementModule (model/bisenet.py:243-260) + the branch sum that follows it (:295/:301)."""
    feat = _cbr(pl, sd, p + "conv.", x)
    pooled = pl.avgpool(feat, 1, 1, fp32_out=True)
    g = pl.linear(pooled, sd[p + "conv_atten.weight"], None, False, p + "conv_atten")
    gs, gb = fold_bn(sd, p + "bn_atten.")
    return pl.gate(feat, g, gs, gb, add_chan=add_chan, add_pix=add_pix, name=p + "gate")


def build_bisenet_phase1(pl: Plan, sd: SD, x_nchw: torch.Tensor, p: str = "", aux: bool = True):
    """BiSeNetV1WithFuse.forward_phase1 (model/bisenet.py:546-563), aux_mode='train'
    -> (out16 NCHW | None, out32 NCHW | None, middle_feat NHWC [N,h8,w8,256])."""
    r = p + "cp.resnet."
    sc, sh = fold_bn(sd, r + "bn1.")
    x = pl.stem(x_nchw, sd[r + "conv1.weight"], sc, sh, name="cp.stem")
    x = pl.maxpool(x)
    x = _basic_block(pl, sd, r + "layer1.0.", x, 1, 1, 1, 1)
    x = _basic_block(pl, sd, r + "layer1.1.", x, 1, 1, 1, 1)
    x = _basic_block(pl, sd, r + "layer2.0.", x, 2, 1, 1, 2)
    f8 = _basic_block(pl, sd, r + "layer2.1.", x, 1, 1, 1, 1)
    x = _basic_block(pl, sd, r + "layer3.0.", f8, 2, 1, 1, 2)
    f16 = _basic_block(pl, sd, r + "layer3.1.", x, 1, 1, 1, 1)
    x = _basic_block(pl, sd, r + "layer4.0.", f16, 2, 1, 1, 2)
    f32 = _basic_block(pl, sd, r + "layer4.1.", x, 1, 1, 1, 1)
    # ContextPath.forward (model/bisenet.py:289-306)
    pooled = pl.avgpool(f32, 1, 1, fp32_out=True)
    sca, sha = fold_bn(sd, p + "cp.conv_avg.bn.")
    w_avg = sd[p + "cp.conv_avg.conv.weight"].reshape(sd[p + "cp.conv_avg.conv.weight"].shape[0], -1) * sca[:, None]
    avg = pl.linear(pooled, w_avg, sha, True, "cp.conv_avg")                      # 1x1 conv + BN + ReLU on [N,512]
    f32s = _arm(pl, sd, p + "cp.arm32.", f32, add_chan=avg)
    N, h32, w32, _ = f32s.shape
    _, h16, w16, _ = f16.shape
    up = pl.resize(f32s, 2 * h32, 2 * w32, L.RESIZE_NEAREST, name="cp.up32")
    up = pl.resize(up, h16, w16, L.RESIZE_BILINEAR_AC, name="cp.up32_fit")
    f32u = _cbr(pl, sd, p + "cp.conv_head32.", up)
    f16s = _arm(pl, sd, p + "cp.arm16.", f16, add_pix=f32u)
    up = pl.resize(f16s, 2 * h16, 2 * w16, L.RESIZE_NEAREST, name="cp.up16")
    h8, w8 = 2 * h16, 2 * w16
    cat = pl.empty((N, h8, w8, 256))
    _cbr(pl, sd, p + "cp.conv_head16.", up, out=cat, coff=128)                     # feat_cp8 -> cat[..., 128:]
    # SpatialPath (model/bisenet.py:326-340)
    sc, sh = fold_bn(sd, p + "sp.conv1.bn.")
    s = pl.stem(x_nchw, sd[p + "sp.conv1.conv.weight"], sc, sh, name="sp.conv1")
    s = _cbr(pl, sd, p + "sp.conv2.", s, stride=2)
    s = _cbr(pl, sd, p + "sp.conv3.", s, stride=2)
    s = _cbr(pl, sd, p + "sp.conv_out.", s, pad=0)
    pl.resize(s, h8, w8, L.RESIZE_BILINEAR_AC, out=cat, coff=0, name="sp.fit")     # feat_sp -> cat[..., :128] (:550)
    # FeatureFusionModule (model/bisenet.py:387-399)
    feat = _cbr(pl, sd, p + "ffm.convblk.", cat, pad=0)
    pooled = pl.avgpool(feat, 1, 1, fp32_out=True)
    g = pl.linear(pooled, sd[p + "ffm.conv.weight"], None, False, "ffm.conv")
    gs, gb = fold_bn(sd, p + "ffm.bn.")
    fuse = pl.gate(feat, g, gs, gb, add_identity=True, name="ffm.gate")
    mid = _cbr(pl, sd, p + "conv_out.conv.", fuse)                                 # feat_conv_out (:554)
    out16 = out32 = None
    if aux:                                                                        # dead work in evaluation (:556-559)
        cp8 = pl.empty((N, h8, w8, 128))
        # feat_cp8 lives in cat[..., 128:]; the aux head needs it dense
        out16 = _bise_out(pl, sd, p + "conv_out16.", cat, 8, src_coff=128, src_c=128, tmp=cp8)
        out32 = _bise_out(pl, sd, p + "conv_out32.", f32u, 16)
    return out16, out32, mid


def _bise_out(pl: Plan, sd: SD, p: str, x, up: int, src_coff: int = 0, src_c: Optional[int] = None, tmp=None):
    """BiSeNetOutput (model/bisenet.py:207-223): ConvBNReLU 3x3 -> 1x1 (+bias) -> xup bilinear (align_corners=False)."""
    if src_c is not None:
        x = _slice_channels(pl, x, src_coff, src_c, tmp)
    t = _cbr(pl, sd, p + "conv.", x)
    t = pl.conv(t, sd[p + "conv_out.weight"], None, sd[p + "conv_out.bias"], name=p + "conv_out")
    N, h, w, _ = t.shape
    t = pl.resize(t, h * up, w * up, L.RESIZE_BILINEAR, name=p + "up")
    return pl.to_nchw(t)


def _slice_channels(pl: Plan, x: torch.Tensor, coff: int, c: int, out: torch.Tensor) -> torch.Tensor:
    """Dense copy of the channel slice x[..., coff:coff+c]: a 1x1 SIMT conv with a 0/1 selection weight
    (exact: x*1 plus zeros).  Only the aux heads use it, which are dead work in evaluation."""
    eye = torch.zeros(c, x.shape[-1], 1, 1)
    eye[torch.arange(c), coff + torch.arange(c), 0, 0] = 1.0
    saved = pl.engine
    pl.engine = L.CONV_SIMT_F32
    try:
        return pl.conv(x, eye, out=out, name="slice[%d:%d]" % (coff, coff + c))
    finally:
        pl.engine = saved
