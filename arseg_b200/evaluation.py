"""Host-side mirror of the reference's evaluation path (evaluation.py:61-87, 148-215).

* `warpFeature(feature, flow)`            -- same name / arguments as evaluation.py:61 (one kernel).
* `nonkey_step(net, imgs, ref_p, flow)`   -- the literal per-frame sequence evaluation.py:176-204 through
                                             the drop-in modules (forward_phase1 / forward_phase2).
* `NonKeyEngine`                          -- the fused throughput path: all non-keyframes of a GOP batched,
                                             MV field consumed in its on-disk int16 quarter-pel format,
                                             warp + CReFF + classifier + argmax in one kernel, whole step
                                             captured in a CUDA graph.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import os

import torch

from . import _lib as L
from . import engine as E
from . import ops

ARCH_INFO = {
    # arch: (feature channels of p, feature stride of p w.r.t. the frame, n_classes)
    "camvid-psp18": (64, 1, 12),
    "camvid-bise18": (256, 8, 12),
    "cityscapes-psp18": (512, 8, 19),
    "cityscapes-bise18": (256, 8, 19),
}


def warpFeature(feature: torch.Tensor, flow: torch.Tensor) -> torch.Tensor:
    """evaluation.py:61-87: backward bilinear warp of `feature` [B,C,H,W] by `flow` [B,H,W,2] (pixels)."""
    return ops.warp_feature(feature, flow)


def resize_flow(flow: torch.Tensor, Hf: int, Wf: int) -> torch.Tensor:
    """evaluation.py:177-180 on the GPU: flow f64 [B,H,W,2] -> [B,Hf,Wf,2], both components scaled by Hf/H.
    (float64 planes are resized in float64 by the caller's torch in the reference; here the field is
    passed un-resized to the fused kernel instead -- this helper exists for the literal API path and runs
    the arithmetic in float64 on the device with torch only when sizes differ.)"""
    B, H, W, _ = flow.shape
    if (H, W) == (Hf, Wf):
        return flow * Hf / H
    f = flow.permute(0, 3, 1, 2) * Hf / H
    f = torch.nn.functional.interpolate(f, [Hf, Wf], mode="bilinear", align_corners=True)
    return f.permute(0, 2, 3, 1).contiguous()


def nonkey_step(net, imgs: torch.Tensor, ref_p: torch.Tensor, flow: torch.Tensor, scale: float = 0.5,
                label_size: Optional[Sequence[int]] = None):
    """evaluation.py:176-204 with `net` = one of the drop-in *WithFuse modules.
    Returns (preds uint8 [N,H,W], logits fp32 [N,ncls,H,W], fused p, lr p)."""
    flow = resize_flow(flow, ref_p.shape[-2], ref_p.shape[-1])                       # :177-180
    warped = warpFeature(ref_p, flow)                                                # :183
    N, _, H, W = imgs.shape
    new_hw = [int(H * scale), int(W * scale)]                                        # :186-187
    x = ops.resize_nchw(imgs, new_hw, L.RESIZE_BILINEAR_AC)                          # :188
    lr_p = net.forward_phase1(x)[-1]                                                 # :190-191
    out, fused = net.forward_phase2(lr_p, warped)                                    # :193
    size = list(label_size) if label_size is not None else [H, W]
    preds, logits = ops.resize_argmax(out, size, L.RESIZE_BILINEAR_AC, want_logits=True)   # :201-204
    return preds, logits, fused, lr_p


def internal_ref_dtype(arch: str, precision: str, k: int = 7) -> torch.dtype:
    """dtype of the keyframe feature in the engines' internal NHWC layout: fp16 where the non-keyframe step runs the tcgen05 CReFF
    engine (f16 plan, C = 64, k <= 7: its MV-warp pre-pass stores fp16 rows anyway and reads an fp16 feature at half the L2
    footprint), fp32 otherwise."""
    C_ = ARCH_INFO[arch][0]
    tc = precision == "f16" and C_ == 64 and k <= 7 and os.environ.get("ARSEG_CREFF_TC", "1") != "0"
    return torch.float16 if tc else torch.float32


class NonKeyEngine:
    """All non-keyframes of a GOP in one captured step.

    Inputs (device-resident static buffers): `imgs` fp32 [N,3,H,W], `mv` int16 [N,H,W,2] quarter-pel
    (dataset/camvid.py:624-626), `ref_p` fp32 NCHW [1,C,Hf,Wf] = the keyframe feature (shared by the GOP).
    Output: `preds` uint8 [N,H,W] (+ `logits` at feature resolution when want_logits).
    """

    def __init__(self, arch: str, sd: Dict[str, torch.Tensor], n_frames: int, H: int, W: int, scale: float = 0.5,
                 precision: str = "tf32", k: int = 7, device="cuda:0", want_logits: bool = False, want_p: bool = False,
                 graph: bool = True, split_keyframe: bool = False, uint8_frames: bool = False,
                 mean=ops.CAMVID_MEAN, std=ops.CAMVID_STD, ref_nhwc: Optional[torch.Tensor] = None):
        """uint8_frames: `imgs` is uint8 HWC [N,H,W,3] (decoded frames); ToTensor + Normalize(mean, std)
        (dataset/camvid.py:182-185) are fused into the LR down-scale kernel -- a quarter of the host->device bytes.
        ref_nhwc: the keyframe feature as an NHWC [1,Hf,Wf,C] device tensor in the internal layout (fp32, or
        internal_ref_dtype(arch, precision, k); e.g. a KeyFrameEngine's `p_nhwc`): the step reads it in place, without the
        per-GOP NCHW -> NHWC transpose launch.  `ref_p` (the API-layout buffer) is then None."""
        if arch not in ARCH_INFO:
            raise KeyError(arch)
        self.arch, self.N, self.H, self.W, self.scale, self.k = arch, n_frames, H, W, scale, k
        C_, stride, ncls = ARCH_INFO[arch]
        self.C, self.ncls = C_, ncls
        self.Hf, self.Wf = H // stride, W // stride
        self.h, self.w = int(H * scale), int(W * scale)
        self.device = torch.device(device)
        with torch.no_grad(), torch.cuda.device(self.device):
            pl = E.Plan(self.device, precision)
            self.plan = pl
            self.uint8_frames = uint8_frames
            self.imgs = pl.empty((n_frames, H, W, 3), torch.uint8) if uint8_frames else pl.empty((n_frames, 3, H, W), torch.float32)
            self.mv = pl.empty((n_frames, H, W, 2), torch.int16)
            self.ref_nhwc = ref_nhwc
            if ref_nhwc is not None:
                ok_dt = (torch.float32, internal_ref_dtype(arch, precision, k))
                if tuple(ref_nhwc.shape) != (1, self.Hf, self.Wf, C_) or ref_nhwc.dtype not in ok_dt or not ref_nhwc.is_contiguous():
                    raise ValueError("ref_nhwc must be a contiguous [1,%d,%d,%d] tensor of %s" % (self.Hf, self.Wf, C_, " or ".join(map(str, set(ok_dt)))))
                self.ref_p = None
            else:
                self.ref_p = pl.empty((1, C_, self.Hf, self.Wf), torch.float32)
                self.ref_p.zero_()
            self.imgs.zero_(); self.mv.zero_()
            if uint8_frames:
                x = pl.frame_ingest_u8(self.imgs, self.h, self.w, mean, std)                              # dataset/camvid.py:182-185 + evaluation.py:186-188
            else:
                x = pl.resize_nchw(self.imgs, self.h, self.w, L.RESIZE_BILINEAR_AC, name="frame_downscale")   # evaluation.py:186-188
            if arch == "camvid-psp18":
                _, p = E.build_psp_phase1(pl, sd, x, "", aux=False)
                fin, logsm = "final_conv.", True
            elif arch == "cityscapes-psp18":
                _, p = E.build_semseg_phase1(pl, sd, x, "")
                fin, logsm = "cls.4.", False
            else:
                _, _, p = E.build_bisenet_phase1(pl, sd, x, "", aux=False)
                fin, logsm = "conv_out.conv_out.", False
            self.lr_p = p
            direct = arch == "camvid-psp18"     # logits already at frame resolution: argmax inside the kernel
            out_p, out_l, out_a = pl.creff(self.ref_p if ref_nhwc is None else ref_nhwc, p, sd, "fuse_attention.", k, flow=self.mv,
                                           hr_shared=True, hr_layout=L.NCHW if ref_nhwc is None else L.NHWC,
                                           lr_layout=L.NHWC, wcls=sd[fin + "weight"], bcls=sd[fin + "bias"],
                                           log_softmax=logsm, want_p=want_p, want_logits=(want_logits or not direct),
                                           want_argmax=direct, hoist_prepass=True)
            self.fused_p, self.logits = out_p, out_l
            if direct:
                self.preds = out_a
            elif arch.endswith("bise18"):
                # forward_phase2's x8 up-sampling (align_corners=False, model/bisenet.py:573) then evaluation.py:201-204
                H8, W8 = self.Hf * 8, self.Wf * 8
                if (H8, W8) == (H, W):
                    self.preds, _ = pl.resize_argmax(out_l, H, W, L.RESIZE_BILINEAR)
                else:
                    up = pl.resize_nchw(out_l, H8, W8, L.RESIZE_BILINEAR, name="out_upsample")
                    self.preds, _ = pl.resize_argmax(up, H, W, L.RESIZE_BILINEAR_AC)
            else:
                self.preds, _ = pl.resize_argmax(out_l, H, W, L.RESIZE_BILINEAR_AC)          # evaluation.py:201-204
            if graph:
                # split_keyframe: phase 1 (needs only the frames) and the launches that read the keyframe feature are two
                # graphs, so a broadcast of the feature (frame-level sharding) can overlap phase 1: step_phase1() / step_phase2()
                first_ref = next(i for i, nm in enumerate(pl.names) if nm.startswith("hr_nchw_to_nhwc") or nm.startswith("creff"))
                pl.capture(split_at=first_ref if split_keyframe else None)
        self.conv_flops_per_frame = pl.conv_flops / n_frames
        self.launches_per_step = pl.n_launches

    def set_inputs(self, imgs: torch.Tensor, mv: torch.Tensor, ref_p: Optional[torch.Tensor]) -> None:
        """ref_p: the keyframe feature fp32 NCHW [1,C,Hf,Wf] (None leaves it as it is; with ref_nhwc it is transposed into that buffer)."""
        self.imgs.copy_(imgs, non_blocking=True)
        self.mv.copy_(mv, non_blocking=True)
        if ref_p is None:
            return
        if self.ref_nhwc is not None:
            self.ref_nhwc.copy_(ref_p.to(self.ref_nhwc.device).permute(0, 2, 3, 1), non_blocking=True)
        else:
            self.ref_p.copy_(ref_p, non_blocking=True)

    def step(self) -> torch.Tensor:
        """One pass over the N non-keyframes with inputs already resident in HBM."""
        self.plan.run()
        return self.preds

    def step_phase1(self) -> None:
        """Everything that does not read the keyframe feature (frame down-scale + LR-branch network); split_keyframe engines."""
        self.plan.run_part(0)

    def step_phase2(self) -> torch.Tensor:
        """Keyframe-feature transpose, MV warp + CReFF + classifier (+ post-processing); split_keyframe engines."""
        self.plan.run_part(1)
        return self.preds

    def step_host(self, imgs_pinned: torch.Tensor, mv_pinned: torch.Tensor, preds_pinned: torch.Tensor) -> torch.Tensor:
        """End-to-end step from pinned host buffers: H2D frames + MV fields, compute, D2H class maps."""
        self.imgs.copy_(imgs_pinned, non_blocking=True)
        self.mv.copy_(mv_pinned, non_blocking=True)
        self.plan.run()
        preds_pinned.copy_(self.preds, non_blocking=True)
        return preds_pinned

    def host_pipeline(self) -> "HostPipeline":
        return HostPipeline(self)


class KeyFrameEngine:
    """The HR keyframe branch of a GOP as one captured step: `highres_net(ref_imgs)[-1]` of
    evaluation.py:173-174 = the feature p the non-keyframes fuse (model/pspnet.py:76-100, model/bisenet.py:438-461,
    model/pspnet_semseg.py:184-219 without the heads evaluation.py never reads).

    Input: `img` fp32 [1,3,H,W] (or uint8 HWC with uint8_frames).  Output: `p` fp32 NCHW [1,C,H/stride,W/stride] -- pass
    `out=` (e.g. a NonKeyEngine's `ref_p`) to have the step write the feature where its consumer reads it."""

    def __init__(self, arch: str, sd: Dict[str, torch.Tensor], H: int, W: int, precision: str = "tf32", device="cuda:0",
                 out: Optional[torch.Tensor] = None, graph: bool = True, uint8_frames: bool = False,
                 mean=ops.CAMVID_MEAN, std=ops.CAMVID_STD, api_layout: bool = True):
        if arch not in ARCH_INFO:
            raise KeyError(arch)
        C_, stride, _ = ARCH_INFO[arch]
        self.arch, self.H, self.W, self.C = arch, H, W, C_
        self.device = torch.device(device)
        with torch.no_grad(), torch.cuda.device(self.device):
            pl = E.Plan(self.device, precision)
            self.plan = pl
            if uint8_frames:
                self.img = pl.empty((1, H, W, 3), torch.uint8)
                self.img.zero_()
                x = pl.frame_ingest_u8(self.img, H, W, mean, std)
            else:
                self.img = pl.empty((1, 3, H, W), torch.float32)
                self.img.zero_()
                x = self.img
            if arch == "camvid-psp18":
                _, p = E.build_psp_phase1(pl, sd, x, "", aux=False)
            elif arch == "cityscapes-psp18":
                _, p = E.build_semseg_phase1(pl, sd, x, "")
            else:
                _, _, p = E.build_bisenet_phase1(pl, sd, x, "", aux=False)
            # the feature in the internal layout: NHWC for CamVid-PSP in every plan (up_3 writes fp32, or fp16 where the f16 plan's
            # non-keyframe step runs the tcgen05 CReFF engine) -> a NonKeyEngine built with ref_nhwc=self.p_nhwc reads it in place;
            # api_layout=False then skips the NCHW copy altogether
            self.p_nhwc = p if p.dtype in (torch.float32, internal_ref_dtype(arch, precision)) and C_ == 64 else None
            self.p = pl.to_nchw(p, name="p_to_nchw", out=out) if (api_layout or self.p_nhwc is None) else None
            if graph:
                pl.capture()
        self.conv_flops = pl.conv_flops
        self.launches_per_step = pl.n_launches

    def step(self) -> torch.Tensor:
        self.plan.run()
        return self.p if self.p is not None else self.p_nhwc


class HostPipeline:
    """Streaming end-to-end path: GOP after GOP from pinned host memory (the evaluation loop of evaluation.py:161-209
    with the DataLoader's batches already pinned).  Step i's host->device copies (frames fp32 + MV int16) run on a copy
    stream into one of two device staging sets while step i-1 computes; the class maps of step i go back on a third
    stream while step i+1 computes.  Every step still moves all its inputs and its result over PCIe -- only the
    serialisation between copy and compute is removed.

        pipe = eng.host_pipeline()
        for imgs, mv, out in batches: pipe.submit(imgs, mv, out)     # asynchronous, in order
        pipe.drain()                                                 # current stream waits for the last D2H
    """

    def __init__(self, eng: NonKeyEngine):
        self.eng = eng
        dev = eng.device
        with torch.cuda.device(dev):
            self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            self.stage = [(torch.empty_like(eng.imgs), torch.empty_like(eng.mv)) for _ in range(2)]
            self.h2d_done = [torch.cuda.Event() for _ in range(2)]
            self.stage_free = [torch.cuda.Event() for _ in range(2)]
            self.computed = torch.cuda.Event()
            self.d2h_done = torch.cuda.Event()
        self.k = 0

    def submit(self, imgs_pinned: torch.Tensor, mv_pinned: torch.Tensor, preds_pinned: torch.Tensor) -> None:
        eng, i = self.eng, self.k & 1
        with torch.cuda.device(eng.device):
            main = torch.cuda.current_stream()
            s_imgs, s_mv = self.stage[i]
            if self.k == 0:
                self.s_in.wait_stream(main)               # ordered after whatever the caller enqueued (timing events too)
            else:
                self.s_in.wait_event(self.stage_free[i])  # step k-2's staging set has been consumed
            with torch.cuda.stream(self.s_in):
                s_imgs.copy_(imgs_pinned, non_blocking=True)
                s_mv.copy_(mv_pinned, non_blocking=True)
                self.h2d_done[i].record(self.s_in)
            main.wait_event(self.h2d_done[i])
            eng.imgs.copy_(s_imgs, non_blocking=True)     # device-to-device into the captured graph's static inputs
            eng.mv.copy_(s_mv, non_blocking=True)
            self.stage_free[i].record(main)
            if self.k > 0:
                main.wait_event(self.d2h_done)            # the previous class maps have left eng.preds
            eng.plan.run()
            self.computed.record(main)
            self.s_out.wait_event(self.computed)
            with torch.cuda.stream(self.s_out):
                preds_pinned.copy_(eng.preds, non_blocking=True)
                self.d2h_done.record(self.s_out)
        self.k += 1

    def drain(self) -> None:
        if self.k:
            torch.cuda.current_stream(self.eng.device).wait_event(self.d2h_done)
