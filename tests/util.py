"""Shared test helpers (golden loading, synthetic inputs, error metrics)."""
import os

import numpy as np
import torch

from arseg_b200 import models, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["camvid_psp18_s05", "camvid_psp18_s07", "camvid_bise18_s05", "cityscapes_psp18_s05"]


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: g[k] for k in g.files}


def case_setup(g):
    """Re-create (arch, state_dict, imgs, ref_p, mv int16, flow f64, scale) of a golden case from its seeds."""
    arch = str(g["arch"])
    H, W, C, stride = int(g["H"]), int(g["W"]), int(g["C"]), int(g["stride"])
    s_frame, s_refp, s_mv, s_w = [int(v) for v in g["seeds"]]
    net = models.models_fuse[arch]()
    sd = synth.synth_state_dict(net.state_dict(), s_w)
    imgs = synth.synth_frame(1, H, W, s_frame)
    ref_p = synth.synth_feature(1, C, H // stride, W // stride, s_refp)
    mv = synth.synth_mv_int16(H, W, s_mv, distance=int(g["mv_distance"]))
    return arch, net, sd, imgs, ref_p, mv, synth.mv_to_flow(mv), float(g["scale"])


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max(1e-6, max|b|)"""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(1e-6, float(b.abs().max())))


def rms_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float(((a - b) ** 2).mean().sqrt() / max(1e-12, float((b ** 2).mean().sqrt())))
