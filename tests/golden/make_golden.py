"""Generate golden vectors by running the UNMODIFIED reference (imported from /root/reference).

Run in the authoring container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

What it does, per case in CASES:
  1. imports the reference's model classes and `evaluation.warpFeature` as they are, with the two
     unavoidable shims: a CPU `localAttention` module (the oracle's restatement of
     the un-vendored CUDA extension) and an offline `model_zoo.load_url`;
  2. loads the name-keyed synthetic state_dict (arseg_b200/synth.py) into the reference module;
  3. executes the literal per-frame sequence of evaluation.py:176-204;
  4. checks the oracle restatement (oracle/arseg_oracle.py) against it (must agree to <=1e-5);
  5. additionally pins oracle.weighting_forward against the reference's f_weighting_cpu
     (model/attention.py:75-85);
  6. writes tests/golden/<case>.npz with the reference outputs (sub-sampled where large) -- inputs
     are re-generated from the recorded seeds by the tests.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import arseg_oracle as O  # noqa: E402
from arseg_b200 import synth  # noqa: E402

CASES = {
    # name: (arch, frame H, W, scale, feature channels, feature stride, n_classes)
    "camvid_psp18_s05": ("camvid-psp18", 64, 96, 0.5, 64, 1, 12),
    "camvid_psp18_s07": ("camvid-psp18", 64, 96, 0.7, 64, 1, 12),
    "camvid_bise18_s05": ("camvid-bise18", 80, 96, 0.5, 256, 8, 12),
    "cityscapes_psp18_s05": ("cityscapes-psp18", 64, 128, 0.5, 512, 8, 19),
}
SEED_FRAME, SEED_REFP, SEED_MV, SEED_W = 0, 1, 3, 4


def import_reference():
    shim = types.ModuleType("localAttention")
    shim.similar_forward = O.similar_forward
    shim.similar_backward = O.similar_backward
    shim.weighting_forward = O.weighting_forward
    shim.weighting_backward_ori = O.weighting_backward_ori
    shim.weighting_backward_weight = O.weighting_backward_weight
    sys.modules["localAttention"] = shim
    import torch.utils.model_zoo as mz
    mz.load_url = lambda *a, **k: {}
    sys.path.insert(0, REF)
    import model.pspnet as psp
    import model.bisenet as bise
    import model.pspnet_semseg as semseg
    import model.attention as att
    bise.modelzoo.load_url = lambda *a, **k: {}
    # evaluation.py imports datasets (cv2/PIL) at module import; only warpFeature is needed.
    import evaluation as ev
    return psp, bise, semseg, att, ev


def build_ref_net(arch, psp, bise, semseg):
    if arch == "camvid-psp18":
        return psp.PSPNetWithFuse(sizes=(1, 2, 3, 6), n_classes=12, psp_size=512, deep_features_size=256,
                                  backend="resnet18", atten_k=7, pretrained=False)
    if arch == "camvid-bise18":
        return bise.BiSeNetV1WithFuse(n_classes=12, backend="resnet18")
    if arch == "cityscapes-psp18":
        return semseg.PSPNetWithFuse(bins=(1, 2, 3, 6), classes=19, feat_dim=512, layers=18, pretrained=False)
    raise KeyError(arch)


def case_inputs(arch, H, W, C, stride):
    imgs = synth.synth_frame(1, H, W, SEED_FRAME)
    ref_p = synth.synth_feature(1, C, H // stride, W // stride, SEED_REFP)
    mv = synth.synth_mv_int16(H, W, SEED_MV, distance=7)
    return imgs, ref_p, mv


def reference_step(net, ev, imgs, ref_p, flow, scale):
    """Literal transcription of evaluation.py:176-204 (module calls exactly as written there)."""
    highres_ref_p = ref_p
    flow = flow.transpose(2, 3).transpose(1, 2)
    flow = flow * highres_ref_p.shape[-2] / flow.shape[-2]
    flow = F.interpolate(flow, [highres_ref_p.shape[-2], highres_ref_p.shape[-1]], mode="bilinear", align_corners=True)
    flow = flow.transpose(1, 2).transpose(2, 3)
    highres_ref_p = ev.warpFeature(highres_ref_p, flow)
    N, C, H, W = imgs.size()
    new_hw = [int(H * scale), int(W * scale)]
    x = F.interpolate(imgs, new_hw, mode="bilinear", align_corners=True)
    phase1_out = net.forward_phase1(x)
    out_p = phase1_out[-1]
    out, fused = net.forward_phase2(out_p, highres_ref_p)
    logits = F.interpolate(out, size=[H, W], mode="bilinear", align_corners=True)
    probs = torch.softmax(logits, dim=1)
    preds = torch.argmax(probs, dim=1)
    return preds, logits, fused, out_p, highres_ref_p


def main():
    psp, bise, semseg, att, ev = import_reference()
    torch.set_grad_enabled(False)

    # (5) pin weighting_forward against the reference's in-repo restatement
    g = torch.Generator().manual_seed(11)
    v = torch.randn(2, 5, 9, 11, generator=g)
    a = torch.softmax(torch.randn(2, 9, 11, 15, generator=g), dim=3)
    ref_w = att.f_weighting_cpu(v, a, 3, 5)
    err = (O.weighting_forward(v, a, 3, 5) - ref_w).abs().max().item()
    assert err < 1e-6, err
    q = torch.randn(2, 5, 9, 11, generator=g)
    np.savez(os.path.join(HERE, "local_attention_ops.npz"),
             v=v.numpy(), a=a.numpy(), q=q.numpy(), weighting_ref=ref_w.numpy(),
             similar_oracle=O.similar_forward(q, v, 3, 5).numpy(), kH=3, kW=5)
    print("f_weighting_cpu pin: max|d| = %.3g" % err)

    for name, (arch, H, W, scale, C, stride, ncls) in CASES.items():
        torch.manual_seed(233)
        net = build_ref_net(arch, psp, bise, semseg).eval()
        sd = synth.synth_state_dict(net.state_dict(), SEED_W)
        net.load_state_dict(sd)
        imgs, ref_p, mv = case_inputs(arch, H, W, C, stride)
        flow = synth.mv_to_flow(mv)
        preds, logits, fused, lr_p, warped = reference_step(net, ev, imgs, ref_p, flow, scale)
        o_preds, o_logits, o_fused, o_lr_p = O.nonkey_step(arch, sd, imgs, ref_p, flow, scale)
        d = {"logits": (o_logits - logits).abs().max().item(), "fused": (o_fused - fused).abs().max().item(),
             "lr_p": (o_lr_p - lr_p).abs().max().item(), "argmax_mismatch": int((o_preds != preds).sum())}
        print(name, "oracle-vs-reference", d, "| |logits|max", logits.abs().max().item())
        assert d["logits"] <= 1e-5 and d["fused"] <= 1e-5 and d["lr_p"] <= 1e-5 and d["argmax_mismatch"] == 0, d
        np.savez(os.path.join(HERE, name + ".npz"),
                 arch=arch, H=H, W=W, scale=scale, C=C, stride=stride, n_classes=ncls,
                 seeds=np.array([SEED_FRAME, SEED_REFP, SEED_MV, SEED_W]), mv_distance=7,
                 logits=logits.numpy(), preds=preds.numpy().astype(np.uint8),
                 lr_p=lr_p.numpy()[:, ::4], lr_p_sum=lr_p.double().sum().item(),
                 fused=fused.numpy()[:, ::8], fused_sum=fused.double().sum().item(),
                 warped=warped.numpy()[:, ::8], warped_sum=warped.double().sum().item(),
                 n_state=len(sd), state_sum=float(sum(t.double().sum().item() for t in sd.values())))
    # state_dict contract: key -> shape for every net of the evaluation.py:24-36 registries
    import json
    specs = {}
    torch.manual_seed(233)
    hr_builders = {
        "camvid-psp18": lambda: psp.PSPNet(sizes=(1, 2, 3, 6), n_classes=12, psp_size=512, deep_features_size=256,
                                           backend="resnet18", pretrained=False),
        "camvid-bise18": lambda: bise.BiSeNetV1(n_classes=12, backend="resnet18"),
        "cityscapes-psp18": lambda: semseg.PSPNetWithFuse(bins=(1, 2, 3, 6), classes=19, feat_dim=512, layers=18, pretrained=False),
        "cityscapes-bise18": lambda: bise.BiSeNetV1(n_classes=19, backend="resnet18"),
    }
    for arch, fn in hr_builders.items():
        specs["hr/" + arch] = {k: list(v.shape) for k, v in fn().state_dict().items()}
    for arch in ("camvid-psp18", "camvid-bise18", "cityscapes-psp18"):
        specs["lr/" + arch] = {k: list(v.shape) for k, v in build_ref_net(arch, psp, bise, semseg).state_dict().items()}
    specs["lr/cityscapes-bise18"] = {k: list(v.shape) for k, v in
                                     bise.BiSeNetV1WithFuse(n_classes=19, backend="resnet18").state_dict().items()}
    with open(os.path.join(HERE, "state_specs.json"), "w") as f:
        json.dump(specs, f, indent=0, sort_keys=True)
    print("goldens written to", HERE)


if __name__ == "__main__":
    main()
