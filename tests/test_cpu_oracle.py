"""CPU suite: the oracle against the committed golden vectors (generated from the unmodified reference by
tests/golden/make_golden.py), the oracle's plain-C twins, and the synthetic-data generators."""
import json
import os
import subprocess

import numpy as np
import pytest
import torch

from arseg_b200 import models, synth
from oracle import arseg_oracle as O
from tests.util import CASES, GOLDEN, case_setup, load_golden

torch.set_grad_enabled(False)


@pytest.fixture(scope="module", autouse=True)
def _build_c_oracle():
    subprocess.check_call(["make", "-C", os.path.dirname(O.__file__)], stdout=subprocess.DEVNULL)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    g = load_golden(name)
    arch, net, sd, imgs, ref_p, mv, flow, scale = case_setup(g)
    assert len(sd) == int(g["n_state"])
    assert abs(sum(t.double().sum().item() for t in sd.values()) - float(g["state_sum"])) < 1e-6
    preds, logits, fused, lr_p = O.nonkey_step(arch, sd, imgs, ref_p, flow, scale)
    assert np.abs(logits.numpy() - g["logits"]).max() <= 1e-5
    assert np.array_equal(preds.numpy().astype(np.uint8), g["preds"])
    assert np.abs(lr_p.numpy()[:, ::4] - g["lr_p"]).max() <= 1e-5
    assert np.abs(fused.numpy()[:, ::8] - g["fused"]).max() <= 1e-5
    assert abs(fused.double().sum().item() - float(g["fused_sum"])) <= 1e-3 * max(1.0, abs(float(g["fused_sum"])))


def test_weighting_forward_pinned_to_reference_restatement():
    g = np.load(os.path.join(GOLDEN, "local_attention_ops.npz"))
    v, a, q = (torch.from_numpy(g[k]) for k in ("v", "a", "q"))
    kH, kW = int(g["kH"]), int(g["kW"])
    # weighting_ref = output of the reference's own f_weighting_cpu (model/attention.py:75-85)
    assert np.abs(O.weighting_forward(v, a, kH, kW).numpy() - g["weighting_ref"]).max() < 1e-6
    assert np.abs(O.similar_forward(q, v, kH, kW).numpy() - g["similar_oracle"]).max() < 1e-6


@pytest.mark.parametrize("k", [(3, 3), (7, 7), (3, 5)])
def test_c_twins_match_python(k):
    g = torch.Generator().manual_seed(5)
    q, kk = torch.randn(2, 6, 7, 9, generator=g), torch.randn(2, 6, 7, 9, generator=g)
    s = O.similar_forward(q, kk, *k)
    assert (O.similar_forward_c(q, kk, *k) - s).abs().max() < 1e-5
    a = torch.softmax(s, 3)
    assert (O.weighting_forward_c(kk, a, *k) - O.weighting_forward(kk, a, *k)).abs().max() < 1e-5


def test_similar_definition_unfold():
    """similar_forward against the unfold definition the reference's debug code documents (attention.py:55-73)."""
    g = torch.Generator().manual_seed(6)
    q, k = torch.randn(1, 4, 6, 8, generator=g), torch.randn(1, 4, 6, 8, generator=g)
    uf = torch.nn.functional.unfold(k, (5, 5), padding=2).view(1, 4, 25, 6, 8)
    ref = (q.unsqueeze(2) * uf).sum(1).permute(0, 2, 3, 1)
    assert (O.similar_forward(q, k, 5, 5) - ref).abs().max() < 1e-5


def test_backward_ops_against_autograd():
    g = torch.Generator().manual_seed(7)
    with torch.enable_grad():
        q = torch.randn(1, 3, 5, 6, generator=g, dtype=torch.float64, requires_grad=True)
        k = torch.randn(1, 3, 5, 6, generator=g, dtype=torch.float64, requires_grad=True)
        s = O.similar_forward(q, k, 3, 3)
        go = torch.randn(s.shape, generator=g, dtype=torch.float64)
        gq, gk = torch.autograd.grad(s, (q, k), go)
    assert (O.similar_backward(k.detach(), go, 3, 3, True) - gq).abs().max() < 1e-10
    assert (O.similar_backward(q.detach(), go, 3, 3, False) - gk).abs().max() < 1e-10
    with torch.enable_grad():
        v = torch.randn(1, 3, 5, 6, generator=g, dtype=torch.float64, requires_grad=True)
        a = torch.rand(1, 5, 6, 9, generator=g, dtype=torch.float64, requires_grad=True)
        o = O.weighting_forward(v, a, 3, 3)
        go = torch.randn(o.shape, generator=g, dtype=torch.float64)
        gv, ga = torch.autograd.grad(o, (v, a), go)
    assert (O.weighting_backward_ori(a.detach(), go, 3, 3) - gv).abs().max() < 1e-10
    assert (O.weighting_backward_weight(v.detach(), go, 3, 3) - ga).abs().max() < 1e-10


def test_zero_flow_is_not_identity():
    """the align_corners mismatch of warpFeature (evaluation.py:80-85)."""
    f = synth.synth_feature(1, 2, 8, 10, 0)
    w = O.warp_feature(f, torch.zeros(1, 8, 10, 2, dtype=torch.float64))
    assert (w - f).abs().max() > 1e-3


def test_lr_size_truncation():
    assert synth.lr_size(720, 960, 0.7) == [503, 672]      # int(0.7*720) = 503 (evaluation.py:186-187)
    assert synth.lr_size(720, 960, 0.5) == [360, 480]


def test_mv_field_properties():
    mv = synth.synth_mv_int16(64, 96, 3, distance=7)
    assert mv.dtype == np.int16 and mv.shape == (64, 96, 2)
    assert (mv % 4 == 0).all()                              # integer-pel multiples (generate..camvid.py:53-54)
    assert (mv[:16, :16] == mv[0, 0]).all()                 # block-constant 16x16
    assert np.array_equal(mv, synth.synth_mv_int16(64, 96, 3, distance=7))


def test_state_dict_contract_matches_reference():
    """Key names and shapes of every drop-in module == the reference's (dumped by make_golden.py)."""
    specs = json.load(open(os.path.join(GOLDEN, "state_specs.json")))
    for arch in models.models_fuse:
        mine = {k: list(v.shape) for k, v in models.models_fuse[arch]().state_dict().items()}
        assert mine == specs["lr/" + arch], arch
        mine = {k: list(v.shape) for k, v in models.models[arch]().state_dict().items()}
        assert mine == specs["hr/" + arch], arch


def test_module_prefixed_checkpoint_loads():
    """evaluation.py:41-46: checkpoints are saved from nn.DataParallel -> 'module.' prefix."""
    net = models.models_fuse["camvid-psp18"]()
    sd = {"module." + k: v for k, v in synth.synth_state_dict(net.state_dict(), 4).items()}
    torch.nn.DataParallel(net).load_state_dict(sd)


def test_merge_motion_oracle_matches_reference_golden():
    """oracle.merge_motion against the golden recorded from the unmodified reference's mergeMotion
    (pre-process/generate_compressed_dataset_camvid.py:6-56; tests/golden/make_merge_motion_golden.py), 720x960, 4 frames."""
    import os
    import zlib
    import numpy as np
    from arseg_b200 import synth
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "merge_motion.npz"))
    Fn, H, W = int(g["F"]), int(g["H"]), int(g["W"])
    out = O.merge_motion(synth.synth_decoder_maps(Fn, H, W, int(g["seed"])))
    assert out.shape == (H, W, Fn + 1, 2) and (out[:, :, 0] == -1).all()
    assert np.array_equal(out[:, :, Fn].astype(np.int16), g["last"])
    for f in range(Fn + 1):
        assert zlib.crc32(np.ascontiguousarray(out[:, :, f].astype(np.int16)).tobytes()) == int(g["crcs"][f])


def test_ingest_u8_oracle_is_totensor_normalize_interpolate():
    import torch
    g = torch.Generator().manual_seed(3)
    fr = torch.randint(0, 256, (1, 10, 12, 3), generator=g, dtype=torch.uint8)
    out = O.ingest_u8(fr, (0.39068785, 0.40521392, 0.41434407), (0.29652068, 0.30514979, 0.30080369), (10, 12))
    want = (fr.permute(0, 3, 1, 2).float() / 255 - torch.tensor([0.39068785, 0.40521392, 0.41434407]).view(1, 3, 1, 1)) / \
        torch.tensor([0.29652068, 0.30514979, 0.30080369]).view(1, 3, 1, 1)
    assert torch.allclose(out, want, atol=1e-6)
