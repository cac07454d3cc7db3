"""Deterministic synthetic weights, frames and motion-vector fields.

There are no checkpoints or datasets offline, so every parity test and the
benchmark run on seeded synthetic data.  The generators here are keyed by
*name* (state_dict key) rather than by construction order, so the reference
model (imported from /root/reference when generating goldens), the oracle and
the B200 model all receive bit-identical parameters from the same seed.

Shapes / statistics follow SURVEY.md §8(d):
  * BN running stats are randomised so that BN folding bugs are visible;
  * MV fields mimic HEVC prediction units: block-constant on a 16x16 grid,
    integer-pel values stored as int16 quarter-pel (dataset/camvid.py:624-626),
    10 % intra (zero) blocks, magnitude growing with keyframe distance.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Mapping, Sequence

import numpy as np
import torch


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))
    return g


# The reference registers some modules twice (model/bisenet.py:490-491 feat_conv_out / final_conv alias
# conv_out.conv / conv_out.conv_out; model/pspnet_semseg.py:169 final_conv aliases cls.4), so their
# state_dicts carry the same tensor under two names.  Aliased names are seeded by the canonical one.
_ALIASES = (("feat_conv_out.", "conv_out.conv."), ("final_conv.", "conv_out.conv_out."), ("final_conv.", "cls.4."))


def canonical_key(k: str, keys) -> str:
    for alias, canon in _ALIASES:
        i = k.find(alias)
        if i >= 0 and (i == 0 or k[i - 1] == "."):
            c = k[:i] + canon + k[i + len(alias):]
            if c in keys:
                return c
    return k


def synth_state_dict(spec: Mapping[str, torch.Tensor], seed: int = 4) -> Dict[str, torch.Tensor]:
    """Return a state_dict with the keys/shapes/dtypes of `spec`, filled deterministically.

    `spec` is any state_dict (only shapes and dtypes are read).
    """
    keys = list(spec.keys())
    bn_prefixes = {k[: -len("running_mean")] for k in keys if k.endswith("running_mean")}
    out: Dict[str, torch.Tensor] = {}
    keyset = set(keys)
    for k0 in keys:
        ref = spec[k0]
        shape = tuple(ref.shape)
        k = canonical_key(k0, keyset)
        g = _gen(seed, k)
        prefix = k[: k.rfind(".") + 1]
        leaf = k[k.rfind(".") + 1:]
        if leaf == "num_batches_tracked":
            t = torch.zeros(shape, dtype=ref.dtype)
        elif leaf == "running_mean":
            t = torch.randn(shape, generator=g) * 0.1
        elif leaf == "running_var":
            t = torch.rand(shape, generator=g) + 0.5
        elif prefix in bn_prefixes and leaf == "weight":
            t = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif prefix in bn_prefixes and leaf == "bias":
            t = torch.randn(shape, generator=g) * 0.1
        elif leaf == "weight" and len(shape) == 1 and shape[0] == 1:
            t = torch.full(shape, 0.25)  # PReLU slope (model/pspnet.py:40)
        elif leaf == "weight" and len(shape) >= 2:
            # LeCun-style gain keeps activations O(1) through the residual trunk; the CReFF q/k/v
            # depthwise filters are scaled so that the k*k attention logits have a spread of a few
            # units (neither uniform nor saturated), which keeps tap-order bugs visible.
            fan_in = int(np.prod(shape[1:]))
            gain = 1.5 * (64.0 / shape[0]) ** 0.25 if "fuse_attention" in k else 1.0
            t = torch.randn(shape, generator=g) * (gain * math.sqrt(1.0 / fan_in))
        elif leaf == "bias":
            t = torch.randn(shape, generator=g) * 0.05
        else:
            t = torch.randn(shape, generator=g) * 0.1
        out[k0] = t.to(ref.dtype)
    return out


def synth_frame(n: int, h: int, w: int, seed: int) -> torch.Tensor:
    """Normalised RGB-like frame batch [n,3,h,w] fp32 (randn, as §8d config 1)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(n, 3, h, w, generator=g)


def synth_feature(n: int, c: int, h: int, w: int, seed: int) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(n, c, h, w, generator=g)


def synth_mv_int16(h: int, w: int, seed: int, distance: int = 11, gop: int = 12,
                   block: int = 16, max_u: int = 32, max_v: int = 16) -> np.ndarray:
    """int16[h,w,2] quarter-pel MV map, channel 0 = x, 1 = y (dataset/camvid.py:624-626)."""
    rng = np.random.default_rng(seed)
    bh, bw = (h + block - 1) // block, (w + block - 1) // block
    s = distance / float(max(gop - 1, 1))
    u = np.rint(rng.integers(-max_u, max_u + 1, size=(bh, bw)) * s).astype(np.int32)
    v = np.rint(rng.integers(-max_v, max_v + 1, size=(bh, bw)) * s).astype(np.int32)
    intra = rng.random((bh, bw)) < 0.10
    u[intra] = 0
    v[intra] = 0
    mv = np.stack([u, v], axis=-1) * 4  # integer-pel -> quarter-pel units
    mv = np.repeat(np.repeat(mv, block, axis=0), block, axis=1)[:h, :w]
    return np.ascontiguousarray(mv.astype(np.int16))


def mv_to_flow(mv: np.ndarray) -> torch.Tensor:
    """int16 quarter-pel -> float64 pixel flow [1,h,w,2], as the reference DataLoader yields it."""
    return torch.from_numpy(mv.astype(np.float64) / 4.0).unsqueeze(0)


def lr_size(h: int, w: int, scale: float) -> Sequence[int]:
    """evaluation.py:186-187 -- int() truncation of the float product (0.7*720 -> 503)."""
    return [int(h * scale), int(w * scale)]


def synth_decoder_maps(F: int, H: int, W: int, seed: int) -> np.ndarray:
    """Per-frame HEVC MV maps as the patched dec265 dumps them (`test_%03d.bin`, short[H][W][3] = mvx, mvy quarter-pel,
    refIdx; pre-process/libde265 de265.cc:927-1040): block-constant on a 16x16 grid with a few 8x8 splits, quarter-pel
    (NOT only integer-pel) vectors, refIdx in {0,1,2}, ~10 % intra blocks (refIdx -1) and a few out-of-range refIdx values."""
    rng = np.random.default_rng(seed)
    out = np.zeros((F, H, W, 3), dtype=np.int16)
    for f in range(F):
        for bs in (16, 8):
            by, bx = (H + bs - 1) // bs, (W + bs - 1) // bs
            mv = np.stack([rng.integers(-96, 97, (by, bx)), rng.integers(-48, 49, (by, bx)), rng.integers(0, 3, (by, bx))], -1)
            intra = rng.random((by, bx)) < 0.10
            mv[intra] = (0, 0, -1)
            mv[rng.random((by, bx)) < 0.02, 2] = 90
            full = np.repeat(np.repeat(mv, bs, 0), bs, 1)[:H, :W]
            if bs == 16:
                out[f] = full
            else:
                split = np.repeat(np.repeat(rng.random((by, bx)) < 0.15, bs, 0), bs, 1)[:H, :W]
                out[f][split] = full[split]
    return out
