"""CPU suite: host-side sharding logic, world_size 2 over gloo (the product kernels are not involved)."""
import os
import subprocess
import sys
import textwrap

import torch

from arseg_b200 import dist as adist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_dealing_is_a_partition():
    for world in (1, 2, 4, 8):
        for owner in range(world):
            shares = [adist.frames_of_rank(11, world, r, owner) for r in range(world)]
            flat = sorted(f for s in shares for f in s)
            assert flat == list(range(11))
            sizes = [len(s) for s in shares]
            assert max(sizes) - min(sizes) <= 1
            assert len(shares[owner]) == min(sizes)       # the keyframe owner takes the short share
    assert [len(adist.frames_of_rank(11, 8, r)) for r in range(8)] == [1, 2, 2, 2, 1, 1, 1, 1]


def test_gop_sharding_is_a_partition():
    for world in (1, 2, 3, 8):
        flat = sorted(g for r in range(world) for g in adist.gops_of_rank(13, world, r))
        assert flat == list(range(13))


def test_miou():
    h = torch.tensor([[5, 1], [2, 7]])
    assert abs(adist.miou_from_hist(h.flatten().view(2, 2)) - (5 / 8 + 7 / 10) / 2) < 1e-12


WORKER = textwrap.dedent("""
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from arseg_b200 import dist as adist
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    ncls = 5
    def step(item):     # stand-in for the engine: a deterministic histogram per work item
        g = torch.Generator().manual_seed(int(item))
        pred = torch.randint(0, ncls, (64,), generator=g); lab = torch.randint(0, ncls, (64,), generator=g)
        return torch.bincount(lab * ncls + pred, minlength=ncls * ncls)
    gops = list(range(7))
    hist = adist.run_sharded(gops, step, ncls, mode="gop")
    ref = sum(step(g) for g in gops)
    assert torch.equal(hist, ref), (rank, hist, ref)
    # frame-level mode: every rank handles its share of every GOP
    def step_frames(item):
        out = torch.zeros(ncls * ncls, dtype=torch.int64)
        for f in adist.frames_of_rank(11, world, rank):
            out += step(item * 100 + f)
        return out
    hist2 = adist.run_sharded(gops, step_frames, ncls, mode="frame")
    ref2 = sum(step(g * 100 + f) for g in gops for f in range(11))
    assert torch.equal(hist2, ref2)
    # keyframe feature broadcast (CPU tensors go through the plain path)
    p = torch.full((1, 4, 6, 8), float(rank + 1))
    adist.broadcast_keyframe_feature(p, src=1)
    assert (p == 2).all()
    dist.destroy_process_group()
    print("ok", rank)
""") % ROOT


def test_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2
