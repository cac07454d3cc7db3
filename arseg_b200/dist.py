"""Multi-GPU sharding of the non-keyframe path (one process per GPU, torch.distributed / NCCL over NVLink).

The path shards naturally: given the keyframe feature p_HR every non-keyframe is independent
(evaluation.py:161-209 carries no cross-iteration state except the additive confusion matrix).

Two levels:
  * GOP level  (throughput mode, bench default, weak scaling): GOP g -> rank g mod world.  No data-path
    collective at all; each rank consumes its own frames / MV fields / keyframe feature.
  * frame level (latency mode, strong scaling): the 11 non-keyframes of ONE GOP are dealt over the ranks;
    the owner of the keyframe broadcasts p_HR once per GOP (`broadcast_keyframe_feature`, ncclBroadcast,
    176.9 MB fp32 for CamVid-PSP) on a side stream so it overlaps LR phase 1, which needs only pixels.
The only reduction is the [n_cls, n_cls] confusion matrix at the end of a run (`allreduce_hist`), mirroring
the reference's dist.all_reduce(hist) (evaluation.py:134-135, 210-211).
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def frames_of_rank(n_frames: int, world: int, rank: int, owner: int = 0) -> List[int]:
    """Deal frame indices 0..n_frames-1 over `world` ranks.  Shares differ by at most one frame and the
    `owner` rank (which also runs the HR keyframe forward, 617 vs 167 GFLOP) is last in line for the extra
    frames: 11 frames over 8 ranks -> owner 1 frame, ranks after it 2,2,2 then 1,1,1,1."""
    if not (0 <= rank < world) or not (0 <= owner < world):
        raise ValueError("rank/owner out of range")
    order = [(owner + 1 + i) % world for i in range(world)]      # owner comes last
    pos = order.index(rank)
    return [f for f in range(n_frames) if f % world == pos]


def gops_of_rank(n_gops: int, world: int, rank: int) -> List[int]:
    """GOP-level sharding: GOP g -> rank g mod world."""
    return list(range(rank, n_gops, world))


def broadcast_keyframe_feature(ref_p: torch.Tensor, src: int, group=None, stream: Optional[torch.cuda.Stream] = None):
    """ncclBroadcast of the keyframe feature p_HR from the GOP owner.  When `stream` is given the collective is
    enqueued there and an event is returned that the consumer (the fused CReFF kernel's stream) must wait on,
    so the transfer hides behind LR phase 1."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    if stream is None or not ref_p.is_cuda:
        dist.broadcast(ref_p, src=src, group=group)
        return None
    stream.wait_stream(torch.cuda.current_stream(ref_p.device))
    with torch.cuda.stream(stream):
        dist.broadcast(ref_p, src=src, group=group)
        ev = torch.cuda.Event()
        ev.record(stream)
    return ev


def allreduce_hist(hist: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the confusion matrix over ranks (evaluation.py:210-211)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


def miou_from_hist(hist: torch.Tensor) -> float:
    """evaluation.py:212-213."""
    h = hist.double()
    ious = h.diag() / (h.sum(dim=0) + h.sum(dim=1) - h.diag())
    return float(ious.mean())


def run_sharded(gops: Sequence, step_fn: Callable[[object], torch.Tensor], n_classes: int, mode: str = "gop",
                device: Optional[torch.device] = None) -> torch.Tensor:
    """Evaluate a list of GOP work items over all ranks and return the globally summed confusion matrix.

    `step_fn(item)` processes one work item on the calling rank and returns its int64 [n_cls*n_cls] histogram.
    mode 'gop': each rank takes whole GOPs (items g with g mod world == rank).
    mode 'frame': every rank sees every GOP; step_fn is expected to process only `frames_of_rank(...)` of it.
    N-rank result == 1-rank result exactly (integer histogram, frames independent)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    hist = torch.zeros(n_classes * n_classes, dtype=torch.int64, device=device)
    todo: Iterable[int] = gops_of_rank(len(gops), world, rank) if mode == "gop" else range(len(gops))
    for g in todo:
        hist += step_fn(gops[g]).to(hist.device)
    return allreduce_hist(hist)
