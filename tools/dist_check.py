"""N-GPU correctness check (run under torchrun, one rank per GPU, NCCL): frame-level sharding of one GOP.

Rank 0 owns the keyframe feature and broadcasts it (ncclBroadcast over NVLink); the 11 non-keyframes are dealt over
the ranks (arseg_b200.dist.frames_of_rank); every rank runs its share through NonKeyEngine; the class maps are
gathered on rank 0 and must equal, bit for bit, the maps rank 0 computes for all 11 frames alone.  The per-rank
confusion matrices are all-reduced and compared with the single-GPU histogram (evaluation.py:205-211).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from arseg_b200 import dist as adist  # noqa: E402
from arseg_b200 import evaluation as ev  # noqa: E402
from arseg_b200 import models, ops, synth  # noqa: E402


def main():
    torch.set_grad_enabled(False)
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    arch, H, W, scale, NF, ncls = "camvid-psp18", 360, 480, 0.5, 11, 12
    sd = synth.synth_state_dict(models.models_fuse[arch]().state_dict(), 4)
    frames = torch.cat([synth.synth_frame(1, H, W, 10 + i) for i in range(NF)])
    mvs = torch.from_numpy(np.stack([synth.synth_mv_int16(H, W, 50 + d, distance=d) for d in range(1, NF + 1)]))
    labels = torch.randint(0, ncls, (NF, H, W), generator=torch.Generator().manual_seed(7))
    # only the owner has the keyframe feature; the others start from garbage and must receive it
    ref_p = (synth.synth_feature(1, 64, H, W, 3) * 0.5).to(dev) if rank == 0 else torch.full((1, 64, H, W), float("nan"), device=dev)
    my = adist.frames_of_rank(NF, world, rank, owner=0)
    eng = ev.NonKeyEngine(arch, sd, max(len(my), 1), H, W, scale, "f16", device=dev)
    side = torch.cuda.Stream(dev)
    e = adist.broadcast_keyframe_feature(ref_p, src=0, stream=side)
    if e is not None:
        torch.cuda.current_stream().wait_event(e)
    hist = torch.zeros(ncls * ncls, dtype=torch.int64, device=dev)
    out = torch.zeros((NF, H, W), dtype=torch.uint8, device=dev)
    if my:
        eng.set_inputs(frames[my].to(dev), mvs[my].to(dev), ref_p)
        preds = eng.step()
        out[my] = preds
        hist += ops.confusion_hist(preds, labels[my].to(dev), ncls).flatten()
    dist.all_reduce(out, op=dist.ReduceOp.SUM)          # disjoint shares: sum == gather
    adist.allreduce_hist(hist)
    if rank == 0:
        full = ev.NonKeyEngine(arch, sd, NF, H, W, scale, "f16", device=dev)
        full.set_inputs(frames.to(dev), mvs.to(dev), ref_p)
        want = full.step()
        want_hist = ops.confusion_hist(want, labels.to(dev), ncls).flatten()
        torch.cuda.synchronize()
        same = bool(torch.equal(out, want))
        print("dist_check world=%d shares=%s: class maps identical=%s, histogram identical=%s, mIoU %.4f" %
              (world, [len(adist.frames_of_rank(NF, world, r)) for r in range(world)], same, bool(torch.equal(hist, want_hist)),
               adist.miou_from_hist(hist.view(ncls, ncls))))
        assert same and torch.equal(hist, want_hist)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
