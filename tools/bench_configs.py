"""BASELINE configs 2-4 on one GPU: non-keyframe frames/s of the fused NonKeyEngine for every architecture.

    python tools/bench_configs.py [--steps 10] [--precision f16] [--out gpurun_out/configs.md]

  config 2  CamVid 720x960     PSPNet-18  AR-0.5x  (p: 64 ch at frame resolution, k=7)      <- bench.py's workload
  config 3  CamVid 720x960     BiSeNet-18 AR-0.5x  (p: 256 ch at 1/8 resolution)
  config 4  Cityscapes 1024x2048 PSPNet-18 AR-0.5x (p: 512 ch at 1/8 resolution, 19 classes)
Each step = the 11 non-keyframes of one GOP, inputs resident, CUDA-graph replay, L2 flushed between steps.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from arseg_b200 import evaluation as ev  # noqa: E402
from arseg_b200 import models, synth  # noqa: E402

CONFIGS = [("config 2", "camvid-psp18", 720, 960), ("config 3", "camvid-bise18", 720, 960), ("config 4", "cityscapes-psp18", 1024, 2048)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--precision", default="f16")
    ap.add_argument("--only", default="", help="substring of the config name / architecture to run alone")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.md"))
    a = ap.parse_args()
    torch.set_grad_enabled(False)
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lines = ["| config | architecture | frame | precision | launches/step | ms/step (11 frames) | frames/s | top kernels |", "|---|---|---|---|---|---|---|---|"]
    for name, arch, H, W in CONFIGS:
        if a.only and a.only not in name and a.only not in arch:
            continue
        C_, stride, ncls = ev.ARCH_INFO[arch]
        sd = synth.synth_state_dict(models.models_fuse[arch]().state_dict(), 4)
        N = 11
        eng = ev.NonKeyEngine(arch, sd, N, H, W, 0.5, a.precision, device=dev)
        frames = torch.cat([synth.synth_frame(1, H, W, i) for i in range(N)])
        mvs = torch.from_numpy(np.stack([synth.synth_mv_int16(H, W, 20 + d, distance=d) for d in range(1, N + 1)]))
        ref_p = synth.synth_feature(1, C_, H // stride, W // stride, 3) * 0.5
        eng.set_inputs(frames.to(dev), mvs.to(dev), ref_p.to(dev))
        for _ in range(3):
            eng.step()
        ts = []
        for _ in range(a.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.step(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sum(ts) / len(ts)
        prof = sorted(eng.plan.profile(iters=2, warmup=1), key=lambda x: -x[1])[:4]
        lines.append("| %s | %s | %dx%d | %s | %d | %.3f | %.0f | %s |" % (name, arch, H, W, a.precision, eng.launches_per_step, ms, N / ms * 1e3,
                                                                       "; ".join("%s %.3f ms" % (n, t) for n, t in prof)))
        print(lines[-1], flush=True)
        del eng
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    open(a.out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
