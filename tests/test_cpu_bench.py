"""CPU suite: bench.py's contract pieces that need no GPU."""
import importlib.util
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_algorithmic_bytes_match_survey_8d():
    """176.9 + 44.2 + 2.8 + 176.9 + 33.2 = 434.0 MB per frame (CamVid-PSP, AR-0.5x); the bytes ONE launch has to
    move count the keyframe feature once per launch (the 11 frames of a GOP share it) and no fused-p write."""
    b = _bench()
    lr_numel, logits_numel = 64 * 360 * 480, 12 * 720 * 960
    assert abs(b.creff_bytes_survey_8d(lr_numel, logits_numel) / 1e6 - 434.0) < 0.5
    one = b.creff_bytes_moved(1, 4, lr_numel, 4, logits_numel, write_p=False)
    assert abs(one / 1e6 - (176.9 + 44.2 + 2.8 + 33.2 + 0.7)) < 0.5
    eleven = b.creff_bytes_moved(11, 4, lr_numel, 4, logits_numel, write_p=False)
    assert abs(eleven / 1e6 - (176.9 + 11 * (44.2 + 2.8 + 33.2 + 0.7))) < 1.0          # 1.07 GB: what ncu measures as DRAM traffic
    assert b.creff_bytes_moved(1, 4, lr_numel, 4, logits_numel, True) - one == 64 * 720 * 960 * 4
    # f16 LR feature (the f16 plan) and f16 keyframe feature (tcgen05 engine) halve those terms
    assert one - b.creff_bytes_moved(1, 2, lr_numel, 2, logits_numel, False) == (64 * 720 * 960 + lr_numel) * 2


def test_workloads_cover_baseline_configs():
    b = _bench()
    assert set(b.WORKLOADS) == {"camvid-psp18", "camvid-bise18", "cityscapes-psp18"}
    b.set_workload("cityscapes-psp18")
    assert (b.H, b.W, b.C_P, b.STRIDE_P, b.N_CLS) == (1024, 2048, 512, 8, 19) and "1024x2048" in b.metric_name()
    b.set_workload("camvid-psp18")
    assert "720x960" in b.metric_name()


def test_peaks_come_from_measured_file_when_present():
    b = _bench()
    hbm, burst, sust, src = b.peaks()
    assert hbm > 1000 and burst >= sust > 100 and src in ("measured", "fallback")


def test_clock_sampler_summary_shapes():
    b = _bench()
    s = b.ClockSampler(0)
    assert s.summary()["reasons"] == ["unsampled"]
    s.sm, s.sm_max, s.bits = [1965.0, 1950.0, 1965.0], 1965.0, 0x4
    out = s.summary()
    assert out["sm_mhz"] == 1965.0 and out["sm_min_mhz"] == 1950.0 and out["reasons"] == ["sw_power_cap"]


def test_product_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback: without CUDA the product arm exits with a message instead of timing anything."""
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
