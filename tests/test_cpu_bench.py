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
    """176.9 + 44.2 + 2.8 + 176.9 + 33.2 = 434.0 MB per frame (CamVid-PSP, AR-0.5x), and the bytes the engine
    has to move when the fused p is not materialised (fp32 LR feature, log-probs + u8 class map out)."""
    b = _bench()
    assert abs(b.CREFF_BYTES_FULL / 1e6 - 434.0) < 0.5
    moved = b.creff_bytes(4, write_p=False, write_logits=True)
    assert abs(moved / 1e6 - (176.9 + 44.2 + 2.8 + 33.2 + 0.7)) < 0.5
    assert b.creff_bytes(4, True, True) - moved == 64 * 720 * 960 * 4


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
