"""The reference's resampler acceptance thresholds, replayed through the CUDA
kernels (soundml/test/resample/resample_quality.ml).

The reference declares the resampler its one exception to value goldens and
gates it on measured decibels instead; these are the same signals, the same
ruler and the same limits -- Q1/Q2 SFDR and THD+N, Q3 out-of-band residual, Q4
passband flatness, Q5 the -3 dB edge against the committed soxr measurement, Q6
the swept-sine worst alias, Q9 the 44.1 -> 48 -> 44.1 round trip, and the x2 / /2
overlap-save classes -- for float32 (planned executors: tcgen05 GEMM and
overlap-save) and float64 (direct kernel).  ``pytest -m gpu``."""
import numpy as np
import pytest
import scipy.fft

from oracle import window_oracle

pytestmark = pytest.mark.gpu

MAINS = [(44100, 48000), (48000, 44100), (44100, 16000)]
TRIM = 0.15                      # resample_quality.ml:70
HALF_WIDTH = 16                  # resample_quality.ml:72


@pytest.fixture(scope="module")
def sb(lib):
    import torch
    assert torch.cuda.is_available()
    return lib


def tone(sr, f, seconds):
    n = int(round(sr * seconds))
    return np.sin(2.0 * np.pi * f * np.arange(n) / sr)


def tone_set(target):
    return [1000.0, 3000.0, 6500.0] if target == 16000 else [1000.0, 5000.0, 10000.0, 17000.0]


def kaiser(beta, n):
    return window_oracle.make("kaiser", n, True, beta)


def convert(sb, cfg, x, dtype):
    return sb.Resample.apply(cfg, x.astype(dtype)).astype(np.float64)


def spectrum(y):                                            # resample_quality.ml:84-102
    n = len(y)
    i0 = int(n * TRIM)
    cut = y[i0:n - i0]
    p = 1
    while p * 2 <= len(cut):
        p *= 2
    cut = cut[:p]
    return np.abs(scipy.fft.rfft(kaiser(30.0, p) * cut))


def sfdr_thdn(mags):                                        # resample_quality.ml:109-131
    p = int(np.argmax(mags))
    lo, hi = max(0, p - HALF_WIDTH), min(len(mags), p + HALF_WIDTH + 1)
    mask = np.ones(len(mags), bool)
    mask[lo:hi] = False
    spur = mags[mask & (np.arange(len(mags)) >= 2)].max()
    fund = (mags[lo:hi] ** 2).sum()
    rest = (mags[mask] ** 2).sum()
    return 20 * np.log10(mags[p] / spur), 10 * np.log10(rest / fund)


def amp_at(sr, f, x):                                       # resample_quality.ml:136-149
    n = len(x)
    i0 = int(n * TRIM)
    ln = n - 2 * i0
    w = kaiser(30.0, ln)
    ph = 2.0 * np.pi * f * (i0 + np.arange(ln)) / sr
    v = w * x[i0:i0 + ln]
    return 2.0 * np.hypot((v * np.cos(ph)).sum(), (v * np.sin(ph)).sum()) / w.sum()


def peak_dbfs(x):
    i0 = int(len(x) * TRIM)
    return 20 * np.log10(max(np.abs(x[i0:len(x) - i0]).max(), np.finfo(float).tiny))


PRECISIONS = [(np.float32, 125.0), (np.float64, 130.0)]


@pytest.mark.parametrize("sr,target", MAINS + [(8000, 48000), (48000, 8000)])
def test_q1_q2_tone_sfdr_and_thdn(sb, sr, target):
    cfg = sb.Resample.Config.create(sample_rate=sr, target=target)
    tones = tone_set(target) if (sr, target) in MAINS else [400.0, 1000.0, 3000.0]
    for dtype, sfdr_min in PRECISIONS:
        for f in tones:
            d, t = sfdr_thdn(spectrum(convert(sb, cfg, tone(sr, f, 2.0), dtype)))
            assert d >= sfdr_min, (sr, target, dtype.__name__, f, d)
            assert t <= -125.0, (sr, target, dtype.__name__, f, t)


def test_q3_out_of_band_tones_vanish(sb):
    cfg = sb.Resample.Config.create(sample_rate=44100, target=16000)
    for dtype, limit in ((np.float32, -125.0), (np.float64, -130.0)):
        for f in (9000.0, 12000.0, 18000.0):
            peak = peak_dbfs(convert(sb, cfg, tone(44100, f, 2.0), dtype))
            assert peak <= limit, (dtype.__name__, f, peak)
    for quality, limit in (("fast", -85.0), ("best", -130.0)):
        c = sb.Resample.Config.create(sample_rate=44100, target=16000, quality=quality)
        for f in (9000.0, 12000.0, 18000.0):
            assert peak_dbfs(convert(sb, c, tone(44100, f, 2.0), np.float64)) <= limit


@pytest.mark.parametrize("sr,target", MAINS)
def test_q4_passband_flatness(sb, sr, target):
    cfg = sb.Resample.Config.create(sample_rate=sr, target=target)
    edge = 0.913 * min(sr, target) / 2.0
    for dtype, limit in ((np.float32, 0.02), (np.float64, 0.01)):
        for f in (100.0, 0.25 * edge, 0.5 * edge, 0.75 * edge, edge):
            y = convert(sb, cfg, tone(sr, f, 1.0), dtype)
            dev = abs(20 * np.log10(amp_at(target, f, y)))
            assert dev <= limit, (sr, target, dtype.__name__, f, dev)


@pytest.mark.parametrize("sr,target", MAINS)
def test_q5_edge_against_soxr(sb, goldens, sr, target):
    expected = float(goldens.arrays[f"resample/soxr_reference/edge_hq_{sr}_{target}"][0])
    cfg = sb.Resample.Config.create(sample_rate=sr, target=target)
    goal = 1.0 / np.sqrt(2.0)

    def gain(f):
        return amp_at(target, f, convert(sb, cfg, tone(sr, f, 1.0), np.float64))
    nyq = min(sr, target) / 2.0
    lo, hi = 0.85 * nyq, 0.9995 * nyq
    assert gain(lo) > goal > gain(hi)
    for _ in range(40):
        mid = 0.5 * (lo + hi)
        if gain(mid) > goal:
            lo = mid
        else:
            hi = mid
    got = 0.5 * (lo + hi)
    assert abs(got - expected) / expected <= 0.01, (got, expected)


@pytest.mark.parametrize("sr,target", MAINS)
def test_q6_sweep_worst_alias(sb, sr, target):
    cfg = sb.Resample.Config.create(sample_rate=sr, target=target)
    seconds, nfft, hop = 30.0, 8192, 4096
    f0, f1 = 1000.0, 0.9 * target / 2.0
    rate = (f1 - f0) / seconds
    t = np.arange(int(sr * seconds)) / sr
    x = np.sin(2.0 * np.pi * (f0 * t + 0.5 * rate * t * t))
    w = window_oracle.make("kaiser", nfft, True, 16.0)
    guard = 3.0 * rate * (nfft / target) + 400.0
    skip = int(0.2 * target)
    bin_hz = target / nfft
    for dtype in (np.float32, np.float64):
        y = convert(sb, cfg, x, dtype)
        worst, start = -np.inf, skip
        while start + nfft <= len(y) - skip:
            f_inst = f0 + rate * ((start + nfft // 2) / target)
            mags = np.abs(scipy.fft.rfft(w * y[start:start + nfft]))
            near = np.abs(np.arange(len(mags)) * bin_hz - f_inst) <= guard
            valid = np.arange(len(mags)) >= 3
            worst = max(worst, 20 * np.log10(mags[valid & ~near].max() / mags[valid & near].max()))
            start += hop
        assert worst <= -100.0, (sr, target, dtype.__name__, worst)


def test_q9_round_trip_snr(sb):
    rng = np.random.default_rng(0x51AB)
    sr = n = 44100
    freqs = 100.0 + rng.uniform(0, 18000.0, 20)
    phases = rng.uniform(0, 2 * np.pi, 20)
    t = np.arange(n) / sr
    x = sum(0.05 * np.sin(2 * np.pi * f * t + ph) for f, ph in zip(freqs, phases))
    up = sb.Resample.Config.create(sample_rate=44100, target=48000)
    down = sb.Resample.Config.create(sample_rate=48000, target=44100)
    for dtype, limit in ((np.float32, 100.0), (np.float64, 110.0)):
        y = convert(sb, down, convert(sb, up, x, dtype), dtype)
        assert len(y) == n
        i0 = n // 10
        snr = 10 * np.log10((x[i0:n - i0] ** 2).sum() / ((x[i0:n - i0] - y[i0:n - i0]) ** 2).sum())
        assert snr >= limit, (dtype.__name__, snr)


# ---- Q10: the standard-rate matrix (resample_quality.ml:462-545) -------------------------
STANDARD_RATES = [8000, 11025, 16000, 22050, 24000, 32000, 44100, 48000, 88200, 96000, 192000]


@pytest.mark.parametrize("sr", STANDARD_RATES)
def test_q10_q1_q4_hold_over_all_110_standard_rate_pairs(sb, sr):
    """Every ordered pair of standard rates: tone SFDR >= 130 dB and THD+N <= -125 dB at
    four scaled positions, gain within 0.01 dB at three (float64 audio, the precision the
    reference's Q10 runs), the bank inside the documented 8 MB creation budget, and the
    planned float32 executors (overlap-save / tensor cores / blocked direct) bound to the
    float64 direct kernel within the resampler's float32 bar."""
    for target in STANDARD_RATES:
        if target == sr:
            continue
        cfg = sb.Resample.Config.create(sample_rate=sr, target=target)
        bank_bytes = cfg.l * (2 * cfg.latency + 1) * 8
        assert bank_bytes <= 8 * 1024 * 1024, (sr, target, bank_bytes)
        nyq = min(sr, target) / 2.0
        q12 = [frac * nyq for frac in (0.045, 0.23, 0.45, 0.79)]
        q4 = [frac * nyq for frac in (0.02, 0.5, 0.913)]
        x = np.stack([tone(sr, f, 1.0) for f in q12 + q4])           # one call per pair
        y = sb.Resample.apply(cfg, x)
        assert y.dtype == np.float64 and y.shape == (7, cfg.output_frames(x.shape[1]))
        for row, f in enumerate(q12):
            d, t = sfdr_thdn(spectrum(y[row]))
            assert d >= 130.0, (sr, target, f, d)
            assert t <= -125.0, (sr, target, f, t)
        for row, f in enumerate(q4, start=len(q12)):
            dev = abs(20 * np.log10(amp_at(target, f, y[row])))
            assert dev <= 0.01, (sr, target, f, dev)
        # the other surface stays bound to this one (the reference compares its GEMM surface
        # with the C kernel in float64, 32 ULP of peak; here the second surface is the planned
        # float32 path, so the bar is the float32 one)
        planned = sb.Resample.apply(cfg, x[1].astype(np.float32)).astype(np.float64)
        peak = np.abs(y[1]).max()
        assert np.abs(planned - y[1]).max() <= 1e-5 * peak, (sr, target)
