"""Oracle: mel filterbank weights and projection.  TEST INFRASTRUCTURE ONLY.

Restates ``soundml/lib/mel.ml:39-164,202-231`` and the mel scale of
``soundml/lib/convert.ml:72-102`` in float64, operation for operation (the
reference documents its bin-frequency and breakpoint arithmetic as
bit-parity-with-librosa, mel.ml:39-60).
"""
import math

import numpy as np

F_SP = 200.0 / 3.0                  # convert.ml:74
MIN_LOG_HZ = 1000.0
MIN_LOG_MEL = MIN_LOG_HZ / F_SP
LOGSTEP = math.log(6.4) / 27.0


def hz_to_mel(f, scale="slaney"):
    f = np.asarray(f, dtype=np.float64)
    if scale == "htk":
        return np.log(f / 700.0 + 1.0) * (2595.0 / math.log(10.0))
    linear = f / F_SP
    with np.errstate(divide="ignore", invalid="ignore"):
        log_branch = np.log(f / MIN_LOG_HZ) / LOGSTEP + MIN_LOG_MEL
    return np.where(f < MIN_LOG_HZ, linear, log_branch)


def mel_to_hz(m, scale="slaney"):
    m = np.asarray(m, dtype=np.float64)
    if scale == "htk":
        return (np.exp(m * (math.log(10.0) / 2595.0)) - 1.0) * 700.0
    linear = m * F_SP
    log_branch = np.exp((m - MIN_LOG_MEL) * LOGSTEP) * MIN_LOG_HZ
    return np.where(m < MIN_LOG_MEL, linear, log_branch)


class MelConfig:
    """``Mel.Config.create`` (mel.ml:119-164)."""

    def __init__(self, n_mels, sample_rate, fft_size, f_min=0.0, f_max=None,
                 scale="slaney", norm="slaney"):
        if n_mels < 1:
            raise ValueError(
                f"create: cannot build {n_mels} mel bands (n_mels must be at least 1)")
        if sample_rate < 1:
            raise ValueError(
                f"create: cannot use a sample rate of {sample_rate} Hz "
                "(sample_rate must be at least 1)")
        if fft_size < 1:
            raise ValueError(
                f"create: cannot use an FFT of size {fft_size} "
                "(fft_size must be at least 1)")
        if not (math.isfinite(f_min) and f_min >= 0.0):
            raise ValueError("create: f_min must be finite and non-negative")
        nyquist = float(sample_rate) / 2.0
        f_max = nyquist if f_max is None else float(f_max)
        if not (math.isfinite(f_max) and f_max > f_min):
            raise ValueError("create: f_max must be finite and greater than f_min")
        if f_max > nyquist:
            raise ValueError("create: f_max must not exceed the Nyquist frequency")
        self.n_mels, self.sample_rate, self.fft_size = n_mels, sample_rate, fft_size
        self.f_min, self.f_max, self.scale, self.norm = f_min, f_max, scale, norm
        self.weights = weights_of(f_min, f_max, scale, norm, n_mels,
                                  sample_rate, fft_size)

    @property
    def bins(self):
        return self.fft_size // 2 + 1


def fft_frequencies(sample_rate, fft_size, bins):
    """mel.ml:39-43 — one reciprocal, one multiply per bin."""
    step = 1.0 / (float(fft_size) * (1.0 / float(sample_rate)))
    return np.arange(bins, dtype=np.float64) * step


def breakpoints(scale, f_min, f_max, count):
    """mel.ml:50-62 — linspace as i*step + min with the endpoint pinned."""
    bounds = hz_to_mel(np.array([f_min, f_max]), scale)
    mel_min, mel_max = float(bounds[0]), float(bounds[1])
    step = (mel_max - mel_min) / float(count - 1)
    mels = np.array([mel_max if i == count - 1 else float(i) * step + mel_min
                     for i in range(count)], dtype=np.float64)
    return mel_to_hz(mels, scale)


def weights_of(f_min, f_max, scale, norm, n_mels, sample_rate, fft_size):
    """mel.ml:69-117."""
    bins = fft_size // 2 + 1
    count = n_mels + 2
    points = breakpoints(scale, f_min, f_max, count)
    steps = points[1:] - points[:-1]
    if np.any(steps <= 0.0):
        raise ValueError(
            f"create: cannot resolve {n_mels} mel bands between {f_min:g} and "
            f"{f_max:g} Hz (adjacent breakpoints collapse in double precision)")
    ramps = points[:, None] - fft_frequencies(sample_rate, fft_size, bins)[None, :]
    lower = (-ramps[:n_mels]) / steps[:n_mels, None]
    upper = ramps[2:count] / steps[1:n_mels + 1, None]
    weights = np.maximum(0.0, np.minimum(lower, upper))
    if np.any(weights.max(axis=-1) <= 0.0):
        raise ValueError(
            f"create: cannot support {n_mels} mel bands with an FFT of size "
            f"{fft_size} (at least one filter spans no FFT bin; raise fft_size "
            "or lower n_mels)")
    if norm == "slaney":
        span = points[2:count] - points[:n_mels]
        weights = weights * (2.0 / span)[:, None]
    return weights


def apply(c, s):
    """``Mel.apply`` (mel.ml:202-231): float64 matmul, one rounding."""
    s = np.asarray(s)
    if s.ndim < 2:
        raise ValueError(
            f"apply: cannot project a rank-{s.ndim} tensor (the mel projection "
            "needs [...; bins; frames])")
    if s.shape[-2] != c.bins:
        raise ValueError(
            f"apply: cannot project {s.shape[-2]} frequency bins through a "
            f"filterbank built for an FFT of size {c.fft_size} ({c.bins} bins)")
    if 0 in s.shape:
        return np.zeros(s.shape[:-2] + (c.n_mels, s.shape[-1]), dtype=s.dtype)
    return np.matmul(c.weights, s.astype(np.float64)).astype(s.dtype)


def mel_spectrogram(stft_cfg, mel_cfg, x, power=2.0, workers=1):
    """``Soundml.mel_spectrogram`` (soundml.ml:12-24)."""
    from . import stft_oracle
    if stft_cfg.fft_size != mel_cfg.fft_size:
        raise ValueError("mel_spectrogram: fft sizes disagree")
    return apply(mel_cfg, stft_oracle.power_spectrum(stft_cfg, x, power, workers))
