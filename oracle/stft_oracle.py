"""Oracle: STFT analysis (framing, padding, windowed rFFT, |X|^p).

TEST INFRASTRUCTURE ONLY.  Restates ``soundml/lib/stft.ml`` of the reference;
the FFT itself lives in the un-vendored ``nx`` dependency
(``dune-project:22-26``, gabyfle/raven pin cec410b0), so ``Nx.stft`` is
restated from its call site (stft.ml:356-364) as: strided frames x float64
window -> batched rfft in double -> one rounding into the complex storage.
Pinned by the reference's librosa goldens (tests/test_oracle_goldens.py).
"""
import numpy as np
import scipy.fft

from . import window_oracle


class StftConfig:
    """``Stft.Config.create`` (stft.ml:61-111)."""

    def __init__(self, fft_size, hop=None, win_length=None, window="hann",
                 window_param=0.0, alignment="centered", pad="reflect",
                 pad_value=0.0, scale="none"):
        if fft_size < 1:
            raise ValueError(
                f"create: cannot use an FFT of size {fft_size} "
                "(fft_size must be at least 1)")
        win_length = fft_size if win_length is None else win_length
        if win_length < 1 or win_length > fft_size:
            raise ValueError(
                f"create: cannot use a {win_length}-point window with an FFT "
                f"of size {fft_size} (win_length must lie in [1, fft_size])")
        hop = max(1, fft_size // 4) if hop is None else hop
        if hop < 1:
            raise ValueError(
                f"create: cannot advance frames by {hop} samples "
                "(hop must be at least 1)")
        self.fft_size, self.hop, self.win_length = fft_size, hop, win_length
        self.window, self.window_param = window, window_param
        self.alignment, self.pad, self.pad_value = alignment, pad, pad_value
        self.scale = scale
        coeff = window_oracle.make(window, win_length, True, window_param)
        left = (fft_size - win_length) // 2          # stft.ml:97-102
        w = np.zeros(fft_size, dtype=np.float64)
        w[left:left + win_length] = coeff
        if scale == "magnitude":                      # stft.ml:103-109
            w = w / w.sum()
        elif scale == "psd":
            w = w / np.sqrt((w * w).sum())
        self.analysis_window = w

    @property
    def bins(self):
        return self.fft_size // 2 + 1

    def left_width(self):                             # stft.ml:132-139
        return {"centered": self.fft_size // 2, "left": 0,
                "right": self.fft_size - 1}[self.alignment]

    def right_width(self):                            # stft.ml:141-142
        return self.fft_size // 2 if self.alignment == "centered" else 0


def frames(c, n):
    """stft.ml:217-223."""
    if n < 0:
        raise ValueError(
            f"frames: cannot analyse a signal of length {n} "
            "(length must be non-negative)")
    if n == 0:
        return 0
    padded = n + c.left_width() + c.right_width()
    if padded < c.fft_size:
        return 0
    return 1 + (padded - c.fft_size) // c.hop


def reflect_index(n, q):
    """stft.ml:300-305."""
    if n == 1:
        return 0
    period = 2 * (n - 1)
    m = ((q % period) + period) % period
    return m if m < n else period - m


def source_indices(c, n):
    """Source index (or -1 for a constant fill) of every padded position:
    ``pad_signal`` (stft.ml:318-338) expressed as one index vector."""
    left, right = c.left_width(), c.right_width()
    q = np.arange(-left, n + right, dtype=np.int64)
    if c.pad == "reflect":
        if n == 1:
            return np.zeros_like(q)
        period = 2 * (n - 1)
        m = np.mod(np.mod(q, period) + period, period)
        return np.where(m < n, m, period - m)
    if c.pad == "edge":
        return np.clip(q, 0, n - 1)
    idx = q.copy()
    idx[(q < 0) | (q >= n)] = -1
    return idx


def pad_signal(c, x):
    n = x.shape[-1]
    idx = source_indices(c, n)
    out = x[..., np.maximum(idx, 0)]
    if c.pad == "constant":
        out = np.where(idx < 0, np.asarray(c.pad_value, dtype=x.dtype), out)
    return out


def transform(c, x, workers=1):
    """``Stft.transform`` (stft.ml:632-650) in its single-shot form
    ``transform_range`` (stft.ml:652-666): pad, frame, float64 window multiply,
    rfft in double, round once (complex64 for float32 audio, stft.ml:681-685).
    Returns ``[..., bins, frames]``."""
    x = np.asarray(x)
    n = x.shape[-1]
    count = frames(c, n)
    lead = x.shape[:-1]
    cdtype = np.complex128 if x.dtype == np.float64 else np.complex64
    if count == 0 or 0 in lead:
        return np.zeros(lead + (c.bins, count), dtype=cdtype)
    padded = pad_signal(c, x).astype(np.float64)      # stft.ml:345-346
    span = (count - 1) * c.hop + c.fft_size
    padded = padded[..., :span]
    fr = np.lib.stride_tricks.sliding_window_view(
        padded, c.fft_size, axis=-1)[..., ::c.hop, :]
    spec = scipy.fft.rfft(fr * c.analysis_window, axis=-1, workers=workers)
    return np.ascontiguousarray(np.swapaxes(spec, -1, -2)).astype(cdtype)


def power_spectrum(c, x, power=2.0, workers=1):
    """``Stft.power_spectrum`` (stft.ml:670-691): |z| lands in the input's
    float dtype first, then the power is taken in that dtype."""
    x = np.asarray(x)
    z = transform(c, x, workers=workers)
    m = np.abs(z).astype(x.dtype)
    if power == 2.0:
        return m * m
    if power == 1.0:
        return m
    return m ** np.asarray(power, dtype=x.dtype)


def times(c, sample_rate, n):
    """stft.ml:245-254."""
    return np.arange(frames(c, n), dtype=np.float64) * float(c.hop) / float(sample_rate)


def frequencies(c, sample_rate):
    """stft.ml:256-261."""
    return np.arange(c.bins, dtype=np.float64) * (float(sample_rate) / float(c.fft_size))
