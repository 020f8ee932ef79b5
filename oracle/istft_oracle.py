"""Oracle: least-squares STFT synthesis (``Stft.invert``).

TEST INFRASTRUCTURE ONLY.  Restates ``soundml/lib/stft.ml:693-939`` of the
reference: inverse real transform of every frame in double (``Nx.irfft`` of the
un-vendored nx dependency, restated from its call site stft.ml:920 as
``scipy.fft.irfft(n=fft_size)``), times the analysis window, overlap-added in
padded coordinates, divided by the overlap-added squared window, trimmed of the
boundary extension and cut / zero-extended to the requested length.
Pinned by the reference's librosa goldens ``soundml/test/istft/vectors``
(tests/test_oracle_goldens.py).
"""
import numpy as np
import scipy.fft


def folded_square_window(c):
    """stft.ml:712-721: one entry per residue class modulo the hop."""
    w = c.analysis_window
    folded = np.zeros(c.hop, dtype=np.float64)
    for j in range(c.fft_size):                       # j ascending, as the reference
        folded[j % c.hop] += w[j] * w[j]
    return folded


def nola(c):
    """stft.ml:731-742."""
    if c.hop > c.fft_size:
        return False
    folded = folded_square_window(c)
    return bool(folded.min() > 1e-10 * max(folded.max(), 0.0))


def check_invertible(op, c):
    if not nola(c):
        raise ValueError(
            f"{op}: cannot invert a {c.win_length}-point window advanced by {c.hop} "
            f"samples inside a {c.fft_size}-point frame (the overlap-added squared "
            "window must stay above 1e-10 of its largest value at every position)")


def output_length(c, frames):
    """stft.ml:790-794."""
    if frames == 0:
        return 0
    return (frames - 1) * c.hop + c.fft_size - c.left_width() - c.right_width()


def _guard(v):
    return 1.0 if v == 0.0 else v


def envelope(c, frames):
    """stft.ml:846-894: overlap-added squared window over the padded span."""
    fft, hop = c.fft_size, c.hop
    w = c.analysis_window
    complete = folded_square_window(c)
    span = (frames - 1) * hop + fft
    head = min(span, fft - hop)
    stop = max(head, min(span, frames * hop))
    env = np.empty(span, dtype=np.float64)

    def partial(q):                                   # stft.ml:855-864, p ascending
        first = max(0, -((-(q - fft + 1)) // hop))
        last = min(frames - 1, q // hop)
        total = 0.0
        for p in range(first, last + 1):
            j = q - p * hop
            total += w[j] * w[j]
        return _guard(total)

    for q in range(0, head):
        env[q] = partial(q)
    for q in range(head, stop):
        env[q] = _guard(complete[q % hop])
    for q in range(stop, span):
        env[q] = partial(q)
    return env


def invert(c, z, length=None, dtype=np.float64):
    """``Stft.invert dtype c ?length z`` (stft.ml:896-939): ``[..., bins, frames]``
    complex -> ``[..., length]`` real."""
    z = np.asarray(z)
    if z.ndim < 2:
        raise ValueError(
            f"invert: cannot invert a rank-{z.ndim} tensor (the bin and frame axes must exist)")
    if z.shape[-2] != c.bins:
        raise ValueError(
            f"invert: cannot invert {z.shape[-2]} frequency bins of a {c.fft_size}-point "
            f"transform (the bin axis must hold fft_size / 2 + 1 = {c.bins} values)")
    if length is not None and length < 0:
        raise ValueError(
            f"invert: cannot synthesise a signal of length {length} (length must be non-negative)")
    check_invertible("invert", c)
    fft, hop, left = c.fft_size, c.hop, c.left_width()
    frames = z.shape[-1]
    lead = z.shape[:-2]
    out_len = output_length(c, frames) if length is None else length
    count = frames if length is None else min(frames, -((-(length + left)) // hop))
    if count == 0 or out_len == 0 or 0 in lead:
        return np.zeros(lead + (out_len,), dtype=dtype)
    zz = np.swapaxes(z[..., :count].astype(np.complex128), -1, -2)     # [..., count, bins]
    y = scipy.fft.irfft(zz, n=fft, axis=-1) * c.analysis_window
    span = (count - 1) * hop + fft
    acc = np.zeros(lead + (span,), dtype=np.float64)
    # the reference adds block plane k = 0, 1, ...: at one position that is
    # frame index descending (stft.ml:806-838)
    blocks = -(-fft // hop)
    for k in range(blocks):
        lo, hi = k * hop, min(fft, (k + 1) * hop)
        for p in range(count):
            acc[..., p * hop + lo:p * hop + hi] += y[..., p, lo:hi]
    acc = acc / envelope(c, count)
    stop = min(span, left + out_len)
    out = np.zeros(lead + (out_len,), dtype=np.float64)
    out[..., :stop - left] = acc[..., left:stop]
    return out.astype(dtype)


def griffin_lim(c, s, n_iter=32, momentum=0.99, init_phase=None, length=None):
    """``Stft.griffin_lim ?n_iter ?momentum ?init ?length c s`` (stft.ml:964-1025):
    fast Griffin-Lim phase reconstruction of a magnitude spectrogram ``[..., bins,
    frames]``; the loop runs in complex128 at the natural synthesis length, the
    result is rounded into the dtype of ``s``."""
    from . import stft_oracle
    s = np.asarray(s)
    if s.ndim < 2:
        raise ValueError(
            f"griffin_lim: cannot invert a rank-{s.ndim} tensor (the bin and frame axes must exist)")
    if s.shape[-2] != c.bins:
        raise ValueError(
            f"griffin_lim: cannot invert {s.shape[-2]} frequency bins of a {c.fft_size}-point "
            f"transform (the bin axis must hold fft_size / 2 + 1 = {c.bins} values)")
    if length is not None and length < 0:
        raise ValueError(
            f"griffin_lim: cannot synthesise a signal of length {length} "
            "(length must be non-negative)")
    check_invertible("griffin_lim", c)
    if n_iter < 1:
        raise ValueError(
            f"griffin_lim: cannot run {n_iter} iterations (n_iter must be at least 1)")
    if momentum < 0:
        raise ValueError(
            f"griffin_lim: cannot use a momentum of {momentum:g} (momentum must be non-negative)")
    mags = s.astype(np.float64).astype(np.complex128)
    if init_phase is None:
        angles = np.ones(s.shape, dtype=np.complex128)
    else:
        p = np.asarray(init_phase, dtype=np.float64)
        if p.shape != s.shape:
            raise ValueError("griffin_lim: the initial phase must have the shape of the magnitudes")
        angles = np.cos(p) + 1j * np.sin(p)
    frames = s.shape[-1]
    beta = momentum / (1.0 + momentum)
    iterate = output_length(c, frames) > 0 and frames > 0 and 0 not in s.shape[:-2]
    previous = None
    tiny = np.finfo(np.float64).tiny                   # Float.min_float
    for _ in range(n_iter if iterate else 0):
        rebuilt = stft_oracle.transform(c, invert(c, mags * angles))[..., :frames]
        extrapolated = rebuilt if previous is None else rebuilt - previous * beta
        angles = extrapolated / (np.abs(extrapolated) + tiny)
        previous = rebuilt
    return invert(c, mags * angles, length=length, dtype=s.dtype)
