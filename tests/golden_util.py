"""Loader for tests/golden/reference_vectors.npz (see make_golden.py) and the
reference's deterministic test signals."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

STFT_SEED = 20250803     # soundml/test/stft/stft_goldens.ml:19-23
MEL_SEED = 20260803      # soundml/test/mel/mel_goldens.ml:34-42

# the reference's gates (stft_goldens.ml:15-17, test/support/tutils.ml:80-86)
F64_RTOL, F64_ATOL = 1e-9, 1e-12
F32_RTOL, F32_ATOL = 1e-6, 1e-7
WINDOW_RTOL, WINDOW_ATOL = 1e-12, 1e-15


def lcg_signal(n, seed, envelope=False):
    """31-bit LCG of the reference suites, bit-exact in float64."""
    state, out = seed, np.empty(n, dtype=np.float64)
    for i in range(n):
        state = (1103515245 * state + 12345) % (1 << 31)
        out[i] = state / float(1 << 30) - 1.0
    if envelope:
        out = out * np.exp(-12.0 * np.arange(n, dtype=np.float64) / float(n))
    return out


ISTFT_SEED_RE, ISTFT_SEED_IM = 20250803, 20250804   # soundml/test/istft/istft_goldens.ml:38-47


def istft_spectrum(fft_size, frames, dtype):
    """The synthetic (inconsistent) spectrum of the reference's synthesis goldens:
    two LCG streams, one per complex component, DC / Nyquist made real
    (istft_goldens.ml:30-47); float32 cases quantise both components."""
    bins = fft_size // 2 + 1
    re = lcg_signal(bins * frames, ISTFT_SEED_RE).reshape(bins, frames)
    im = lcg_signal(bins * frames, ISTFT_SEED_IM).reshape(bins, frames)
    im[0, :] = 0.0
    if fft_size % 2 == 0:
        im[bins - 1, :] = 0.0
    z = re + 1j * im
    return z.astype(np.complex64 if dtype == "float32" else np.complex128)


def griffin_lim_magnitudes(fft_size, frames, dtype):
    """Magnitudes of the reference's Griffin-Lim goldens: LCG + 1, so strictly
    positive (soundml/test/istft/gl_goldens.ml:35-37)."""
    bins = fft_size // 2 + 1
    m = (lcg_signal(bins * frames, ISTFT_SEED_RE) + 1.0).reshape(bins, frames)
    return m.astype(np.float32 if dtype == "float32" else np.float64)


class Goldens:
    def __init__(self):
        z = np.load(os.path.join(HERE, "golden", "reference_vectors.npz"))
        self.index = json.loads(bytes(z["__index__"]).decode())
        self.arrays = z

    def cases(self, suite, stem=None):
        for key, entry in self.index.items():
            s, f, name = key.split("/")
            if s == suite and (stem is None or f == stem):
                yield key, f, name, entry

    def values(self, key):
        entry = self.index[key]
        return self.arrays[key].reshape(entry["shape"])


def window_spec(params):
    name = params["window"]
    for p in ("beta", "std", "taper"):
        if p in params:
            return (name, params[p])
    return name


def assert_close(got, expected, rtol, atol, msg=""):
    got = np.asarray(got, dtype=np.float64)
    expected = np.asarray(expected, dtype=np.float64)
    assert got.shape == expected.shape, f"{msg}: shape {got.shape} != {expected.shape}"
    tol = atol + rtol * np.abs(expected)
    bad = np.abs(got - expected) > tol
    if bad.any():
        i = np.argmax(np.abs(got - expected) / tol)
        raise AssertionError(
            f"{msg}: {bad.sum()} of {bad.size} outside tolerance; worst at flat index {i}: "
            f"got {got.flat[i]!r}, expected {expected.flat[i]!r}")


def peak_rel_err(got, ref):
    """max |got - ref| / max |ref| — the yardstick of BASELINE.md for float32."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    peak = np.abs(ref).max()
    return float(np.abs(got - ref).max() / peak) if peak > 0 else float(np.abs(got - ref).max())
