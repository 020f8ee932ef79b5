"""Oracle: decibel conversions and MFCC.  TEST INFRASTRUCTURE ONLY.

Restates ``soundml/lib/convert.ml:20-56`` (``to_db``: floor at amin, log in the
input's own dtype, scale, offset, optional whole-tensor ``top_db`` clamp) and
``soundml/lib/soundml.ml:26-95`` (``mfcc``: log-mel with the 80 dB clamp, raw
type-II DCT along the mel axis, orthonormal row scales, sinusoidal lifter, all
in double, one rounding).  Pinned by the reference's ``db/vectors/*.json`` and
``mel/vectors/mfcc.json`` goldens (tests/test_oracle_goldens.py).
"""
import math

import numpy as np
import scipy.fft

DECADE = 10.0 / math.log(10.0)          # convert.ml:27


def _check(fn, reference, amin, top_db):
    if not (math.isfinite(reference) and reference > 0.0):
        raise ValueError(f"Soundml.Convert.{fn}: reference must be finite and positive")
    if not (math.isfinite(amin) and amin > 0.0):
        raise ValueError(f"Soundml.Convert.{fn}: amin must be finite and positive")
    if top_db is not None and not (math.isfinite(top_db) and top_db >= 0.0):
        raise ValueError(f"Soundml.Convert.{fn}: top_db must be finite and non-negative")


def _to_db(gain, magnitude, reference, amin, top_db, s):
    s = np.asarray(s)
    if s.size == 0:
        return s.copy()
    dt = s.dtype.type
    scale = gain / 10.0 * DECADE
    if magnitude:
        s = np.abs(s)
    floored = np.maximum(s, dt(amin))
    offset = scale * math.log(max(amin, reference))
    db = np.log(floored) * dt(scale) - dt(offset)       # every op in the input dtype
    if top_db is None:
        return db
    maximum = float(db.max())
    return np.maximum(db, dt(maximum - top_db))


def power_to_db(s, reference=1.0, amin=1e-10, top_db=None):
    _check("power_to_db", reference, amin, top_db)
    return _to_db(10.0, False, reference, amin, top_db, s)


def amplitude_to_db(s, reference=1.0, amin=1e-5, top_db=None):
    _check("amplitude_to_db", reference, amin, top_db)
    return _to_db(20.0, True, reference, amin, top_db, s)


def mfcc(stft_cfg, mel_cfg, x, n_mfcc=20, lifter=None, workers=1):
    """``Soundml.mfcc`` (soundml.ml:50-95)."""
    from . import mel_oracle
    n_mels = mel_cfg.n_mels
    if n_mfcc < 1 or n_mfcc > n_mels:
        raise ValueError(
            f"mfcc: cannot keep {n_mfcc} cepstral coefficients of {n_mels} mel bands "
            "(n_mfcc must lie in [1, n_mels])")
    if lifter is not None and not (math.isfinite(lifter) and lifter >= 0.0):
        raise ValueError(
            f"mfcc: cannot lifter with a coefficient of {lifter:g} (lifter must be finite and "
            "non-negative)")
    x = np.asarray(x)
    mel = mel_oracle.mel_spectrogram(stft_cfg, mel_cfg, x, 2.0, workers)
    if mel.size == 0:
        return np.zeros(mel.shape[:-2] + (n_mfcc, mel.shape[-1]), dtype=x.dtype)
    db = power_to_db(mel.astype(np.float64), top_db=80.0)
    raw = scipy.fft.dct(db, type=2, axis=-2)[..., :n_mfcc, :]
    scales = np.array([1.0 / math.sqrt(4.0 * n_mels) if k == 0 else 1.0 / math.sqrt(2.0 * n_mels)
                       for k in range(n_mfcc)])[:, None]
    cep = raw * scales
    if lifter is not None and lifter > 0.0:
        w = np.array([1.0 + lifter / 2.0 * math.sin(math.pi * (k + 1) / lifter)
                      for k in range(n_mfcc)])[:, None]
        cep = cep * w
    return cep.astype(x.dtype)
