"""Oracle: window generation in float64.  TEST INFRASTRUCTURE ONLY.

Follows ``soundml/lib/window.ml`` of the reference:

* ``fill_window`` (window.ml:362-364): a one-point window is 1; otherwise the
  periodic window of length n is the symmetric window of m = n + 1 points with
  the last sample dropped.
* ``cosine_fill`` (window.ml:147-166): theta_i = (2i - (m-1)) * pi/(m-1),
  value sum_k a_k T_k(cos theta) by the Chebyshev recurrence, first half
  evaluated and mirrored.
* ``bartlett_fill`` (:170-176), ``gaussian_fill`` (:180-188),
  ``tukey_fill`` (:195-207), ``kaiser_fill`` (:312-332).

Kaiser: the reference evaluates I0 through two polynomial branches that are a
speed optimisation of the same function (documented error <= 35 ulp,
window.ml:300-305); this oracle sums the defining power series
(window.ml:104-111), which agrees far inside the reference's own 1e-12 gate.
"""
import math

import numpy as np

COSINE = {
    "hann": (0.5, 0.5),
    "hamming": (0.54, 0.46),
    "blackman": (0.42, 0.5, 0.08),
    "blackman_harris": (0.35875, 0.48829, 0.14128, 0.01168),
    "nuttall": (0.3635819, 0.4891775, 0.1365995, 0.0106411),
    "flat_top": (0.21557895, 0.41663158, 0.277263158, 0.083578947, 0.006947368),
}


def bessel_i0(x):
    """window.ml:104-111 — power series, stop at term <= 1e-17 * sum."""
    q = 0.25 * x * x
    if q == 0.0:
        return 1.0
    term, total, k = 1.0, 1.0, 1
    while True:
        term = term * q / float(k * k)
        total += term
        if term <= 1e-17 * total:
            return total
        k += 1


def _put(buf, i, v):
    if i < len(buf):
        buf[i] = v


def _cosine_fill(buf, coeffs, m):
    step = math.pi / float(m - 1)
    for i in range((m - 1) // 2 + 1):
        c = math.cos(float(2 * i - (m - 1)) * step)
        acc = coeffs[0] + coeffs[1] * c
        prev, cur = 1.0, c
        for k in range(2, len(coeffs)):
            t = 2.0 * c * cur - prev
            acc = acc + coeffs[k] * t
            prev, cur = cur, t
        _put(buf, i, acc)
        _put(buf, m - 1 - i, acc)


def _fill(buf, kind, param, m):
    n = len(buf)
    if kind == "rectangular":
        buf[:] = 1.0
    elif kind in COSINE:
        _cosine_fill(buf, COSINE[kind], m)
    elif kind == "bartlett":
        last = float(m - 1)
        for i in range((m - 1) // 2 + 1):
            v = 2.0 * float(i) / last
            _put(buf, i, v)
            _put(buf, m - 1 - i, v)
    elif kind == "gaussian":
        half = float(m - 1) / 2.0
        scale = -1.0 / (2.0 * param * param)
        for i in range((m - 1) // 2 + 1):
            x = float(i) - half
            v = math.exp(x * x * scale)
            _put(buf, i, v)
            _put(buf, m - 1 - i, v)
    elif kind == "tukey":
        if param <= 0.0:
            buf[:] = 1.0
        elif param >= 1.0:
            _cosine_fill(buf, COSINE["hann"], m)
        else:
            last = float(m - 1)
            width = int(math.floor(param * last / 2.0))
            step = 2.0 / param / last
            for i in range(width + 1):
                v = 0.5 * (1.0 + math.cos(math.pi * (-1.0 + step * float(i))))
                _put(buf, i, v)
                _put(buf, m - 1 - i, v)
            for i in range(width + 1, m - width - 1):
                _put(buf, i, 1.0)
    elif kind == "kaiser":
        alpha = float(m - 1) / 2.0
        denom = bessel_i0(param)
        for i in range((m - 1) // 2 + 1):
            r = (float(i) - alpha) / alpha
            v = bessel_i0(param * math.sqrt(max(0.0, 1.0 - r * r))) / denom
            _put(buf, i, v)
            _put(buf, m - 1 - i, v)
    else:
        raise ValueError(f"unknown window {kind!r}")
    assert n in (m, m - 1)


def make(kind, n, periodic=True, param=0.0):
    """``Window.make Nx.float64 ~periodic kind n`` (window.ml:374-401)."""
    if n < 1:
        raise ValueError(
            f"make: cannot make a {n}-point window (length must be at least 1)")
    buf = np.zeros(n, dtype=np.float64)
    if n == 1:
        buf[0] = 1.0
        return buf
    _fill(buf, kind, param, n + 1 if periodic else n)
    return buf


def cola(kind, length, hop, param=0.0):
    """``Window.cola`` (window.ml:409-434)."""
    w = make(kind, length, True, param)
    sums = np.zeros(hop)
    for i in range(length):
        sums[i % hop] += w[i]
    mean = sums.sum() / hop
    return bool(mean > 0 and np.all(np.abs(sums - mean) <= 1e-10 * mean))
