"""Oracle: polyphase resampler / FIR in float64.  TEST INFRASTRUCTURE ONLY.

Restates the reference's filter design (``soundml/lib/resample.ml:105-179``) and
evaluates a stage the way the reference's *independent evaluator* does
(``soundml/test/resample/resample_kernel.ml:144-163``):

    out[i] = sum_{tt=0}^{2K} h[p + tt L] * x[base - tt],
    p = (i M) mod L,  base = (i M) div L + K,  x = 0 outside [0, n)

Cascades follow ``Kernel.drain_run`` (resample.ml:1819-1842): stage 1 emits
ceil(n L1/M1) samples, stage 2 runs over them with zeros beyond and the result
is cut to ceil(n L/M).  The reference declares the resampler "the one
deliberate exception to librosa bit-parity" (test/README.md:112-121), so parity
is pinned here by (a) the pinned plan strings and latencies
(resample_config.ml:92-146), reproduced by the library's planner, and (b) the
reference's own C executor compiled from its source (oracle/_ref, see
ref_executor.py), which this evaluator matches to ~1e-15 of peak in float64.
"""
import math

import numpy as np


def kaiser_beta(att):                                   # resample.ml:105-109
    if att > 50.0:
        return 0.1102 * (att - 8.7)
    if att > 21.0:
        return 0.5842 * (att - 21.0) ** 0.4 + 0.07886 * (att - 21.0)
    return 0.0


def kaiser_numtaps(att, width):                         # resample.ml:113-116
    n = math.ceil((att - 7.95) / 2.285 / (math.pi * width) + 1.0)
    return n + 1 if n % 2 == 0 else n


def bessel_i0(x):                                       # resample.ml:126-137
    hx2 = 0.25 * x * x
    term, total, k = 1.0, 1.0, 1
    while True:
        term = term * hx2 * (1.0 / float(k * k)) if k < 129 else term * hx2 / float(k * k)
        total += term
        if term <= np.finfo(np.float64).eps * total or k > 1000:
            return total
        k += 1


def design_prototype(l, k, fc, beta):                   # resample.ml:145-163
    mid = k * l
    n = 2 * mid + 1
    i0b = bessel_i0(beta)
    h = np.zeros(n, dtype=np.float64)
    for i in range(mid, n):
        z = float(i - mid)
        s = fc if i == mid else math.sin(math.pi * fc * z) / (math.pi * z)
        r = z / float(mid)
        v = s * (bessel_i0(beta * math.sqrt(1.0 - r * r)) / i0b)
        h[i] = v
        h[n - 1 - i] = v
    total = 0.0
    for v in h:                                         # Array.fold_left ( +. )
        total += v
    return h * (float(l) / total)


def bank_of_prototype(l, k, h):                         # resample.ml:169-179
    taps = 2 * k + 1
    b = np.zeros((l, taps), dtype=np.float64)
    for p in range(l):
        for s in range(taps):
            idx = p + (taps - 1 - s) * l
            if idx < h.size:
                b[p, s] = h[idx]
    return b


def single_stage_design(sample_rate, target, attenuation=126.0, passband=0.913):
    """The single-stage geometry of Config.create (resample.ml:900-940)."""
    g = math.gcd(sample_rate, target)
    l, m = target // g, sample_rate // g
    width = (1.0 - passband) / float(max(l, m))
    ntaps = kaiser_numtaps(attenuation, width)
    k = max(1, int(math.ceil((ntaps - 1.0) / (2.0 * l))))
    fc = (1.0 + passband) / (2.0 * float(max(l, m)))
    return dict(l=l, m=m, k=k, fc=fc, beta=kaiser_beta(attenuation))


def stage_apply(x, h, l, m, k, n_out):
    """One stage over ``x [..., n]`` in float64 (evaluator form)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[-1]
    i = np.arange(n_out, dtype=np.int64)
    t = i * m
    p = t % l
    base = t // l + k
    out = np.zeros(x.shape[:-1] + (n_out,), dtype=np.float64)
    for tt in range(2 * k + 1):
        idx = base - tt
        hidx = p + tt * l
        ok = (idx >= 0) & (idx < n) & (hidx < h.size)
        if not ok.any():
            continue
        coeff = np.where(ok, h[np.minimum(hidx, h.size - 1)], 0.0)
        out += coeff * x[..., np.clip(idx, 0, max(n - 1, 0))]
    return out


def ceil_div(a, b):
    return 0 if a <= 0 else (a - 1) // b + 1


def apply_plan(x, stages, l_total, m_total):
    """Offline ``Resample.apply`` for a one- or two-stage plan; ``stages`` is a
    list of dicts with keys l, m, k, proto."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[-1]
    total = ceil_div(n * l_total, m_total)
    if len(stages) == 1:
        s = stages[0]
        return stage_apply(x, s["proto"], s["l"], s["m"], s["k"], total)
    s1, s2 = stages
    n1 = ceil_div(n * s1["l"], s1["m"])
    mid = stage_apply(x, s1["proto"], s1["l"], s1["m"], s1["k"], n1)
    return stage_apply(mid, s2["proto"], s2["l"], s2["m"], s2["k"], total)


def fir_apply(x, h):
    """y[i] = sum_t h[t] x[i + K - t], zeros outside — the direct stage at
    L = M = 1."""
    h = np.asarray(h, dtype=np.float64)
    k = (h.size - 1) // 2
    return stage_apply(x, h, 1, 1, k, np.asarray(x).shape[-1])
