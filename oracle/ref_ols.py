"""The reference's overlap-save executor, offline, with ITS OWN shaping code.

TEST INFRASTRUCTURE ONLY.  The spectrum shaping -- the product with the plan spectrum,
imaged for an xL stage, alias-folded for a /M stage -- is the reference's
``soundml_resample_shape`` (``resample_stubs.c:329-408``) called from
``oracle/_ref/libsoundml_ref.so`` (the file compiled unmodified); everything around it
restates ``resample.ml``: the block rule and phase alignment (``ols_geom`` :279-300), the
plan spectrum with its folded 1/M and 1/W (:856-867), the grid with its lead of
2K + delta zeros, the kept span of every block and where its outputs sit in the
circular result (``ols_hi`` :1313-1315, ``ols_run`` :1456-1598).  The transforms are
numpy's complex128 rfft / irfft, as the reference's are Nx's.
"""
import ctypes as C

import numpy as np

from . import ref_executor

CEILING_MS = 130                                            # resample.ml:273


def ols_geom(rate, l, m, k):                                # resample.ml:279-300
    f_div = m if l == 1 else 1
    target = max(64, 10 * k)
    n = 3 if f_div % 3 == 0 else 1
    while n < target:
        n *= 2
    if n * 1000 > CEILING_MS * rate:
        return None
    b = (n - 2 * k) // f_div * f_div
    delta = (f_div - (3 * k) % f_div) % f_div
    return (n, b, delta) if b >= 1 else None


def plan_spectrum(proto, n, l, m):                          # resample.ml:856-867
    length = n * l if l > 1 else n
    w = n * l if l > 1 else n // m
    padded = np.zeros(length)
    padded[:len(proto)] = proto
    spec = np.fft.rfft(padded)
    folds = (w & (w - 1)) == 0                              # ols_folds_inverse :309
    scale = (1.0 / m if m > 1 else 1.0) * (1.0 / w if folds else 1.0)
    return spec * scale, w, folds


def shape(x_spec, h_spec, n, sl, sm, w):
    """[lines, n/2+1] complex128 -> [lines, w/2+1] through the reference's C."""
    lines = x_spec.shape[0]
    xs = np.ascontiguousarray(x_spec, dtype=np.complex128)
    hs = np.ascontiguousarray(h_spec, dtype=np.complex128)
    ys = np.zeros((lines, w // 2 + 1), dtype=np.complex128)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = ref_executor.lib().ref_resample_shape(
        p(xs), C.c_int64(xs.size), p(hs), C.c_int64(hs.size), p(ys), C.c_int64(ys.size),
        C.c_int64(lines), C.c_int64(n), C.c_int64(sl), C.c_int64(sm))
    if rc:
        raise RuntimeError(ref_executor.lib().ref_last_error().decode())
    return ys


def stage_apply(x, proto, l, m, k, n_out, geom):
    """One stage over a whole signal [n] (float64) by overlap-save: n_out outputs."""
    assert l == 1 or m == 1
    n, b, delta = geom
    h_spec, w, folds = plan_spectrum(np.asarray(proto, dtype=np.float64), n, l, m)
    x = np.asarray(x, dtype=np.float64)
    lead = 2 * k + delta

    def hi(blk):                                            # ols_hi
        if l > 1:
            return l * (blk * b + n - 3 * k) - 1
        return (blk * b + n - 3 * k - delta - 1) // m

    blocks = 0
    while blocks == 0 or hi(blocks - 1) < n_out - 1:
        blocks += 1
    grid = np.zeros(lead + max(len(x), (blocks - 1) * b + n))   # the drain feeds zeros
    grid[lead:lead + len(x)] = x
    frames = np.stack([grid[j * b:j * b + n] for j in range(blocks)])
    shaped = shape(np.fft.rfft(frames, axis=-1), h_spec, n, l, m, w)
    r = np.fft.irfft(shaped, n=w, axis=-1)
    if folds:
        r = r * w                                           # norm `Forward: the 1/W is in the plan spectrum
    out = np.zeros(n_out)
    done = 0
    for j in range(blocks):
        cnt = min(hi(j) + 1 - done, n_out - done)
        if cnt <= 0:
            continue
        if l > 1:
            pos = done + l * (3 * k - j * b)
        else:
            pos = (done * m + 3 * k + delta - j * b) // m
        out[done:done + cnt] = r[j, pos:pos + cnt]
        done += cnt
    assert done == n_out
    return out
