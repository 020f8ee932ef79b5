"""The reference's own polyphase executor, run from oracle/_ref.

TEST INFRASTRUCTURE ONLY.  ``oracle/_ref/libsoundml_ref.so`` is the reference's
``soundml/lib/resample_stubs.c`` compiled unmodified (oracle/Makefile).  This
module replays ``Resample.apply``'s orchestration of the direct executor
(``resample.ml:1435-1441, 1786-1842, 1913-1936``): one step over the whole
signal, then the drain over K virtual zeros, with the exact integer
bookkeeping (fed / emitted / row0 / s0) of ``direct_run``.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "_ref", "libsoundml_ref.so")
L1_EDGE_BYTES = 128 * 1024                      # resample.ml:235


def available():
    return os.path.exists(PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(PATH)
        _lib.ref_last_error.restype = C.c_char_p
    return _lib


def _ceil_pos(a, b):
    return 0 if a <= 0 else (a - 1) // b + 1


class DirectStage:
    """State of one direct-executor stage (Kernel.prepare, resample.ml:1386-1400)."""

    def __init__(self, bank64, l, m, k, dtype, channels, max_in):
        self.l, self.m, self.k, self.channels = l, m, k, channels
        self.dtype = np.dtype(dtype)
        taps = 2 * k + 1
        self.visit = l * taps * self.dtype.itemsize > L1_EDGE_BYTES
        bank = np.asarray(bank64, dtype=np.float64).reshape(l, taps)
        if self.visit:                           # visit_bank, resample.ml:197-205
            bank = bank[[(j * m) % l for j in range(l)]]
        self.bank = np.ascontiguousarray(bank.astype(self.dtype)).ravel()
        self.hist = np.zeros(channels * 2 * k, dtype=self.dtype)
        self.scratch = np.zeros(2 * k + max(max_in, k), dtype=self.dtype)
        self.fed = 0
        self.emitted = 0

    def ready(self, fed):                        # resample.ml:1300
        return _ceil_pos((fed - self.k) * self.l, self.m)

    def _call(self, x, y, n, n_out, y_off, y_stride, is_flush):
        t = self.emitted * self.m                # direct_run, resample.ml:1435-1441
        row0 = self.emitted % self.l
        s0 = t // self.l + self.k - self.fed
        kind = 0 if self.dtype == np.float32 else 1
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = lib().ref_resample_step(
            kind, p(self.bank), C.c_int64(self.bank.size), p(self.hist), C.c_int64(self.hist.size),
            p(self.scratch), C.c_int64(self.scratch.size), p(x), C.c_int64(x.size), p(y),
            C.c_int64(y.size), C.c_int64(n), C.c_int64(n_out), C.c_int64(self.channels),
            C.c_int64(self.k), C.c_int64(self.l), C.c_int64(self.m), C.c_int64(row0),
            C.c_int64(s0), C.c_int64(y_off), C.c_int64(y_stride), int(self.visit), int(is_flush))
        if rc:
            raise RuntimeError(lib().ref_last_error().decode())
        self.emitted += n_out

    def run(self, x, y, n, y_off, y_stride):
        n_out = self.ready(self.fed + n) - self.emitted
        self._call(x, y if n_out else self.scratch, n, n_out, y_off if n_out else 0,
                   y_stride if n_out else 0, False)
        self.fed += n
        return n_out

    def drain(self, y, n_out, y_off, y_stride):
        self._call(self.hist, y, self.k, n_out, y_off, y_stride, True)


def apply_single(x, bank64, l, m, k):
    """``Resample.apply`` for a single direct stage; x [channels, n]."""
    x = np.ascontiguousarray(x)
    channels, n = x.shape
    total = _ceil_pos(n * l, m)
    out = np.zeros((channels, total), dtype=x.dtype)
    if n == 0:
        return out
    st = DirectStage(bank64, l, m, k, x.dtype, channels, n)
    flat = out.reshape(-1)
    stepped = st.run(x.reshape(-1), flat, n, 0, total)
    rest = _ceil_pos(st.fed * l, m) - st.emitted
    assert rest == total - stepped
    if rest > 0:
        st.drain(flat, rest, stepped, total)
    return out


def apply_cascade(x, stage1, stage2, l_total, m_total):
    """Two direct stages (run + drain_run, resample.ml:1786-1842).  Each stage
    is (bank64, l, m, k)."""
    x = np.ascontiguousarray(x)
    channels, n = x.shape
    b1, l1, m1, k1 = stage1
    b2, l2, m2, k2 = stage2
    n1 = _ceil_pos(n * l1, m1)
    mid = apply_single(x, b1, l1, m1, k1)          # stage 1 emits its exact ceil tail
    total = _ceil_pos(n * l_total, m_total)
    full = apply_single(mid, b2, l2, m2, k2)       # ceil(n1 l2/m2) >= total
    assert mid.shape[-1] == n1 and full.shape[-1] >= total
    return np.ascontiguousarray(full[:, :total])
