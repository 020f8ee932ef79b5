"""``Soundml.Mel`` mirror (reference: soundml/lib/mel.ml)."""
import ctypes as C
import math

import numpy as np

from . import _lib


class Config:
    """``Mel.Config.t``.  Build with :meth:`create`."""

    def __init__(self, handle, **params):
        self._h = handle
        self.__dict__.update(params)

    @classmethod
    def create(cls, *, n_mels, sample_rate, fft_size, f_min=0.0, f_max=None,
               scale="slaney", norm="slaney"):
        """``Mel.Config.create ?f_min ?f_max ?scale ?norm ~n_mels ~sample_rate
        ~fft_size ()`` (mel.ml:119-164)."""
        if scale not in _lib.MEL_SCALES:
            raise ValueError(f"create: unknown scale {scale!r}")
        if norm not in _lib.MEL_NORMS:
            raise ValueError(f"create: unknown norm {norm!r}")
        h = C.c_void_p()
        _lib.check(_lib.lib.smb_mel_plan_create(
            C.byref(h), int(n_mels), int(sample_rate), int(fft_size), float(f_min),
            math.nan if f_max is None else float(f_max),
            _lib.MEL_SCALES[scale], _lib.MEL_NORMS[norm]))
        return cls(h, sample_rate=sample_rate, f_min=f_min, f_max=f_max, scale=scale, norm=norm)

    @classmethod
    def of_weights(cls, weights, fft_size):
        """A projection through caller-supplied ``[n_mels, bins]`` weights."""
        w = np.ascontiguousarray(weights, dtype=np.float64)
        if w.ndim != 2 or w.shape[1] != fft_size // 2 + 1:
            raise ValueError("of_weights: weights must be [n_mels, fft_size/2 + 1]")
        h = C.c_void_p()
        _lib.check(_lib.lib.smb_mel_plan_create_with_weights(
            C.byref(h), int(w.shape[0]), int(fft_size), w.ctypes.data_as(C.POINTER(C.c_double))))
        return cls(h)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and getattr(_lib, "lib", None) is not None:      # not during interpreter teardown
            _lib.lib.smb_mel_plan_destroy(h)

    n_mels = property(lambda self: int(_lib.lib.smb_mel_n_mels(self._h)))
    bins = property(lambda self: int(_lib.lib.smb_mel_bins(self._h)))
    fft_size = property(lambda self: int(_lib.lib.smb_mel_fft_size(self._h)))


def filterbank(c, dtype=np.float64):
    """``Mel.filterbank dtype c`` (mel.ml:198-200): a copy of the weights."""
    out = np.zeros((c.n_mels, c.bins), dtype=np.float64)
    _lib.check(_lib.lib.smb_mel_filterbank(c._h, out.ctypes.data_as(C.POINTER(C.c_double))))
    return out.astype(dtype, copy=False)


def apply(c, s):
    """``Mel.apply c s`` (mel.ml:202-231): ``[..., bins, frames]`` ->
    ``[..., n_mels, frames]``."""
    if s.ndim < 2:
        raise ValueError(f"apply: cannot project a rank-{s.ndim} tensor (the mel "
                         "projection needs [...; bins; frames])")
    if s.shape[-2] != c.bins:
        raise ValueError(
            f"apply: cannot project {s.shape[-2]} frequency bins through a filterbank "
            f"built for an FFT of size {c.fft_size} ({c.bins} bins)")
    s = _lib.contiguous(s)
    lead = tuple(int(d) for d in s.shape[:-2])
    count = int(s.shape[-1])
    batch = int(np.prod(lead, dtype=np.int64)) if lead else 1
    out = _lib.empty_like_kind(s, lead + (c.n_mels, count))
    if batch == 0 or count == 0:
        return out
    ptr, mem, dtype = _lib.describe(s)
    stream = _lib.current_stream(s)
    if stream is not None:
        _lib.check(_lib.lib.smb_mel_plan_set_stream(c._h, stream))
    _lib.check(_lib.lib.smb_mel_apply(c._h, ptr, batch, count, dtype, _lib.out_pointer(out), mem))
    return out
