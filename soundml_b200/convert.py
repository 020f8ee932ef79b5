"""``Soundml.Convert`` mirror — decibel conversions (reference:
soundml/lib/convert.ml:20-56).  The logarithm, scale, offset and the optional
whole-tensor ``top_db`` clamp run on the GPU in the input's own dtype."""
import math

from . import _lib


def _to_db(fn, s, reference, amin, top_db):
    s = _lib.contiguous(s)
    ptr, mem, dtype = _lib.describe(s)
    out = _lib.empty_like_kind(s, tuple(s.shape))
    count = 1
    for d in s.shape:
        count *= int(d)
    _lib.check(fn(ptr, count, dtype, float(reference), float(amin),
                  math.nan if top_db is None else float(top_db), _lib.out_pointer(out), mem,
                  _lib.current_stream(s)))
    return out


def power_to_db(s, reference=1.0, amin=1e-10, top_db=None):
    """``Convert.power_to_db ?reference ?amin ?top_db s`` (convert.ml:46-50)."""
    return _to_db(_lib.lib.smb_power_to_db, s, reference, amin, top_db)


def amplitude_to_db(s, reference=1.0, amin=1e-5, top_db=None):
    """``Convert.amplitude_to_db`` (convert.ml:52-56)."""
    return _to_db(_lib.lib.smb_amplitude_to_db, s, reference, amin, top_db)
