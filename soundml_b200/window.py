"""``Soundml.Window`` mirror (reference: soundml/lib/window.ml, window.mli).

Window generation is host-side design work in the reference too; here it is
done by the native library (``smb_window_make``), in double precision.
"""
import ctypes as C

import numpy as np

from . import _lib


def parse(window):
    """A window spec is a name or ``(name, parameter)`` — ``"hann"``,
    ``("kaiser", 8.6)``, ``("gaussian", std)``, ``("tukey", taper)`` — the
    Python spelling of ``Window.t`` (window.ml:22-33)."""
    if isinstance(window, str):
        name, param = window, 0.0
    else:
        name, param = window[0], float(window[1])
    if name not in _lib.WINDOWS:
        raise ValueError(f"make: unknown window {name!r}")
    return _lib.WINDOWS[name], param


def make(window, n, periodic=True, dtype=np.float64):
    """``Window.make dtype ?periodic window n`` (window.ml:374-401): generated
    in float64 and rounded once on the way out."""
    kind, param = parse(window)
    out = np.zeros(max(int(n), 0), dtype=np.float64)
    _lib.check(_lib.lib.smb_window_make(kind, param, int(bool(periodic)), int(n),
                                        out.ctypes.data_as(C.POINTER(C.c_double))))
    return out.astype(dtype, copy=False)
