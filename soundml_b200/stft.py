"""``Soundml.Stft`` mirror — analysis half (reference: soundml/lib/stft.ml:48-691).

Same names, argument meaning and error behaviour as the OCaml module:
``Config.create``, ``frames``, ``transform``, ``power_spectrum``, ``times``,
``frequencies``.  The arithmetic runs in libsoundml_b200.so on the GPU; numpy
arrays are treated as host buffers (copied in and out by the library), torch
CUDA tensors are used in place on torch's current stream.
"""
import ctypes as C

import numpy as np

from . import _lib
from . import window as _window


class Config:
    """``Stft.Config.t``.  Build with :meth:`create`."""

    def __init__(self, handle, window, alignment, pad, pad_value, scale, win_length=None):
        self._h = handle
        self.window, self.alignment, self.pad = window, alignment, pad
        self.pad_value, self.scale = pad_value, scale
        self._win_length = win_length

    @classmethod
    def create(cls, *, fft_size, window="hann", win_length=None, hop=None,
               alignment="centered", pad="reflect", scale="none"):
        """``Stft.Config.create ?window ?win_length ?hop ?alignment ?pad ?scale
        ~fft_size ()`` (stft.ml:61-111).  ``pad`` is ``"reflect"``, ``"edge"``
        or ``("constant", v)``.  Raises ValueError with the reference's
        Invalid_argument messages."""
        kind, param = _window.parse(window)
        pad_value = 0.0
        if not isinstance(pad, str):
            pad, pad_value = pad[0], float(pad[1])
        for table, key, what in ((_lib.ALIGNMENTS, alignment, "alignment"),
                                 (_lib.PADS, pad, "pad"), (_lib.SCALES, scale, "scale")):
            if key not in table:
                raise ValueError(f"create: unknown {what} {key!r}")
        h = C.c_void_p()
        _lib.check(_lib.lib.smb_stft_plan_create(
            C.byref(h), int(fft_size),
            _lib.DEFAULT if hop is None else int(hop),
            _lib.DEFAULT if win_length is None else int(win_length),
            kind, param, _lib.ALIGNMENTS[alignment], _lib.PADS[pad], pad_value,
            _lib.SCALES[scale]))
        return cls(h, window, alignment, pad, pad_value, scale, win_length)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and getattr(_lib, "lib", None) is not None:      # not during interpreter teardown
            _lib.lib.smb_stft_plan_destroy(h)

    fft_size = property(lambda self: int(_lib.lib.smb_stft_fft_size(self._h)))
    win_length = property(lambda self: self.fft_size if self._win_length is None
                          else int(self._win_length))
    hop = property(lambda self: int(_lib.lib.smb_stft_hop(self._h)))
    bins = property(lambda self: int(_lib.lib.smb_stft_bins(self._h)))

    @property
    def analysis_window(self):
        out = np.zeros(self.fft_size, dtype=np.float64)
        _lib.check(_lib.lib.smb_stft_analysis_window(
            self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def set_path(self, path):
        """Testing hook: force the generic (double interior) kernel, the fused
        fft-2048 CUDA-core kernel (``"fast"``), the fused tcgen05 kernel
        (``"tensor"``) or the frame-pair kernel (``"pair"``: mel output of fft 2048).
        ``"auto"`` picks the frame-pair kernel for mel spectrograms it covers, else
        the fused CUDA-core kernel when it applies, else the generic one."""
        code = {"auto": _lib.PATH_AUTO, "generic": _lib.PATH_GENERIC,
                "fast": _lib.PATH_FAST, "tensor": _lib.PATH_TENSOR,
                "pair": _lib.PATH_PAIR}[path]
        _lib.check(_lib.lib.smb_stft_plan_set_path(self._h, code))
        return self

    def source_indices(self, n):
        """Source sample each padded position reads (-1 = constant fill): the
        boundary-extension contract of ``pad_signal`` (stft.ml:318-338)."""
        total = C.c_int64()
        _lib.check(_lib.lib.smb_stft_source_indices(self._h, int(n), None, C.byref(total)))
        out = np.zeros(total.value, dtype=np.int64)
        _lib.check(_lib.lib.smb_stft_source_indices(
            self._h, int(n), out.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(total)))
        return out


def frames(c, n):
    """``Stft.frames c ~n`` (stft.ml:217-223)."""
    r = _lib.lib.smb_stft_frames(c._h, int(n))
    if r < 0:
        raise ValueError(_lib.last_error())
    return int(r)


def _check_rank(op, x):
    if x.ndim < 1:
        raise ValueError(
            f"{op}: cannot analyse a rank-zero tensor (the time axis must exist)")


def _run(op, c, x, complex_out, power, out=None):
    _check_rank(op, x)
    x = _lib.contiguous(x)
    n = int(x.shape[-1])
    lead = tuple(int(d) for d in x.shape[:-1])
    batch = int(np.prod(lead, dtype=np.int64)) if lead else 1
    count = frames(c, n)
    out = _lib.empty_like_kind(x, lead + (c.bins, count), complex_out, out=out)
    ptr, mem, dtype = _lib.describe(x)
    if batch == 0 or count == 0:
        return out
    stream = _lib.current_stream(x)
    if stream is not None:
        _lib.check(_lib.lib.smb_stft_plan_set_stream(c._h, stream))
    if complex_out:
        _lib.check(_lib.lib.smb_stft_transform(c._h, ptr, batch, n, dtype,
                                               _lib.out_pointer(out), mem))
    else:
        _lib.check(_lib.lib.smb_stft_power_spectrum(c._h, ptr, batch, n, dtype, float(power),
                                                    _lib.out_pointer(out), mem))
    return out


def transform(c, x, out=None):
    """``Stft.transform cdtype c x`` (stft.ml:632-650): ``[..., n]`` ->
    complex ``[..., bins, frames]`` (complex64 for float32 audio)."""
    return _run("transform", c, x, True, 1.0, out)


def transform_range(c, x, p0, p1):
    """``Stft.transform_range cdtype c ~p0 ~p1 x`` (stft.ml:652-666): frames
    ``[p0, p1)`` of ``transform c x`` without evaluating the others."""
    _check_rank("transform_range", x)
    x = _lib.contiguous(x)
    n = int(x.shape[-1])
    lead = tuple(int(d) for d in x.shape[:-1])
    batch = int(np.prod(lead, dtype=np.int64)) if lead else 1
    total = frames(c, n)
    if p0 < 0 or p0 > p1 or p1 > total:
        raise ValueError(
            f"transform_range: cannot take frames [{p0}, {p1}) of a {total}-frame transform "
            "(the range must satisfy 0 <= p0 <= p1 <= frames)")
    out = _lib.empty_like_kind(x, lead + (c.bins, p1 - p0), True)
    if batch == 0 or p1 == p0:
        return out
    ptr, mem, dtype = _lib.describe(x)
    stream = _lib.current_stream(x)
    if stream is not None:
        _lib.check(_lib.lib.smb_stft_plan_set_stream(c._h, stream))
    _lib.check(_lib.lib.smb_stft_transform_range(c._h, ptr, batch, n, dtype, int(p0), int(p1),
                                                 _lib.out_pointer(out), mem))
    return out


def power_spectrum(c, x, power=2.0, out=None):
    """``Stft.power_spectrum ?power c x`` (stft.ml:687-691)."""
    return _run("power_spectrum", c, x, False, power, out)


def nola(c):
    """``Stft.nola c`` (stft.ml:731-742): the overlap-added squared window clears
    1e-10 of its maximum at every position, so synthesis is defined."""
    return bool(_lib.lib.smb_stft_nola(c._h))


def output_length(c, frames):
    """``Stft.output_length c ~frames`` (stft.ml:790-794)."""
    r = _lib.lib.smb_stft_output_length(c._h, int(frames))
    if r < 0:
        raise ValueError(_lib.last_error())
    return int(r)


def invert(c, z, length=None, dtype=None):
    """``Stft.invert dtype c ?length z`` (stft.ml:693-939): least-squares
    synthesis of ``[..., bins, frames]`` complex frames into ``[..., length]``
    real samples (``output_length`` when no length is named).  The interior is
    double; ``dtype`` (default: the component width of ``z``) is the one
    rounding at the end."""
    if z.ndim < 2:
        raise ValueError(
            f"invert: cannot invert a rank-{z.ndim} tensor (the bin and frame axes must exist)")
    bins, count = int(z.shape[-2]), int(z.shape[-1])
    if bins != c.bins:
        raise ValueError(
            f"invert: cannot invert {bins} frequency bins of a {c.fft_size}-point transform "
            f"(the bin axis must hold fft_size / 2 + 1 = {c.bins} values)")
    z = _lib.contiguous(z)
    lead = tuple(int(d) for d in z.shape[:-2])
    batch = int(np.prod(lead, dtype=np.int64)) if lead else 1
    if _lib.is_torch(z):
        import torch
        if not z.is_cuda:
            raise ValueError("torch tensors must live on a CUDA device; pass numpy for host data")
        codes = {torch.complex64: _lib.F32, torch.complex128: _lib.F64}
        outs = {_lib.F32: torch.float32, _lib.F64: torch.float64}
        names = {torch.float32: _lib.F32, torch.float64: _lib.F64, None: None}
    else:
        codes = {np.dtype(np.complex64): _lib.F32, np.dtype(np.complex128): _lib.F64}
        outs = {_lib.F32: np.float32, _lib.F64: np.float64}
        names = {None: None}
        if dtype is not None:
            names[dtype] = {np.dtype(np.float32): _lib.F32,
                            np.dtype(np.float64): _lib.F64}.get(np.dtype(dtype))
    in_code = codes.get(z.dtype)
    if in_code is None:
        raise ValueError(f"unsupported dtype {z.dtype} (complex64 and complex128 are carried)")
    out_code = in_code if dtype is None else names.get(dtype)
    if out_code is None:
        raise ValueError(f"unsupported dtype {dtype} (float32 and float64 are carried)")
    if length is not None and length < 0:
        raise ValueError(
            f"invert: cannot synthesise a signal of length {length} (length must be non-negative)")
    if not nola(c):
        _lib.check(_lib.lib.smb_stft_invert(c._h, None, 0, 0, in_code, 0, 0, out_code, None,
                                            _lib.MEM_HOST))
    out_len = output_length(c, count) if length is None else int(length)
    if _lib.is_torch(z):
        import torch
        out = torch.zeros(lead + (out_len,), dtype=outs[out_code], device=z.device)
        mem, zp, op = _lib.MEM_DEVICE, z.data_ptr(), out.data_ptr()
    else:
        out = np.zeros(lead + (out_len,), dtype=outs[out_code])
        mem, zp, op = _lib.MEM_HOST, z.ctypes.data, out.ctypes.data
    if batch == 0 or out_len == 0 or count == 0:
        return out
    stream = _lib.current_stream(z)
    if stream is not None:
        _lib.check(_lib.lib.smb_stft_plan_set_stream(c._h, stream))
    _lib.check(_lib.lib.smb_stft_invert(c._h, zp, batch, count, in_code,
                                        0 if length is None else 1,
                                        0 if length is None else int(length), out_code, op, mem))
    return out


def griffin_lim(c, s, n_iter=32, momentum=0.99, init=None, length=None):
    """``Stft.griffin_lim ?n_iter ?momentum ?init ?length c s`` (stft.ml:964-1025):
    fast Griffin-Lim phase reconstruction of a magnitude spectrogram ``[..., bins,
    frames]`` (float32 or float64) into ``[..., length]`` samples of the same
    dtype.  ``init`` is ``None`` (zero phase) or a tensor of starting phases with
    the shape of ``s``."""
    if s.ndim < 2:
        raise ValueError(
            f"griffin_lim: cannot invert a rank-{s.ndim} tensor (the bin and frame axes must exist)")
    bins, count = int(s.shape[-2]), int(s.shape[-1])
    if bins != c.bins:
        raise ValueError(
            f"griffin_lim: cannot invert {bins} frequency bins of a {c.fft_size}-point transform "
            f"(the bin axis must hold fft_size / 2 + 1 = {c.bins} values)")
    s = _lib.contiguous(s)
    ptr, mem, dtype = _lib.describe(s)
    pptr = None
    if init is not None:
        if tuple(init.shape) != tuple(s.shape):
            shape = lambda t: "; ".join(str(int(d)) for d in t.shape)
            raise ValueError(
                f"griffin_lim: cannot start from a [{shape(init)}] phase for a [{shape(s)}] "
                "spectrogram (the initial phase must have the shape of the magnitudes)")
        init = _lib.contiguous(init)
        pptr, pmem, pdtype = _lib.describe(init)
        if pmem != mem or pdtype != dtype:
            raise ValueError("griffin_lim: the initial phase must live where the magnitudes "
                             "live and share their dtype")
    lead = tuple(int(d) for d in s.shape[:-2])
    batch = int(np.prod(lead, dtype=np.int64)) if lead else 1
    # precondition order of the reference: length, invertibility, n_iter, momentum
    probe = _lib.lib.smb_stft_griffin_lim(c._h, None, 0, 0, dtype, int(n_iter), float(momentum),
                                          None, 0 if length is None else 1,
                                          0 if length is None else int(length), None, mem)
    _lib.check(probe)
    out_len = output_length(c, count) if length is None else int(length)
    out = _lib.empty_like_kind(s, lead + (out_len,))
    if batch == 0 or out_len == 0:
        return out
    stream = _lib.current_stream(s)
    if stream is not None:
        _lib.check(_lib.lib.smb_stft_plan_set_stream(c._h, stream))
    _lib.check(_lib.lib.smb_stft_griffin_lim(
        c._h, ptr, batch, count, dtype, int(n_iter), float(momentum), pptr,
        0 if length is None else 1, 0 if length is None else int(length),
        _lib.out_pointer(out), mem))
    return out


# ---- streaming analysis ------------------------------------------------------

def _last(t):
    return int(t.shape[-1])


def _cat(parts):
    if len(parts) == 1:
        return parts[0]
    if _lib.is_torch(parts[0]):
        import torch
        return torch.cat(parts, dim=-1)
    return np.concatenate(parts, axis=-1)


def _copy(t):
    return t.clone() if _lib.is_torch(t) else np.array(t, copy=True)


def _take(t, idx):
    """``take_last`` (stft.ml:307-312): gather along the time axis."""
    if _lib.is_torch(t):
        import torch
        return t.index_select(-1, torch.as_tensor(idx, dtype=torch.int64, device=t.device))
    return np.take(t, np.asarray(idx, dtype=np.int64), axis=-1)


def _full(t, count, value):
    shape = tuple(t.shape[:-1]) + (count,)
    if _lib.is_torch(t):
        import torch
        return torch.full(shape, value, dtype=t.dtype, device=t.device)
    return np.full(shape, value, dtype=t.dtype)


class Kernel:
    """``Stft.Kernel`` (stft.ml:375-622): the chunked form of ``transform``.

    ``step`` feeds a chunk ``[..., m]`` and returns the frames it completed
    (``[..., bins, frames]``) or ``None``; ``flush`` installs the right boundary
    extension and returns the remaining frames.  Concatenating everything
    returned, for any partition of the signal into chunks, equals
    ``transform c x`` -- bit for bit here too, because a frame's arithmetic on
    the GPU does not depend on which call carries it.

    The state is the reference's: the padded stream not yet consumed
    (``pending``), the raw prelude kept until the left extension is computable
    (``left + 1`` samples under reflection, one otherwise), the last
    ``right + 1`` raw samples for the right extension, and ``skip`` for hops
    wider than the frame.  Frames are evaluated by a left-aligned twin of the
    configuration (same analysis window) over the pending stream.
    """

    def __init__(self, c, channels, max_block, _analyse=None):
        if channels < 1:
            raise ValueError(
                f"prepare: cannot analyse {channels} channels (channels must be at least 1)")
        if max_block < 1:
            raise ValueError(
                f"prepare: cannot accept blocks of {max_block} samples "
                "(max_block must be at least 1)")
        self.cfg = c
        self.fft, self.hop = c.fft_size, c.hop
        self.left = {"centered": self.fft // 2, "left": 0, "right": self.fft - 1}[c.alignment]
        self.right = self.fft // 2 if c.alignment == "centered" else 0
        self._analyse = _analyse
        self._twin = None
        self.reset()

    @classmethod
    def prepare(cls, c, *, channels, max_block):
        """``Kernel.prepare cdtype c dtype ~channels ~max_block``; the dtypes are
        those of the chunks fed."""
        return cls(c, channels, max_block)

    def reset(self):
        """``Kernel.reset`` (stft.ml:403-411)."""
        self.started = self.drained = False
        self.received = 0
        self.prelude, self.pending = [], []
        self.pending_len = 0
        self.tail = None
        self.skip = 0

    # frames of a left-aligned stream: the twin plan shares the analysis window
    def _frames_of(self, samples, count):
        span = (count - 1) * self.hop + self.fft
        if _last(samples) != span:
            samples = samples[..., :span]
        if self._analyse is not None:
            return self._analyse(samples, count)
        if self._twin is None:
            h = C.c_void_p()
            w = self.cfg.analysis_window
            _lib.check(_lib.lib.smb_stft_plan_create_with_window(
                C.byref(h), self.fft, self.hop, _lib.ALIGNMENTS["left"], _lib.PADS["constant"],
                0.0, w.ctypes.data_as(C.POINTER(C.c_double))))
            self._twin = Config(h, self.cfg.window, "left", "constant", 0.0, self.cfg.scale)
        return transform(self._twin, samples)

    def _process(self, extra, extra_len):                     # stft.ml:417-446
        total = self.pending_len + extra_len
        count = 0 if total < self.fft else 1 + (total - self.fft) // self.hop
        if count == 0:
            self.pending += [_copy(t) for t in extra if _last(t) > 0]
            self.pending_len = total
            return None
        samples = _cat(self.pending + list(extra))
        out = self._frames_of(samples, count)
        next_start = count * self.hop
        if next_start >= total:
            self.skip += next_start - total
            self.pending, self.pending_len = [], 0
        else:
            self.pending = [_copy(samples[..., next_start:total])]
            self.pending_len = total - next_start
        return out

    def _install_threshold(self):                              # stft.ml:452-458
        return self.left + 1 if self.cfg.pad == "reflect" else 1

    def _left_pad(self, x):                                    # stft.ml:463-480
        if self.left == 0:
            return None
        if self.cfg.pad == "constant":
            return _full(x, self.left, self.cfg.pad_value)
        if self.cfg.pad == "reflect":
            return _take(x, [self.left - j for j in range(self.left)])
        return _take(x, [0] * self.left)

    def _install(self, x):                                     # stft.ml:485-497
        n = _last(x)
        if self.right > 0:
            keep = min(self.right + 1, n)
            self.tail = _copy(x[..., n - keep:n])
        self.started = True
        self.prelude = []
        lp = self._left_pad(x)
        if lp is None:
            return self._process([x], n)
        return self._process([lp, x], self.left + n)

    def _update_tail(self, chunk):                             # stft.ml:501-513
        keep = self.right + 1
        m = _last(chunk)
        if m >= keep:
            self.tail = _copy(chunk[..., m - keep:m])
        else:
            combined = chunk if self.tail is None else _cat([self.tail, chunk])
            cm = _last(combined)
            self.tail = _copy(combined[..., max(0, cm - keep):cm])

    def _right_pad(self):                                      # stft.ml:517-530
        tl = _last(self.tail)
        if self.cfg.pad == "constant":
            return _full(self.tail, self.right, self.cfg.pad_value)
        if self.cfg.pad == "reflect":
            return _take(self.tail, [tl - 2 - i for i in range(self.right)])
        return _take(self.tail, [tl - 1] * self.right)

    def step(self, chunk):
        """``Kernel.step k chunk`` (stft.ml:532-569)."""
        if self.drained:
            raise ValueError("step: cannot feed a drained kernel (flush consumed the tail; "
                             "reset before reusing)")
        if any(int(d) == 0 for d in chunk.shape[:-1]):
            raise ValueError("step: cannot analyse a chunk with a zero-size leading axis "
                             "(channels must be at least 1)")
        m = _last(chunk)
        if m == 0:
            return None
        self.received += m
        if not self.started:
            if self.received >= self._install_threshold():
                return self._install(_cat(self.prelude + [chunk]))
            self.prelude.append(_copy(chunk))
            return None
        if self.right > 0:
            self._update_tail(chunk)
        if self.skip >= m:
            self.skip -= m
            return None
        dropped, self.skip = self.skip, 0
        if dropped:
            chunk = chunk[..., dropped:m]
        return self._process([chunk], m - dropped)

    def flush(self):
        """``Kernel.flush k`` (stft.ml:571-606)."""
        if self.drained:
            return None
        self.drained = True
        out = None
        if not self.started:
            if self.received > 0:
                x = _cat(self.prelude)
                idx = self.cfg.source_indices(_last(x))        # pad_signal, stft.ml:318-338
                padded = _take(x, np.maximum(idx, 0))
                if (idx < 0).any():                            # constant extension
                    fill = np.nonzero(idx < 0)[0]
                    if _lib.is_torch(padded):
                        import torch
                        fill = torch.as_tensor(fill, dtype=torch.int64, device=padded.device)
                    padded[..., fill] = self.cfg.pad_value
                self.started, self.prelude = True, []
                out = self._process([padded], _last(padded))
        elif self.right > 0:
            rp = self._right_pad()
            r = _last(rp)
            if self.skip >= r:
                self.skip -= r
            else:
                dropped, self.skip = self.skip, 0
                if dropped:
                    rp = rp[..., dropped:r]
                out = self._process([rp], r - dropped)
        self.pending, self.pending_len = [], 0
        return out


def synthesis_latency(c):
    """``Config.synthesis_latency`` (stft.ml:152-153): output samples the streaming
    synthesis trails by -- the head trim plus the tail holdback."""
    fft, hop = c.fft_size, c.hop
    left = {"centered": fft // 2, "left": 0, "right": fft - 1}[c.alignment]
    right = fft // 2 if c.alignment == "centered" else 0
    return left + max(0, hop + right - fft)


class Synthesis:
    """``Stft.Synthesis`` (stft.ml:1027-1296): the chunked form of ``invert``.

    ``step`` feeds frames ``[..., bins, m]`` and returns the samples they
    completed (``hop`` per frame once the head trim is paid) or ``None``;
    ``flush`` releases the positions the last frame reaches.  Concatenating
    everything returned equals ``invert`` at its default length on the
    concatenated frames, for every partition of the frame sequence -- bit for bit,
    because every position is computed from the same frames in the same order
    with the same envelope value.

    Padded position q is final once frame q / hop has arrived.  The state is the
    spectra of the last frames that still reach an unreleased position (the
    reference keeps their windowed inverses instead); each step inverts a window
    of them with a left-aligned twin of the configuration, started early enough
    that the released positions lie in the fully covered region of the window,
    where the envelope is the same periodic tile as in the whole signal.
    """

    def __init__(self, c, channels, max_block, dtype=None, _invert=None):
        if channels < 1:
            raise ValueError(
                f"prepare: cannot synthesise {channels} channels (channels must be at least 1)")
        if max_block < 1:
            raise ValueError(
                f"prepare: cannot accept blocks of {max_block} frames "
                "(max_block must be at least 1)")
        if not nola(c):
            raise ValueError(
                f"prepare: cannot invert a {c.win_length}-point window advanced by {c.hop} "
                f"samples inside a {c.fft_size}-point frame (the overlap-added squared window "
                "must stay above 1e-10 of its largest value at every position)")
        self.cfg, self.dtype = c, dtype
        self.fft, self.hop = c.fft_size, c.hop
        self.left = {"centered": self.fft // 2, "left": 0, "right": self.fft - 1}[c.alignment]
        self.right = self.fft // 2 if c.alignment == "centered" else 0
        self.hold = max(0, self.hop + self.right - self.fft)
        self._invert = _invert
        self._twin = None
        self.reset()

    @classmethod
    def prepare(cls, c, *, channels, max_block, dtype=None):
        """``Synthesis.prepare dtype c cdtype ~channels ~max_block``; ``dtype`` is
        the output dtype (default: the component width of the frames fed)."""
        return cls(c, channels, max_block, dtype)

    def reset(self):
        self.frames = 0            # frames fed
        self.released = 0          # padded positions released so far
        self.first = 0             # index of the first retained frame
        self.hist = []             # spectra of frames [first, frames)
        self.drained = False

    def _window(self, lo, hi, final):
        """Padded positions [lo, hi) from the retained frames."""
        n, h = self.fft, self.hop
        p0 = max(self.first, max(0, (lo - (n - h)) // h)) if lo >= n - h else 0
        z = _cat(self.hist)
        self.hist = [z]
        zw = z[..., p0 - self.first:]
        base = p0 * h
        length = None if final else hi - base
        if self._invert is not None:
            y = self._invert(zw, length)
        else:
            if self._twin is None:
                hdl = C.c_void_p()
                w = self.cfg.analysis_window
                _lib.check(_lib.lib.smb_stft_plan_create_with_window(
                    C.byref(hdl), n, h, _lib.ALIGNMENTS["left"], _lib.PADS["constant"], 0.0,
                    w.ctypes.data_as(C.POINTER(C.c_double))))
                self._twin = Config(hdl, self.cfg.window, "left", "constant", 0.0, self.cfg.scale)
            y = invert(self._twin, zw, length=length, dtype=self.dtype)
        out = y[..., lo - base:hi - base]
        return out.contiguous() if _lib.is_torch(out) else np.ascontiguousarray(out)

    def step(self, z):
        """``Synthesis.step k z`` (stft.ml:1168-1230)."""
        if self.drained:
            raise ValueError("step: cannot feed a drained kernel (flush consumed the tail; "
                             "reset before reusing)")
        if z.ndim < 2:
            raise ValueError(
                f"step: cannot invert a rank-{z.ndim} tensor (the bin and frame axes must exist)")
        if int(z.shape[-2]) != self.cfg.bins:
            raise ValueError(
                f"step: cannot invert {int(z.shape[-2])} frequency bins of a {self.fft}-point "
                f"transform (the bin axis must hold fft_size / 2 + 1 = {self.cfg.bins} values)")
        if any(int(d) == 0 for d in z.shape[:-2]):
            raise ValueError("step: cannot synthesise frames with a zero-size leading axis "
                             "(channels must be at least 1)")
        m = _last(z)
        if m == 0:
            return None
        self.hist.append(_copy(z))
        self.frames += m
        # positions released after F frames = F H - max 0 (H + R - N)   (stft.mli:489-494)
        upto = max(0, self.frames * self.hop - self.hold)
        lo = max(self.released, self.left)
        out = self._window(lo, upto, False) if upto > lo else None
        self.released = max(self.released, upto)
        # frames that still reach an unreleased position (or anchor the next window)
        keep = max(0, (max(self.released, self.left) - (self.fft - self.hop)) // self.hop)
        keep = min(keep, self.frames)
        if keep > self.first:
            zall = _cat(self.hist)
            self.hist = [_copy(zall[..., keep - self.first:])]
            self.first = keep
        return out

    def flush(self):
        """``Synthesis.flush k`` (stft.ml:1232-1270): the positions the last frame
        reaches, less the trailing trim."""
        if self.drained:
            return None
        self.drained = True
        out = None
        if self.frames > 0:
            span = (self.frames - 1) * self.hop + self.fft
            lo, hi = max(self.released, self.left), span - self.right
            if hi > lo:
                out = self._window(lo, hi, True)
        self.hist = []
        return out


def times(c, sample_rate, n, dtype=np.float64):
    """``Stft.times`` (stft.ml:245-254)."""
    if sample_rate < 1:
        raise ValueError(f"times: cannot use a sample rate of {sample_rate} Hz "
                         "(sample_rate must be at least 1)")
    count = frames(c, n)
    return (np.arange(count, dtype=np.float64) * float(c.hop) / float(sample_rate)).astype(dtype)


def frequencies(c, sample_rate, dtype=np.float64):
    """``Stft.frequencies`` (stft.ml:256-261)."""
    if sample_rate < 1:
        raise ValueError(f"frequencies: cannot use a sample rate of {sample_rate} Hz "
                         "(sample_rate must be at least 1)")
    return (np.arange(c.bins, dtype=np.float64) *
            (float(sample_rate) / float(c.fft_size))).astype(dtype)
