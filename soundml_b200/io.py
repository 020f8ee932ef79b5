"""Device ingest for soundml-io's fused decode + resample path.

The reference decodes with libsndfile into an interleaved staging block, lays
the block out planar (or downmixes it) and feeds ``Resample.Kernel.step`` block
by block, so the native-rate signal never exists as a whole buffer
(``soundml-io/lib/soundml_io.ml:742-807`` ``decode_step`` / ``resolve_eof``,
``soundml_io_stubs.c:832-872`` the layout pass).  The codec stays where it is,
on the CPU behind libsndfile; what moves to the device is everything after
``sf_readf_*``: the block goes up from (pinned) host memory, the layout pass
runs as a kernel, and the streaming resampler consumes it on arrival.

``layout`` is the layout pass alone; ``Ingest`` is the ``read ~sample_rate``
loop over decoded blocks on the library's reader object (pinned double-buffered staging,
upload and kernels of one block under the decode of the next).
"""
import numpy as np

from . import _lib
from . import resample as _resample

MODES = {"planar": _lib.INGEST_PLANAR, "mono": _lib.INGEST_DOWNMIX}


def decode_block_frames(channels, elt, advertised=0):
    """``decode_block_frames ~channels ~elt ~advertised`` (soundml_io.ml:532-536)."""
    r = int(_lib.lib.smb_ingest_block_frames(int(channels), int(elt), int(advertised)))
    if r < 0:
        raise ValueError("decode_block_frames: channels and elt must be positive")
    return r


def layout(block, mode="planar", device="cuda", out=None, out_off=0):
    """Interleaved ``[frames, channels]`` (numpy = host, torch CUDA = device) ->
    planar ``[channels, frames]`` or, with ``mode="mono"``, the downmix
    ``[1, frames]`` (``soundml_io_read_planar_*``).  The result lives on the
    device unless ``device="host"``.  ``out`` / ``out_off`` place the block inside a
    longer planar destination ``[width, total]``, as the reference's readers do."""
    if block.ndim != 2:
        raise ValueError("layout: expected an interleaved block [frames, channels]")
    block = _lib.contiguous(block)
    ptr, mem_in, dtype = _lib.describe(block)
    frames, channels = int(block.shape[0]), int(block.shape[1])
    width = channels if mode == "planar" else 1
    code = MODES[mode]
    stream = None
    if out is None:
        if device == "host":
            out = np.zeros((width, frames), dtype=np.float32 if dtype == _lib.F32 else np.float64)
        else:
            import torch
            dev = block.device if _lib.is_torch(block) else torch.device(device)
            out = torch.zeros((width, frames), device=dev,
                              dtype=torch.float32 if dtype == _lib.F32 else torch.float64)
    optr, mem_out, odtype = _lib.describe(out)
    if odtype != dtype or out.ndim != 2 or int(out.shape[0]) != width:
        raise ValueError(f"layout: out must be [{width}, total] of the block's dtype")
    for t in (block, out):
        if _lib.is_torch(t):
            stream = _lib.current_stream(t)
    _lib.check(_lib.lib.smb_ingest_layout(ptr, frames, channels, code, dtype, optr, int(out.shape[1]),
                                          int(out_off), mem_in, mem_out, stream))
    return out


class Ingest:
    """``Soundml_io.read ~sample_rate`` over decoded blocks, on the library's native reader
    (``smb_ingest_*``): two pinned staging blocks, the upload of block i and its layout pass
    + resampler step enqueued behind it while the caller decodes block i + 1 straight into
    the other staging block (``staging()``; ``feed`` copies a block there for callers that
    already hold one).  ``finish`` drains the resampler's tail (``resolve_eof``).
    ``channels`` / ``sample_rate`` describe the source, ``target`` the delivered rate
    (``None`` = native), ``mode`` ``"planar"`` or ``"mono"`` (the reader's ``~mono``
    downmix)."""

    def __init__(self, *, channels, sample_rate, target=None, mode="planar", quality="high",
                 max_block=None, device="cuda", dtype=np.float32):
        import ctypes as C
        if channels < 1:
            raise ValueError("ingest: channels must be at least 1")
        if quality not in _lib.QUALITIES:
            raise ValueError(f"ingest: unknown quality {quality!r}")
        self.channels, self.mode, self.device = int(channels), mode, device
        self.width = self.channels if mode == "planar" else 1
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise ValueError("ingest: dtype must be float32 or float64")
        h = C.c_void_p()
        _lib.check(_lib.lib.smb_ingest_create(
            C.byref(h), self.channels, int(sample_rate), int(target) if target else 0, MODES[mode],
            _lib.QUALITIES[quality], int(max_block) if max_block else 0,
            _lib.F32 if self.dtype == np.float32 else _lib.F64))
        self._h = h
        self.max_block = int(_lib.lib.smb_ingest_max_block(self._h))
        self.src_pos = 0
        self.finished = False

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.lib.smb_ingest_destroy(h)

    def staging(self):
        """The pinned block ``[max_block, channels]`` the decoder fills next (a numpy view;
        valid until the submit after next)."""
        import ctypes as C
        p = C.c_void_p()
        _lib.check(_lib.lib.smb_ingest_staging(self._h, C.byref(p)))
        ct = C.c_float if self.dtype == np.float32 else C.c_double
        buf = (ct * (self.max_block * self.channels)).from_address(p.value)
        return np.frombuffer(buf, dtype=self.dtype).reshape(self.max_block, self.channels)

    def _torch_dtype(self):
        import torch
        return torch.float32 if self.dtype == np.float32 else torch.float64

    def submit(self, frames):
        """The first ``frames`` frames of the current staging block are decoded: enqueue
        their upload, layout and resampling; returns the frames they release, planar on the
        device (or ``None``), ordered on ``stream()``."""
        import torch
        if self.finished:
            raise ValueError("ingest: cannot feed a finished reader")
        frames = int(frames)
        if frames < 0 or frames > self.max_block:
            raise ValueError(f"ingest: cannot feed a {frames}-frame block (max_block is {self.max_block})")
        if frames == 0:
            return None
        released = int(_lib.lib.smb_ingest_submit_frames(self._h, frames))
        out = torch.empty((self.width, released), device=self.device, dtype=self._torch_dtype())
        self._ext(out).wait_stream(torch.cuda.current_stream(out.device))   # `out` belongs to that stream
        _lib.check(_lib.lib.smb_ingest_submit(self._h, frames, out.data_ptr()))
        self.src_pos += frames
        self._order(out)
        return out if released else None

    def _ext(self, out):
        import torch
        return torch.cuda.ExternalStream(int(_lib.lib.smb_ingest_stream(self._h)), device=out.device)

    def _order(self, out):
        # the caller's stream continues behind the reader's compute stream
        import torch
        # (no record_stream: `out` is only ever used behind this wait, and the reader's stream
        # may be gone by the time the tensor is freed)
        torch.cuda.current_stream(out.device).wait_stream(self._ext(out))

    def feed(self, block):
        """One decoded block ``[frames, channels]`` (host) -> the frames it releases, planar on
        the device (or ``None``).  The block is copied into the pinned staging block; decode
        into ``staging()`` and call ``submit`` to skip that copy."""
        if self.finished:
            raise ValueError("ingest: cannot feed a finished reader")
        if block.ndim != 2 or int(block.shape[1]) != self.channels:
            raise ValueError(f"ingest: expected interleaved blocks [frames, {self.channels}]")
        frames = int(block.shape[0])
        if frames > self.max_block:
            raise ValueError(f"ingest: cannot feed a {frames}-frame block (max_block is {self.max_block})")
        if frames == 0:
            return None
        if _lib.is_torch(block):
            block = block.detach().cpu().numpy()
        self.staging()[:frames] = np.asarray(block, dtype=self.dtype)
        return self.submit(frames)

    def finish(self):
        """Decoder EOF: the resampler's one flush."""
        import torch
        if self.finished:
            return None
        self.finished = True
        tail = int(_lib.lib.smb_ingest_finish_frames(self._h))
        out = torch.empty((self.width, tail), device=self.device, dtype=self._torch_dtype())
        self._ext(out).wait_stream(torch.cuda.current_stream(out.device))
        _lib.check(_lib.lib.smb_ingest_finish(self._h, out.data_ptr()))
        self._order(out)
        return out if tail else None

    def sync(self):
        _lib.check(_lib.lib.smb_ingest_sync(self._h))

    def read(self, blocks):
        """All of it: ``[width, ceil(frames * L / M)]`` on the device."""
        import torch
        pieces = [p for p in (self.feed(b) for b in blocks) if p is not None]
        tail = self.finish()
        if tail is not None:
            pieces.append(tail)
        if not pieces:
            return torch.zeros((self.width, 0), device=self.device, dtype=self._torch_dtype())
        return torch.cat(pieces, dim=-1)
