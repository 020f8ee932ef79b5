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
loop over caller-supplied decoded blocks.
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
    """``Soundml_io.read ~sample_rate`` over decoded blocks: every block is laid
    out on the device and fed to the streaming resampler on arrival; ``finish``
    drains the resampler's tail (``resolve_eof``).  ``channels`` / ``sample_rate``
    describe the source, ``target`` the delivered rate (``None`` = native),
    ``mode`` ``"planar"`` or ``"mono"`` (the reader's ``~mono`` downmix)."""

    def __init__(self, *, channels, sample_rate, target=None, mode="planar", quality="high",
                 max_block=None, device="cuda", dtype=np.float32):
        if channels < 1:
            raise ValueError("ingest: channels must be at least 1")
        self.channels, self.mode, self.device = int(channels), mode, device
        self.width = self.channels if mode == "planar" else 1
        self.dtype = np.dtype(dtype)
        self.max_block = int(max_block) if max_block else decode_block_frames(
            self.channels, self.dtype.itemsize, 0)
        self.kernel = None
        if target is not None and int(target) != int(sample_rate):
            cfg = _resample.Config.create(sample_rate=int(sample_rate), target=int(target),
                                          quality=quality)
            self.kernel = _resample.Kernel.prepare(cfg, channels=self.width, max_block=self.max_block)
        self.src_pos = 0
        self.finished = False

    def feed(self, block):
        """One decoded block ``[frames, channels]`` -> the frames it releases,
        planar on the device (or ``None``)."""
        if self.finished:
            raise ValueError("ingest: cannot feed a finished reader")
        if block.ndim != 2 or int(block.shape[1]) != self.channels:
            raise ValueError(f"ingest: expected interleaved blocks [frames, {self.channels}]")
        if int(block.shape[0]) > self.max_block:
            raise ValueError(f"ingest: cannot feed a {int(block.shape[0])}-frame block "
                             f"(max_block is {self.max_block})")
        if int(block.shape[0]) == 0:
            return None
        planar = layout(block, self.mode, self.device)
        self.src_pos += int(block.shape[0])
        return planar if self.kernel is None else self.kernel.step(planar)

    def finish(self):
        """Decoder EOF: the resampler's one flush."""
        if self.finished:
            return None
        self.finished = True
        return None if self.kernel is None else self.kernel.flush()

    def read(self, blocks):
        """All of it: ``[width, ceil(frames * L / M)]`` on the device."""
        import torch
        pieces = [p for p in (self.feed(b) for b in blocks) if p is not None]
        tail = self.finish()
        if tail is not None:
            pieces.append(tail)
        if not pieces:
            return torch.zeros((self.width, 0), device=self.device,
                               dtype=torch.float32 if self.dtype == np.float32 else torch.float64)
        return torch.cat(pieces, dim=-1)
