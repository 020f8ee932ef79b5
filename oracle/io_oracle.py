"""TEST INFRASTRUCTURE — CPU restatement of soundml-io's layout pass and block sizing.

Follows ``soundml-io/lib/soundml_io_stubs.c:832-872`` (``soundml_io_read_planar_*``:
planar scatter ``out[c * total + i] = fr[c]``; downmix = channels added in order in
the sample type, then one multiply by ``1 / channels``) and
``soundml-io/lib/soundml_io.ml:532-536`` (``decode_block_frames``).  Integer / copy
work and one rounded multiply: the CUDA path must match bit for bit.  libsndfile is
absent here, so there is no decoder in this oracle: blocks are the arrays
``sf_readf_*`` would deliver.  Parity pinned by the reference's own law that the
chunked read equals the whole read (``soundml-io/test/io_law.ml``) — restated in
tests/test_gpu_io.py as chunking invariance — not by golden vectors (the reference
holds none for this pass)."""
import numpy as np


def layout(block, mode="planar"):
    """``[frames, channels]`` interleaved -> ``[channels, frames]`` or ``[1, frames]``."""
    block = np.asarray(block)
    frames, channels = block.shape
    if mode == "planar":
        return np.ascontiguousarray(block.T)
    t = block.dtype.type
    acc = np.zeros(frames, dtype=block.dtype)
    for c in range(channels):                 # in order, in the sample type
        acc = (acc + block[:, c]).astype(block.dtype)
    inv = t(1) / t(channels)
    return (acc * inv).astype(block.dtype)[None, :]


def decode_block_frames(channels, elt, advertised=0):
    budget = 4194304 // (channels * elt)
    block = min(1048576, max(4096, budget))
    return min(block, max(4096, advertised)) if advertised > 0 else block
