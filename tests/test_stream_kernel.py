"""Stft.Kernel (streaming analysis) host logic: the partition law of
stft.mli:436-470 with the oracle standing in for the GPU transform.  CPU only;
tests/test_gpu_stream.py replays it through the CUDA kernels."""
import numpy as np
import pytest

from golden_util import STFT_SEED, lcg_signal
from oracle import stft_oracle

CASES = [
    dict(fft_size=64, hop=16, alignment="centered", pad="reflect"),
    dict(fft_size=64, hop=16, alignment="left", pad="reflect"),
    dict(fft_size=64, hop=16, alignment="right", pad="edge"),
    dict(fft_size=32, hop=7, win_length=20, alignment="centered", pad=("constant", 0.25)),
    dict(fft_size=16, hop=40, alignment="centered", pad="reflect"),       # hop wider than the frame
    dict(fft_size=31, hop=5, alignment="centered", pad="edge"),
]


def oracle_twin(case):
    kw = dict(case)
    pad = kw.pop("pad")
    pad_value = 0.0
    if not isinstance(pad, str):
        pad, pad_value = pad
    full = stft_oracle.StftConfig(kw.pop("fft_size"), pad=pad, pad_value=pad_value, **kw)
    left = stft_oracle.StftConfig(full.fft_size, hop=full.hop, win_length=full.win_length,
                                  alignment="left", pad="constant")
    return full, left


def chunkings(n, rng):
    yield [n]
    yield [1] * n if n <= 200 else [n // 2, n - n // 2]
    for _ in range(4):
        cuts = np.sort(rng.integers(0, n + 1, size=rng.integers(1, 8)))
        yield list(np.diff(np.concatenate([[0], cuts, [n]])))


@pytest.mark.parametrize("case", CASES)
def test_partition_law_host_logic(lib, case):
    c = lib.Stft.Config.create(**case)
    full, left = oracle_twin(case)
    rng = np.random.default_rng(7)
    for n in (0, 1, 2, 17, 33, 100, 517):
        x = lcg_signal(max(n, 1) * 2, STFT_SEED).reshape(2, -1)[:, :n]
        want = stft_oracle.transform(full, x)
        for sizes in chunkings(n, rng):
            k = lib.Stft.Kernel(c, 2, 4096,
                                _analyse=lambda s, count: stft_oracle.transform(left, s))
            outs, at = [], 0
            for m in sizes:
                o = k.step(x[:, at:at + m])
                at += m
                if o is not None:
                    outs.append(o)
            o = k.flush()
            if o is not None:
                outs.append(o)
            assert k.flush() is None
            got = np.concatenate(outs, axis=-1) if outs else np.zeros((2, full.bins, 0), complex)
            assert got.shape == want.shape, (n, sizes)
            assert np.array_equal(got, want), (n, sizes)
            with pytest.raises(ValueError, match="drained kernel"):
                k.step(x[:, :1])
            k.reset()
            assert k.step(x[:, :0]) is None


def test_prepare_errors(lib):
    c = lib.Stft.Config.create(fft_size=64)
    with pytest.raises(ValueError, match="channels must be at least 1"):
        lib.Stft.Kernel.prepare(c, channels=0, max_block=16)
    with pytest.raises(ValueError, match="max_block must be at least 1"):
        lib.Stft.Kernel.prepare(c, channels=1, max_block=0)
    k = lib.Stft.Kernel.prepare(c, channels=1, max_block=16)
    with pytest.raises(ValueError, match="zero-size leading axis"):
        k.step(np.zeros((0, 5)))
