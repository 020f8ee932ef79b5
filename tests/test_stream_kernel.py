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


# ---- Resample.Kernel ----------------------------------------------------------

from oracle import resample_oracle as R                      # noqa: E402
from test_resample_oracle import RATE_PAIRS, oracle_stages   # noqa: E402


@pytest.mark.parametrize("sr,target", RATE_PAIRS + [(22050, 22050)])
def test_resample_kernel_partition_law_host_logic(lib, sr, target):
    """step/flush over the oracle's apply: concatenated chunks == offline apply,
    ceil(n L / M) samples, for every partition (resample.mli:296-317)."""
    cfg = lib.Resample.Config.create(sample_rate=sr, target=target)
    st = oracle_stages(cfg) if sr != target else []
    offline = (lambda x: R.apply_plan(x, st, cfg.l, cfg.m)) if st else (lambda x: x.copy())
    rng = np.random.default_rng(sr + target)
    for n in (0, 1, 5, 700, 4001):
        x = rng.uniform(-1, 1, (2, n))
        want = offline(x)
        assert want.shape[-1] == cfg.output_frames(n)
        for sizes in chunkings(n, rng):
            k = lib.Resample.Kernel(cfg, 2, 8192, _apply=offline)
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
            got = np.concatenate(outs, axis=-1) if outs else np.zeros((2, 0))
            assert got.shape == want.shape, (n, sizes)
            np.testing.assert_allclose(got, want, rtol=0, atol=1e-13)
            with pytest.raises(ValueError, match="drained kernel"):
                k.step(x[:, :1])


def test_resample_kernel_errors(lib):
    cfg = lib.Resample.Config.create(sample_rate=44100, target=22050)
    with pytest.raises(ValueError, match="channels must be at least 1"):
        lib.Resample.Kernel.prepare(cfg, channels=0, max_block=16)
    with pytest.raises(ValueError, match="max_block must be at least 1"):
        lib.Resample.Kernel.prepare(cfg, channels=1, max_block=0)
    k = lib.Resample.Kernel.prepare(cfg, channels=2, max_block=16)
    with pytest.raises(ValueError, match=r"17-sample chunk \(max_block is 16\)"):
        k.step(np.zeros((2, 17), np.float32))
    with pytest.raises(ValueError, match="3-channel chunks"):
        k.step(np.zeros((3, 4), np.float32))


# ---- Stft.Synthesis -------------------------------------------------------------

from oracle import istft_oracle                                 # noqa: E402

SYNTH_CASES = [
    dict(fft_size=64, hop=16, alignment="centered"),
    dict(fft_size=64, hop=16, alignment="left"),
    dict(fft_size=64, hop=16, alignment="right"),
    dict(fft_size=32, hop=7, win_length=20, alignment="centered"),
    dict(fft_size=64, hop=64, window="rectangular", alignment="centered"),   # hold = hop + R - N > 0
    dict(fft_size=31, hop=5, alignment="centered"),
]


@pytest.mark.parametrize("case", SYNTH_CASES)
def test_synthesis_partition_law_host_logic(lib, case):
    """Chunked synthesis == invert at its default length, bit for bit, for every
    partition of the frame sequence (stft.mli:473-517), with the oracle standing in
    for the GPU inverse."""
    c = lib.Stft.Config.create(**case)
    kw = dict(case)
    fft = kw.pop("fft_size")
    full = stft_oracle.StftConfig(fft, **kw)
    kw["alignment"] = "left"
    left = stft_oracle.StftConfig(fft, **kw)
    rng = np.random.default_rng(3)
    assert lib.Stft.synthesis_latency(c) == full.left_width() + max(
        0, full.hop + full.right_width() - fft)
    for frames in (0, 1, 2, 3, 9, 40):
        z = rng.standard_normal((2, full.bins, frames)) + 1j * rng.standard_normal((2, full.bins, frames))
        want = istft_oracle.invert(full, z)
        for sizes in chunkings(frames, rng):
            k = lib.Stft.Synthesis(c, 2, 4096,
                                   _invert=lambda zw, length: istft_oracle.invert(left, zw, length=length))
            outs, at, emitted = [], 0, 0
            for m in sizes:
                o = k.step(z[..., at:at + m])
                at += m
                if o is not None:
                    outs.append(o)
                    emitted += o.shape[-1]
                # samples emitted after F frames = F H - synthesis_latency (stft.mli:489-494)
                assert emitted == max(0, at * full.hop - lib.Stft.synthesis_latency(c))
            o = k.flush()
            if o is not None:
                outs.append(o)
            assert k.flush() is None
            got = np.concatenate(outs, axis=-1) if outs else np.zeros((2, 0))
            assert got.shape == want.shape, (frames, sizes, got.shape, want.shape)
            assert np.array_equal(got, want), (frames, sizes)
            with pytest.raises(ValueError, match="drained kernel"):
                k.step(z[..., :1])


def test_synthesis_prepare_errors(lib):
    c = lib.Stft.Config.create(fft_size=64, hop=16)
    with pytest.raises(ValueError, match="cannot synthesise 0 channels"):
        lib.Stft.Synthesis.prepare(c, channels=0, max_block=4)
    with pytest.raises(ValueError, match="blocks of 0 frames"):
        lib.Stft.Synthesis.prepare(c, channels=1, max_block=0)
    bad = lib.Stft.Config.create(fft_size=64, hop=65)
    with pytest.raises(ValueError, match="prepare: cannot invert a 64-point window advanced by 65"):
        lib.Stft.Synthesis.prepare(bad, channels=1, max_block=4)
    k = lib.Stft.Synthesis.prepare(c, channels=1, max_block=4)
    with pytest.raises(ValueError, match="frequency bins"):
        k.step(np.zeros((32, 2), complex))
    with pytest.raises(ValueError, match="zero-size leading axis"):
        k.step(np.zeros((0, 33, 2), complex))
