"""Stft.Kernel through the CUDA kernels: chunked analysis equals the offline
transform bit for bit, on host arrays and on device tensors.
``pytest -m gpu``."""
import numpy as np
import pytest

from golden_util import STFT_SEED, lcg_signal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb(lib):
    import torch
    assert torch.cuda.is_available()
    return lib


def run_chunks(k, x, sizes, cat):
    outs, at = [], 0
    for m in sizes:
        o = k.step(x[..., at:at + m])
        at += m
        if o is not None:
            outs.append(o)
    o = k.flush()
    if o is not None:
        outs.append(o)
    return cat(outs) if outs else None


@pytest.mark.parametrize("case", [
    dict(fft_size=64, hop=16, alignment="centered", pad="reflect"),
    dict(fft_size=100, hop=30, win_length=80, alignment="right", pad="edge"),
    dict(fft_size=16, hop=40, alignment="centered", pad=("constant", 0.5)),
    dict(fft_size=2048, hop=512, alignment="centered", pad="reflect"),
    dict(fft_size=2048, hop=500, win_length=1200, alignment="left", pad="reflect"),
])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_partition_law_on_gpu(sb, case, dtype):
    c = sb.Stft.Config.create(**case)
    rng = np.random.default_rng(11)
    big = case["fft_size"] == 2048
    for n in ((1, 700, 5000, 20011) if big else (1, 33, 517, 4000)):
        x = lcg_signal(2 * n, STFT_SEED).reshape(2, n).astype(dtype)
        want = sb.Stft.transform(c, x)
        for _ in range(3):
            cuts = np.sort(rng.integers(0, n + 1, size=rng.integers(0, 6)))
            sizes = list(np.diff(np.concatenate([[0], cuts, [n]])))
            k = sb.Stft.Kernel.prepare(c, channels=2, max_block=n)
            got = run_chunks(k, x, sizes, lambda o: np.concatenate(o, axis=-1))
            if want.shape[-1] == 0:
                assert got is None
            else:
                assert got.dtype == want.dtype and got.shape == want.shape, (n, sizes)
                assert np.array_equal(got, want), (n, sizes)


def test_device_chunks_match_offline_transform(sb):
    import torch
    from soundml_b200 import synth
    c = sb.Stft.Config.create(fft_size=2048, hop=512)
    x = synth.clips_torch(8, 220500, device="cuda")
    want = sb.Stft.transform(c, x)
    k = sb.Stft.Kernel.prepare(c, channels=8, max_block=65536)
    sizes = [1000, 24, 65536, 50000, 1, 65536, 220500 - 1000 - 24 - 65536 - 50000 - 1 - 65536]
    got = run_chunks(k, x, sizes, lambda o: torch.cat(o, dim=-1))
    assert got.is_cuda and got.shape == want.shape
    assert torch.equal(got, want)
    k.reset()
    again = run_chunks(k, x, [220500], lambda o: torch.cat(o, dim=-1))
    assert torch.equal(again, want)


# ---- Resample.Kernel ----------------------------------------------------------

@pytest.mark.parametrize("sr,target,executor", [
    (44100, 16000, "direct"), (44100, 16000, "planned"), (44100, 22050, "direct"),
    (44100, 22050, "planned"), (22050, 44100, "planned"), (48000, 44100, "direct"),
    (8000, 48000, "planned"), (44100, 48000, "direct"),
])
def test_resample_kernel_partition_law_on_gpu(sb, sr, target, executor):
    """Chunked resampling equals apply: bit for bit with the direct executor,
    to 1e-5 of peak where block executors (overlap-save, tcgen05) are planned."""
    cfg = sb.Resample.Config.create(sample_rate=sr, target=target).set_executor(executor)
    rng = np.random.default_rng(sr)
    for n in (1, 900, 30011):
        x = rng.uniform(-1, 1, (2, n)).astype(np.float32)
        want = sb.Resample.apply(cfg, x)
        for _ in range(2):
            cuts = np.sort(rng.integers(0, n + 1, size=rng.integers(0, 5)))
            sizes = list(np.diff(np.concatenate([[0], cuts, [n]])))
            k = sb.Resample.Kernel.prepare(cfg, channels=2, max_block=n)
            got = run_chunks(k, x, sizes, lambda o: np.concatenate(o, axis=-1))
            assert got is not None and got.shape == want.shape, (n, sizes)
            if executor == "direct":
                assert np.array_equal(got, want), (n, sizes)
            else:
                assert np.abs(got - want).max() <= 1e-5 * max(np.abs(want).max(), 1e-3), (n, sizes)


def test_resample_kernel_device_stream(sb):
    import torch
    cfg = sb.Resample.Config.create(sample_rate=44100, target=22050)
    x = torch.rand((4, 200000), device="cuda") * 2 - 1
    want = sb.Resample.apply(cfg, x)
    k = sb.Resample.Kernel.prepare(cfg, channels=4, max_block=70000)
    got = run_chunks(k, x, [65536, 1, 65536, 200000 - 2 * 65536 - 1], lambda o: torch.cat(o, dim=-1))
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= 1e-5 * want.abs().max().item()


# ---- Stft.Synthesis -------------------------------------------------------------

@pytest.mark.parametrize("case,cdtype", [
    (dict(fft_size=64, hop=16, alignment="centered"), np.complex128),
    (dict(fft_size=100, hop=30, win_length=80, alignment="right"), np.complex128),
    (dict(fft_size=64, hop=64, window="rectangular", alignment="centered"), np.complex64),
    (dict(fft_size=2048, hop=512, alignment="centered"), np.complex64),
    (dict(fft_size=2048, hop=500, win_length=1200, alignment="left"), np.complex64),
])
def test_synthesis_partition_law_on_gpu(sb, case, cdtype):
    """Chunked synthesis equals invert at its default length bit for bit, on the
    double-interior kernel and on the fft-2048 float32 kernel."""
    c = sb.Stft.Config.create(**case)
    rng = np.random.default_rng(21)
    for frames in (1, 5, 37, 150):
        z = (rng.standard_normal((2, c.bins, frames)) +
             1j * rng.standard_normal((2, c.bins, frames))).astype(cdtype)
        want = sb.Stft.invert(c, z)
        for _ in range(3):
            cuts = np.sort(rng.integers(0, frames + 1, size=rng.integers(0, 5)))
            sizes = list(np.diff(np.concatenate([[0], cuts, [frames]])))
            k = sb.Stft.Synthesis.prepare(c, channels=2, max_block=frames)
            got = run_chunks(k, z, sizes, lambda o: np.concatenate(o, axis=-1))
            if want.shape[-1] == 0:
                assert got is None
            else:
                assert got.shape == want.shape and got.dtype == want.dtype, (frames, sizes)
                assert np.array_equal(got, want), (frames, sizes)


def test_streaming_round_trip_on_device(sb):
    """Analysis kernel -> synthesis kernel, chunk by chunk on device tensors,
    reconstructs the signal (less the synthesis latency, made up at flush)."""
    import torch
    from soundml_b200 import synth
    c = sb.Stft.Config.create(fft_size=2048, hop=512)
    x = synth.clips_torch(4, 100000, device="cuda")
    ana = sb.Stft.Kernel.prepare(c, channels=4, max_block=30000)
    syn = sb.Stft.Synthesis.prepare(c, channels=4, max_block=64)
    outs = []
    for start in range(0, 100000, 30000):
        z = ana.step(x[..., start:start + 30000])
        if z is not None:
            y = syn.step(z)
            if y is not None:
                outs.append(y)
    z = ana.flush()
    if z is not None:
        y = syn.step(z)
        if y is not None:
            outs.append(y)
    y = syn.flush()
    if y is not None:
        outs.append(y)
    got = torch.cat(outs, dim=-1)
    assert got.shape[-1] == sb.Stft.output_length(c, sb.Stft.frames(c, 100000))
    m = min(got.shape[-1], 100000)
    err = (got[..., :m] - x[..., :m]).abs().max().item() / x.abs().max().item()
    assert err <= 1e-5, err
