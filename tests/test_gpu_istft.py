"""Stft.invert on the GPU against the reference's librosa goldens, the oracle and
the round-trip law.  ``pytest -m gpu``."""
import numpy as np
import pytest

from golden_util import (F32_ATOL, F32_RTOL, F64_ATOL, F64_RTOL, STFT_SEED, assert_close,
                         griffin_lim_magnitudes, istft_spectrum, lcg_signal)
from oracle import istft_oracle, stft_oracle

pytestmark = pytest.mark.gpu

FAST_TOL = 2e-6      # float32 synthesis kernel, max |got - want| / max |want| per call


@pytest.fixture(scope="module")
def sb(lib):
    import torch
    assert torch.cuda.is_available()
    return lib


def test_istft_goldens_on_gpu(sb, goldens):
    """All 88 synthesis goldens at the reference's own gates
    (soundml/test/istft/istft_goldens.ml:26-28, 79-86)."""
    n = 0
    for key, stem, name, e in goldens.cases("istft"):
        p = e["params"]
        c = sb.Stft.Config.create(fft_size=p["fft_size"], hop=p["hop"], win_length=p["win_length"],
                                  alignment=p["alignment"], pad=("constant", 0.0))
        z = istft_spectrum(p["fft_size"], p["frames"], p["dtype"])
        f32 = p["dtype"] == "float32"
        got = sb.Stft.invert(c, z, length=p.get("length"))
        assert got.dtype == (np.float32 if f32 else np.float64)
        assert_close(got, goldens.values(key), F32_RTOL if f32 else F64_RTOL,
                     F32_ATOL if f32 else F64_ATOL, key)
        n += 1
    assert n == 88


@pytest.mark.parametrize("fft,hop,win,alignment", [
    (256, 64, None, "centered"), (256, 64, None, "left"), (256, 64, None, "right"),
    (100, 30, 80, "centered"), (2048, 512, None, "centered"), (2048, 500, 1200, "left"),
    (31, 5, None, "centered"), (64, 1, None, "centered"), (512, 512, None, "centered"),
    (2048, 512, None, "right"), (2048, 1024, None, "centered"), (2048, 150, None, "centered"),
    (2048, 300, 2000, "left"), (1024, 256, None, "centered"), (1024, 200, 800, "right"),
    (128, 32, None, "left"),
])
def test_invert_matches_oracle(sb, fft, hop, win, alignment):
    window = "rectangular" if hop == fft else "hann"
    c = sb.Stft.Config.create(fft_size=fft, hop=hop, win_length=win, alignment=alignment,
                              window=window)
    oc = stft_oracle.StftConfig(fft, hop=hop, win_length=win, alignment=alignment, window=window)
    rng = np.random.default_rng(fft + hop)
    bins = fft // 2 + 1
    for frames in (1, 2, 7, 40) + ((3000 // hop + 20,) if fft == 2048 else ()):
        z = rng.standard_normal((2, 3, bins, frames)) + 1j * rng.standard_normal((2, 3, bins, frames))
        for length in (None, 0, 1, hop * frames // 2 + 3, istft_oracle.output_length(oc, frames) + 50):
            want = istft_oracle.invert(oc, z, length=length)
            got = sb.Stft.invert(c, z, length=length)
            assert got.shape == want.shape and got.dtype == np.float64
            if want.size:
                scale = max(np.abs(want).max(), 1e-300)
                assert np.abs(got - want).max() / scale <= 1e-12, (frames, length)
        z32 = z.astype(np.complex64)
        want32 = istft_oracle.invert(oc, z32, dtype=np.float32)
        c.set_path("generic")                    # double interior: the reference's f32 gate
        got32 = sb.Stft.invert(c, z32)
        assert got32.dtype == np.float32
        assert_close(got32, want32, F32_RTOL, F32_ATOL, "f32")
        c.set_path("auto")                       # fft 2048: float32 register-FFT kernel
        auto32 = sb.Stft.invert(c, z32)
        if want32.size:
            peak = max(np.abs(want32).max(), 1e-30)
            assert np.abs(auto32 - want32).max() / peak <= FAST_TOL, (frames,)
        mixed = sb.Stft.invert(c, z32, dtype=np.float64)
        assert mixed.dtype == np.float64


def test_fast_and_generic_kernels_agree_on_long_batches(sb):
    """fft 2048 / hop 512 complex64: the register-FFT kernel against the double
    interior kernel over many runs per clip, explicit lengths and a ragged tail."""
    import torch
    c = sb.Stft.Config.create(fft_size=2048, hop=512)
    g = torch.Generator(device="cuda").manual_seed(3)
    z = torch.randn((5, 1025, 431), dtype=torch.complex64, device="cuda", generator=g)
    for length in (None, 220500, 100000, 230000, 1):
        c.set_path("fast")
        fast = sb.Stft.invert(c, z, length=length)
        c.set_path("generic")
        ref = sb.Stft.invert(c, z, length=length)
        assert fast.shape == ref.shape
        err = (fast - ref).abs().max().item() / ref.abs().max().item()
        assert err <= FAST_TOL, (length, err)
    c.set_path("auto")


def test_round_trip_at_headline_geometry(sb):
    """invert(transform(x)) == x (istft_law.ml): fft 2048 / hop 512, 64 clips of
    10 s through the device path, float32 end to end."""
    import torch
    from soundml_b200 import synth
    c = sb.Stft.Config.create(fft_size=2048, hop=512)
    x = synth.clips_torch(64, 220500, device="cuda")
    z = sb.Stft.transform(c, x)
    y = sb.Stft.invert(c, z, length=x.shape[-1])
    assert y.shape == x.shape and y.dtype == torch.float32
    err = (y - x).abs().max().item() / x.abs().max().item()
    assert err <= 1e-5, err
    # frames past the requested length are never read (stft.ml:907-915)
    short = sb.Stft.invert(c, z, length=5000)
    assert torch.equal(short, sb.Stft.invert(c, z[..., :12].contiguous(), length=5000))


def test_invert_device_and_host_agree_and_errors(sb):
    import torch
    c = sb.Stft.Config.create(fft_size=128, hop=32)
    rng = np.random.default_rng(5)
    z = (rng.standard_normal((4, 65, 20)) + 1j * rng.standard_normal((4, 65, 20))).astype(np.complex64)
    host = sb.Stft.invert(c, z)
    dev = sb.Stft.invert(c, torch.from_numpy(z).cuda())
    assert np.array_equal(host, dev.cpu().numpy())
    assert sb.Stft.invert(c, np.zeros((0, 65, 3), np.complex64)).shape == (0, sb.Stft.output_length(c, 3))
    assert sb.Stft.invert(c, np.zeros((65, 0), np.complex128)).shape == (0,)
    assert np.array_equal(sb.Stft.invert(c, np.zeros((65, 0), np.complex128), length=7), np.zeros(7))
    with pytest.raises(ValueError, match="frequency bins"):
        sb.Stft.invert(c, np.zeros((64, 3), np.complex64))
    with pytest.raises(ValueError, match="rank-1"):
        sb.Stft.invert(c, np.zeros(65, np.complex64))
    with pytest.raises(ValueError, match="length must be non-negative"):
        sb.Stft.invert(c, z, length=-1)
    bad = sb.Stft.Config.create(fft_size=128, hop=129)
    assert not sb.Stft.nola(bad)
    with pytest.raises(ValueError, match="overlap-added squared window"):
        sb.Stft.invert(bad, z)


def test_griffin_lim_goldens_on_gpu(sb, goldens):
    """All 42 Griffin-Lim goldens through the CUDA path at the reference's gates."""
    n = 0
    for key, stem, name, e in goldens.cases("griffinlim"):
        p = e["params"]
        c = sb.Stft.Config.create(fft_size=p["fft_size"], hop=p["hop"], win_length=p["win_length"],
                                  alignment=p["alignment"], pad=("constant", 0.0))
        mags = griffin_lim_magnitudes(p["fft_size"], p["frames"], p["dtype"])
        f32 = p["dtype"] == "float32"
        got = sb.Stft.griffin_lim(c, mags, n_iter=p["n_iter"], momentum=p["momentum"])
        assert got.dtype == mags.dtype
        assert_close(got, goldens.values(key), F32_RTOL if f32 else F64_RTOL,
                     F32_ATOL if f32 else F64_ATOL, key)
        n += 1
    assert n == 42


def test_griffin_lim_matches_oracle_and_errors(sb):
    import torch
    c = sb.Stft.Config.create(fft_size=128, hop=32)
    oc = stft_oracle.StftConfig(128, hop=32)
    rng = np.random.default_rng(9)
    mags = rng.uniform(0.1, 2.0, (3, 65, 12))
    phase = rng.uniform(-3, 3, mags.shape)
    for kw in (dict(n_iter=3, momentum=0.99), dict(n_iter=2, momentum=0.0, length=200),
               dict(n_iter=4, momentum=0.5, init=phase)):
        okw = {("init_phase" if k == "init" else k): v for k, v in kw.items()}
        want = istft_oracle.griffin_lim(oc, mags, **okw)
        got = sb.Stft.griffin_lim(c, mags, **kw)
        assert got.shape == want.shape and got.dtype == np.float64
        assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max(), kw
    dev = sb.Stft.griffin_lim(c, torch.from_numpy(mags).cuda(), n_iter=3, momentum=0.99)
    assert np.array_equal(dev.cpu().numpy(), sb.Stft.griffin_lim(c, mags, n_iter=3, momentum=0.99))
    # reconstruction of a real signal's magnitudes is consistent: re-analysis matches them
    x = lcg_signal(2000, STFT_SEED)
    s = np.abs(sb.Stft.transform(c, x))
    y = sb.Stft.griffin_lim(c, s, n_iter=32)
    s2 = np.abs(sb.Stft.transform(c, y))[..., :s.shape[-1]]
    assert np.linalg.norm(s2 - s) / np.linalg.norm(s) < 0.2
    with pytest.raises(ValueError, match="n_iter must be at least 1"):
        sb.Stft.griffin_lim(c, mags, n_iter=0)
    with pytest.raises(ValueError, match="momentum must be non-negative"):
        sb.Stft.griffin_lim(c, mags, momentum=-0.1)
    with pytest.raises(ValueError, match="initial phase must have the shape"):
        sb.Stft.griffin_lim(c, mags, init=phase[:, :, :5])
    with pytest.raises(ValueError, match="frequency bins"):
        sb.Stft.griffin_lim(c, mags[:, :64])
