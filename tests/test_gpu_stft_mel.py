"""Parity of the CUDA path (through the C ABI) against the oracle and the
reference's golden vectors.  Run on the B200 box: ``pytest -m gpu``.

Tolerances: the generic kernels keep the reference's double interior, so they
are held to the reference's own gates (f64 rtol 1e-9 / atol 1e-12, f32 rtol
1e-6 / atol 1e-7).  The fused fft-2048 kernel is float32 inside and is held to
BASELINE.json's bar: max |got - ref| / max |ref| <= 1e-4 per clip.
"""
import numpy as np
import pytest

from golden_util import (F32_ATOL, F32_RTOL, F64_ATOL, F64_RTOL, MEL_SEED, STFT_SEED,
                         assert_close, lcg_signal, peak_rel_err)
from oracle import mel_oracle, stft_oracle

pytestmark = pytest.mark.gpu

SPECTRUM_TOL = 1e-4      # BASELINE.json north_star: spectrogram max rel err


@pytest.fixture(scope="module")
def sb(lib):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return lib


@pytest.fixture(params=["fast", "tensor"])
def fused(request):
    """The two fused fft-2048 kernels: the CUDA-core register FFT ("fast") and the
    tcgen05 one ("tensor"); both are held to the same bar."""
    return request.param


def _signal(n, seed=7):
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    return (0.6 * np.sin(2 * np.pi * 440.0 * t / 22050.0) + 0.1 * rng.uniform(-1, 1, n)).astype(np.float32)


# ---------------------------------------------------------------- goldens ----

def test_stft_goldens_on_gpu(sb, goldens):
    launches = sb.kernel_launch_count()
    n = 0
    for key, stem, name, e in goldens.cases("stft"):
        if stem == "coordinates":
            continue
        p = e["params"]
        c = sb.Stft.Config.create(fft_size=p["fft_size"], hop=p["hop"], win_length=p["win_length"],
                                  alignment=p["alignment"])
        x = lcg_signal(p["length"], STFT_SEED)
        if p["dtype"] == "float32":
            x = x.astype(np.float32)
        if p["kind"] in ("magnitude", "power"):
            got = sb.Stft.power_spectrum(c, x, 1.0 if p["kind"] == "magnitude" else 2.0)
        else:
            z = sb.Stft.transform(c, x)
            got = z.real if p["kind"] == "real" else z.imag
        rtol, atol = (F64_RTOL, F64_ATOL) if p["dtype"] == "float64" else (F32_RTOL, F32_ATOL)
        assert_close(got, goldens.values(key), rtol, atol, key)
        n += 1
    assert n == 66
    assert sb.kernel_launch_count() - launches >= 66      # the CUDA path did the work


def test_mel_spectrogram_goldens_on_gpu(sb, goldens):
    for key, stem, name, e in goldens.cases("mel", "mel_spectrogram"):
        p = e["params"]
        sc = sb.Stft.Config.create(fft_size=p["fft_size"], hop=p["hop"], alignment=p["alignment"])
        mc = sb.Mel.Config.create(n_mels=p["n_mels"], sample_rate=p["sample_rate"],
                                  fft_size=p["fft_size"], f_min=p["f_min"], f_max=p["f_max"],
                                  scale=p["scale"], norm=p["norm"])
        x = lcg_signal(p["length"], MEL_SEED, p["envelope"])
        if p["dtype"] == "float32":
            x = x.astype(np.float32)
        got = sb.mel_spectrogram(sc, mc, x, power=p["power"])
        rtol, atol = (F64_RTOL, F64_ATOL) if p["dtype"] == "float64" else (F32_RTOL, F32_ATOL)
        assert_close(got, goldens.values(key), rtol, atol, key)


# ------------------------------------------------ config 1: one 10 s clip ----

@pytest.mark.parametrize("path", ["fast", "tensor", "generic"])
def test_config1_clip_against_oracle(sb, path):
    from soundml_b200 import synth
    x = synth.clips_numpy(1, 220500)[0]
    c = sb.Stft.Config.create(fft_size=2048, hop=512).set_path(path)
    o = stft_oracle.StftConfig(2048, 512)
    assert sb.Stft.frames(c, x.size) == 431
    ref_z = stft_oracle.transform(o, x)
    z = sb.Stft.transform(c, x)
    assert z.shape == (1025, 431) and z.dtype == np.complex64
    assert peak_rel_err(z.view(np.float32), ref_z.view(np.float32)) <= SPECTRUM_TOL
    ref_p = stft_oracle.power_spectrum(o, x)
    p = sb.Stft.power_spectrum(c, x)
    assert peak_rel_err(p, ref_p) <= SPECTRUM_TOL
    m1 = sb.Stft.power_spectrum(c, x, power=1.0)
    assert peak_rel_err(m1, stft_oracle.power_spectrum(o, x, 1.0)) <= SPECTRUM_TOL
    mc = sb.Mel.Config.create(n_mels=128, sample_rate=22050, fft_size=2048)
    mo = mel_oracle.MelConfig(128, 22050, 2048)
    mel = sb.mel_spectrogram(c, mc, x)
    assert mel.shape == (128, 431)
    assert peak_rel_err(mel, mel_oracle.apply(mo, ref_p)) <= SPECTRUM_TOL


def test_fast_kernel_error_is_float32_class(sb, fused):
    """The fused kernels are float32 inside (the tensor-core one through split
    fp16 operands, 22 significant bits); report how close they land."""
    x = _signal(50000)
    c = sb.Stft.Config.create(fft_size=2048, hop=512).set_path(fused)
    o = stft_oracle.StftConfig(2048, 512)
    err = peak_rel_err(sb.Stft.power_spectrum(c, x), stft_oracle.power_spectrum(o, x))
    assert err <= (2e-6 if fused == "fast" else 6e-6), err


@pytest.mark.parametrize("amp", [1e-30, 1e-12, 1e-3, 37.0, 32768.0, 1e12])
def test_tensor_kernel_is_scale_free(sb, amp):
    """fp16 operands have 5 exponent bits: the kernel rescales every tile by a
    power of two, so the error does not depend on the signal's level."""
    rng = np.random.default_rng(11)
    x = (rng.uniform(-1, 1, 20000) * amp).astype(np.float32)
    c = sb.Stft.Config.create(fft_size=2048, hop=512).set_path("tensor")
    o = stft_oracle.StftConfig(2048, 512)
    z, ref = sb.Stft.transform(c, x), stft_oracle.transform(o, x)
    assert peak_rel_err(z.view(np.float32), ref.view(np.float32)) <= 3e-6
    p = sb.Stft.power_spectrum(c, x, power=1.0)
    assert peak_rel_err(p, stft_oracle.power_spectrum(o, x, 1.0)) <= 3e-6
    assert not np.abs(sb.Stft.power_spectrum(c, np.zeros(5000, np.float32))).any()


# ------------------------------------------- geometry sweep, fast kernel ----

@pytest.mark.parametrize("hop", [512, 500, 333, 128, 1, 600])
@pytest.mark.parametrize("alignment", ["centered", "left", "right"])
def test_fast_kernel_geometries(sb, hop, alignment, fused):
    x = np.stack([_signal(9000, 1), _signal(9000, 2), _signal(9000, 3)])
    if hop == 1:
        x = x[:, :2600]
    c = sb.Stft.Config.create(fft_size=2048, hop=hop, alignment=alignment).set_path(fused)
    o = stft_oracle.StftConfig(2048, hop, alignment=alignment)
    ref = stft_oracle.power_spectrum(o, x)
    got = sb.Stft.power_spectrum(c, x)
    assert got.shape == ref.shape
    for b in range(x.shape[0]):
        assert peak_rel_err(got[b], ref[b]) <= SPECTRUM_TOL, (hop, alignment, b)
    refz = stft_oracle.transform(o, x)
    gotz = sb.Stft.transform(c, x)
    assert peak_rel_err(gotz.view(np.float32), refz.view(np.float32)) <= SPECTRUM_TOL


@pytest.mark.parametrize("fft,hop,wl,n_mels", [(1024, 256, None, 80), (1024, 160, 800, 128),
                                               (512, 128, None, 80), (512, 200, 400, 40),
                                               (256, 64, None, 40), (128, 32, None, 20)])
@pytest.mark.parametrize("alignment,pad", [("centered", "reflect"), ("left", "edge"),
                                           ("right", ("constant", 0.25))])
def test_shorter_frames_ride_the_fused_kernel(sb, fft, hop, wl, n_mels, alignment, pad, fused):
    """fft 1024 / 512 / 256 / 128 run zero-padded inside the fft-2048 kernel; power,
    complex and mel outputs against the oracle, and the forced fast path must
    accept them."""
    name = pad if isinstance(pad, str) else pad[0]
    val = 0.0 if isinstance(pad, str) else pad[1]
    c = sb.Stft.Config.create(fft_size=fft, hop=hop, win_length=wl, alignment=alignment,
                              pad=pad).set_path(fused)
    o = stft_oracle.StftConfig(fft, hop, win_length=wl, alignment=alignment, pad=name,
                               pad_value=val)
    mc = sb.Mel.Config.create(n_mels=n_mels, sample_rate=16000, fft_size=fft)
    mo = mel_oracle.MelConfig(n_mels, 16000, fft)
    for n in (1, fft // 2, fft + 1, 7001):
        x = np.stack([_signal(n, 1), _signal(n, 2) + 0.1])
        ref = stft_oracle.power_spectrum(o, x)
        got = sb.Stft.power_spectrum(c, x)
        assert got.shape == ref.shape, (n, got.shape, ref.shape)
        if ref.size == 0:
            continue
        for b in range(2):
            assert peak_rel_err(got[b], ref[b]) <= SPECTRUM_TOL, (fft, hop, alignment, n, b)
        refz = stft_oracle.transform(o, x)
        gotz = sb.Stft.transform(c, x)
        assert peak_rel_err(gotz.view(np.float32), refz.view(np.float32)) <= SPECTRUM_TOL
        refm = mel_oracle.mel_spectrogram(o, mo, x)
        gotm = sb.mel_spectrogram(c, mc, x)
        assert gotm.shape == refm.shape
        for b in range(2):
            assert peak_rel_err(gotm[b], refm[b]) <= SPECTRUM_TOL, (fft, hop, alignment, n, b)
        c.set_path("generic")
        assert peak_rel_err(sb.mel_spectrogram(c, mc, x), gotm) <= SPECTRUM_TOL
        c.set_path(fused)


@pytest.mark.parametrize("pad", ["reflect", "edge", ("constant", 0.5)])
@pytest.mark.parametrize("n", [1, 2, 700, 1024, 1025, 2047, 2048, 2049, 5000])
def test_fast_kernel_short_signals_and_pads(sb, pad, n, fused):
    """n <= fft/2 exercises multi-reflection (stft.ml:297-305)."""
    x = _signal(n, 5) + 0.25
    name = pad if isinstance(pad, str) else pad[0]
    val = 0.0 if isinstance(pad, str) else pad[1]
    c = sb.Stft.Config.create(fft_size=2048, hop=512, pad=pad).set_path(fused)
    o = stft_oracle.StftConfig(2048, 512, pad=name, pad_value=val)
    ref = stft_oracle.power_spectrum(o, x)
    got = sb.Stft.power_spectrum(c, x)
    assert got.shape == ref.shape == (1025, 1 + n // 512)
    assert peak_rel_err(got, ref) <= SPECTRUM_TOL


@pytest.mark.parametrize("window,win_length,scale", [
    ("hamming", None, "none"), (("kaiser", 8.6), 1200, "magnitude"), ("blackman", 2000, "psd"),
    (("tukey", 0.25), None, "none"), ("rectangular", 512, "none")])
def test_fast_kernel_windows(sb, window, win_length, scale, fused):
    x = _signal(12000, 9)
    c = sb.Stft.Config.create(fft_size=2048, hop=500, window=window, win_length=win_length,
                              scale=scale).set_path(fused)
    name, param = (window, 0.0) if isinstance(window, str) else window
    o = stft_oracle.StftConfig(2048, 500, win_length, window=name, window_param=param, scale=scale)
    assert peak_rel_err(sb.Stft.power_spectrum(c, x), stft_oracle.power_spectrum(o, x)) <= SPECTRUM_TOL


def test_fast_mel_matches_oracle_on_batch(sb, fused):
    from soundml_b200 import synth
    x = synth.clips_numpy(5, 30000, first_clip=57)
    c = sb.Stft.Config.create(fft_size=2048, hop=512).set_path(fused)
    o = stft_oracle.StftConfig(2048, 512)
    for n_mels, kw in [(128, {}), (40, dict(scale="htk", norm="none")), (80, dict(f_min=300.0, f_max=8000.0))]:
        mc = sb.Mel.Config.create(n_mels=n_mels, sample_rate=22050, fft_size=2048, **kw)
        mo = mel_oracle.MelConfig(n_mels, 22050, 2048, **kw)
        ref = mel_oracle.mel_spectrogram(o, mo, x)
        got = sb.mel_spectrogram(c, mc, x)
        assert got.shape == ref.shape
        for b in range(x.shape[0]):
            assert peak_rel_err(got[b], ref[b]) <= SPECTRUM_TOL, (n_mels, b)
        # power = 1 (magnitude mel)
        assert peak_rel_err(sb.mel_spectrogram(c, mc, x, power=1.0),
                            mel_oracle.mel_spectrogram(o, mo, x, 1.0)) <= SPECTRUM_TOL


def test_dense_custom_weights_take_generic_mel(sb):
    rng = np.random.default_rng(3)
    w = rng.uniform(0, 1, (12, 1025))
    x = _signal(8000)
    c = sb.Stft.Config.create(fft_size=2048, hop=512)
    mc = sb.Mel.Config.of_weights(w, 2048)
    o = stft_oracle.StftConfig(2048, 512)
    ref = (w @ stft_oracle.power_spectrum(o, x).astype(np.float64)).astype(np.float32)
    assert peak_rel_err(sb.mel_spectrogram(c, mc, x), ref) <= SPECTRUM_TOL


# ------------------------------------------------------ generic kernels ----

@pytest.mark.parametrize("fft,hop,wl", [(16, 4, None), (31, 5, None), (64, 17, 40), (100, 30, None),
                                         (512, 128, None), (1024, 256, 800), (4096, 1024, None),
                                         (16, 40, None), (1, 1, None)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_generic_kernel_any_geometry(sb, fft, hop, wl, dtype):
    x = np.stack([_signal(3000, 11), _signal(3000, 12)]).astype(dtype)
    c = sb.Stft.Config.create(fft_size=fft, hop=hop, win_length=wl)
    o = stft_oracle.StftConfig(fft, hop, wl)
    rtol, atol = (F64_RTOL, F64_ATOL * 100) if dtype == np.float64 else (1e-5, 1e-6)
    ref = stft_oracle.power_spectrum(o, x)
    got = sb.Stft.power_spectrum(c, x)
    assert got.dtype == dtype and got.shape == ref.shape
    assert_close(got, ref, rtol, atol * max(1.0, float(np.abs(ref).max())), (fft, hop))
    refz = stft_oracle.transform(o, x)
    gotz = sb.Stft.transform(c, x)
    assert gotz.dtype == refz.dtype
    assert peak_rel_err(gotz.view(dtype), refz.view(dtype)) <= (1e-12 if dtype == np.float64 else 1e-6)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_mel_apply_standalone(sb, dtype):
    rng = np.random.default_rng(5)
    s = rng.uniform(0, 10, (2, 3, 257, 50)).astype(dtype)
    mc = sb.Mel.Config.create(n_mels=40, sample_rate=22050, fft_size=512)
    mo = mel_oracle.MelConfig(40, 22050, 512)
    got = sb.Mel.apply(mc, s)
    ref = mel_oracle.apply(mo, s)
    assert got.shape == (2, 3, 40, 50) and got.dtype == dtype
    assert_close(got, ref, 1e-12 if dtype == np.float64 else 1e-6, 1e-12 if dtype == np.float64 else 1e-6)
    with pytest.raises(ValueError, match=r"apply: cannot project 100 frequency bins"):
        sb.Mel.apply(mc, np.zeros((100, 4), dtype))
    assert sb.Mel.apply(mc, np.zeros((257, 0), dtype)).shape == (40, 0)


# ----------------------------------------------------- laws and edge cases ----

def test_leading_axes_equal_standalone_calls_bitwise(sb):
    """stft_grid.ml:180-205: a batch slice equals the standalone call."""
    x = np.stack([_signal(7000, s) for s in range(6)]).reshape(2, 3, 7000)
    c = sb.Stft.Config.create(fft_size=2048, hop=512)
    mc = sb.Mel.Config.create(n_mels=128, sample_rate=22050, fft_size=2048)
    whole = sb.mel_spectrogram(c, mc, x)
    assert whole.shape == (2, 3, 128, 14)
    for i in range(2):
        for j in range(3):
            assert np.array_equal(whole[i, j], sb.mel_spectrogram(c, mc, x[i, j]))
    wz = sb.Stft.transform(c, x)
    assert np.array_equal(wz[1, 2], sb.Stft.transform(c, x[1, 2]))


def test_empty_inputs(sb):
    c = sb.Stft.Config.create(fft_size=2048, hop=512)
    mc = sb.Mel.Config.create(n_mels=128, sample_rate=22050, fft_size=2048)
    assert sb.Stft.power_spectrum(c, np.zeros(0, np.float32)).shape == (1025, 0)
    assert sb.Stft.transform(c, np.zeros((3, 0), np.float32)).shape == (3, 1025, 0)
    assert sb.mel_spectrogram(c, mc, np.zeros((0, 5000), np.float32)).shape == (0, 128, 10)
    left = sb.Stft.Config.create(fft_size=2048, hop=512, alignment="left")
    assert sb.Stft.power_spectrum(left, np.zeros(2047, np.float32)).shape == (1025, 0)
    with pytest.raises(ValueError, match="rank-zero"):
        sb.Stft.power_spectrum(c, np.float32(1.0) * np.ones((), np.float32))
    m512 = sb.Mel.Config.create(n_mels=40, sample_rate=22050, fft_size=512)
    with pytest.raises(ValueError, match=r"mel_spectrogram: cannot project a 2048-point STFT through a "
                                         r"filterbank built for an FFT of size 512"):
        sb.mel_spectrogram(c, m512, np.zeros(4000, np.float32))


def test_device_tensors_run_in_place_on_torch_stream(sb):
    import torch
    from soundml_b200 import synth
    xh = synth.clips_numpy(4, 20000)
    xd = torch.from_numpy(xh).cuda()
    c = sb.Stft.Config.create(fft_size=2048, hop=512)
    mc = sb.Mel.Config.create(n_mels=128, sample_rate=22050, fft_size=2048)
    host = sb.mel_spectrogram(c, mc, xh)
    dev = sb.mel_spectrogram(c, mc, xd)
    torch.cuda.synchronize()
    assert dev.is_cuda and dev.shape == (4, 128, 40)
    assert np.array_equal(dev.cpu().numpy(), host)           # same kernel, same bits
    z = sb.Stft.transform(c, xd)
    torch.cuda.synchronize()
    assert z.dtype == torch.complex64
    assert np.array_equal(z.cpu().numpy(), sb.Stft.transform(c, xh))


def test_parseval_at_scale(sb):
    """Size-independent check at a batch the oracle cannot reach quickly:
    sum_k c_k |X_k|^2 == N * sum_j (w_j x_j)^2 for every frame."""
    import torch
    from soundml_b200 import synth
    xd = synth.clips_torch(96, 220500, "cuda")
    c = sb.Stft.Config.create(fft_size=2048, hop=512)
    p = sb.Stft.power_spectrum(c, xd)
    torch.cuda.synchronize()
    assert p.shape == (96, 1025, 431)
    w = torch.from_numpy(c.analysis_window).cuda()
    padded = torch.nn.functional.pad(xd.double()[:, None, :], (1024, 1024), mode="reflect")[:, 0]
    fr = padded.unfold(-1, 2048, 512) * w                     # [96, 431, 2048]
    energy = 2048.0 * (fr * fr).sum(-1)                       # [96, 431]
    weights = torch.full((1025,), 2.0, dtype=torch.float64, device="cuda")
    weights[0] = weights[-1] = 1.0
    lhs = (p.double() * weights[None, :, None]).sum(1)        # [96, 431]
    rel = ((lhs - energy).abs() / energy).max().item()
    assert rel <= 2e-6, rel


@pytest.mark.parametrize("fft,hop,path", [(2048, 512, "fast"), (2048, 300, "fast"), (64, 16, "generic")])
def test_transform_range_tiles_the_full_transform(sb, fft, hop, path):
    """stft_grid.ml:32-73: adjacent ranges reassemble the transform exactly."""
    x = np.stack([_signal(20000, 21), _signal(20000, 22)])
    c = sb.Stft.Config.create(fft_size=fft, hop=hop).set_path(path)
    full = sb.Stft.transform(c, x)
    total = sb.Stft.frames(c, 20000)
    cuts = [0, 1, 7, 8, 9, total // 2, total - 1, total]
    parts = [sb.Stft.transform_range(c, x, a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(np.concatenate(parts, axis=-1), full)
    assert sb.Stft.transform_range(c, x, 5, 5).shape == (2, fft // 2 + 1, 0)
    with pytest.raises(ValueError, match=r"transform_range: cannot take frames \[3, 2\)"):
        sb.Stft.transform_range(c, x, 3, 2)
    with pytest.raises(ValueError, match=r"the range must satisfy 0 <= p0 <= p1 <= frames"):
        sb.Stft.transform_range(c, x, 0, total + 1)


@pytest.mark.parametrize("frames", [431, 100, 101, 102, 104, 57, 9, 8, 3])
@pytest.mark.parametrize("fft", [2048, 512])
def test_bin_major_write_out_carries_sectors_across_tiles(sb, frames, fft, monkeypatch):
    """The power / complex write-out of the fused kernel parks the frames that do
    not fill a 32-byte sector of an output row and writes them with the next tile
    of the same clip (stft2048.cu, write-out).  Batches large enough that a group
    walks several tiles and crosses clip boundaries, row lengths of every phase
    mod 8: same bits as the plain write-out, and the oracle's values."""
    hop = fft // 4
    n = (frames - 1) * hop + 3
    batch = 19 if frames > 200 else 90            # > 296 tiles whenever frames > 8*4
    x = np.stack([_signal(n, 100 + s) for s in range(batch)])
    c = sb.Stft.Config.create(fft_size=fft, hop=hop).set_path("fast")
    assert sb.Stft.frames(c, n) == frames
    got_p = sb.Stft.power_spectrum(c, x)
    got_z = sb.Stft.transform(c, x)
    monkeypatch.setenv("SMB_NO_CARRY", "1")
    plain_p = sb.Stft.power_spectrum(c, x)
    plain_z = sb.Stft.transform(c, x)
    monkeypatch.delenv("SMB_NO_CARRY")
    assert np.array_equal(got_p, plain_p)
    assert np.array_equal(got_z.view(np.float32), plain_z.view(np.float32))
    o = stft_oracle.StftConfig(fft, hop)
    ref = stft_oracle.power_spectrum(o, x[:8])
    for b in range(8):
        assert peak_rel_err(got_p[b], ref[b]) <= SPECTRUM_TOL, b
    refz = stft_oracle.transform(o, x[-3:])
    assert peak_rel_err(got_z[-3:].view(np.float32), refz.view(np.float32)) <= SPECTRUM_TOL
