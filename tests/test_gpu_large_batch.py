"""Batches past the 65 535 limit of a grid's second dimension (and past one launch's tile
counters): every launcher that walks the batch in launches must hand clip 65 535 and its
neighbours to the right place.  The reference has no such limit (its batch is a leading
axis of an Nx tensor, stft.mli:216-218); a clip's result must equal the standalone call on
that clip, bit for bit (stft_grid.ml:180-205).  ``pytest -m gpu``."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BATCH = 66000
PICK = [0, 1, 65534, 65535, 65536, BATCH - 1]


@pytest.fixture(scope="module")
def sb(lib):
    import torch
    assert torch.cuda.is_available()
    return lib


def clips(n, seed=0):
    import torch
    g = torch.Generator("cuda").manual_seed(seed)
    return torch.rand((BATCH, n), device="cuda", generator=g) * 2 - 1


@pytest.mark.parametrize("sr,target", [(48000, 16000), (44100, 22050), (22050, 44100), (44100, 48000),
                                       (48000, 8000), (3, 2)])
def test_resample_batches_past_the_grid_limit(sb, sr, target):
    import torch
    x = clips(700, sr + target)
    cfg = sb.Resample.Config.create(sample_rate=sr, target=target)
    y = sb.Resample.apply(cfg, x)
    assert y.shape == (BATCH, cfg.output_frames(700))
    for b in PICK:
        alone = sb.Resample.apply(cfg, x[b:b + 1])
        assert torch.equal(alone[0], y[b]), (sr, target, b)
    direct = sb.Resample.apply(cfg.set_executor("direct"), x)
    scale = float(y.abs().max())
    assert float((direct - y).abs().max()) <= 1e-5 * scale


@pytest.mark.parametrize("method", ["direct", "ols"])
def test_fir_batches_past_the_grid_limit(sb, method):
    import torch
    x = clips(900, 7)
    fir = sb.Fir.lowpass(k=20, cutoff=0.3)
    y = fir.apply(x, method=method)
    for b in PICK:
        assert torch.equal(fir.apply(x[b:b + 1], method=method)[0], y[b]), (method, b)


def test_stft_family_batches_past_the_grid_limit(sb):
    import torch
    x = clips(3000, 11)
    sc = sb.Stft.Config.create(fft_size=2048, hop=512)
    mc = sb.Mel.Config.create(n_mels=128, sample_rate=22050, fft_size=2048)
    m = sb.mel_spectrogram(sc, mc, x)
    assert m.shape == (BATCH, 128, sb.Stft.frames(sc, 3000))
    p = sb.Stft.power_spectrum(sc, x[:, :2048 + 512])         # 1025 x 6 floats per clip: keep it small
    for b in PICK:
        assert torch.equal(sb.mel_spectrogram(sc, mc, x[b:b + 1])[0], m[b]), b
        assert torch.equal(sb.Stft.power_spectrum(sc, x[b:b + 1, :2048 + 512])[0], p[b]), b
    small = sb.Stft.Config.create(fft_size=64, hop=16)          # the generic (double interior) kernels
    z = sb.Stft.transform(small, x[:, :400])
    back = sb.Stft.invert(small, z, length=400)
    for b in PICK:
        assert torch.equal(sb.Stft.transform(small, x[b:b + 1, :400])[0], z[b]), b
    assert float((back - x[:, :400]).abs().max()) <= 1e-5


def test_empty_and_tiny_inputs_through_the_newer_entry_points(sb):
    """Zero-size leading axes, zero-length signals and signals shorter than anything the block
    executors need (resample_kernel.ml:28-52: output length is ceil(n L / M) also for tiny inputs;
    stft.ml:217-223: no frames when the padded signal is shorter than a frame)."""
    import torch
    sc = sb.Stft.Config.create(fft_size=2048, hop=512)
    mc = sb.Mel.Config.create(n_mels=128, sample_rate=22050, fft_size=2048)
    assert sb.log_mel_spectrogram(sc, mc, np.zeros((0, 5000), np.float32)).shape == (0, 128, 10)
    assert sb.log_mel_spectrogram(sc, mc, np.zeros((3, 0), np.float32)).shape == (3, 128, 0)
    one = sb.log_mel_spectrogram(sc, mc, np.ones((1, 1), np.float32))       # one sample, reflect-extended
    assert one.shape == (1, 128, 1) and np.isfinite(one).all()
    for sr, target in [(44100, 16000), (44100, 48000), (44100, 32000), (44100, 22050), (48000, 8000)]:
        cfg = sb.Resample.Config.create(sample_rate=sr, target=target)
        for n in (0, 1, 2, 17):
            x = torch.ones((2, n), device="cuda")
            y = sb.Resample.apply(cfg, x)
            assert y.shape == (2, -(-n * cfg.l // cfg.m)), (sr, target, n)
            if n:
                ref = sb.Resample.apply(cfg.set_executor("direct"), x)
                assert float((y - ref).abs().max()) <= 1e-5, (sr, target, n)
                cfg.set_executor("planned")
        assert sb.Resample.apply(cfg, torch.ones((0, 300), device="cuda")).shape == (0, cfg.output_frames(300))
    rd = sb.Io.Ingest(channels=2, sample_rate=44100, target=16000, max_block=1000)
    assert tuple(rd.read([]).shape) == (2, 0)
    rd = sb.Io.Ingest(channels=2, sample_rate=44100, target=16000, max_block=1000)
    out = rd.read([np.ones((1, 2), np.float32)])                           # a one-frame file
    assert tuple(out.shape) == (2, 1)
