"""One signal of more than 2^31 samples (8.8 GB of float32): every sample and frame index past
the 32-bit range must still land in the right place.  The reference's indices are OCaml ints
(63 bits; `output_frames` guards n * L against overflow, resample.ml:1038-1051).  Checked on
windows near the end of the signal against the same operation on a cropped segment -- bit for bit
where an output's arithmetic does not depend on the call it sits in (direct executors, frames),
within the resampler's bar where a block grid shifts (overlap-save, tensor-core tiles).
``pytest -m gpu``; skipped when the device has less than 60 GB free."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 2_200_000_000                    # > 2^31 samples


@pytest.fixture(scope="module")
def sb(lib):
    import torch
    assert torch.cuda.is_available()
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * 2 ** 30:
        pytest.skip("needs 60 GB of free device memory")
    return lib


@pytest.fixture(scope="module")
def signal(sb):
    import torch
    g = torch.Generator("cuda").manual_seed(3)
    x = torch.empty((1, N), device="cuda")
    step = 1 << 28
    for i in range(0, N, step):                      # (one rand call of 2.2e9 elements is not needed)
        x[0, i:i + step] = torch.rand((min(step, N - i),), device="cuda", generator=g) * 2 - 1
    yield x
    del x
    torch.cuda.empty_cache()


def test_fir_past_2_31_samples(sb, signal):
    import torch
    k = 20
    fir = sb.Fir.lowpass(k=k, cutoff=0.3)
    y = fir.apply(signal, method="direct")
    assert y.shape == (1, N)
    for i0 in (0, (1 << 31) - 500, (1 << 31) + 12345, N - 4000):
        i1 = min(N, i0 + 3000)
        lo, hi = max(0, i0 - k), min(N, i1 + k)
        seg = fir.apply(signal[:, lo:hi].contiguous(), method="direct")
        assert torch.equal(seg[0, i0 - lo:i0 - lo + (i1 - i0)], y[0, i0:i1]), i0
    del y
    z = fir.apply(signal, method="ols")
    i0 = (1 << 31) + 777
    seg = fir.apply(signal[:, i0 - 5000:i0 + 8000].contiguous(), method="direct")
    assert float((seg[0, 5000:8000 + 5000 - 100] - z[0, i0:i0 + 8000 - 100]).abs().max()) <= 1e-5
    del z
    torch.cuda.empty_cache()


@pytest.mark.parametrize("sr,target", [(44100, 22050), (3, 2), (44100, 48000)])
def test_resample_past_2_31_samples(sb, signal, sr, target):
    import torch
    cfg = sb.Resample.Config.create(sample_rate=sr, target=target)
    y = sb.Resample.apply(cfg, signal)
    assert y.shape == (1, cfg.output_frames(N)) and cfg.output_frames(N) == -(-N * cfg.l // cfg.m)
    # a segment that starts on a whole number of input cycles maps onto output o0 = s0 L / M
    reach = 4 * cfg.latency + 64
    for s_mid in ((1 << 31) + 5000, N - 300000):
        s0 = (s_mid // cfg.m) * cfg.m
        seg = sb.Resample.apply(cfg, signal[:, s0:s0 + 200000].contiguous())
        o0 = s0 // cfg.m * cfg.l
        skip = reach * cfg.l // cfg.m + 8                      # the segment's own zero-extended edges
        a = seg[0, skip:seg.shape[1] - skip]
        b = y[0, o0 + skip:o0 + seg.shape[1] - skip]
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()), (sr, target, s_mid)
    del y
    torch.cuda.empty_cache()


def test_mel_spectrogram_past_2_31_samples(sb, signal):
    import torch
    sc = sb.Stft.Config.create(fft_size=2048, hop=512)
    mc = sb.Mel.Config.create(n_mels=128, sample_rate=22050, fft_size=2048)
    m = sb.mel_spectrogram(sc, mc, signal)
    frames = sb.Stft.frames(sc, N)
    assert m.shape == (1, 128, frames) and frames == 1 + N // 512
    left = sb.Stft.Config.create(fft_size=2048, hop=512, alignment="left")
    for p0 in ((1 << 31) // 512 - 3, (1 << 31) // 512 + 1000, frames - 600):
        s0 = p0 * 512 - 1024                                    # centred frame p covers source [p hop - 1024, + 2048)
        seg = sb.mel_spectrogram(left, mc, signal[:, s0:s0 + 2048 + 199 * 512].contiguous())
        assert seg.shape == (1, 128, 200)
        assert torch.equal(seg[0], m[0, :, p0:p0 + 200]), p0
    del m
    torch.cuda.empty_cache()
