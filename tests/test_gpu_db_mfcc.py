"""dB conversion and MFCC on the GPU against the reference's goldens and the
oracle.  ``pytest -m gpu``."""
import numpy as np
import pytest

from golden_util import MEL_SEED, assert_close, lcg_signal
from oracle import convert_oracle, mel_oracle, stft_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb(lib):
    import torch
    assert torch.cuda.is_available()
    return lib


def test_db_goldens_on_gpu(sb, goldens):
    n = 0
    for key, stem, name, e in goldens.cases("db"):
        p = e["params"]
        x = goldens.arrays[key + "#input"].reshape(e["shape"]).astype(p["dtype"])
        fn = sb.Convert.power_to_db if p["function"] == "power_to_db" else sb.Convert.amplitude_to_db
        got = fn(x, reference=p["reference"], amin=p["amin"], top_db=p["top_db"])
        assert got.dtype == x.dtype and got.shape == x.shape
        tol = 1e-10 if p["dtype"] == "float64" else 1e-4
        assert_close(got, goldens.values(key), tol, tol, key)
        n += 1
    assert n == 18


def test_mfcc_goldens_on_gpu(sb, goldens):
    n = 0
    for key, stem, name, e in goldens.cases("mel", "mfcc"):
        p = e["params"]
        sc = sb.Stft.Config.create(fft_size=p["fft_size"], hop=p["hop"], alignment=p["alignment"])
        mc = sb.Mel.Config.create(n_mels=p["n_mels"], sample_rate=p["sample_rate"],
                                  fft_size=p["fft_size"], f_min=p["f_min"], f_max=p["f_max"],
                                  scale=p["scale"], norm=p["norm"])
        x = lcg_signal(p["length"], MEL_SEED, p["envelope"])
        if p["dtype"] == "float32":
            x = x.astype(np.float32)
        got = sb.mfcc(sc, mc, x, n_mfcc=p["n_mfcc"], lifter=p["lifter"] if p["lifter"] > 0 else None)
        tol = (1e-9, 1e-9) if p["dtype"] == "float64" else (1e-4, 1e-4)
        assert_close(got, goldens.values(key), tol[0], tol[1], key)
        n += 1
    assert n == 9


def test_mfcc_fused_path_batch_and_device(sb):
    """fft 2048 takes the fused mel kernel; the 80 dB clamp is global over the batch."""
    import torch
    from soundml_b200 import synth
    x = synth.clips_numpy(6, 30000, first_clip=3)
    x[4] *= 1e-4                                      # a quiet clip: the global clamp bites
    sc = sb.Stft.Config.create(fft_size=2048, hop=512)
    mc = sb.Mel.Config.create(n_mels=128, sample_rate=22050, fft_size=2048)
    want = convert_oracle.mfcc(stft_oracle.StftConfig(2048, 512),
                               mel_oracle.MelConfig(128, 22050, 2048), x, 20, 22.0)
    got = sb.mfcc(sc, mc, x, n_mfcc=20, lifter=22.0)
    assert got.shape == want.shape == (6, 20, 59)
    assert_close(got, want, 1e-4, 1e-4, "mfcc fused")
    dev = sb.mfcc(sc, mc, torch.from_numpy(x).cuda(), n_mfcc=20, lifter=22.0)
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), got)


def test_db_large_and_errors(sb):
    rng = np.random.default_rng(0)
    x = (rng.uniform(0, 1, (3, 128, 777)) ** 8).astype(np.float32)
    got = sb.Convert.power_to_db(x, top_db=80.0)
    want = convert_oracle.power_to_db(x, top_db=80.0)
    assert_close(got, want, 1e-5, 1e-4, "power_to_db")
    assert got.max() - got.min() <= 80.0 + 1e-3
    assert sb.Convert.power_to_db(np.zeros((0, 4), np.float32)).shape == (0, 4)
    with pytest.raises(ValueError, match=r"Soundml.Convert.power_to_db: amin must be finite and positive"):
        sb.Convert.power_to_db(x, amin=0.0)
    with pytest.raises(ValueError, match=r"Soundml.Convert.amplitude_to_db: top_db must be finite and non-negative"):
        sb.Convert.amplitude_to_db(x, top_db=-1.0)
    sc = sb.Stft.Config.create(fft_size=512, hop=128)
    mc = sb.Mel.Config.create(n_mels=40, sample_rate=22050, fft_size=512)
    with pytest.raises(ValueError, match=r"mfcc: cannot keep 41 cepstral coefficients of 40 mel bands"):
        sb.mfcc(sc, mc, np.zeros(1000, np.float32), n_mfcc=41)
    with pytest.raises(ValueError, match=r"mfcc: cannot lifter with a coefficient of -1"):
        sb.mfcc(sc, mc, np.zeros(1000, np.float32), lifter=-1.0)
    assert sb.mfcc(sc, mc, np.zeros((2, 0), np.float32), n_mfcc=13).shape == (2, 13, 0)


@pytest.mark.parametrize("fft,hop,top_db", [(2048, 512, 80.0), (2048, 512, None), (2048, 500, 60.0),
                                            (1024, 256, 80.0), (400, 160, 80.0)])
def test_log_mel_spectrogram_is_the_composition(sb, fft, hop, top_db):
    """Convert.power_to_db (Soundml.mel_spectrogram ...) in one call: for fft 2048 the
    frame-pair kernel leaves the whole-tensor maximum behind (one decibel pass in place),
    other geometries reduce separately -- the same numbers as the two calls, held to the
    oracle's composition like them."""
    import torch
    from soundml_b200 import synth
    x = synth.clips_numpy(5, 30000, first_clip=11)
    x[2] *= 1e-5                                      # a quiet clip: the global clamp bites
    n_mels = 128 if fft == 2048 else 40
    sc = sb.Stft.Config.create(fft_size=fft, hop=hop)
    mc = sb.Mel.Config.create(n_mels=n_mels, sample_rate=22050, fft_size=fft)
    mel = mel_oracle.mel_spectrogram(stft_oracle.StftConfig(fft, hop),
                                     mel_oracle.MelConfig(n_mels, 22050, fft), x).astype(np.float32)
    want = convert_oracle.power_to_db(mel, top_db=top_db)
    got = sb.log_mel_spectrogram(sc, mc, x, top_db=top_db)
    assert got.shape == want.shape and got.dtype == np.float32
    assert_close(got, want, 1e-5, 2e-4, "log mel")        # decibels: 2e-4 dB absolute
    two = sb.Convert.power_to_db(sb.mel_spectrogram(sc, mc, x), top_db=top_db)
    assert np.abs(got - two).max() <= 1e-5                # one float32 rounding of the clamp floor apart at most
    before = sb.kernel_launch_count()
    dev = sb.log_mel_spectrogram(sc, mc, torch.from_numpy(x).cuda(), top_db=top_db)
    torch.cuda.synchronize()
    if fft == 2048 and hop == 512:
        assert sb.kernel_launch_count() - before == 2     # mel kernel (+ maximum) and one decibel pass
    assert np.array_equal(dev.cpu().numpy(), got)
    with pytest.raises(ValueError, match=r"Soundml.Convert.power_to_db: top_db must be finite and non-negative"):
        sb.log_mel_spectrogram(sc, mc, x, top_db=-3.0)


def test_mfcc_host_batches_go_up_in_slices(sb):
    """Host audio larger than one 64 MB slice: the mel values of every slice stay on the device
    until the whole-tensor maximum is known, so the 80 dB clamp is still global over the batch --
    bit for bit the device-memory call (which takes the batch at once)."""
    import torch
    from soundml_b200 import synth
    x = synth.clips_numpy(40, 500000, first_clip=11)           # 80 MB: two slices of 33 and 7 clips
    x[2] *= 1e-4                                                # quiet clip in the first slice ...
    x[37] *= 4.0                                                # ... the maximum in the second
    sc = sb.Stft.Config.create(fft_size=2048, hop=512)
    mc = sb.Mel.Config.create(n_mels=128, sample_rate=22050, fft_size=2048)
    host = sb.mfcc(sc, mc, x, n_mfcc=13)
    dev = sb.mfcc(sc, mc, torch.from_numpy(x).cuda(), n_mfcc=13)
    torch.cuda.synchronize()
    assert host.shape == (40, 13, 977)
    assert np.array_equal(dev.cpu().numpy(), host)
    # the clamp really is global: clip 2's floor sits 80 dB under clip 37's peak, not its own
    alone = sb.mfcc(sc, mc, x[2:3], n_mfcc=13)
    assert not np.array_equal(alone[0], host[2])
