"""CUDA resampler / FIR against the float64 oracle (``pytest -m gpu``).

Tolerance: BASELINE.json's bar for FIR/resample output, max |got - ref| /
max |ref| <= 1e-5 per signal (the reference's own surfaces differ by up to
32 float32 units of peak, resample_gemm.ml:80-100).  Output lengths and the
phase/index arithmetic are exact integers and must match exactly.
"""
import numpy as np
import pytest

from golden_util import peak_rel_err
from oracle import resample_oracle as R
from test_resample_oracle import RATE_PAIRS, oracle_stages

pytestmark = pytest.mark.gpu

RESAMPLE_TOL = 1e-5


@pytest.fixture(scope="module")
def sb(lib):
    import torch
    assert torch.cuda.is_available()
    return lib


def noise(shape, seed):
    return np.random.default_rng(seed).uniform(-1, 1, shape).astype(np.float32)


@pytest.mark.parametrize("sr,target", RATE_PAIRS)
def test_apply_matches_oracle(sb, sr, target):
    cfg = sb.Resample.Config.create(sample_rate=sr, target=target)
    st = oracle_stages(cfg)
    x = noise((3, 6000), sr * 7 + target)
    got = sb.Resample.apply(cfg, x)
    want = R.apply_plan(x, st, cfg.l, cfg.m)
    assert got.dtype == np.float32 and got.shape == want.shape == (3, cfg.output_frames(6000))
    for c in range(3):
        assert peak_rel_err(got[c], want[c]) <= RESAMPLE_TOL, (sr, target, c)


# resample_kernel.ml:28-52 — output length is ceil(n L / M), also for tiny inputs
@pytest.mark.parametrize("sr,target", [(44100, 32000), (22050, 48000), (44100, 96000),
                                       (11025, 48000), (16000, 44100)])
def test_wide_phase_count_stages_run_on_the_tensor_cores(sb, sr, target):
    """L = 320, 441, 640: the tcgen05 executor cuts the stage into column groups
    of <= 160 phases; results must match the direct kernel and the oracle."""
    cfg = sb.Resample.Config.create(sample_rate=sr, target=target)
    assert any(s["exec"] == "gemm" and s["l"] > 160 for s in cfg.stages()), cfg.pp()
    st = oracle_stages(cfg)
    for n in (1, 500, 40013):
        x = noise((2, n), n + target)
        want = R.apply_plan(x, st, cfg.l, cfg.m)
        planned = sb.Resample.apply(cfg.set_executor("planned"), x)
        direct = sb.Resample.apply(cfg.set_executor("direct"), x)
        assert planned.shape == direct.shape == want.shape
        scale = max(np.abs(want).max(), 1e-3)
        assert np.abs(planned - want).max() / scale <= RESAMPLE_TOL, (sr, target, n)
        assert np.abs(direct - want).max() / scale <= RESAMPLE_TOL, (sr, target, n)


@pytest.mark.parametrize("sr,target,l,m", [(44100, 48000, 160, 147), (48000, 44100, 147, 160),
                                            (44100, 16000, 160, 441), (44100, 22050, 1, 2),
                                            (22050, 44100, 2, 1), (3, 2, 2, 3)])
def test_output_lengths_and_short_inputs(sb, sr, target, l, m):
    cfg = sb.Resample.Config.create(sample_rate=sr, target=target)
    st = oracle_stages(cfg)
    for n in [0, 1, 2, 3, 17, 100, 147, 160, 1000]:
        x = noise((n,), n + 1)
        y = sb.Resample.apply(cfg, x)
        assert y.shape == (-(-n * l // m),)
        if n:
            want = R.apply_plan(x, st, cfg.l, cfg.m)
            assert np.abs(y - want).max() <= RESAMPLE_TOL * max(np.abs(want).max(), 1e-3)


@pytest.mark.parametrize("sr,target,quality,j,expected", [
    (44100, 48000, "high", 294, 320), (48000, 44100, "high", 320, 294),
    (44100, 22050, "high", 500, 250), (22050, 44100, "best", 250, 500),
    (44100, 16000, "high", 882, 320), (48000, 8000, "high", 600, 100),
    (8000, 48000, "high", 100, 600)])
def test_impulse_landing(sb, sr, target, quality, j, expected):
    cfg = sb.Resample.Config.create(sample_rate=sr, target=target, quality=quality)
    x = np.zeros(1000, np.float32)
    x[j] = 1.0
    assert int(np.argmax(np.abs(sb.Resample.apply(cfg, x)))) == expected


def test_channels_never_mix_and_match_standalone_calls(sb):
    """resample_kernel.ml:199-237: leading axes are independent signals."""
    cfg = sb.Resample.Config.create(sample_rate=44100, target=16000)
    x = noise((2, 3, 5000), 3)
    whole = sb.Resample.apply(cfg, x)
    assert whole.shape == (2, 3, cfg.output_frames(5000))
    for i in range(2):
        for j in range(3):
            assert np.array_equal(whole[i, j], sb.Resample.apply(cfg, x[i, j]))


def test_identity_and_flat_api(sb):
    x = noise((2, 777), 1)
    cfg = sb.Resample.Config.create(sample_rate=48000, target=48000)
    assert np.array_equal(sb.Resample.apply(cfg, x), x)
    y = sb.resample(x, sample_rate=44100, target=22050)
    assert y.shape == (2, 389)
    assert np.array_equal(sb.Resample.apply(cfg, x.astype(np.float64)), x.astype(np.float64))
    with pytest.raises(ValueError, match="float32 audio only"):
        sb.Fir(np.ones(3)).apply(x.astype(np.float64))


@pytest.mark.parametrize("sr,target", [(44100, 48000), (44100, 16000), (48000, 8000), (22050, 44100)])
def test_float64_audio_matches_oracle_tightly(sb, sr, target):
    """The reference resamples float32 and float64 (resample.ml:72-84)."""
    cfg = sb.Resample.Config.create(sample_rate=sr, target=target)
    x = np.random.default_rng(sr).uniform(-1, 1, (2, 4000))
    got = sb.Resample.apply(cfg, x)
    want = R.apply_plan(x, oracle_stages(cfg), cfg.l, cfg.m)
    assert got.dtype == np.float64 and got.shape == want.shape
    assert np.abs(got - want).max() / np.abs(want).max() <= 1e-13


def test_device_tensors(sb):
    import torch
    cfg = sb.Resample.Config.create(sample_rate=44100, target=16000)
    x = noise((4, 44100), 9)
    host = sb.Resample.apply(cfg, x)
    dev = sb.Resample.apply(cfg, torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), host)


def test_tone_quality_44k_to_16k(sb):
    """A passband tone comes through at unit gain and an out-of-band tone is
    rejected by the designed stop band (resample_quality.ml Q3/Q4 in spirit)."""
    cfg = sb.Resample.Config.create(sample_rate=44100, target=16000)
    t = np.arange(88200) / 44100.0
    keep = sb.Resample.apply(cfg, np.sin(2 * np.pi * 3000.0 * t).astype(np.float32))
    mid = keep[4000:-4000]
    assert abs(np.abs(mid).max() - 1.0) < 2e-3
    gone = sb.Resample.apply(cfg, np.sin(2 * np.pi * 12000.0 * t).astype(np.float32))
    assert np.abs(gone[4000:-4000]).max() < 1e-5           # < -100 dBFS in float32


# ----------------------------------------------------------------- FIR ----

@pytest.mark.parametrize("k", [1, 7, 50, 255])
def test_fir_direct_matches_oracle(sb, k):
    fir = sb.Fir.lowpass(k=k, cutoff=0.25)
    assert fir.taps.shape == (2 * k + 1,)
    np.testing.assert_allclose(fir.taps, R.design_prototype(1, k, 0.25, R.kaiser_beta(126.0)),
                               rtol=1e-12, atol=1e-18)
    x = noise((2, 2, 9000), k)
    got = fir.apply(x, method="direct")
    want = R.fir_apply(x, fir.taps)
    assert got.shape == x.shape
    assert peak_rel_err(got, want) <= RESAMPLE_TOL
    # the default picks overlap-save from 17 taps up, the direct kernel below
    auto = fir.apply(x)
    assert np.array_equal(auto, fir.apply(x, method="ols" if k >= 8 else "direct"))
    assert peak_rel_err(auto, want) <= RESAMPLE_TOL
    for n in (1, 2, k, 2 * k + 1):
        xs = noise((n,), n)
        assert peak_rel_err(fir.apply(xs, method="direct"), R.fir_apply(xs, fir.taps)) <= RESAMPLE_TOL
        assert peak_rel_err(fir.apply(xs), R.fir_apply(xs, fir.taps)) <= RESAMPLE_TOL


def test_fir_arbitrary_taps_and_errors(sb):
    h = np.array([0.25, 0.5, 0.25])
    x = noise((100,), 4)
    got = sb.Fir(h).apply(x)
    want = np.convolve(x.astype(np.float64), h, mode="same")
    assert peak_rel_err(got, want) <= RESAMPLE_TOL
    with pytest.raises(ValueError, match="taps must be odd"):
        sb.Fir(np.ones(4))


# --------------------------------------------------- overlap-save executors ----

@pytest.mark.parametrize("k", [7, 50, 255, 600])
def test_fir_overlap_save_matches_oracle_and_direct(sb, k):
    fir = sb.Fir.lowpass(k=k, cutoff=0.3)
    x = noise((3, 20011), k + 100)
    want = R.fir_apply(x, fir.taps)
    ols = fir.apply(x, method="ols")
    direct = fir.apply(x, method="direct")
    assert ols.shape == x.shape
    assert peak_rel_err(ols, want) <= RESAMPLE_TOL
    assert peak_rel_err(ols, direct) <= RESAMPLE_TOL
    for n in (1, 2, k, 2 * k + 1, 4097):
        xs = noise((n,), n + 7)
        assert peak_rel_err(fir.apply(xs, method="ols"), R.fir_apply(xs, fir.taps)) <= RESAMPLE_TOL


@pytest.mark.parametrize("sr,target", [(44100, 22050), (22050, 44100), (48000, 8000), (8000, 48000),
                                       (48000, 12000), (12000, 48000), (44100, 88200),
                                       (48000, 16000), (16000, 48000), (96000, 32000)])
def test_planned_ols_stages_match_direct_and_oracle(sb, sr, target):
    """Stages the planner tags for overlap-save (x2, x4, /2, /4) run the FFT
    kernel by default; forcing the direct kernel must give the same signal."""
    cfg = sb.Resample.Config.create(sample_rate=sr, target=target)
    assert any(s["exec"] == "ols" for s in cfg.stages()), cfg.pp()
    st = oracle_stages(cfg)
    for n in (1, 37, 5000, 30000):
        x = noise((2, n), n + sr)
        want = R.apply_plan(x, st, cfg.l, cfg.m)
        planned = sb.Resample.apply(cfg.set_executor("planned"), x)
        direct = sb.Resample.apply(cfg.set_executor("direct"), x)
        assert planned.shape == direct.shape == want.shape
        scale = max(np.abs(want).max(), 1e-3)
        assert np.abs(planned - want).max() / scale <= RESAMPLE_TOL, (sr, target, n)
        assert np.abs(direct - want).max() / scale <= RESAMPLE_TOL, (sr, target, n)


# ---- BASELINE.json configs[3] and configs[4] at their full clip lengths ----------------

def test_config_length_resample_44k1_to_16k(sb):
    """One 30 s clip of configs[3] (1 323 000 samples, 44.1 -> 16 kHz, the tcgen05
    executor) against the float64 oracle at the FIR/resample bar."""
    cfg = sb.Resample.Config.create(sample_rate=44100, target=16000)
    n = 30 * 44100
    x = noise((2, n), 4416)
    got = sb.Resample.apply(cfg, x)
    want = R.apply_plan(x, oracle_stages(cfg), cfg.l, cfg.m)
    assert got.shape == want.shape == (2, 480000)
    for c in range(2):
        assert peak_rel_err(got[c], want[c]) <= RESAMPLE_TOL, c


def test_config5_resample_then_mel_spectrogram(sb):
    """configs[4] composed: Resample.apply(44.1 -> 22.05 kHz) -> Soundml.mel_spectrogram on
    four 10 s clips (441 000 samples) against oracle resampler -> oracle STFT + mel, at the
    spectrogram bar (max |got - ref| / max |ref| <= 1e-4 per clip); soundml.thumper:151,161
    is the reference's closest benchmark of the pair."""
    from oracle import mel_oracle, stft_oracle
    from soundml_b200 import synth
    x = synth.clips_numpy(4, 441000, sample_rate=44100, first_clip=21)
    cfg = sb.Resample.Config.create(sample_rate=44100, target=22050)
    sc = sb.Stft.Config.create(fft_size=2048, hop=512)
    mc = sb.Mel.Config.create(n_mels=128, sample_rate=22050, fft_size=2048)
    mid = sb.Resample.apply(cfg, x)
    assert mid.shape == (4, 220500)
    got = sb.mel_spectrogram(sc, mc, mid)
    want_mid = R.apply_plan(x, oracle_stages(cfg), cfg.l, cfg.m)
    for c in range(4):
        assert peak_rel_err(mid[c], want_mid[c]) <= RESAMPLE_TOL, c
    # the reference's Resample.apply returns float32 audio: the oracle chain rounds there too
    want = mel_oracle.mel_spectrogram(stft_oracle.StftConfig(2048, 512),
                                      mel_oracle.MelConfig(128, 22050, 2048),
                                      want_mid.astype(np.float32))
    assert got.shape == want.shape == (4, 128, 431)
    for c in range(4):
        assert peak_rel_err(got[c], want[c]) <= 1e-4, c


# resample.ml:1608-1698 (gemm_run) — the row form of the tensor-core stage (one persistent
# CTA per SM walking many tiles) against the window form, at a size where every CTA runs
# dozens of tiles back to back: the hand-overs between tiles (accumulators drained and
# zeroed by two warps per lane quarter, the borrowed B stage) only show under load.
@pytest.mark.parametrize("a_tiles", ["default", "shared", "tensor"])
@pytest.mark.parametrize("sr,target", [(44100, 16000), (44100, 48000), (48000, 44100), (44100, 32000)])
def test_row_form_of_the_tensor_core_stage_at_scale(sb, sr, target, a_tiles, monkeypatch):
    import torch
    # where the A tiles live is decided when the plan's device tables are built (first apply)
    monkeypatch.delenv("SMB_ROWS_SMEM_A", raising=False)
    monkeypatch.delenv("SMB_ROWS_TMEM_A", raising=False)
    if a_tiles == "shared":
        monkeypatch.setenv("SMB_ROWS_SMEM_A", "1")
    elif a_tiles == "tensor":
        monkeypatch.setenv("SMB_ROWS_TMEM_A", "1")
    cfg = sb.Resample.Config.create(sample_rate=sr, target=target)
    assert any(s["exec"] == "gemm" for s in cfg.stages()), cfg.pp()
    clips, n = 96, 12 * sr
    x = torch.rand((clips, n), device="cuda", generator=torch.Generator("cuda").manual_seed(sr + target)) * 2 - 1
    rows = torch.empty((clips, cfg.output_frames(n)), device="cuda")
    windows = torch.empty_like(rows)
    monkeypatch.delenv("SMB_GEMM_WINDOWS", raising=False)
    for _ in range(3):                                     # a race does not show every time
        rows.zero_()
        sb.Resample.apply(cfg, x, out=rows)
        torch.cuda.synchronize()
        monkeypatch.setenv("SMB_GEMM_WINDOWS", "1")
        sb.Resample.apply(cfg, x, out=windows)
        torch.cuda.synchronize()
        monkeypatch.delenv("SMB_GEMM_WINDOWS")
        peak = float(windows.abs().max())
        assert float((rows - windows).abs().max()) / peak <= RESAMPLE_TOL, (sr, target)
    # and both against the oracle on a few clips (float64, CPU)
    st = oracle_stages(cfg)
    pick = [0, clips // 2, clips - 1]
    want = R.apply_plan(x[pick].cpu().numpy(), st, cfg.l, cfg.m)
    got = rows[pick].cpu().numpy()
    for c in range(len(pick)):
        assert peak_rel_err(got[c], want[c]) <= RESAMPLE_TOL, (sr, target, c)
