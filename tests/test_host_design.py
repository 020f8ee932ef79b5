"""Host-side design code of the library (no GPU): the C ABI loads and exports
every declared symbol, and windows / mel weights / frame grid / boundary
indices / resampler plans match the reference's goldens and pinned plans."""
import ctypes
import os
import re

import numpy as np
import pytest

from golden_util import (F64_ATOL, F64_RTOL, WINDOW_ATOL, WINDOW_RTOL, assert_close,
                         window_spec)
from oracle import mel_oracle, stft_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "soundml_b200.h")).read()
    declared = set(re.findall(r"\b(smb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 50
    cdll = ctypes.CDLL(lib._lib.LIB_PATH)
    missing = [name for name in sorted(declared) if not hasattr(cdll, name)]
    assert not missing, missing
    # and the Python binding declares a signature for each of them
    assert declared == set(lib._lib.SIGNATURES), declared ^ set(lib._lib.SIGNATURES)


def test_window_goldens_through_c_abi(lib, goldens):
    for key, stem, name, e in goldens.cases("window"):
        if stem == "cola":
            continue
        p = e["params"]
        got = lib.Window.make(window_spec(p), p["n"], periodic=p["periodic"])
        assert_close(got, goldens.values(key), WINDOW_RTOL, WINDOW_ATOL, key)


def test_window_errors_use_reference_wording(lib):
    with pytest.raises(ValueError, match=r"make: cannot make a 0-point window \(length must be at least 1\)"):
        lib.Window.make("hann", 0)
    with pytest.raises(ValueError, match=r"make: cannot use a kaiser window with beta -1 "):
        lib.Window.make(("kaiser", -1.0), 8)
    with pytest.raises(ValueError, match=r"standard deviation 0 "):
        lib.Window.make(("gaussian", 0.0), 8)
    with pytest.raises(ValueError, match=r"taper 1.5 \(taper must lie in \[0, 1\]\)"):
        lib.Window.make(("tukey", 1.5), 8)


def test_stft_config_matches_oracle(lib):
    for fft, hop, wl, scale in [(16, 4, None, "none"), (32, 8, 20, "none"), (64, 16, 40, "magnitude"),
                                (2048, 512, 1200, "psd"), (31, 5, None, "none")]:
        c = lib.Stft.Config.create(fft_size=fft, hop=hop, win_length=wl, scale=scale)
        o = stft_oracle.StftConfig(fft, hop, wl, scale=scale)
        np.testing.assert_allclose(c.analysis_window, o.analysis_window, rtol=1e-15, atol=1e-17)
        assert c.bins == o.bins
    c = lib.Stft.Config.create(fft_size=2048)
    assert c.hop == 512                                   # default hop = fft/4 (stft.ml:75)
    assert lib.Stft.Config.create(fft_size=3).hop == 1


def test_stft_config_errors(lib):
    S = lib.Stft
    with pytest.raises(ValueError, match=r"create: cannot use an FFT of size 0 \(fft_size must be at least 1\)"):
        S.Config.create(fft_size=0)
    with pytest.raises(ValueError, match=r"create: cannot use a 17-point window with an FFT of size 16"):
        S.Config.create(fft_size=16, win_length=17)
    with pytest.raises(ValueError, match=r"create: cannot advance frames by 0 samples \(hop must be at least 1\)"):
        S.Config.create(fft_size=16, hop=0)
    with pytest.raises(ValueError, match=r"create: cannot use a kaiser window with beta -2"):
        S.Config.create(fft_size=16, window=("kaiser", -2.0))
    c = S.Config.create(fft_size=16, hop=4)
    with pytest.raises(ValueError, match=r"frames: cannot analyse a signal of length -1"):
        S.frames(c, -1)


@pytest.mark.parametrize("alignment", ["centered", "left", "right"])
def test_frame_grid_bit_matches_oracle(lib, alignment):
    for fft, hop in [(16, 4), (16, 3), (16, 5), (32, 7), (2048, 512), (2048, 500), (16, 40)]:
        c = lib.Stft.Config.create(fft_size=fft, hop=hop, alignment=alignment)
        o = stft_oracle.StftConfig(fft, hop, alignment=alignment)
        for n in [0, 1, 2, 7, 15, 16, 17, 61, 127, 128, 1000, 220500]:
            assert lib.Stft.frames(c, n) == stft_oracle.frames(o, n), (fft, hop, n)


@pytest.mark.parametrize("pad", ["reflect", "edge", ("constant", 0.25)])
def test_boundary_indices_bit_match_oracle(lib, pad):
    name = pad if isinstance(pad, str) else pad[0]
    for alignment in ("centered", "right", "left"):
        for fft in (16, 64, 2048):
            c = lib.Stft.Config.create(fft_size=fft, hop=max(1, fft // 4), alignment=alignment, pad=pad)
            o = stft_oracle.StftConfig(fft, max(1, fft // 4), alignment=alignment, pad=name)
            for n in (1, 2, 3, 7, 9, 100, 5000):      # includes n <= fft/2: multi-reflection
                assert np.array_equal(c.source_indices(n), stft_oracle.source_indices(o, n)), (alignment, fft, n)


def test_mel_filterbank_goldens_through_c_abi(lib, goldens):
    for key, stem, name, e in goldens.cases("mel", "filterbank"):
        p = e["params"]
        c = lib.Mel.Config.create(n_mels=p["n_mels"], sample_rate=p["sample_rate"],
                                  fft_size=p["fft_size"], f_min=p["f_min"], f_max=p["f_max"],
                                  scale=p["scale"], norm=p["norm"])
        assert_close(lib.Mel.filterbank(c), goldens.values(key), F64_RTOL, F64_ATOL, key)


def test_mel_weights_match_oracle_at_bench_geometry(lib):
    c = lib.Mel.Config.create(n_mels=128, sample_rate=22050, fft_size=2048)
    o = mel_oracle.MelConfig(128, 22050, 2048)
    np.testing.assert_allclose(lib.Mel.filterbank(c), o.weights, rtol=1e-9, atol=1e-12)
    # each bin feeds at most two triangles
    assert ((o.weights != 0).sum(axis=0) <= 2).all()


def test_mel_config_errors(lib):
    M = lib.Mel
    with pytest.raises(ValueError, match=r"create: cannot build 0 mel bands"):
        M.Config.create(n_mels=0, sample_rate=22050, fft_size=512)
    with pytest.raises(ValueError, match=r"f_max must not exceed the Nyquist frequency 11025"):
        M.Config.create(n_mels=8, sample_rate=22050, fft_size=512, f_max=12000.0)
    with pytest.raises(ValueError, match=r"at least one filter spans no FFT bin"):
        M.Config.create(n_mels=128, sample_rate=22050, fft_size=64)
    with pytest.raises(ValueError, match=r"f_max must be finite and greater than f_min"):
        M.Config.create(n_mels=8, sample_rate=22050, fft_size=512, f_min=500.0, f_max=400.0)


# soundml/test/resample/resample_config.ml:115-146 — the plans of record
PINNED_PLANS = [
    (44100, 48000, "resample(44100 -> 48000 Hz, quality=high, L/M=160/147, taps=191(gemm), latency=95)"),
    (48000, 44100, "resample(48000 -> 44100 Hz, quality=high, L/M=147/160, taps=207(gemm), latency=103)"),
    (44100, 16000, "resample(44100 -> 16000 Hz, quality=high, L/M=160/441, taps=523(gemm), latency=261)"),
    (16000, 44100, "resample(16000 -> 44100 Hz, quality=high, L/M=441/160, taps=191(gemm), latency=95)"),
    (8000, 48000, "resample(8000 -> 48000 Hz, quality=high, L/M=6/1, stages=2/1:401(ols,N=2048) >> 3/1:21, latency=105)"),
    (48000, 8000, "resample(48000 -> 8000 Hz, quality=high, L/M=1/6, stages=1/3:51 >> 1/2:399(ols,N=2048), latency=622)"),
    (44100, 22050, "resample(44100 -> 22050 Hz, quality=high, L/M=1/2, taps=381(ols,N=2048), latency=190)"),
    (22050, 44100, "resample(22050 -> 44100 Hz, quality=high, L/M=2/1, taps=381(ols,N=2048), latency=95)"),
]


@pytest.mark.parametrize("sr,target,expected", PINNED_PLANS)
def test_resample_plans_of_record(lib, sr, target, expected):
    assert lib.Resample.Config.create(sample_rate=sr, target=target).pp() == expected


def test_resample_identity_and_output_length(lib):
    c = lib.Resample.Config.create(sample_rate=48000, target=48000)
    assert c.latency == 0 and (c.l, c.m) == (1, 1)
    assert c.pp() == "resample(48000 Hz, identity)"
    c = lib.Resample.Config.create(sample_rate=44100, target=16000)
    for n in (0, 1, 2, 440, 441, 442, 1323000):
        assert c.output_frames(n) == -(-n * 160 // 441)
    assert c.output_frames(1323000) == 480000


def test_resample_config_errors(lib):
    R = lib.Resample
    with pytest.raises(ValueError, match=r"create: cannot resample from 0 Hz"):
        R.Config.create(sample_rate=0, target=8000)
    with pytest.raises(ValueError, match=r"create: cannot resample to -1 Hz"):
        R.Config.create(sample_rate=8000, target=-1)
    with pytest.raises(ValueError, match=r"attenuation must be finite, in \[40, 200\]"):
        R.Config.create(sample_rate=8000, target=16000, quality=("custom", 30.0, 0.9))
    with pytest.raises(ValueError, match=r"passband must be finite, in \[0.5, 0.99\]"):
        R.Config.create(sample_rate=8000, target=16000, quality=("custom", 100.0, 0.995))
    with pytest.raises(ValueError, match=r"no two-stage split brings it under"):
        R.Config.create(sample_rate=44100, target=44101)
    c = R.Config.create(sample_rate=44100, target=16000)
    with pytest.raises(ValueError, match=r"output_frames: cannot resample a signal of length -1"):
        c.output_frames(-1)


def test_compute_without_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    c = lib.Stft.Config.create(fft_size=16, hop=4)
    with pytest.raises(lib.SoundmlError, match="no CPU fallback"):
        lib.Stft.power_spectrum(c, np.zeros(64, np.float32))
