"""The oracle against every golden vector the reference commits for this path
(SURVEY.md 8c).  CPU only.  Tolerances are the reference's own."""
import numpy as np
import pytest

from golden_util import (F32_ATOL, F32_RTOL, F64_ATOL, F64_RTOL, MEL_SEED, STFT_SEED,
                         WINDOW_ATOL, WINDOW_RTOL, assert_close, griffin_lim_magnitudes,
                         istft_spectrum, lcg_signal, window_spec)
from oracle import istft_oracle, mel_oracle, stft_oracle, window_oracle


def _param(spec):
    return spec[1] if not isinstance(spec, str) else 0.0


def _name(spec):
    return spec if isinstance(spec, str) else spec[0]


def test_window_goldens(goldens):
    n = 0
    for key, stem, name, e in goldens.cases("window"):
        p = e["params"]
        spec = window_spec(p)
        if stem == "cola":
            assert window_oracle.cola(_name(spec), p["length"], p["hop"], _param(spec)) == e["expected"], key
        else:
            got = window_oracle.make(_name(spec), p["n"], p["periodic"], _param(spec))
            assert_close(got, goldens.values(key), WINDOW_RTOL, WINDOW_ATOL, key)
        n += 1
    assert n == 325


def test_stft_spectrum_goldens(goldens):
    n = 0
    for key, stem, name, e in goldens.cases("stft"):
        if stem == "coordinates":
            continue
        p = e["params"]
        c = stft_oracle.StftConfig(p["fft_size"], p["hop"], p["win_length"], alignment=p["alignment"])
        x = lcg_signal(p["length"], STFT_SEED)
        if p["dtype"] == "float32":
            x = x.astype(np.float32)
        if p["kind"] in ("magnitude", "power"):
            got = stft_oracle.power_spectrum(c, x, 1.0 if p["kind"] == "magnitude" else 2.0)
        else:
            z = stft_oracle.transform(c, x)
            got = z.real if p["kind"] == "real" else z.imag
        rtol, atol = (F64_RTOL, F64_ATOL) if p["dtype"] == "float64" else (F32_RTOL, F32_ATOL)
        assert_close(got, goldens.values(key), rtol, atol, key)
        n += 1
    assert n == 66


def test_stft_coordinate_goldens(goldens):
    for key, stem, name, e in goldens.cases("stft", "coordinates"):
        p = e["params"]
        if p["kind"] == "frequencies":
            c = stft_oracle.StftConfig(p["fft_size"], p["hop"])
            got = stft_oracle.frequencies(c, p["sample_rate"])
        else:
            c = stft_oracle.StftConfig(p["fft_size"], p["hop"], alignment=p["alignment"])
            got = stft_oracle.times(c, p["sample_rate"], p["length"])
        assert_close(got, goldens.values(key), 1e-12, 1e-15, key)


def test_mel_filterbank_goldens(goldens):
    for key, stem, name, e in goldens.cases("mel", "filterbank"):
        p = e["params"]
        c = mel_oracle.MelConfig(p["n_mels"], p["sample_rate"], p["fft_size"], p["f_min"],
                                 p["f_max"], p["scale"], p["norm"])
        assert_close(c.weights, goldens.values(key), F64_RTOL, F64_ATOL, key)


def test_mel_spectrogram_goldens(goldens):
    for key, stem, name, e in goldens.cases("mel", "mel_spectrogram"):
        p = e["params"]
        sc = stft_oracle.StftConfig(p["fft_size"], p["hop"], alignment=p["alignment"])
        mc = mel_oracle.MelConfig(p["n_mels"], p["sample_rate"], p["fft_size"], p["f_min"],
                                  p["f_max"], p["scale"], p["norm"])
        x = lcg_signal(p["length"], MEL_SEED, p["envelope"])
        if p["dtype"] == "float32":
            x = x.astype(np.float32)
        got = mel_oracle.mel_spectrogram(sc, mc, x, p["power"])
        rtol, atol = (F64_RTOL, F64_ATOL) if p["dtype"] == "float64" else (F32_RTOL, F32_ATOL)
        assert_close(got, goldens.values(key), rtol, atol, key)


# frame-count table of soundml/test/stft/stft_grid.ml:125-143
@pytest.mark.parametrize("fft,hop,alignment,expected", [
    (16, 4, "centered", {0: 0, 1: 1, 2: 1, 7: 2, 16: 5, 17: 5, 61: 16}),
    (16, 4, "left", {0: 0, 1: 0, 2: 0, 7: 0, 16: 1, 17: 1, 61: 12}),
    (16, 4, "right", {0: 0, 1: 1, 2: 1, 7: 2, 16: 4, 17: 5, 61: 16}),
])
def test_frame_counts(fft, hop, alignment, expected):
    c = stft_oracle.StftConfig(fft, hop, alignment=alignment)
    for n, want in expected.items():
        assert stft_oracle.frames(c, n) == want, (alignment, n)


def test_reflect_index_matches_numpy_pad():
    for n in (1, 2, 3, 5, 17):
        x = np.arange(n, dtype=np.float64)
        c = stft_oracle.StftConfig(16, 4)
        padded = stft_oracle.pad_signal(c, x)
        if n > 1:
            want = np.pad(x, (8, 8), mode="reflect") if n > 8 else None
            if want is not None:
                assert np.array_equal(padded, want)
        assert padded.shape[-1] == n + 16


def test_db_goldens(goldens):
    """convert.ml to_db against the reference's librosa vectors
    (soundml/test/db/test_golden.ml:52-57: f32 1e-4/1e-4, f64 1e-10/1e-10)."""
    from oracle import convert_oracle
    n = 0
    for key, stem, name, e in goldens.cases("db"):
        p = e["params"]
        x = goldens.arrays[key + "#input"].reshape(e["shape"]).astype(p["dtype"])
        fn = convert_oracle.power_to_db if p["function"] == "power_to_db" else convert_oracle.amplitude_to_db
        got = fn(x, p["reference"], p["amin"], p["top_db"])
        tol = 1e-10 if p["dtype"] == "float64" else 1e-4
        assert_close(got, goldens.values(key), tol, tol, key)
        n += 1
    assert n == 18


def test_mfcc_goldens(goldens):
    """soundml.ml mfcc (mel_goldens.ml:134-157: f64 rtol 1e-9 / atol 1e-9, f32 1e-4 / 1e-4)."""
    from oracle import convert_oracle
    n = 0
    for key, stem, name, e in goldens.cases("mel", "mfcc"):
        p = e["params"]
        sc = stft_oracle.StftConfig(p["fft_size"], p["hop"], alignment=p["alignment"])
        mc = mel_oracle.MelConfig(p["n_mels"], p["sample_rate"], p["fft_size"], p["f_min"],
                                  p["f_max"], p["scale"], p["norm"])
        x = lcg_signal(p["length"], MEL_SEED, p["envelope"])
        if p["dtype"] == "float32":
            x = x.astype(np.float32)
        got = convert_oracle.mfcc(sc, mc, x, p["n_mfcc"], p["lifter"] if p["lifter"] > 0 else None)
        tol = (1e-9, 1e-9) if p["dtype"] == "float64" else (1e-4, 1e-4)
        assert_close(got, goldens.values(key), tol[0], tol[1], key)
        n += 1
    assert n == 9


def test_istft_goldens(goldens):
    """Stft.invert against librosa's least-squares synthesis
    (soundml/test/istft/istft_goldens.ml:62-88), incl. explicit lengths."""
    n = 0
    for key, stem, name, e in goldens.cases("istft"):
        p = e["params"]
        c = stft_oracle.StftConfig(p["fft_size"], hop=p["hop"], win_length=p["win_length"],
                                   alignment=p["alignment"], pad="constant", pad_value=0.0)
        z = istft_spectrum(p["fft_size"], p["frames"], p["dtype"])
        f32 = p["dtype"] == "float32"
        got = istft_oracle.invert(c, z, length=p.get("length"),
                                  dtype=np.float32 if f32 else np.float64)
        assert_close(got, goldens.values(key), F32_RTOL if f32 else F64_RTOL,
                     F32_ATOL if f32 else F64_ATOL, key)
        n += 1
    assert n == 88


def test_istft_round_trip_and_errors():
    """invert(transform(x)) == x wherever the envelope is conditioned
    (istft_law.ml), and the reference's precondition messages."""
    x = lcg_signal(3000, STFT_SEED)
    for alignment in ("centered", "left"):
        c = stft_oracle.StftConfig(256, hop=64, alignment=alignment)
        z = stft_oracle.transform(c, x)
        y = istft_oracle.invert(c, z, length=len(x))
        lo, hi = (0, len(x)) if alignment == "centered" else (256, len(x) - 256)
        np.testing.assert_allclose(y[lo:hi], x[lo:hi], rtol=0, atol=1e-12)
    with pytest.raises(ValueError, match="overlap-added squared"):
        istft_oracle.invert(stft_oracle.StftConfig(64, hop=65), np.zeros((33, 2), complex))
    with pytest.raises(ValueError, match="frequency bins"):
        istft_oracle.invert(stft_oracle.StftConfig(64, hop=16), np.zeros((32, 2), complex))
    with pytest.raises(ValueError, match="length must be non-negative"):
        istft_oracle.invert(stft_oracle.StftConfig(64, hop=16), np.zeros((33, 2), complex), length=-1)
    c = stft_oracle.StftConfig(64, hop=16)
    assert istft_oracle.output_length(c, 0) == 0 and istft_oracle.output_length(c, 9) == 128
    assert istft_oracle.invert(c, np.zeros((33, 0), complex)).shape == (0,)


def test_griffin_lim_goldens(goldens):
    """Stft.griffin_lim against librosa (soundml/test/istft/gl_goldens.ml:49-78):
    1 to 32 iterations, momentum 0 / 0.5 / 0.99, three geometries."""
    n = 0
    for key, stem, name, e in goldens.cases("griffinlim"):
        p = e["params"]
        c = stft_oracle.StftConfig(p["fft_size"], hop=p["hop"], win_length=p["win_length"],
                                   alignment=p["alignment"], pad="constant", pad_value=0.0)
        mags = griffin_lim_magnitudes(p["fft_size"], p["frames"], p["dtype"])
        f32 = p["dtype"] == "float32"
        got = istft_oracle.griffin_lim(c, mags, n_iter=p["n_iter"], momentum=p["momentum"])
        assert got.dtype == mags.dtype
        assert_close(got, goldens.values(key), F32_RTOL if f32 else F64_RTOL,
                     F32_ATOL if f32 else F64_ATOL, key)
        n += 1
    assert n == 42


def test_io_layout_oracle_and_block_sizing():
    """soundml-io layout pass restated (soundml_io_stubs.c:832-872): planar scatter and
    the in-order downmix; decode_block_frames (soundml_io.ml:532-536) known answers."""
    import numpy as np
    from oracle import io_oracle
    b = np.array([[1.0, 2.0, 4.0], [0.5, -0.5, 1.0]], dtype=np.float32)
    assert np.array_equal(io_oracle.layout(b, "planar"), b.T)
    third = np.float32(1) / np.float32(3)
    want = np.array([[np.float32(7.0) * third, np.float32(1.0) * third]], dtype=np.float32)
    assert np.array_equal(io_oracle.layout(b, "mono"), want)
    s = np.array([[0.25, 0.5]], dtype=np.float64)
    assert io_oracle.layout(s, "mono")[0, 0] == 0.375
    assert io_oracle.decode_block_frames(2, 4) == 524288
    assert io_oracle.decode_block_frames(1, 4) == 1048576
    assert io_oracle.decode_block_frames(64, 8) == 8192
    assert io_oracle.decode_block_frames(2, 4, 100) == 4096
    assert io_oracle.decode_block_frames(2, 4, 50000) == 50000
