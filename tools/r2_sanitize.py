"""Small round-2 workloads for compute-sanitizer (memcheck / racecheck / synccheck): the
frame-pair kernel, the row-form tensor-core resampler stage in its three shapes (TMEM A
tiles, shared-memory A tiles, split accumulators with two issuers) and the reader.
compute-sanitizer --tool memcheck python tools/r2_sanitize.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
import soundml_b200 as sb
from soundml_b200 import synth

x = synth.clips_numpy(3, 30000)
mc = sb.Mel.Config.create(n_mels=128, sample_rate=22050, fft_size=2048)
for hop in (512, 256, 334):
    c = sb.Stft.Config.create(fft_size=2048, hop=hop).set_path("pair")
    m = sb.mel_spectrogram(c, mc, x)
    print("pair", hop, m.shape, float(np.abs(m).max()))
rng = np.random.default_rng(0)
for sr, target in ((44100, 48000), (48000, 44100), (44100, 16000), (44100, 32000), (16000, 44100)):
    cfg = sb.Resample.Config.create(sample_rate=sr, target=target)
    y = sb.Resample.apply(cfg, rng.uniform(-1, 1, (2, 40000)).astype(np.float32))
    print("rows", sr, target, y.shape, float(np.abs(y).max()))
rd = sb.Io.Ingest(channels=2, sample_rate=44100, target=16000, mode="mono", max_block=4096)
sig = rng.uniform(-1, 1, (20000, 2)).astype(np.float32)
out = rd.read(sig[i:i + 4096] for i in range(0, 20000, 4096))
print("reader", tuple(out.shape))
