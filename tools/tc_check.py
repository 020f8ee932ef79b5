"""Development check of the tensor-core fft-2048 kernel against the oracle and the
CUDA-core fused kernel (run on the GPU box)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import soundml_b200 as sb
from soundml_b200 import synth
from oracle import stft_oracle, mel_oracle


def perr(a, b):
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))) / max(np.max(np.abs(b)), 1e-300))


x = synth.clips_numpy(3, 220500)
o = stft_oracle.StftConfig(2048, 512)
mo = mel_oracle.MelConfig(128, 22050, 2048)
mc = sb.Mel.Config.create(n_mels=128, sample_rate=22050, fft_size=2048)
refz = stft_oracle.transform(o, x)
refp = stft_oracle.power_spectrum(o, x)
refm = mel_oracle.apply(mo, refp)
for path in ("fast", "tensor"):
    c = sb.Stft.Config.create(fft_size=2048, hop=512).set_path(path)
    z = sb.Stft.transform(c, x)
    p = sb.Stft.power_spectrum(c, x)
    m = sb.mel_spectrogram(c, mc, x)
    print(path, "complex", perr(z.view(np.float32), refz.view(np.float32)), "power", perr(p, refp),
          "mel", perr(m, refm), flush=True)
# scale robustness
rng = np.random.default_rng(1)
base = rng.uniform(-1, 1, 20000).astype(np.float32)
for amp in (1e-30, 1e-12, 1e-3, 1.0, 37.0, 32768.0, 1e12):
    xs = (base * np.float32(amp)).astype(np.float32)
    c = sb.Stft.Config.create(fft_size=2048, hop=512).set_path("tensor")
    z = sb.Stft.transform(c, xs)
    rz = stft_oracle.transform(o, xs)
    print("amp", amp, "complex err", perr(z.view(np.float32), rz.view(np.float32)), flush=True)
xs = np.zeros(5000, np.float32)
c = sb.Stft.Config.create(fft_size=2048, hop=512).set_path("tensor")
print("zeros ->", float(np.abs(sb.Stft.power_spectrum(c, xs)).max()))
# odd hops / partial tiles
for hop in (500, 333, 1, 600):
    n = 2600 if hop == 1 else 9000
    xs = synth.clips_numpy(2, n)
    oo = stft_oracle.StftConfig(2048, hop)
    c = sb.Stft.Config.create(fft_size=2048, hop=hop).set_path("tensor")
    print("hop", hop, perr(sb.Stft.power_spectrum(c, xs), stft_oracle.power_spectrum(oo, xs)), flush=True)
