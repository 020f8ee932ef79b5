"""Small tensor-core-kernel workload for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import soundml_b200 as sb
from soundml_b200 import synth

x = synth.clips_numpy(3, 9000)
mc = sb.Mel.Config.create(n_mels=128, sample_rate=22050, fft_size=2048)
for path in ("tensor", "fast"):
    for hop in (512, 333):
        c = sb.Stft.Config.create(fft_size=2048, hop=hop).set_path(path)
        m = sb.mel_spectrogram(c, mc, x)
        p = sb.Stft.power_spectrum(c, x)
        z = sb.Stft.transform(c, x)
        print(path, hop, m.shape, float(np.abs(m).max()), float(np.abs(p).max()), float(np.abs(z).max()))

# bin-major write-out with the sector carry: enough tiles (> 2 * 148 groups) that a
# group walks consecutive tiles of a clip and crosses clip boundaries
x = synth.clips_numpy(40, 100 * 512 + 3)
c = sb.Stft.Config.create(fft_size=2048, hop=512).set_path("fast")
p = sb.Stft.power_spectrum(c, x)
z = sb.Stft.transform(c, x)
print("carry", p.shape, float(np.abs(p).max()), float(np.abs(z).max()))
