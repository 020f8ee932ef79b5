import torch, time, sys
sys.path.insert(0, '.')
import soundml_b200 as sb
from soundml_b200 import synth
x = synth.clips_torch(256, 220500, device="cuda")
for fft, hop, nm in ((2048, 512, 128), (1024, 256, 80), (512, 128, 80), (400, 160, 80)):
    sc = sb.Stft.Config.create(fft_size=fft, hop=hop)
    mc = sb.Mel.Config.create(n_mels=nm, sample_rate=22050, fft_size=fft)
    out = sb.mel_spectrogram(sc, mc, x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        sb.mel_spectrogram(sc, mc, x, out=out)
    e1.record(); torch.cuda.synchronize()
    print(fft, hop, nm, "mel_spectrogram 256 clips: %.2f ms" % (e0.elapsed_time(e1) / 3))
