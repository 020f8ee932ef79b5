#!/usr/bin/env python
"""Frame-pair kernel, headline workload: the full kernel, the kernel without its mel
phase (SMB_PAIR_SKIP_MEL=1: split and power rows, nothing written) and the FFT-only
ceiling (stage + window + 2 x fft32 + transposition), CUDA events, same process."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import soundml_b200 as sb            # noqa: E402
from soundml_b200 import _lib, synth  # noqa: E402

B, N = 1024, 220500
x = synth.clips_torch(B, N, "cuda:0")
sc = sb.Stft.Config.create(fft_size=2048, hop=512).set_path("pair")
mc = sb.Mel.Config.create(n_mels=128, sample_rate=22050, fft_size=2048)
out = torch.empty((B, 128, 431), dtype=torch.float32, device="cuda:0")
scratch = torch.empty(_lib.lib.smb_stft_fft_ceiling_scratch_bytes(sc._h, B, N) // 4,
                      dtype=torch.float32, device="cuda:0")


def timed(fn, steps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


res = {}
res["mel_ms"] = timed(lambda: sb.mel_spectrogram(sc, mc, x, out=out))
os.environ["SMB_PAIR_SKIP_MEL"] = "1"
res["no_mel_phase_ms"] = timed(lambda: sb.mel_spectrogram(sc, mc, x, out=out))
del os.environ["SMB_PAIR_SKIP_MEL"]
st = torch.cuda.current_stream().cuda_stream
_lib.check(_lib.lib.smb_stft_plan_set_stream(sc._h, st))
res["fft_ceiling_ms"] = timed(lambda: _lib.check(_lib.lib.smb_stft_fft_ceiling(
    sc._h, x.data_ptr(), B, N, scratch.data_ptr())))
os.environ["SMB_PAIR_FREERUN"] = "1"
res["fft_ceiling_free_running_ms"] = timed(lambda: _lib.check(_lib.lib.smb_stft_fft_ceiling(
    sc._h, x.data_ptr(), B, N, scratch.data_ptr())))
del os.environ["SMB_PAIR_FREERUN"]
# timing only (results are wrong): what the group barrier before the mel phase costs
os.environ["SMB_PAIR_NO_B"] = "1"
res["mel_without_group_barrier_ms"] = timed(lambda: sb.mel_spectrogram(sc, mc, x, out=out))
res["fft_ceiling_without_group_barrier_ms"] = timed(lambda: _lib.check(_lib.lib.smb_stft_fft_ceiling(
    sc._h, x.data_ptr(), B, N, scratch.data_ptr())))
del os.environ["SMB_PAIR_NO_B"]
res["fast_kernel_ms"] = timed(lambda: sb.mel_spectrogram(sc.set_path("fast"), mc, x, out=out))
res["hbm_floor_ms"] = 1e3 * 1129136128 / 6542.7e9
print(json.dumps(res))
