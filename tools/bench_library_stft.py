#!/usr/bin/env python
"""Library yardstick for the headline workload: the same 1024 x 10 s clips through
torch.stft (cuFFT) -> |X|^2 -> torch.matmul with the mel weights (cuBLAS), timed
beside the fused kernel on the same box.  Not a product path and not an oracle:
it answers "what would stock library kernels do here" for DESIGN.md 4.1.

    python tools/bench_library_stft.py [--clips 1024] [--steps 20]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, steps, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    import torch
    import soundml_b200 as sb
    from soundml_b200 import synth

    dev = "cuda:0"
    n = 220500
    x = synth.clips_torch(args.clips, n, dev)
    sc = sb.Stft.Config.create(fft_size=2048, hop=512)
    mc = sb.Mel.Config.create(n_mels=128, sample_rate=22050, fft_size=2048)
    w = torch.tensor(sb.Mel.filterbank(mc), dtype=torch.float32, device=dev)          # [128, 1025]
    win = torch.tensor(sc.analysis_window, dtype=torch.float32, device=dev)

    def library():
        z = torch.stft(x, 2048, hop_length=512, window=win, center=True, pad_mode="reflect",
                       return_complex=True)                                  # [B, 1025, 431]
        p = z.real * z.real + z.imag * z.imag
        return torch.matmul(w, p)                                            # [B, 128, 431]

    def fused():
        return sb.mel_spectrogram(sc, mc, x)

    ref = library()
    got = fused()
    err = float(((got - ref).abs().amax(dim=(1, 2)) / ref.abs().amax(dim=(1, 2))).max())
    del ref, got
    t_lib = timed(library, args.steps)
    t_fused = timed(fused, args.steps)

    # the complex spectrum alone: torch.stft (cuFFT) against Stft.transform
    def library_z():
        return torch.stft(x, 2048, hop_length=512, window=win, center=True, pad_mode="reflect",
                          return_complex=True)
    zout = sb.Stft.transform(sc, x)
    zref = library_z()
    zerr = float((torch.view_as_real(zout) - torch.view_as_real(zref)).abs().amax() /
                 torch.view_as_real(zref).abs().amax())
    del zref
    t_lib_z = timed(library_z, args.steps)
    t_fused_z = timed(lambda: sb.Stft.transform(sc, x, out=zout), args.steps)
    print(json.dumps({
        "workload": f"mel_spectrogram 128 bands, {args.clips} x 10 s 22.05 kHz clips, n_fft=2048 hop=512",
        "library_ms": t_lib, "library": "torch.stft (cuFFT) + elementwise |X|^2 + torch.matmul (cuBLAS, fp32)",
        "fused_ms": t_fused, "speedup": t_lib / t_fused,
        "max_rel_diff_per_clip_fused_vs_library": err,
        "transform_library_ms": t_lib_z, "transform_fused_ms": t_fused_z,
        "transform_speedup": t_lib_z / t_fused_z, "transform_max_rel_diff": zerr,
        "note": "library intermediates (complex spectrum 3.6 GB, power 1.8 GB) go through HBM; "
                "torch's allow_tf32 left at its default (False) for the matmul",
    }))


if __name__ == "__main__":
    main()
