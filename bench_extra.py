#!/usr/bin/env python
"""Secondary measurements: BASELINE.json configs[2..4] on one GPU.

These are parity-test cases, not the headline bench line (bench.py); the numbers
land in profiles/ for DESIGN.md.  Each line carries the same roofline fields.

    python bench_extra.py [--scale 1.0] [--steps 5]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


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
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the configured batch")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import torch
    import soundml_b200 as sb
    from bench import peak_hbm
    peak, kind = peak_hbm()
    dev = torch.device("cuda", 0)

    def report(name, workload, ms, algo_bytes, audio_seconds, launches, extra=None):
        ach = algo_bytes / (ms * 1e-3) / 1e9
        line = {"config": name, "workload": workload, "ms_per_step": ms,
                "audio_seconds_per_second": audio_seconds / (ms * 1e-3),
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                             "frac": ach / peak, "peak_source": kind,
                             "algorithmic_bytes_per_step": algo_bytes},
                "gpu_launches_per_step": launches}
        if extra:
            line.update(extra)
        print(json.dumps(line), flush=True)

    # ---- config 3: 511-tap FIR lowpass over 256 x 60 s stereo 48 kHz clips
    if not args.only or "fir" in args.only:
        lines = max(1, int(256 * 2 * args.scale))
        n = 60 * 48000
        x = torch.rand((lines, n), device=dev) * 2 - 1
        out = torch.empty_like(x)
        fir = sb.Fir.lowpass(k=255, cutoff=0.25)
        for method in ("ols", "direct"):
            c0 = sb.kernel_launch_count()
            ms = timed(lambda: fir.apply(x, method=method, out=out), args.steps, 2)
            per = (sb.kernel_launch_count() - c0) // (args.steps + 2)
            report(f"fir511_{method}", f"{lines} lines x 60 s @48 kHz, 511 taps ({method})",
                   ms, 2 * lines * n * 4, lines * 60.0, per)
        del x, out

    # ---- config 4: 44.1 -> 16 kHz polyphase, 4096 x 30 s over 8 GPUs -> 512 clips per GPU
    if not args.only or "resample" in args.only:
        clips = max(1, int(512 * args.scale))
        n = 30 * 44100
        x = torch.rand((clips, n), device=dev) * 2 - 1
        cfg = sb.Resample.Config.create(sample_rate=44100, target=16000)
        total = cfg.output_frames(n)
        out = torch.empty((clips, total), device=dev)
        c0 = sb.kernel_launch_count()
        ms = timed(lambda: sb.Resample.apply(cfg, x, out=out), max(1, args.steps // 2), 1)
        per = (sb.kernel_launch_count() - c0) // (max(1, args.steps // 2) + 1)
        report("resample_44k_16k", f"{clips} clips x 30 s, {cfg.pp()}", ms,
               clips * (n + total) * 4, clips * 30.0, per)
        del x, out

    # ---- config 5: resample 44.1 -> 22.05 kHz + STFT + mel, 8192 x 10 s clips per GPU share
    if not args.only or "e2e" in args.only:
        clips = max(1, int(8192 * args.scale))
        n = 10 * 44100
        x = torch.rand((clips, n), device=dev) * 2 - 1
        cfg = sb.Resample.Config.create(sample_rate=44100, target=22050)
        mid = torch.empty((clips, cfg.output_frames(n)), device=dev)
        sc = sb.Stft.Config.create(fft_size=2048, hop=512)
        mc = sb.Mel.Config.create(n_mels=128, sample_rate=22050, fft_size=2048)
        frames = sb.Stft.frames(sc, mid.shape[1])
        out = torch.empty((clips, 128, frames), device=dev)

        def step():
            sb.Resample.apply(cfg, x, out=mid)
            sb.mel_spectrogram(sc, mc, mid, out=out)
        c0 = sb.kernel_launch_count()
        ms = timed(step, args.steps, 2)
        per = (sb.kernel_launch_count() - c0) // (args.steps + 2)
        report("resample_stft_mel", f"{clips} clips x 10 s @44.1 kHz -> 22.05 kHz -> mel 128",
               ms, clips * (n * 4 + 128 * frames * 4), clips * 10.0, per)

    # ---- config 1 outputs at batch scale: complex spectrum and power spectrogram, 1024 clips
    if not args.only or "spectrum" in args.only:
        clips = max(1, int(1024 * args.scale))
        n = 220500
        from soundml_b200 import synth
        x = synth.clips_torch(clips, n, device=dev)
        sc = sb.Stft.Config.create(fft_size=2048, hop=512)
        frames = sb.Stft.frames(sc, n)
        for name, fn, width in (("stft_power", sb.Stft.power_spectrum, 4),
                                ("stft_complex", sb.Stft.transform, 8)):
            out = fn(sc, x)
            c0 = sb.kernel_launch_count()
            ms = timed(lambda: fn(sc, x, out=out), args.steps, 2)
            per = (sb.kernel_launch_count() - c0) // (args.steps + 2)
            report(name, f"{clips} clips x 10 s @22.05 kHz, fft 2048 hop 512 -> [1025, {frames}]",
                   ms, clips * (n * 4 + 1025 * frames * width), clips * 10.0, per)
            del out
        del x

    # ---- SURVEY 8f rank 1 and the standalone Mel.apply: 1024 clips
    if not args.only or "epilogue" in args.only:
        clips = max(1, int(1024 * args.scale))
        n = 220500
        from soundml_b200 import synth
        x = synth.clips_torch(clips, n, device=dev)
        sc = sb.Stft.Config.create(fft_size=2048, hop=512)
        mc = sb.Mel.Config.create(n_mels=128, sample_rate=22050, fft_size=2048)
        frames = sb.Stft.frames(sc, n)
        for name, fn, out_rows in (("mfcc20", lambda: sb.mfcc(sc, mc, x, n_mfcc=20), 20),
                                   ("logmel_db", lambda: sb.log_mel_spectrogram(sc, mc, x, top_db=80.0), 128),
                                   ("logmel_db_two_calls", lambda: sb.Convert.power_to_db(
                                       sb.mel_spectrogram(sc, mc, x), top_db=80.0), 128)):
            c0 = sb.kernel_launch_count()
            ms = timed(fn, args.steps, 2)
            per = (sb.kernel_launch_count() - c0) // (args.steps + 2)
            report(name, f"{clips} clips x 10 s @22.05 kHz -> [{out_rows}, {frames}]", ms,
                   clips * (n * 4 + out_rows * frames * 4), clips * 10.0, per)
        s_pow = sb.Stft.power_spectrum(sc, x)
        del x
        c0 = sb.kernel_launch_count()
        ms = timed(lambda: sb.Mel.apply(mc, s_pow), args.steps, 2)
        per = (sb.kernel_launch_count() - c0) // (args.steps + 2)
        report("mel_apply", f"{clips} x [1025, {frames}] f32 power spectrogram -> [128, {frames}]", ms,
               clips * frames * (1025 + 128) * 4, clips * 10.0, per)
        del s_pow

    # ---- analysis + least-squares synthesis round trip, 1024 x 10 s clips
    if not args.only or "istft" in args.only:
        clips = max(1, int(1024 * args.scale))
        n = 220500
        from soundml_b200 import synth
        x = synth.clips_torch(clips, n, device=dev)
        sc = sb.Stft.Config.create(fft_size=2048, hop=512)
        z = sb.Stft.transform(sc, x)
        c0 = sb.kernel_launch_count()
        ms = timed(lambda: sb.Stft.invert(sc, z, length=n), max(1, args.steps // 2), 1)
        per = (sb.kernel_launch_count() - c0) // (max(1, args.steps // 2) + 1)
        y = sb.Stft.invert(sc, z, length=n)
        err = ((y - x).abs().max() / x.abs().max()).item()
        report("istft", f"{clips} clips x 10 s @22.05 kHz, fft 2048 hop 512, complex64 -> f32",
               ms, z.numel() * 8 + clips * n * 4, clips * 10.0, per,
               {"round_trip_peak_rel_err": err})


if __name__ == "__main__":
    main()
