"""Throughput of Resample.apply over common rate pairs (one GPU, device-resident,
128 clips x 30 s): shows which executor each plan runs and how far from the HBM
bound it lands.  python tools/bench_resample_pairs.py"""
import json
import sys

import torch

sys.path.insert(0, ".")
import soundml_b200 as sb  # noqa: E402
from bench import peak_hbm  # noqa: E402

PAIRS = [(44100, 16000), (44100, 48000), (48000, 44100), (48000, 16000), (16000, 48000),
         (44100, 22050), (22050, 44100), (48000, 8000), (8000, 48000), (32000, 16000),
         (96000, 48000), (44100, 32000)]


def main():
    peak, _ = peak_hbm()
    clips = 128
    for sr, target in PAIRS:
        n = 30 * sr
        x = torch.rand((clips, n), device="cuda") * 2 - 1
        cfg = sb.Resample.Config.create(sample_rate=sr, target=target)
        out = torch.empty((clips, cfg.output_frames(n)), device="cuda")
        sb.Resample.apply(cfg, x, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            sb.Resample.apply(cfg, x, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        gbs = (x.numel() + out.numel()) * 4 / (ms * 1e-3) / 1e9
        print(json.dumps({"pair": f"{sr}->{target}", "plan": cfg.pp(), "ms": round(ms, 3),
                          "audio_s_per_s": round(clips * 30 / (ms * 1e-3)),
                          "hbm_frac": round(gbs / peak, 3)}), flush=True)
        del x, out


if __name__ == "__main__":
    main()
