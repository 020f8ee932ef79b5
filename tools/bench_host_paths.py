"""Host-memory calls (pinned buffers) of the resampler and the FIR against the copy
floor of the same buffers: cudaMemcpyAsync of input and output alone, both directions
at once.  python tools/bench_host_paths.py"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import soundml_b200 as sb  # noqa: E402


def pinned(shape):
    return torch.empty(shape, dtype=torch.float32).pin_memory()


def timed(fn, steps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / steps


def copy_floor(xh, yh):
    d_in = torch.empty(xh.shape, device="cuda")
    d_out = torch.empty(yh.shape, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def go():
        with torch.cuda.stream(s1):
            d_in.copy_(xh, non_blocking=True)
        with torch.cuda.stream(s2):
            yh.copy_(d_out, non_blocking=True)
    return timed(go)


def main():
    # BASELINE.json configs[3], one GPU's share: 512 clips x 30 s, 44.1 -> 16 kHz
    clips, n = 512, 30 * 44100
    cfg = sb.Resample.Config.create(sample_rate=44100, target=16000)
    xh = pinned((clips, n))
    xh.uniform_(-1, 1)
    yh = pinned((clips, cfg.output_frames(n)))
    ms = timed(lambda: sb.Resample.apply(cfg, xh.numpy(), out=yh.numpy()))
    floor = copy_floor(xh, yh)
    dev = sb.Resample.apply(cfg, xh.cuda()).cpu()
    print(json.dumps({"case": "resample 44.1->16 kHz, 512 x 30 s, pinned host buffers", "ms": round(ms, 2),
                      "copy_floor_ms": round(floor, 2), "over_floor": round(ms / floor, 3),
                      "equals_device_call": bool(torch.equal(dev, yh))}))
    del xh, yh
    lines, n = 128, 60 * 48000
    fir = sb.Fir.lowpass(k=255, cutoff=0.25)
    xh = pinned((lines, n))
    xh.uniform_(-1, 1)
    yh = pinned((lines, n))
    ms = timed(lambda: fir.apply(xh.numpy(), method="ols", out=yh.numpy()))
    floor = copy_floor(xh, yh)
    print(json.dumps({"case": "FIR 511 taps (ols), 128 lines x 60 s, pinned host buffers", "ms": round(ms, 2),
                      "copy_floor_ms": round(floor, 2), "over_floor": round(ms / floor, 3)}))


if __name__ == "__main__":
    main()
