"""Row form against window form of the tensor-core resampler stage on random input, small and at
scale, three times over (a race does not show every time): max difference and where the bad
values sit (clip, block-row, column).  python tools/check_gemm_forms.py"""
import os, sys, torch, numpy as np
sys.path.insert(0, ".")
import soundml_b200 as sb
for sr, target, clips, secs in [(44100, 48000, 4, 2), (44100, 48000, 128, 30), (44100, 16000, 4, 2)]:
    n = secs * sr
    torch.manual_seed(0)
    x = torch.rand((clips, n), device="cuda") * 2 - 1
    cfg = sb.Resample.Config.create(sample_rate=sr, target=target)
    out = torch.empty((clips, cfg.output_frames(n)), device="cuda")
    os.environ.pop("SMB_GEMM_WINDOWS", None)
    for rep in range(3):
        sb.Resample.apply(cfg, x, out=out)
        torch.cuda.synchronize()
        got = out.clone()
        os.environ["SMB_GEMM_WINDOWS"] = "1"
        sb.Resample.apply(cfg, x, out=out)
        torch.cuda.synchronize()
        os.environ.pop("SMB_GEMM_WINDOWS", None)
        d = (got - out).abs()
        bad = (d > 1e-4).nonzero()
        print(sr, target, clips, secs, "rep", rep, "max", float(d.max()), "bad", bad.shape[0])
        if bad.shape[0]:
            b = bad.cpu().numpy()
            L = 160
            rows = b[:, 1] // L; cols = b[:, 1] % L
            print(" clips", np.unique(b[:, 0])[:10], "rows", np.unique(rows)[:40], "n rows", len(np.unique(rows)), "cols", np.unique(cols)[:40])
            print(" rows mod 120:", np.unique(rows % 120)[:60])
            break
