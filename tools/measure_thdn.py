"""SFDR and THD+N (resample_quality.ml Q1/Q2: gate 125 dB / -125 dB in float32) of the
float32 planned path for the tensor-core pairs: the row form as shipped, the row form with
the A tiles in shared memory, and round 1's window form.  One JSON-ish line per (pair, form):
(tone Hz, SFDR dB, THD+N dB).  python tools/measure_thdn.py   (SMB_ROWS_SPLIT=0 shows the
out-of-gate single-accumulator variant)"""
import os, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import soundml_b200 as sb
import test_gpu_resample_quality as Q
for sr, target in [(44100, 16000), (44100, 48000), (48000, 44100), (44100, 32000)]:
    for form in ("rows", "rows_smemA", "windows"):
        os.environ.pop("SMB_GEMM_WINDOWS", None); os.environ.pop("SMB_ROWS_SMEM_A", None)
        if form == "windows": os.environ["SMB_GEMM_WINDOWS"] = "1"
        if form == "rows_smemA": os.environ["SMB_ROWS_SMEM_A"] = "1"
        cfg = sb.Resample.Config.create(sample_rate=sr, target=target)
        res = []
        for f in Q.tone_set(target):
            d, t = Q.sfdr_thdn(Q.spectrum(Q.convert(sb, cfg, Q.tone(sr, f, 2.0), np.float32)))
            res.append((f, round(d, 1), round(t, 1)))
        print(sr, target, form, res, flush=True)
