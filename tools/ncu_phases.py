#!/usr/bin/env python
"""Instruction / sample share per line range of one file in an ncu source page.
    python tools/ncu_phases.py src.csv file.cu  a-b:name  c-d:name ...   (frames for per-frame counts via FRAMES env)"""
import csv, os, sys
def main(path, fname, ranges):
    rows = list(csv.reader(open(path)))
    hdr, cur = None, ""
    data = []
    for r in rows:
        if r and r[0] == "File Path": cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No": hdr = r
        elif hdr and r and r[0].isdigit() and len(r) == len(hdr):
            def g(n):
                try: return int(float(r[hdr.index(n)]))
                except ValueError: return 0
            data.append((cur, int(r[0]), g("Instructions Executed"), g("# Samples"), g("L1 Wavefronts Shared")))
    ti = sum(d[2] for d in data); ts = sum(d[3] for d in data)
    frames = float(os.environ.get("FRAMES", "441344"))
    print(f"total inst {ti} ({ti/frames:.0f}/frame) samples {ts}")
    other_i = ti; other_s = ts
    for spec in ranges:
        rng, name = spec.split(":")
        a, b = map(int, rng.split("-"))
        i = sum(d[2] for d in data if d[0].startswith(fname[:12]) and a <= d[1] <= b)
        s = sum(d[3] for d in data if d[0].startswith(fname[:12]) and a <= d[1] <= b)
        w = sum(d[4] for d in data if d[0].startswith(fname[:12]) and a <= d[1] <= b)
        other_i -= i; other_s -= s
        print(f"{name:>14}: inst {i/frames:7.1f}/frame ({100*i/ti:4.1f}%)  samples {100*s/ts:4.1f}%  sh.wavefronts {w/frames:6.1f}/frame")
    print(f"{'other files':>14}: inst {other_i/frames:7.1f}/frame ({100*other_i/ti:4.1f}%)  samples {100*other_s/ts:4.1f}%")
    if len(ranges) == 0:
        pass
main(sys.argv[1], sys.argv[2], sys.argv[3:])
