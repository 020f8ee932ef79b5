#!/usr/bin/env python
"""Summarise `ncu --page source --print-source cuda,sass --csv` per CUDA line.

    ncu -i prof.ncu-rep --page source --print-source cuda,sass --csv > src.csv
    python tools/ncu_lines.py src.csv [top]
"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    data, hdr, fname = [], None, ""
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and r and r[0] not in ("", "Function Name") and r[0].isdigit():
            if len(r) != len(hdr):
                continue

            def g(name):
                v = r[hdr.index(name)]
                try:
                    return int(float(v))
                except ValueError:
                    return 0
            data.append(dict(file=fname, line=int(r[0]), src=r[1].strip(), samples=g("# Samples"),
                             inst=g("Instructions Executed"), shwf=g("L1 Wavefronts Shared"),
                             shex=g("L1 Wavefronts Shared Excessive"),
                             long_sb=g("stall_long_sb"), short_sb=g("stall_short_sb"),
                             mio=g("stall_mio"), barrier=g("stall_barrier"), wait=g("stall_wait")))
    ts = sum(d["samples"] for d in data) or 1
    ti = sum(d["inst"] for d in data) or 1
    print(f"total samples {ts}  total warp instructions {ti}")
    print(" smp%  inst%   sh.wavefronts  excess   long  short   mio   bar  line  source")
    for d in sorted(data, key=lambda d: -d["samples"])[:top]:
        print(f"{100*d['samples']/ts:5.1f} {100*d['inst']/ti:6.1f} {d['shwf']:>14} {d['shex']:>10} "
              f"{d['long_sb']:>6} {d['short_sb']:>6} {d['mio']:>5} {d['barrier']:>5} "
              f"{d['file'][:12]}:{d['line']:<4} {d['src'][:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
