#!/usr/bin/env python
"""Phase breakdown of the frame-pair kernel from `ncu --page source --csv`:
samples, instructions per frame and top stall reasons between landmarks of the
SASS stream (group barriers, the transposition stores, the split's shuffles).

    ncu -i prof.ncu-rep --page source --csv > src.csv
    python tools/ncu_phases2.py src.csv [frames]
"""
import csv
import sys


def main(path, frames=441344):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {n: i for i, n in enumerate(hdr)}
    keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    data = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        f = lambda k: int(float(r[ix[k]] or 0))
        src = r[ix["Source"]].strip()
        if src.startswith("@"):
            src = src.split(None, 1)[1]
        data.append((src, f("# Samples"), f("Instructions Executed"), {k: f(k) for k in keys},
                     f("L1 Wavefronts Shared")))
    ts = sum(d[1] for d in data)

    def find(prefix, start=0, last=False):
        hits = [i for i, d in enumerate(data) if d[0].startswith(prefix) and i >= start]
        return (hits[-1] if last else hits[0]) if hits else None
    # the two group barriers of the tile loop are the named ones (register operand)
    bars = [i for i, d in enumerate(data) if d[0].startswith("BAR.SYNC") and " R" in d[0]]
    b1, b2 = bars[0], bars[1]
    s0, s1 = find("STS.128", b1), find("STS.128", b1, last=True)
    h0 = find("SHFL", b1)
    regions = [("setup", 0, b1), ("pass 1: loads, window, fft32", b1, s0),
               ("twiddle + transposition stores", s0, s1 + 1),
               ("pass 2: loads, fft32", s1 + 1, h0), ("split, |X|^2, power rows", h0, b2),
               ("staging + mel + write-out", b2, len(data))]
    print(f"{'phase':34s} {'samples':>8s} {'inst/frame':>10s} {'sh.wf/frame':>11s}  top stalls")
    for name, a, b in regions:
        s = sum(d[1] for d in data[a:b])
        n = sum(d[2] for d in data[a:b])
        w = sum(d[4] for d in data[a:b])
        st = {}
        for d in data[a:b]:
            for k, v in d[3].items():
                st[k] = st.get(k, 0) + v
        top = ", ".join(f"{k[6:]} {100 * v / max(1, s):.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
        print(f"{name:34s} {100 * s / ts:7.1f}% {n / frames:10.1f} {w / frames:11.1f}  {top}")
    print(f"{'total':34s} {100.0:7.1f}% {sum(d[2] for d in data) / frames:10.1f} {sum(d[4] for d in data) / frames:11.1f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 441344)
