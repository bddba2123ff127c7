#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python tools/launch_summary.py gpurun_out/launches.csv [--skip N] [--detail PATTERN]"""
import argparse
import collections
import csv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--skip", type=int, default=0, help="ignore the first N launches (warm-up)")
    ap.add_argument("--detail", default="", help="print every launch whose kernel name contains this")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ni, ui, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Unit"), hdr.index("Grid Size"), hdr.index("Block Size")
    d = collections.defaultdict(lambda: [0, 0.0])
    k = 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        k += 1
        if k <= a.skip:
            continue
        v = float(r[-1].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
        d[r[ni]][0] += 1
        d[r[ni]][1] += v
        if a.detail and a.detail in r[ni]:
            print(f"  #{r[0]:>5s} {v:9.1f} us grid {r[gi]} block {r[bi]} {r[ni][:70]}")
    tot = sum(v[1] for v in d.values())
    print(f"{'us':>10s} {'calls':>5s} {'share':>6s}  kernel")
    for name, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:10.1f} {v[0]:5d} {100 * v[1] / tot:5.1f}%  {name[:110]}")
    print(f"{tot:10.1f} total over {sum(v[0] for v in d.values())} launches")


if __name__ == "__main__":
    main()
