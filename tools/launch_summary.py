#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/launch_summary.py gpurun_out/launches.csv [steps_in_capture]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    lines = [l for l in open(path) if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"{sum(c for c, _ in agg.values())} launches, {tot:.1f} us total, {tot / steps:.1f} us per step ({steps:g} steps)")
    print(f"{'us/step':>10} {'launches/step':>13} {'share':>6}  kernel")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / steps:10.1f} {c / steps:13.1f} {t / tot * 100:5.1f}%  {k[:90]}")


if __name__ == "__main__":
    main()
