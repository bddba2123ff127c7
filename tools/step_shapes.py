#!/usr/bin/env python
"""Per-(kernel, grid) timing inside one step of an ncu launch list (between two ddpm_step launches).
usage: python tools/step_shapes.py gpurun_out/launches.csv"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
idx = [i for i, r in enumerate(rows) if "ddpm_step" in r["Kernel Name"]]
a, b = idx[0], idx[1]
agg = collections.OrderedDict()
for r in rows[a + 1:b + 1]:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void dmme::", "").replace("dmme::", "")
    key = (name, r["Grid Size"])
    v = float(r["Metric Value"].replace(",", "")) / 1e3
    agg.setdefault(key, [0, 0.0])
    agg[key][0] += 1
    agg[key][1] += v
tot = sum(v[1] for v in agg.values())
print("step total us", round(tot, 1))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:8.1f} us  x{c:2d}  avg {t / c:7.1f}  {k[0][:40]:40s} grid {k[1]}")
