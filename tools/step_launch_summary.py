#!/usr/bin/env python
"""Per-kernel summary of ONE graph-replayed sampling step inside an ncu launch list of bench.py (the step = the launches
between the last two add_i64_kernel launches, i.e. between two `t -= 1`).
usage: python tools/step_launch_summary.py gpurun_out/r2_step_launches.csv [measured ms per step]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ni, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
    marks = [i for i, r in enumerate(data) if "add_i64_kernel" in r[ni]]
    seg = data[marks[-2] + 1:marks[-1] + 1] if len(marks) >= 2 else data  # tools/one_step.py lists exactly one step
    d = collections.defaultdict(lambda: [0, 0.0])
    for r in seg:
        v = float(r[-1].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
        name = r[ni].replace("void ", "").replace("dmme::", "")
        name = name.split("(")[0][:44]
        key = (name, r[gi].strip())
        d[key][0] += 1
        d[key][1] += v
    tot = sum(v[1] for v in d.values())
    print("# one DDPM sampling step (temb -> UNet forward -> DDPM update in the output conv's epilogue), default UNet, batch 256,")
    print("# bf16, one B200; ncu --metrics gpu__time_duration.sum --clock-control none of bench.py (last graph replay) or of")
    print("# tools/one_step.py with --profile-from-start off (exactly one replay)")
    print("# per-launch times are cold-cache and serialised: compare SHARES" +
          (f"; the graph-replayed step measures {sys.argv[2]} ms (bench.py)" if len(sys.argv) > 2 else ""))
    print(f"step total us {tot:.1f}   launches {len(seg)}")
    for (name, grid), v in sorted(d.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:8.1f} us  x{v[0]:2d}  avg {v[1] / v[0]:7.1f}  {name:44s} grid {grid:14s} {100 * v[1] / tot:5.1f}%")


if __name__ == "__main__":
    main()
