#!/usr/bin/env python
"""Per-kernel summary of the LAST graph-replayed training step in an ncu launch list of tools/prof_train.py --graph --reps 1
(the step = the launches between the last two adam_ema_step_kernel launches).
usage: python tools/train_launch_summary.py gpurun_out/r2_train_launches.csv [measured ms]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ni, ui = hdr.index("Kernel Name"), hdr.index("Metric Unit")
    data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
    adam = [i for i, r in enumerate(data) if "adam_ema_step" in r[ni]]
    seg = data[adam[-2] + 1:adam[-1] + 1]
    d = collections.defaultdict(lambda: [0, 0.0])
    for r in seg:
        v = float(r[-1].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
        d[r[ni][:100]][0] += 1
        d[r[ni][:100]][1] += v
    tot = sum(v[1] for v in d.values())
    print("# one graph-replayed IDDPM training step (default UNet, batch 128, bf16): ncu --metrics gpu__time_duration.sum "
          "--clock-control none python tools/prof_train.py --graph --reps 1")
    print("# cold-cache serialised times: compare SHARES" + (f"; the replayed step measures {sys.argv[2]} ms" if len(sys.argv) > 2 else ""))
    print(f"total {tot:.0f} us over {len(seg)} launches")
    for k, v in sorted(d.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{v[1]:9.1f} us x{v[0]:4d} {100 * v[1] / tot:5.1f}%  {k}")


if __name__ == "__main__":
    main()
