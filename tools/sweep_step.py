#!/usr/bin/env python
"""Graph-replayed sampling step time at several per-GPU batches (the strong-scaling shards of BASELINE config #2).
usage: python tools/sweep_step.py [--batches 256,128,64,32] [--flavour ddpm]   (A/B switches come from the environment)"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-made-easy_b200"))
import torch  # noqa: E402

from dmme_b200 import DDPM, IDDPM, ops  # noqa: E402
from dmme_b200.models import ddpm as m_ddpm, iddpm as m_iddpm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="256,128,64,32")
    ap.add_argument("--flavour", default="ddpm", choices=["ddpm", "iddpm"])
    ap.add_argument("--reps", type=int, default=30)
    args = ap.parse_args()
    dev = torch.device("cuda")
    torch.manual_seed(0)
    out = []
    for b in [int(v) for v in args.batches.split(",")]:
        if args.flavour == "ddpm":
            ddpm = DDPM(m_ddpm.UNet().eval(), 1000).to(dev)
        else:
            ddpm = IDDPM(m_iddpm.UNet().eval(), 1000).to(dev)
        x = torch.randn(b, 3, 32, 32, device=dev)
        counter = torch.full((1,), 1000, dtype=torch.int64, device=dev)
        for _ in range(2):
            ddpm._graph_step(x, counter, 1)
        torch.cuda.synchronize()
        ops.reset_launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            ddpm._graph_step(x, counter, 1)
        launches = ops.launch_count()
        for _ in range(5):
            g.replay()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / args.reps)
        out.append(f"{b}: {best:.3f} ms ({launches} launches)")
        del g, ddpm
    print(f"{args.flavour} " + "  ".join(out), flush=True)


if __name__ == "__main__":
    main()
