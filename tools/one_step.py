#!/usr/bin/env python
"""Exactly ONE graph-replayed sampling step between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv`: the launch list of one step at
a given per-GPU batch (cold-cache, serialised times: shares, not absolutes).
usage: ncu --profile-from-start off ... python tools/one_step.py [--batch 256] [--flavour ddpm]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-made-easy_b200"))
import torch  # noqa: E402

from dmme_b200 import DDPM, IDDPM  # noqa: E402
from dmme_b200.models import ddpm as m_ddpm, iddpm as m_iddpm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--flavour", default="ddpm", choices=["ddpm", "iddpm"])
    ap.add_argument("--replays", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda")
    torch.manual_seed(0)
    if args.flavour == "ddpm":
        ddpm = DDPM(m_ddpm.UNet().eval(), 1000).to(dev)
    else:
        ddpm = IDDPM(m_iddpm.UNet().eval(), 1000).to(dev)
    x = torch.randn(args.batch, 3, 32, 32, device=dev)
    counter = torch.full((1,), 1000, dtype=torch.int64, device=dev)
    for _ in range(2):
        ddpm._graph_step(x, counter, 1)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ddpm._graph_step(x, counter, 1)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"# {args.flavour} batch {args.batch}: {e0.elapsed_time(e1) / 20:.4f} ms per graph replay", flush=True)
    torch.cuda.cudart().cudaProfilerStart()
    for _ in range(args.replays):
        g.replay()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
