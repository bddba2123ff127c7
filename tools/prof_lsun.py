#!/usr/bin/env python
"""LSUN-256 UNet (configs/ddpm/lsun_bedroom.yaml:78-90) sampling step at batch B: ms per graph-replayed step, achieved
TFLOP/s (42.4 GFLOP per 64x64 image scale -> measured FLOPs from the conv descriptors), launches per step.
usage: python tools/prof_lsun.py [--batch 2] [--steps 10]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-made-easy_b200"))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    from dmme_b200 import DDPM, ops
    from dmme_b200.models import _engine
    from dmme_b200.models.ddpm import UNet
    torch.manual_seed(0)
    m = UNet(dropout=0.0, channels_per_depth=(128, 128, 256, 256, 512, 512), attention_depths=(5,)).eval()
    d = DDPM(m).cuda()
    x = torch.randn(a.batch, 3, 256, 256, device="cuda")
    counter = torch.full((1,), 1000, dtype=torch.int64, device="cuda")
    flops = []
    orig = ops.conv2d_launch

    def counted(desc, *args, **kw):
        ho, wo = ops.conv_out_hw(desc)
        k = desc.ksize * desc.ksize * (desc.c0 + desc.c1) + desc.rc0 + desc.rc1
        flops.append(2.0 * desc.n * ho * wo * desc.cout * k)
        return orig(desc, *args, **kw)

    _engine.ops.conv2d_launch = counted
    d._graph_step(x, counter, 1)
    _engine.ops.conv2d_launch = orig
    torch.cuda.synchronize()
    ops.reset_launch_count()
    d._graph_step(x, counter, 1)
    torch.cuda.synchronize()
    launches = ops.launch_count()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        d._graph_step(x, counter, 1)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    tf = sum(flops) / (ms * 1e-3) / 1e12
    print(f"lsun-256 batch {a.batch}: {ms:.3f} ms/step, {sum(flops) / 1e9:.1f} conv GFLOP/step, {tf:.1f} TFLOP/s, "
          f"{launches} launches/step, finite={bool(torch.isfinite(x).all())}")


if __name__ == "__main__":
    main()
