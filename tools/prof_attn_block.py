#!/usr/bin/env python
"""Times the one-launch attention block (csrc/attention_block.cu) against the four-launch path it replaces (GroupNorm |
1x1 qkv | attention core | 1x1 proj + x) on the 16x16 x 256-channel site of the default DDPM UNet.
usage: python tools/prof_attn_block.py [--batch 256] [--reps 20]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-made-easy_b200"))
import torch  # noqa: E402

from dmme_b200.models import ddpm as m_ddpm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--hw", type=int, default=16, help="16: the 16x16 sites, 4: the 4x4 middle block")
    ap.add_argument("--c", type=int, default=256, help="channels of the site (256 or 128 at 16x16)")
    args = ap.parse_args()
    dev = torch.device("cuda")
    torch.manual_seed(0)
    unet = m_ddpm.UNet().eval().to(dev)
    eng = unet.engine
    blk = next(m for _, m in eng.resblocks() if not isinstance(m.attention, torch.nn.Identity)
               and m.conv1[2].weight.shape[0] == args.c)
    att = blk.attention
    n = args.batch
    x = (torch.randn(n, args.hw, args.hw, args.c, device=dev) * 1.5).to(torch.bfloat16)
    v = x.double().reshape(n, args.hw * args.hw, args.c // 4, 4)
    st = (torch.stack([v.sum((1, 3)), (v * v).sum((1, 3))], dim=-1) * 2 ** 20).round().to(torch.int64).reshape(-1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    outs = {}
    for fused in (True, False):
        eng.fuse_attn = fused
        times = []
        for rep in range(args.reps + 3):
            eng._begin_stats(dev)
            eng._stats[x.data_ptr()] = st
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y = eng.attention_block("prof", att, x)
            e1.record()
            torch.cuda.synchronize()
            if rep >= 3:
                times.append(e0.elapsed_time(e1) * 1e3)
        outs[fused] = y.float().clone()
        times.sort()
        print(f"batch {n} {'one launch ' if fused else 'four launches'}: median {times[len(times) // 2]:.1f} us  min {times[0]:.1f} us")
    d = (outs[True] - outs[False]).norm() / outs[False].norm()
    print(f"rel-L2 between the two paths: {float(d):.3e}")


if __name__ == "__main__":
    main()
