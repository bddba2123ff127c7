#!/usr/bin/env python
"""Stand-alone timing of the conv kernels on the UNet's dominant shapes (batch 256), CUDA events, L2-cold-ish
(a 256 MB scratch write between repetitions).  usage: python tools/prof_conv.py [--reps 10] [--only halo|tc]"""
import argparse
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-made-easy_b200"))
import torch  # noqa: E402

from dmme_b200 import _lib as L  # noqa: E402
from dmme_b200 import ops  # noqa: E402

SHAPES = [  # (name, n, h, c0, c1, cout, rc0, rc1): the 3x3 signatures of the default DDPM UNet's step
    ("128->128 @32", 256, 32, 128, 0, 128, 0, 0),
    ("256->128 @32", 256, 32, 128, 128, 128, 0, 0),
    ("128->128 @32 +res256", 256, 32, 128, 0, 128, 128, 128),
    ("256->256 @16", 256, 16, 256, 0, 256, 0, 0),
    ("128->256 @16", 256, 16, 128, 0, 256, 0, 0),
    ("512->256 @16", 256, 16, 256, 256, 256, 0, 0),
    ("256->256 @16 +res128", 256, 16, 256, 0, 256, 128, 0),
    ("256->256 @16 +res512", 256, 16, 256, 0, 256, 256, 256),
    ("128->128 @16 +res256", 256, 16, 128, 0, 128, 128, 128),
    ("256->256 @8", 256, 8, 256, 0, 256, 0, 0),
    ("256->256 @8 +res512", 256, 8, 256, 0, 256, 256, 256),
    ("512->256 @8", 256, 8, 256, 256, 256, 0, 0),
    ("256->256 @4", 256, 4, 256, 0, 256, 0, 0),
    ("512->256 @4", 256, 4, 256, 256, 256, 0, 0),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default="")
    ap.add_argument("--flush", type=int, default=1)
    ap.add_argument("--gn", type=int, default=0, help="1: fused GroupNorm+SiLU of the input inside the halo kernel")
    args = ap.parse_args()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, n, h, c0, c1, cout, rc0, rc1 in SHAPES:
        res = rc0 + rc1 > 0
        cin = c0 + c1
        s0 = torch.randn(n, h, h, c0, device=dev, generator=g).bfloat16()
        s1 = torch.randn(n, h, h, c1, device=dev, generator=g).bfloat16() if c1 else None
        w = torch.randn(cout, cin, 3, 3, device=dev, generator=g) / math.sqrt(9 * cin)
        r0 = torch.randn(n, h, h, rc0, device=dev, generator=g).bfloat16() if rc0 else None
        r1 = torch.randn(n, h, h, rc1, device=dev, generator=g).bfloat16() if rc1 else None
        wr = torch.randn(cout, rc0 + rc1, 1, 1, device=dev, generator=g) / math.sqrt(rc0 + rc1) if res else None
        wp = ops.pack_conv_weight(w, wr, True)
        bias = torch.randn(cout, device=dev, generator=g)
        temb = torch.randn(1, cout, device=dev, generator=g)
        out = torch.empty(n, h, h, cout, device=dev, dtype=torch.bfloat16)
        st = torch.zeros(n * cout // 4 * 2, dtype=torch.int64, device=dev)
        flop = 2.0 * n * h * h * cout * (9 * cin + rc0 + rc1)
        for kname, kernel in (("halo", L.CONV_HALO), ("tc", L.CONV_TC)):
            if args.only and args.only != kname:
                continue
            d = ops.make_conv_desc(s0, s1, cout, 3, 1, False, r0, r1, False,
                                   L.OUT_NHWC, torch.bfloat16, kernel)
            if not ops.conv_uses_tc(d):
                continue
            ab = None
            if args.gn:
                if kname != "halo" or not ops.conv_fuses_gn(d):
                    continue
                ab = torch.randn(n, cin, 2, device=dev, generator=g)
            times = []
            for r in range(args.reps + 2):
                if args.flush:
                    flush.fill_(r)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.conv2d_launch(d, wp, bias, out, temb, None, stats=st, gn_ab=ab)
                e1.record()
                torch.cuda.synchronize()
                if r >= 2:
                    times.append(e0.elapsed_time(e1))
            ms = sorted(times)[len(times) // 2]
            print(f"{name + (' +gn' if ab is not None else ''):26s} {kname:5s} {ms * 1e3:8.1f} us  {flop / ms / 1e9:7.1f} TFLOP/s (useful)  best {min(times) * 1e3:.1f} us")


if __name__ == "__main__":
    main()
