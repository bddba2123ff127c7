#!/usr/bin/env python
"""Stand-alone timing of the 1x1 convolutions of the attention blocks (qkv projection, output projection with the
residual addend) at batch 256, CUDA events, L2 flushed between repetitions.  Feature knobs isolate the epilogue cost.
usage: python tools/prof_1x1.py [--reps 10]"""
import argparse
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-made-easy_b200"))
import torch  # noqa: E402

from dmme_b200 import _lib as L  # noqa: E402
from dmme_b200 import ops  # noqa: E402

CASES = [  # (name, n, h, cin, cout, layout, addend, stats)
    ("qkv 256->768 @16", 256, 16, 256, 768, "qkv", False, False),
    ("nhwc 256->768 @16", 256, 16, 256, 768, "nhwc", False, False),
    ("proj 256->256 @16 +x +stats", 256, 16, 256, 256, "nhwc", True, True),
    ("proj 256->256 @16 +I.x +stats", 256, 16, 256, 256, "nhwc", "eye", True),
    ("proj 128->128 @16 +I.x +stats", 256, 16, 128, 128, "nhwc", "eye", True),
    ("proj 256->256 @16 +x", 256, 16, 256, 256, "nhwc", True, False),
    ("proj 256->256 @16 +stats", 256, 16, 256, 256, "nhwc", False, True),
    ("proj 256->256 @16 plain", 256, 16, 256, 256, "nhwc", False, False),
    ("qkv 128->384 @16", 256, 16, 128, 384, "qkv", False, False),
    ("proj 128->128 @16 +x +stats", 256, 16, 128, 128, "nhwc", True, True),
    ("qkv 256->768 @4", 256, 4, 256, 768, "nhwc", False, False),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--flush", type=int, default=1)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, n, h, cin, cout, layout, use_add, use_stats in CASES:
        if args.only and args.only not in name:
            continue
        x = torch.randn(n, h, h, cin, device=dev, generator=g).bfloat16()
        w = torch.randn(cout, cin, 1, 1, device=dev, generator=g) / math.sqrt(cin)
        eye = use_add == "eye"
        wres = torch.eye(cout, device=dev).view(cout, cout, 1, 1).contiguous() if eye else None
        xres = torch.randn(n, h, h, cout, device=dev, generator=g).bfloat16() if eye else None
        wp = ops.pack_conv_weight(w, wres, True)
        bias = torch.randn(cout, device=dev, generator=g)
        lay = L.OUT_QKV if layout == "qkv" else L.OUT_NHWC
        d = ops.make_conv_desc(x, None, cout, 1, 1, False, xres, None, False, lay, torch.bfloat16, L.CONV_AUTO)
        if layout == "qkv":
            c = cout // 3
            out = torch.empty(n, h * h, c, device=dev, dtype=torch.bfloat16)
            out2 = torch.empty_like(out)
            out3 = torch.empty(n, c, h * h, device=dev, dtype=torch.bfloat16)
        else:
            out = torch.empty(n, h, h, cout, device=dev, dtype=torch.bfloat16)
            out2 = out3 = None
        addend = torch.randn(n, h, h, cout, device=dev, generator=g).bfloat16() if (use_add and not eye) else None
        st = torch.zeros(n * cout // 4 * 2, dtype=torch.int64, device=dev) if use_stats else None
        flop = 2.0 * n * h * h * cout * cin
        byts = 2.0 * n * h * h * (cin + cout * (2 if use_add else 1))
        times = []
        for r in range(args.reps + 2):
            if args.flush:
                flush.fill_(r)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv2d_launch(d, wp, bias, out, None, addend, out2, out3, stats=st)
            e1.record()
            torch.cuda.synchronize()
            if r >= 2:
                times.append(e0.elapsed_time(e1))
        ms = sorted(times)[len(times) // 2]
        print(f"{name:30s} {ms * 1e3:8.1f} us  {flop / ms / 1e9:7.1f} TFLOP/s  {byts / ms / 1e6:7.1f} GB/s  best {min(times) * 1e3:.1f} us",
              flush=True)


if __name__ == "__main__":
    main()
