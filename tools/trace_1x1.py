#!/usr/bin/env python
"""Per-role timeline of CTA 0 of the transposed tcgen05 conv kernel (clock64 timestamps written by the kernel when a
trace buffer is set): when the producer issued each k-block, when its operands landed, when each accumulator was ready
and when the tile was stored.  usage: python tools/trace_1x1.py [cin cout h addend(0/1)]"""
import ctypes as C
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-made-easy_b200"))
import torch  # noqa: E402

from dmme_b200 import _lib as L  # noqa: E402
from dmme_b200 import ops  # noqa: E402

cin, cout, h, use_add, ks = (int(a) for a in (sys.argv[1:6] + ["256", "256", "16", "0", "1"][len(sys.argv) - 1:]))
n, dev = 256, "cuda"
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(n, h, h, cin, device=dev, generator=g).bfloat16()
w = torch.randn(cout, cin, ks, ks, device=dev, generator=g) / math.sqrt(cin * ks * ks)
wp = ops.pack_conv_weight(w, None, True)
bias = torch.randn(cout, device=dev, generator=g)
d = ops.make_conv_desc(x, None, cout, ks, 1, False, None, None, False, L.OUT_NHWC, torch.bfloat16, L.CONV_TC)
out = torch.empty(n, h, h, cout, device=dev, dtype=torch.bfloat16)
addend = torch.randn(n, h, h, cout, device=dev, generator=g).bfloat16() if use_add else None
st = torch.zeros(n * cout // 4 * 2, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
trace = torch.zeros(3 * 512, dtype=torch.int64, device=dev)
lib = L.load()
lib.dmme_debug_set_conv_trace.argtypes = [C.c_void_p]
lib.dmme_debug_set_conv_trace.restype = None
for rep in range(3):
    flush.fill_(rep)
    trace.zero_()
    lib.dmme_debug_set_conv_trace(trace.data_ptr())
    ops.conv2d_launch(d, wp, bias, out, None, addend, stats=st)
    torch.cuda.synchronize()
lib.dmme_debug_set_conv_trace(None)
t = trace.cpu().view(3, 512)
t0 = int(t[0, 0])
ev = []
for role, name in enumerate(("issue", "landed", "epi")):
    for i in range(512):
        v = int(t[role, i])
        if v:
            ev.append((v - t0, name, i))
ev.sort()
for dt, name, i in ev[:200]:
    tag = f"kb{i - 1}" if name == "issue" and i else (f"kb{i}" if name == "landed" else (f"tile{i // 2} {'ready' if i % 2 == 0 else 'stored'}" if name == "epi" else "start"))
    print(f"{dt:8d} clk  {name:7s} {tag}")
