#!/usr/bin/env python
"""Per-role timeline of CTA 0 of the halo conv kernel (clock64 timestamps written by the kernel when a trace buffer is
set): per chunk, when the stage was free and its loads issued, when the tile landed, when the fused GroupNorm transform
finished, when the MMA warp saw it and when it had issued the chunk's last tap; per unit, epilogue start / end.
usage: python tools/trace_halo.py [c0 c1 cout h gn(0/1) rc0 rc1]"""
import ctypes as C
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-made-easy_b200"))
import torch  # noqa: E402

from dmme_b200 import _lib as L  # noqa: E402
from dmme_b200 import ops  # noqa: E402

c0, c1, cout, h, gn, rc0, rc1 = (int(a) for a in (sys.argv[1:8] + ["128", "0", "128", "32", "1", "0", "0"][len(sys.argv) - 1:]))
res = rc0 + rc1 > 0
n, dev = 256, "cuda"
cin = c0 + c1
g = torch.Generator(device=dev).manual_seed(0)
s0 = torch.randn(n, h, h, c0, device=dev, generator=g).bfloat16()
s1 = torch.randn(n, h, h, c1, device=dev, generator=g).bfloat16() if c1 else None
w = torch.randn(cout, cin, 3, 3, device=dev, generator=g) / math.sqrt(9 * cin)
r0 = torch.randn(n, h, h, rc0, device=dev, generator=g).bfloat16() if rc0 else None
r1 = torch.randn(n, h, h, rc1, device=dev, generator=g).bfloat16() if rc1 else None
wr = torch.randn(cout, rc0 + rc1, 1, 1, device=dev, generator=g) / math.sqrt(rc0 + rc1) if res else None
wp = ops.pack_conv_weight(w, wr, True)
bias = torch.randn(cout, device=dev, generator=g)
d = ops.make_conv_desc(s0, s1, cout, 3, 1, False, r0, r1, False, L.OUT_NHWC,
                       torch.bfloat16, L.CONV_HALO)
out = torch.empty(n, h, h, cout, device=dev, dtype=torch.bfloat16)
ab = torch.randn(n, cin, 2, device=dev, generator=g) if gn else None
st = torch.zeros(n * cout // 4 * 2, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
trace = torch.zeros(7 * 256, dtype=torch.int64, device=dev)
lib = L.load()
lib.dmme_debug_set_halo_trace.argtypes = [C.c_void_p]
lib.dmme_debug_set_halo_trace.restype = None
for rep in range(3):
    flush.fill_(rep)
    trace.zero_()
    lib.dmme_debug_set_halo_trace(trace.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv2d_launch(d, wp, bias, out, None, None, stats=st, gn_ab=ab)
    e1.record()
    torch.cuda.synchronize()
lib.dmme_debug_set_halo_trace(None)
print(f"# conv 3x3 {cin}->{cout} @{h} gn={gn} res={rc0 + rc1}: {e0.elapsed_time(e1) * 1e3:.1f} us; clock64 ticks relative to the first event")
t = trace.cpu().view(7, 256)
t0 = int(t[t > 0].min())
nchunks = int((t[0] > 0).sum())
print("chunk   free    landed  xf_done  mma_saw  mma_issued | load   xf     wait   mma_issue")
for i in range(nchunks):
    fr, ld, xd, ms, mi = (int(t[r, i]) - t0 if int(t[r, i]) else -1 for r in range(5))
    print(f"{i:4d} {fr:8d} {ld:8d} {xd:8d} {ms:8d} {mi:8d} | {ld - fr:6d} {xd - ld:6d} {ms - xd:6d} {mi - ms:6d}")
print("unit  acc_full  epi_done  epi")
for i in range(int((t[5] > 0).sum())):
    a, b = int(t[5, i]) - t0, int(t[6, i]) - t0
    print(f"{i:4d} {a:8d} {b:8d} {b - a:6d}")
