#!/usr/bin/env python
"""clock64 timeline of the one-launch attention block (csrc/attention_block.cu), leader CTA of cluster 0: when the worker
warps enter / leave each phase and when the MMA thread's waits complete.  usage: python tools/trace_attn_block.py [batch]"""
import ctypes as C
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-made-easy_b200"))
import torch  # noqa: E402

from dmme_b200 import _lib as L  # noqa: E402
from dmme_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
hw = int(sys.argv[3]) if len(sys.argv) > 3 else 16   # 4: the 16-token kernel (events are numbered, see attention_block.cu)
dev, c, seq = "cuda", 256, hw * hw
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(n, hw, hw, c, device=dev, generator=g).bfloat16()
ab = torch.stack([1 + 0.1 * torch.randn(n, c, device=dev, generator=g), 0.1 * torch.randn(n, c, device=dev, generator=g)], dim=-1).contiguous()
wqkv = ops.pack_conv_weight(torch.randn(3 * c, c, 1, 1, device=dev, generator=g) / math.sqrt(c), None, True)
wproj = ops.pack_conv_weight(torch.randn(c, c, 1, 1, device=dev, generator=g) / math.sqrt(c), None, True)
bqkv, bproj = torch.randn(3 * c, device=dev, generator=g), torch.randn(c, device=dev, generator=g)
out = torch.empty_like(x)
st = torch.zeros(n * c // 4 * 2, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
trace = torch.zeros(3 * 256, dtype=torch.int64, device=dev)
lib = L.load()
lib.dmme_debug_set_attn_block_trace.argtypes = [C.c_void_p]
lib.dmme_debug_set_attn_block_trace.restype = None
for rep in range(3):
    if len(sys.argv) > 2:
        flush.fill_(rep)
    trace.zero_()
    lib.dmme_debug_set_attn_block_trace(trace.data_ptr())
    ops.attention_block(x, ab, wqkv, bqkv, wproj, bproj, c ** -0.5, out, st)
    torch.cuda.synchronize()
lib.dmme_debug_set_attn_block_trace(None)
t = trace.cpu().view(3, 256)
if hw == 4:
    W16 = ["start", "x landed", "H done", "K product done", "Q, K drained", "V^T done", "V^T drained", "S done", "P written",
           "O done", "O drained", "out^T done", "stored"]
    M16 = ["H ready", "Q issued", "K issued", "Q, K drained -> V, S", "V issued", "P ready -> P V", "O drained -> proj", "proj issued"]
    ev = [(int(t[0, i]), "worker " + W16[i]) for i in range(13) if int(t[0, i])] + [(int(t[1, i]), "   mma " + M16[i]) for i in range(8) if int(t[1, i])]
    ev.sort()
    for v, name in ev:
        print(f"{v - ev[0][0]:8d} clk  {name}")
    sys.exit(0)
WN = ["x landed", "H done", "Q ready", "Q drained", "K ready", "K drained", "V ready", "V drained", "S ready", "P done",
      "O ready", "O drained", "D ready", "out stored"]
MN = ["H+Wq -> Q GEMM", "Wk -> K GEMM", "Q drained -> V GEMM", "Wv kb0", "Wv kb1", "Wv kb2", "Wv kb3", "K drained -> S GEMM",
      "P ready -> PV GEMM", "O drained+Wp -> proj GEMM"]
DN = ["Q ready", "Q chunk 0", "Q chunk 1", "Q chunk 2", "Q chunk 3", "epilogue: issue residual loads", "loads issued", "D ready",
      "out chunk 0", "out chunk 1", "out chunk 2", "out chunk 3", "", "", "", ""]
ev = []
for i in range(256):
    if int(t[0, i]):
        ev.append((int(t[0, i]), f"worker img{i // 14} {WN[i % 14]}"))
    if int(t[1, i]):
        ev.append((int(t[1, i]), f"   mma img{i // 10} {MN[i % 10]}"))
    if int(t[2, i]):
        ev.append((int(t[2, i]), f"      .. img{i // 16} {DN[i % 16]}"))
ev.sort()
t0 = ev[0][0]
prev = t0
for v, name in ev:
    print(f"{v - t0:8d} clk (+{v - prev:6d})  {name}")
    prev = v
