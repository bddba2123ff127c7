#!/usr/bin/env python
"""Per-role clock64 timeline of CTA 0 of the chain kernel (two 256->256 ResBlock-like ops + one op with streamed chunks).
usage: python tools/trace_chain.py [hw] [n]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-made-easy_b200"))
import torch

from dmme_b200 import ops, _lib as L

DEV = "cuda"


def main():
    hw = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    Cc = 256
    lib = L.load()
    lib.dmme_debug_set_chain_trace.argtypes = [C.c_void_p]
    lib.dmme_debug_set_chain_trace.restype = None
    x = torch.randn(n, hw, hw, Cc, device=DEV).to(torch.bfloat16)
    sk = torch.randn(n, hw, hw, Cc, device=DEV).to(torch.bfloat16)
    w = [ops.pack_conv_weight(torch.randn(Cc, Cc, 3, 3, device=DEV) * 0.02, None, True) for _ in range(2)]
    w2 = ops.pack_conv_weight(torch.randn(Cc, 2 * Cc, 3, 3, device=DEV) * 0.02, None, True)
    b = torch.zeros(Cc, device=DEV)
    g = torch.ones(Cc, device=DEV)
    out = torch.empty_like(x)
    chain = [ops.chain_op(x, None, w[0], b, out_norms=[ops.out_norm(None, g, b, 8, True)], keep=0),
             ops.chain_op(None, None, w[1], b, c0=Cc, addend=x, out=out, out_norms=[ops.out_norm(None, g, b, 16, True)], keep=0),
             ops.chain_op(None, sk, w2, b, c0=Cc, out_norms=[ops.out_norm(None, g, b, 8, True)], keep=0),
             ops.chain_op(None, None, w[1], b, c0=Cc, out=out)]
    for _ in range(3):
        ops.conv_chain(chain, n, hw, hw)
    torch.cuda.synchronize()
    tr = torch.zeros(6 * 1024, dtype=torch.int64, device=DEV)
    lib.dmme_debug_set_chain_trace(tr.data_ptr())
    ops.conv_chain(chain, n, hw, hw)
    torch.cuda.synchronize()
    lib.dmme_debug_set_chain_trace(None)
    t = tr.view(6, 1024).cpu()
    names = ["w issue", "w landed", "epi start", "epi end", "chunk issue", "chunk landed"]
    ev = []
    for r in range(6):
        for i in range(1024):
            if t[r, i]:
                ev.append((int(t[r, i]), names[r], i))
    ev.sort()
    t0 = ev[0][0]
    last = {}
    for tt, nm, i in ev:
        if nm.startswith("w ") and not (i < 24 or i % 36 == 0):
            continue
        print(f"{tt - t0:8d} clk  {nm:12s} {i}")
    wl = [int(t[1, i]) for i in range(1024) if t[1, i]]
    d = [b - a for a, b in zip(wl, wl[1:])]
    print("tiles", len(wl), "median clk between landed tiles", sorted(d)[len(d) // 2], "mean", sum(d) / len(d))


if __name__ == "__main__":
    main()
