#!/usr/bin/env python
"""Stage-by-stage check of the chain kernel (debugging aid): one op raw, one op + norm to global, two ops through the
resident operand."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-made-easy_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F

from dmme_b200 import ops, _lib as L
from test_chain_gpu import bf, nhwc, nchw, ref_norm
from helpers import rel_l2

DEV = "cuda"


def main():
    for hw, n, ipc in ((8, 1, 1), (8, 3, 2), (4, 1, 1), (4, 7, 3)):
        g = torch.Generator().manual_seed(hw + n)
        C = 256
        x = bf(torch.randn(n, C, hw, hw, generator=g))
        w = torch.randn(C, C, 3, 3, generator=g) * (9 * C) ** -0.5
        b = torch.randn(C, generator=g) * 0.1
        gam, bet = 1 + 0.2 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g)
        w2 = torch.randn(C, C, 3, 3, generator=g) * (9 * C) ** -0.5
        want = bf(F.conv2d(x, bf(w), padding=1) + b.view(1, C, 1, 1))
        wantn = ref_norm(want, gam, bet, 8, True, 1e-5)
        want2 = bf(F.conv2d(wantn, bf(w2), padding=1) + b.view(1, C, 1, 1))
        xd = nhwc(x)
        wp, wp2 = ops.pack_conv_weight(w.to(DEV), None, True), ops.pack_conv_weight(w2.to(DEV), None, True)
        bd, gd, bed = b.to(DEV), gam.to(DEV), bet.to(DEV)
        L.load().dmme_set_conv_chain_ipc(ipc)
        out = torch.full((n, hw, hw, C), float("nan"), dtype=torch.bfloat16, device=DEV)
        outn = torch.full_like(out, float("nan"))
        ops.conv_chain([ops.chain_op(xd, None, wp, bd, out=out, out_norms=[ops.out_norm(outn, gd, bed, 8, True)])], n, hw, hw)
        torch.cuda.synchronize()
        print(f"hw {hw} n {n} ipc {ipc}: one op raw {rel_l2(nchw(out), want):.3e}  norm->global {rel_l2(nchw(outn), wantn):.3e}")
        o = nchw(out)
        if rel_l2(o, want) > 1e-2:
            err = (o - want).abs()
            print("   per-image err", [float(err[i].mean()) for i in range(n)])
            print("   per-row err img0", [float(err[0, :, y].mean()) for y in range(hw)])
            print("   per-col err img0", [float(err[0, :, :, xx].mean()) for xx in range(hw)])
            print("   per-64ch err img0", [float(err[0, c:c + 64].mean()) for c in range(0, C, 64)])
        out2 = torch.full_like(out, float("nan"))
        ops.conv_chain([ops.chain_op(xd, None, wp, bd, out_norms=[ops.out_norm(None, gd, bed, 8, True)], keep=0),
                        ops.chain_op(None, None, wp2, bd, c0=C, out=out2)], n, hw, hw)
        torch.cuda.synchronize()
        print(f"   two ops through the resident operand {rel_l2(nchw(out2), want2):.3e}")
        o = nchw(out2)
        if rel_l2(o, want2) > 1e-2:
            err = (o - want2).abs()
            print("   per-image err", [float(err[i].mean()) for i in range(n)])
            print("   per-row err img0", [float(err[0, :, y].mean()) for y in range(hw)])
            print("   per-col err img0", [float(err[0, :, :, xx].mean()) for xx in range(hw)])
            print("   per-64ch err img0", [float(err[0, c:c + 64].mean()) for c in range(0, C, 64)])
    L.load().dmme_set_conv_chain_ipc(0)


if __name__ == "__main__":
    main()
