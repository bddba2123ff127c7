#!/usr/bin/env python
"""In-stream timing of every launch of one sampling step (eager, same stream, warm L2 as inside the real step): CUDA
events around each ops.* call, grouped by signature.  Complements the ncu launch list (cold, serialised, boost clocks).
Event pairs add ~2 us per launch and disable the programmatic-dependent-launch overlap, so small launches read high.
usage: python tools/prof_step.py [--batch 256] [--reps 5]"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-made-easy_b200"))
import torch  # noqa: E402

from dmme_b200 import DDPM, IDDPM, ops  # noqa: E402
from dmme_b200.models import _engine  # noqa: E402
from dmme_b200.models import ddpm as m_ddpm, iddpm as m_iddpm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--flavour", default="ddpm", choices=["ddpm", "iddpm"])
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
    # the whole step replayed from a CUDA graph (what generate() runs)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        ddpm._graph_step(x, counter, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        g.replay()
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"# {args.flavour} sampling step, batch {args.batch}: {e0.elapsed_time(e1) / 20:.3f} ms per graph replay")
    counter.fill_(1000)

    records = []

    def wrap(name, fn, sig):
        def timed(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            records.append((sig(*a, **k), e0, e1))
            return out
        return timed

    def conv_sig(desc, weight, bias, out, temb=None, addend=None, out2=None, out3=None, stats=None, gn_ab=None, gn_silu=True,
                 splitk_ws=None, out_norms=None, sampler=None):
        ho, wo = ops.conv_out_hw(desc)
        return (f"conv {desc.ksize}x{desc.ksize} s{desc.stride} {desc.c0 + desc.c1}->{desc.cout} @{desc.h_in}"
                + (f" +res{desc.rc0 + desc.rc1}" if desc.rc0 + desc.rc1 else "") + (" +add" if addend is not None else "")
                + (" qkv" if out2 is not None else "") + (" +gn" if gn_ab is not None else "")
                + (f" splitk+{len(out_norms or [])}norms" if splitk_ws is not None else "") + (" +sampler" if sampler is not None else ""))

    def gn_sig(src0, src1, *a, **k):
        c = src0.shape[3] + (src1.shape[3] if src1 is not None else 0)
        return f"gn {c} @{src0.shape[1]}"

    def attn_sig(q, k, v, n, heads, seq, dh, *a, **kw):
        return f"attn L{seq} d{dh}"

    patches = [("conv2d_launch", conv_sig), ("groupnorm", gn_sig), ("attention", attn_sig),
               ("upsample2x", lambda x, out=None: f"upsample {x.shape[3]} @{x.shape[1]}"),
               ("temb_mlp", lambda *a, **k: "temb_mlp"), ("temb_proj", lambda *a, **k: "temb_proj"),
               ("ddpm_step_", lambda *a, **k: "ddpm_step"), ("add_i64_", lambda *a, **k: "add_i64")]
    saved = {n: getattr(ops, n) for n, _ in patches}
    for n, s in patches:
        setattr(ops, n, wrap(n, saved[n], s))
    best = None
    try:
        for _ in range(args.reps):
            records.clear()
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record()
            ddpm._graph_step(x, counter, 1)
            w1.record()
            torch.cuda.synchronize()
            agg = collections.OrderedDict()
            for sig, a, b in records:
                g = agg.setdefault(sig, [0, 0.0])
                g[0] += 1
                g[1] += a.elapsed_time(b) * 1e3
            tot = sum(g[1] for g in agg.values())
            if best is None or tot < best[0]:
                best = (tot, agg, w0.elapsed_time(w1) * 1e3)
    finally:
        for n in saved:
            setattr(ops, n, saved[n])
    tot, agg, wall = best
    print(f"# batch {args.batch}: sum of launches {tot:.1f} us, eager step wall {wall:.1f} us, {sum(g[0] for g in agg.values())} launches")
    for sig, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us:8.1f} us  x{cnt:2d}  avg {us / cnt:7.1f}  {sig}")


if __name__ == "__main__":
    main()
