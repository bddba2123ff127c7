#!/usr/bin/env python
"""The honest GPU competitor: the reference's algorithm -- the oracle restatement, plain torch ops (cuDNN /
cuBLAS kernels, eager mode) -- timed on the same B200 for the bench workload: one DDPM sampling step on 256 images, fp32
(TF32 off / on) and under bf16 autocast.  Measurement only (it lives under tests/ because only tests/, smoke() and bench.py may use the oracle); nothing in the product path imports it.
usage: python tests/ref_gpu_eager.py [--batch 256] [--reps 5]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-made-easy_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402

import dmme_oracle as O  # noqa: E402
from dmme_b200.models.ddpm import UNet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    dev = "cuda"
    torch.manual_seed(0)
    sd = {k: v.to(dev) for k, v in UNet().eval().state_dict().items()}
    tables = [t.to(dev) for t in O.linear_tables(1000)]
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(a.batch, 3, 32, 32, device=dev, generator=g)
    z = torch.randn(a.batch, 3, 32, 32, device=dev, generator=g)
    t = torch.tensor([500], device=dev)

    def step():
        return O.ddpm_step(x, t, O.unet_forward(sd, x, t), z, tables)

    for name, tf32, autocast in (("fp32 (TF32 off)", False, False), ("fp32 (TF32 on)", True, False), ("bf16 autocast", True, True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                step()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        print(f"reference algorithm, torch eager on GPU, {name:16s}: {ms:8.2f} ms per step at batch {a.batch} = "
              f"{a.batch / ms:6.1f} samples/s at 1000 steps", flush=True)


if __name__ == "__main__":
    main()
