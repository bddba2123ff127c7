#!/usr/bin/env python
"""torchrun check of sharded sampling: N ranks each denoise their slice (no communication until the final gather);
rank 0 compares the gathered batch with the same generate() run unsharded on one GPU.  x_T and every step's noise are
bit-identical by construction (Philox keyed by the global element index); the UNet output of an image depends on its
position in the batch only through the fp32 rounding of the GroupNorm partial sums (tile boundaries move with the
image index), which 20 bf16 steps amplify to ~1e-4: the bar is 2e-3, not bit equality.
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-made-easy_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from dmme_b200 import DDPM, ops, parallel  # noqa: E402
from dmme_b200.models.ddpm import UNet  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = UNet(pos_dim=32, emb_dim=64, channels_per_depth=(64, 128), num_blocks=1).eval()
    ddpm = DDPM(model, timesteps=20).to(dev)
    size, seed = (10, 3, 32, 32), 1234
    got = parallel.generate_sharded(ddpm, size, seed=seed, dst=0)
    if rank == 0:
        x_T = ops.philox_normal(size, seed, 1 << 62, dev)
        want = ddpm.generate(size, x_T=x_T, seed=seed)
        err = float((got - want).norm() / want.norm())
        print(f"sharded x{dist.get_world_size()} vs unsharded: rel-L2 {err:.3e}, bit-identical {bool(torch.equal(got, want))}")
        assert err < 2e-3
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
