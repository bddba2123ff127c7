#!/usr/bin/env python
"""Times one training step (forward + loss + backward [+ Adam]) of the default UNets on synthetic data.
usage: python tools/prof_train.py [--flavour iddpm|ddpm] [--batch 128] [--reps 3] [--precision bf16]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "diffusion-models-made-easy_b200"))
import torch  # noqa: E402

from dmme_b200 import DDPM, IDDPM, ops  # noqa: E402
from dmme_b200.models import ddpm, iddpm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--flavour", default="iddpm")
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--dropout", type=float, default=None)
    ap.add_argument("--graph", action="store_true", help="capture forward + loss + backward in one CUDA graph and replay it")
    ap.add_argument("--optim", default="fused", choices=["fused", "torch"],
                    help="fused: dmme_b200.optim.FusedAdamEMA (clip + Adam + warm-up + EMA, two launches); "
                         "torch: clip_grad_norm_ + torch.optim.Adam + foreach EMA")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    kw = {} if a.dropout is None else {"dropout": a.dropout}
    if a.flavour == "iddpm":
        dm = IDDPM(iddpm.UNet(precision=a.precision, **kw)).to(dev).train()
    else:
        dm = DDPM(ddpm.UNet(precision=a.precision, **kw)).to(dev).train()
    if world > 1:
        from dmme_b200 import parallel
        parallel.broadcast_parameters(dm)
        parallel.enable_gradient_sync(dm.model)
    params = [p for p in dm.parameters() if p.requires_grad]
    if a.optim == "fused":
        from dmme_b200.optim import FusedAdamEMA
        opt = FusedAdamEMA(params, lr=2e-4, warmup=5000, max_grad_norm=1.0, ema_decay=0.9999)
    else:
        opt = torch.optim.Adam(params, lr=2e-4)
        ema = [p.detach().clone() for p in params]
    torch.manual_seed(100 + rank)  # every rank trains on its own shard of the (synthetic) batch
    x0 = (torch.rand(a.batch, 3, 32, 32, device=dev) * 2 - 1)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    if a.graph:
        # forward + loss + backward replayed from one CUDA graph (dmme_b200.training), optimizer eager after the replay
        from dmme_b200.training import GraphedTrainingStep
        step = GraphedTrainingStep(dm, None)
        step._capture(x0)
        for r in range(a.reps + 1):
            ev[0].record()
            loss = step(x0)
            ev[2].record()
            opt.step()
            ev[3].record()
            torch.cuda.synchronize()
            if rank == 0:
                print(f"[graph] rep {r}: loss {float(loss.detach()):.5f}  fwd+bwd {ev[0].elapsed_time(ev[2]):8.2f} ms  "
                      f"optim({a.optim}) {ev[2].elapsed_time(ev[3]):6.2f} ms", flush=True)
        return
    for r in range(a.reps + 1):
        opt.zero_grad(set_to_none=True)
        ops.reset_launch_count()
        ev[0].record()
        loss = dm.training_step(x0)
        ev[1].record()
        loss.backward()
        ev[2].record()
        if a.optim == "fused":
            opt.step()
        else:
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            with torch.no_grad():
                torch._foreach_mul_(ema, 0.9999)
                torch._foreach_add_(ema, [p.detach() for p in params], alpha=1e-4)
        ev[3].record()
        torch.cuda.synchronize()
        if rank == 0:
          print(f"[{world} rank(s)] rep {r}: loss {float(loss):.5f}  fwd {ev[0].elapsed_time(ev[1]):8.2f} ms  bwd {ev[1].elapsed_time(ev[2]):8.2f} ms  "
              f"optim({a.optim}) {ev[2].elapsed_time(ev[3]):6.2f} ms  launches {ops.launch_count()}  "
              f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    if world > 1:
        w0 = next(dm.parameters()).detach().flatten()[:1000].double().sum()
        ws = [torch.zeros_like(w0) for _ in range(world)]
        dist.all_gather(ws, w0)
        if rank == 0:
            print("weights identical across ranks after the steps:", all(float(v) == float(ws[0]) for v in ws),
                  "| buckets:", len(dm.model.train_engine.last_buckets))
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
