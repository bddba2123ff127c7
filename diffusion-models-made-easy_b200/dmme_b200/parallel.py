"""Multi-GPU plumbing for the two shardings the hot path has (SURVEY par. 8e), one process per GPU over
``torch.distributed`` (NCCL over NVLink on the B200 box; the same code runs on ``gloo`` for the CPU tests).

* Sampling shards by image: every rank denoises its own slice of the batch with a replica of the weights and
  **no data-path collective**; one gather of the finished samples at the end (``generate_sharded``).  Per-step
  noise is keyed by (seed, t, global element index), so the gathered result does not depend on the world size.
* Training is data parallel: replicated model, per-rank batch, gradients averaged by one all-reduce per bucket.
  ``GradBucketer`` works on the flat fp32 gradient arena the training executor fills in backward order: a bucket is
  a contiguous slice, so it is all-reduced in place (no flatten/unflatten copies) as soon as the backward pass has
  moved past it, overlapping the reduction with the rest of the backward kernels.

The reference itself has no distributed code (SURVEY par. 2.2: Lightning's DDP would add bucketed NCCL all-reduce).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

Tensor = torch.Tensor


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of ``total`` items owned by ``rank``; the first ``total % world`` ranks get one extra."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_shards(local: Tensor, total: int, group=None, dst: Optional[int] = None) -> Optional[Tensor]:
    """Concatenates the per-rank slices (dim 0, sizes from ``shard_bounds``) into the full batch.
    ``dst=None``: every rank gets the result (all_gather); otherwise only rank ``dst`` does (gather)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    padded = local
    if local.shape[0] != biggest:  # ragged last shards: pad to a common size, trim after the collective
        padded = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    padded = padded.contiguous()
    if dst is None:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
    else:
        parts = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
        dist.gather(padded, parts, dst=dst, group=group)
        if rank != dst:
            return None
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)


@torch.no_grad()
def generate_sharded(diffusion, img_size: Sequence[int], *, seed: int, x_T: Optional[Tensor] = None, group=None,
                     dst: Optional[int] = None, graph: bool = True) -> Optional[Tensor]:
    """``diffusion.generate(img_size)`` with the sample batch split over the ranks of ``group``.

    Args:
        diffusion: a ``dmme_b200`` DDPM / DDIM / IDDPM module on this rank's GPU
        img_size: (N, C, H, W) of the WHOLE batch
        seed: Philox seed shared by all ranks (x_T and the per-step noise are functions of the global element index)
        x_T: optional whole-batch starting noise (every rank passes the same tensor); default: Philox normals
        dst: rank that receives the gathered samples (None: all ranks)
    """
    from . import ops

    n, c, h, w = (int(v) for v in img_size)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(n, rank, world)
    per_image = c * h * w
    dev = diffusion.beta.device
    if x_T is not None:
        x0 = x_T[lo:hi].to(dev).float().contiguous()
    else:
        # stream id 2^62 is never a timestep: x_T does not collide with any step's noise
        x0 = ops.philox_normal((hi - lo, c, h, w), seed, 1 << 62, dev, noise_offset=lo * per_image)
    local = diffusion.generate((hi - lo, c, h, w), x_T=x0, seed=seed, graph=graph, noise_offset=lo * per_image)
    return gather_shards(local, n, group, dst)


class GradBucketer:
    """In-place bucketed all-reduce (average) over a flat gradient arena that is filled front to back."""

    def __init__(self, arena: Tensor, group=None, bucket_bytes: int = 32 << 20) -> None:
        if arena.dim() != 1:
            raise ValueError("the gradient arena must be one flat tensor")
        self.arena = arena
        self.group = group
        self.bucket_elems = max(1, bucket_bytes // arena.element_size())
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self._start = 0
        self._handles: List = []
        self.buckets: List[Tuple[int, int]] = []  # (lo, hi) element ranges reduced so far (for tests / logging)

    def _launch(self, lo: int, hi: int) -> None:
        if hi <= lo:
            return
        self.buckets.append((lo, hi))
        if self.world > 1:
            self._handles.append(dist.all_reduce(self.arena[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def mark(self, cursor: int) -> None:
        """The arena is final up to ``cursor``: launch every bucket that is now complete."""
        while cursor - self._start >= self.bucket_elems:
            self._launch(self._start, self._start + self.bucket_elems)
            self._start += self.bucket_elems

    def finish(self, cursor: int) -> None:
        """Reduce the tail, wait for all buckets and turn the sums into means."""
        self.mark(cursor)
        self._launch(self._start, cursor)
        self._start = cursor
        for h in self._handles:
            h.wait()
        self._handles.clear()
        if self.world > 1:
            self.arena[:cursor].div_(self.world)


def allreduce_gradients(params: Sequence[Tensor], group=None, bucket_bytes: int = 32 << 20) -> None:
    """Average ``p.grad`` over the ranks for modules whose gradients do not live in an arena (flatten, reduce,
    scatter back).  The training executor path uses ``GradBucketer`` directly instead."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    b = GradBucketer(flat, group, bucket_bytes)
    b.finish(flat.numel())
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def enable_gradient_sync(unet, group=None, bucket_mb: int = 32) -> None:
    """Data-parallel training of a ``dmme_b200`` UNet: from now on its backward pass averages the parameter gradients
    over the ranks of ``group`` (bucketed all-reduce overlapped with the backward kernels).  Call after
    ``torch.distributed.init_process_group``; every rank must hold the same weights."""
    unet.train_engine.grad_sync = {"group": group, "bucket_bytes": int(bucket_mb) << 20}


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank ``src``'s weights and buffers."""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
