"""Graph-captured training step.

The eager training step issues ~950 kernel launches from Python (forward tape, fused loss, backward tape); at ~30 us of
host time per launch the host, not the GPU, bounds it (IDDPM default UNet, batch 128: 38.5 ms eager against 33.6 ms of
GPU work).  ``GraphedTrainingStep`` captures forward + loss + backward once into a CUDA graph and replays it on a static
input buffer; the optimizer (``optim.FusedAdamEMA``: two launches whose scalars change every step) runs eagerly after the
replay.  Random draws inside the captured region (timesteps, noise, dropout masks) come from torch's CUDA generator,
which advances its Philox offset per replay, so every step sees fresh draws exactly as in eager mode.
Data parallel (``parallel.enable_gradient_sync``): the bucketed NCCL all-reduces of the flat gradient arena are issued from
inside the backward pass and are captured with it (NCCL >= 2.9.6 collectives are graph-capturable); each replay then runs
backward kernels and all-reduces concurrently on the captured stream fork, with no host work per bucket.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
from torch import Tensor


class GraphedTrainingStep:
    """``loss = step(x_0)``: one ``diffusion.training_step(x_0)`` + ``loss.backward()`` (+ ``optimizer.step()``).

    Args:
        diffusion: a ``dmme_b200`` ``DDPM`` / ``IDDPM`` module in training mode, on a CUDA device
        optimizer: optional optimizer with ``step()`` / ``zero_grad(set_to_none=...)`` (e.g. ``optim.FusedAdamEMA``)
        warmup: eager iterations before capture (they size every workspace and populate the gradient buffers)
    """

    def __init__(self, diffusion, optimizer=None, warmup: int = 2) -> None:
        self.diffusion = diffusion
        self.optimizer = optimizer
        self.warmup = max(1, int(warmup))
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._x: Optional[Tensor] = None
        self._loss: Optional[Tensor] = None
        self._grads = []  # (parameter, gradient buffer inside the graph's private pool) of the captured backward
        self._pack = None  # ops.PackBatch replayed at the head of the graph (keeps its table and tensors alive)
        self.batch_packs = os.environ.get("DMME_BATCH_PACKS", "1") != "0"

    def _params(self):
        return [p for p in self.diffusion.parameters() if p.requires_grad]

    def _zero(self) -> None:
        for p in self._params():
            p.grad = None

    def _capture(self, x_0: Tensor) -> None:
        synced = getattr(self.diffusion.model.train_engine, "grad_sync", None) is not None
        self._x = x_0.detach().clone()
        side = torch.cuda.Stream(device=x_0.device)
        side.wait_stream(torch.cuda.current_stream(x_0.device))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self._zero()
                self.diffusion.training_step(self._x).backward()
        torch.cuda.current_stream(x_0.device).wait_stream(side)
        self._zero()  # gradients are (re)allocated inside the graph's private pool: stable addresses across replays
        self._graph = torch.cuda.CUDAGraph()
        eng = self.diffusion.model.train_engine
        eng.always_repack = True  # the bf16 weight packing must be replayed too: the weights change between replays
        # ... as ONE launch at the head of the graph for every tensor-core pack the warm-up made (172 pack launches, 0.9 ms of
        # the step, when each conv re-packs its own); the few other derived buffers still rebuild in place in the graph
        self._pack = eng.pack_batch() if self.batch_packs else None
        try:
            # thread-local capture mode with collectives inside: NCCL's watchdog thread polls CUDA events concurrently, which
            # the default (global) mode would treat as a capture violation
            with torch.cuda.graph(self._graph, capture_error_mode="thread_local" if synced else "global"):
                if self._pack is not None:
                    self._pack.launch()
                self._loss = self.diffusion.training_step(self._x)
                self._loss.backward()
        finally:
            eng.always_repack = False
            eng._batched_packs = set()
        self._grads = [(p, p.grad) for p in self._params() if p.grad is not None]

    def __call__(self, x_0: Tensor) -> Tensor:
        if not x_0.is_cuda:
            raise RuntimeError("dmme_b200 training runs on CUDA only; there is no CPU path")
        if self._graph is None or x_0.shape != self._x.shape or x_0.device != self._x.device:
            self._capture(x_0)
        self._x.copy_(x_0, non_blocking=True)
        self._graph.replay()
        # the replay writes into the buffers captured above whatever p.grad points to now: a caller-side
        # zero_grad(set_to_none=True) (or anything else that rebinds p.grad) must not silently disconnect the optimizer
        for p, g in self._grads:
            if p.grad is not g:
                p.grad = g
        if self.optimizer is not None:
            self.optimizer.step()
        return self._loss
