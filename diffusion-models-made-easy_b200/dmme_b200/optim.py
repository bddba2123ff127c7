"""Fused optimizer tail of the reference's training recipe.

The reference trains with ``Adam(lr)`` + ``WarmupLR(warmup)`` (lit_modules/ddpm.py:128-134, lr_scheduler/warmup.py:4-19),
Lightning's ``gradient_clip_val: 1.0`` (configs/ddpm/cifar10.yaml:24: ``clip_grad_norm_`` over all parameters) and the EMA
callback ``ema = decay * ema + (1 - decay) * w`` after every optimizer step (callbacks/ema.py:169-176).  ``FusedAdamEMA``
runs all four as two multi-tensor CUDA launches (csrc/optim.cu) over the parameters' own storage: the ``state_dict``
layout of the model is untouched, Adam moments and the EMA copy live in flat fp32 arenas owned by the optimizer.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch
from torch import Tensor

from . import _lib as L


def warmup_lr(base_lr: float, step: int, warmup: float) -> float:
    """Learning rate of the ``step``-th optimizer step (1-based) under the reference's ``WarmupLR``:
    ``base_lr * step / warmup`` while ``step < warmup``, ``base_lr`` afterwards (lr_scheduler/warmup.py:10-19; the
    scheduler evaluates ``optimizer._step_count + 1`` right after the previous step)."""
    if warmup and step < warmup:
        return base_lr * (step / warmup)
    return base_lr


class FusedAdamEMA:
    """Adam + WarmupLR + global-norm gradient clipping + EMA in one pass over the parameters.

    Args:
        params: parameters to optimise (CUDA, fp32, contiguous)
        lr, betas, eps: as ``torch.optim.Adam`` (no weight decay, no amsgrad: the reference uses the defaults)
        warmup: ``WarmupLR`` steps (0 disables the ramp)
        max_grad_norm: Lightning ``gradient_clip_val`` (``None`` / 0 disables clipping)
        ema_decay: EMA decay (``None`` disables the EMA copy)
    """

    def __init__(self, params: Iterable[Tensor], lr: float = 2e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 warmup: float = 0.0, max_grad_norm: Optional[float] = 1.0, ema_decay: Optional[float] = 0.9999) -> None:
        self.params: List[Tensor] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedAdamEMA got no trainable parameters")
        L.require_cuda(*self.params)
        for p in self.params:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise ValueError("FusedAdamEMA needs contiguous fp32 parameters")
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.warmup = float(warmup)
        self.max_grad_norm = float(max_grad_norm) if max_grad_norm else 0.0
        self.ema_decay = ema_decay
        self.step_count = 0
        dev = self.params[0].device
        lib = L.load()
        self._chunk = lib.dmme_optim_chunk()
        if lib.dmme_optim_table_entry_bytes() != 56:
            raise RuntimeError("dmme_b200: optimizer table layout mismatch")
        total = sum(p.numel() for p in self.params)
        # moments (and the EMA copy) in flat arenas: one allocation each, per-parameter views for inspection
        self._m = torch.zeros(total, dtype=torch.float32, device=dev)
        self._v = torch.zeros(total, dtype=torch.float32, device=dev)
        self._ema = torch.empty(total, dtype=torch.float32, device=dev) if ema_decay is not None else None
        self.exp_avg, self.exp_avg_sq, self.ema = [], [], []
        off = 0
        for p in self.params:
            n = p.numel()
            self.exp_avg.append(self._m[off:off + n].view_as(p))
            self.exp_avg_sq.append(self._v[off:off + n].view_as(p))
            if self._ema is not None:
                e = self._ema[off:off + n].view_as(p)
                e.copy_(p.detach())
                self.ema.append(e)
            off += n
        self._items = sum((p.numel() + self._chunk - 1) // self._chunk for p in self.params)
        self._grid = min(max(1, self._items), 148 * 8)
        self._partial = torch.zeros(self._grid, dtype=torch.float32, device=dev)
        self._norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self._table_host = torch.zeros(len(self.params) * 7, dtype=torch.int64).pin_memory()
        self._table_dev = torch.zeros(len(self.params) * 7, dtype=torch.int64, device=dev)
        self._grad_ptrs: Optional[List[int]] = None

    # ------------------------------------------------------------------------------------------
    def _refresh_table(self) -> None:
        ptrs = [p.grad.data_ptr() if p.grad is not None else 0 for p in self.params]
        if not any(ptrs):
            # e.g. zero_grad(set_to_none=True) after a graph-captured backward rebinds p.grad away from the captured buffers
            raise RuntimeError("FusedAdamEMA.step(): no parameter has a gradient -- nothing would be updated")
        if ptrs == self._grad_ptrs:
            return
        t = self._table_host.view(-1, 7)
        item = 0
        for i, p in enumerate(self.params):
            g = p.grad
            if g is not None and (g.dtype != torch.float32 or not g.is_contiguous() or g.device != p.device):
                raise ValueError("FusedAdamEMA needs contiguous fp32 gradients on the parameter's device")
            t[i, 0] = p.data_ptr()
            t[i, 1] = ptrs[i]
            t[i, 2] = self.exp_avg[i].data_ptr()
            t[i, 3] = self.exp_avg_sq[i].data_ptr()
            t[i, 4] = self.ema[i].data_ptr() if self.ema else 0
            t[i, 5] = p.numel()
            t[i, 6] = item
            item += (p.numel() + self._chunk - 1) // self._chunk
        self._table_dev.copy_(self._table_host, non_blocking=True)
        self._grad_ptrs = ptrs

    @torch.no_grad()
    def step(self) -> None:
        """One optimizer step on the current ``p.grad`` values (parameters without a gradient are left untouched)."""
        self._refresh_table()
        self.step_count += 1
        lr = warmup_lr(self.lr, self.step_count, self.warmup)
        with torch.cuda.device(self._table_dev.device):
            L.check(L.load().dmme_adam_ema_step(self._table_dev.data_ptr(), len(self.params), self._items, lr,
                                                self.betas[0], self.betas[1], self.eps, self.step_count, self.max_grad_norm,
                                                self.ema_decay if self.ema_decay is not None else 0.0,
                                                self._partial.data_ptr(), self._grid, self._norm.data_ptr(), L.stream_ptr()),
                    "adam_ema_step")
        # the kernels write the weights through raw pointers: tell the executors' packed-weight caches (models/_engine.py
        # keys them on (data_ptr, _version, _dmme_gen)) that every parameter changed
        for p in self.params:
            p._dmme_gen = self.step_count

    def zero_grad(self, set_to_none: bool = False) -> None:
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    @property
    def grad_norm(self) -> Tensor:
        """Global L2 norm of the gradients seen by the last ``step`` (before clipping); a device scalar."""
        return self._norm

    def swap_ema(self) -> None:
        """Exchange the weights with their EMA copy in place (evaluate / checkpoint with EMA weights, then swap back),
        like ``EMAOptimizer.swap_model_weights`` of callbacks/ema.py."""
        if self._ema is None:
            raise RuntimeError("FusedAdamEMA was built without an EMA copy")
        with torch.no_grad():
            for p, e in zip(self.params, self.ema):
                tmp = p.detach().clone()
                p.copy_(e)
                e.copy_(tmp)

    def state_dict(self) -> Dict:
        return {"step": self.step_count, "exp_avg": self._m.clone(), "exp_avg_sq": self._v.clone(),
                "ema": self._ema.clone() if self._ema is not None else None}

    def load_state_dict(self, sd: Dict) -> None:
        self.step_count = int(sd["step"])
        self._m.copy_(sd["exp_avg"])
        self._v.copy_(sd["exp_avg_sq"])
        if self._ema is not None and sd.get("ema") is not None:
            self._ema.copy_(sd["ema"])
