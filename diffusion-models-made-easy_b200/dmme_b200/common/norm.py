"""``norm`` / ``denorm`` of the reference (common/norm.py) for CUDA image batches."""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from .. import _lib as L


def norm(x: Tensor) -> Tensor:
    r"""Normalize input to :math:`[-1, 1]` linearly (common/norm.py:4-6; dataset-side, plain torch)."""
    return (x - 0.5) * 2


def denorm(x: Tensor, out: Optional[Tensor] = None) -> Tensor:
    r"""Denormalize input normalized to :math:`[-1, 1]` linearly back to :math:`[0, 1]` (common/norm.py:9-11):
    ``clip((x + 1) / 2, 0, 1)`` in one fused pass, bit-exact."""
    L.require_cuda(x, out)
    x = x.detach().float().contiguous()
    y = torch.empty_like(x) if out is None else out
    L.check(L.load().dmme_denorm(x.data_ptr(), y.data_ptr(), None, x.numel(), L.stream_ptr()), "denorm")
    return y


def denorm_uint8(x: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """``round(255 * denorm(x))`` as uint8, same shape: the form image loggers and FID feature extractors consume."""
    L.require_cuda(x, out)
    x = x.detach().float().contiguous()
    y = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if out is None else out
    L.check(L.load().dmme_denorm(x.data_ptr(), None, y.data_ptr(), x.numel(), L.stream_ptr()), "denorm")
    return y
