"""RNG entry points and schedule padding (src/dmme/common/noise.py:4-23).  These are host-side
plumbing around torch's generators; per-step sampling noise inside captured graphs comes from the
Philox kernel in csrc/sampler.cu instead."""
import torch


def gaussian(shape, dtype=None, device=None):
    """Standard normal tensor of the given shape."""
    return torch.randn(shape, dtype=dtype, device=device)


def gaussian_like(x):
    """Standard normal tensor shaped like ``x``."""
    return torch.randn_like(x)


def uniform_int(min, max, count=1, device=None):
    """``count`` integers uniform in [min, max) -- ``max`` is exclusive, so training never draws
    t = T (reference behaviour)."""
    return torch.randint(min, max, size=(count,), device=device)


def pad(x: torch.Tensor, value: float = 0) -> torch.Tensor:
    """Prepends one entry of ``value`` so that index t addresses step t."""
    head = torch.ones_like(x[0:1]) * value
    return torch.cat([head, x], dim=0)
