"""Cosine schedule (src/dmme/equations/iddpm/iddpm.py:6-20) and the beta table derived from it
(src/dmme/diffusion_models/iddpm.py:46-58)."""
import math

import torch

from ..common.noise import pad


def cosine_schedule(timesteps: int = 4000, offset: float = 0.008):
    r""":math:`\bar\alpha_t = f(t)/f(0)`, :math:`f(t)=\cos^2\!\big(\tfrac{t/T+s}{1+s}\tfrac{\pi}{2}\big)`, t = 0..T."""
    def f(t):
        return torch.cos((t / timesteps + offset) / (1 + offset) * math.pi / 2) ** 2

    return f(torch.arange(0, timesteps + 1)) / f(torch.tensor([0], dtype=torch.float32))


def cosine_tables(timesteps: int, offset: float = 0.008):
    """(beta, alpha, alpha_bar): beta clipped to [0, 0.999] and padded with beta_0 = 1."""
    alpha_bar = cosine_schedule(timesteps, offset)
    beta = pad(torch.clip(1 - alpha_bar[1:] / alpha_bar[:-1], 0, 0.999), value=1)
    return beta, 1 - beta, alpha_bar
