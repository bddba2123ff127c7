"""Linear variance schedule (src/dmme/equations/ddpm/ddpm.py:9-21)."""
import torch
from torch import Tensor

from ..common.noise import pad


def linear_schedule(timesteps: int, start: float = 0.0001, end: float = 0.02) -> Tensor:
    r""":math:`\beta_t` for t = 0..T with :math:`\beta_0 = 0` (1-D tensor of length T + 1)."""
    return pad(torch.linspace(start, end, timesteps))


def schedule_tables(beta: Tensor):
    r"""(beta, alpha, alpha_bar) with :math:`\alpha = 1-\beta`, :math:`\bar\alpha = \mathrm{cumprod}(\alpha)`
    (src/dmme/diffusion_models/ddpm.py:44-47)."""
    alpha = 1 - beta
    return beta, alpha, torch.cumprod(alpha, dim=0)
