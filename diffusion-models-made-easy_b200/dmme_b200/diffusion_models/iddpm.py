"""Improved DDPM: cosine schedule and learned variance (drop-in for ``dmme.diffusion_models.IDDPM``,
src/dmme/diffusion_models/iddpm.py:16-164).  The variance interpolation uses the raw network output
``v`` (no squashing), as the reference does (equations/iddpm/losses.py:34-37)."""
from __future__ import annotations

from collections import namedtuple
from typing import Optional

import torch
from torch import nn, Tensor

from .. import ops
from ..equations import iddpm as eq_iddpm
from .ddpm import DDPM

NoiseVariance = namedtuple("NoiseVariance", ["noise", "variance"])


class IDDPM(DDPM):
    r"""Improved DDPM with cosine variance schedule and learned variance

    Args:
        model: model predicting noise and variance coefficient (2C output channels)
        timesteps: total timesteps :math:`T`
        loss_type: "hybrid", "vlb" or "simple"
        gamma: :math:`\gamma` in hybrid loss
        schedule: "linear" or "cosine"
        offset: cosine schedule offset
        start: linear variance schedule start value
        end: linear variance schedule end value
    """

    def __init__(self, model: nn.Module, timesteps: int = 1000, loss_type="hybrid", gamma=0.001,
                 schedule: str = "cosine", offset=0.008, start: float = 0.0001, end: float = 0.02) -> None:
        super().__init__(model, timesteps, start, end)
        self.loss_type = loss_type
        self.gamma = gamma
        if schedule == "cosine":
            self._register_tables(*eq_iddpm.cosine_tables(timesteps, offset))
        elif schedule != "linear":
            raise NotImplementedError

    def _update_(self, x: Tensor, model_out: Tensor, noise: Optional[Tensor], t: Tensor, seed: int) -> Tensor:
        return ops.iddpm_step_(x, model_out, noise, self.beta, self.alpha, self.alpha_bar, t, seed)
