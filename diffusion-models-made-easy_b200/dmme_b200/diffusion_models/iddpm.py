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
        return ops.iddpm_step_(x, model_out, noise, self.beta, self.alpha, self.alpha_bar, t, seed,
                               getattr(self, "_noise_offset", 0))

    _sampler_kind = 3  # dmme_b200._lib.SAMPLER_IDDPM

    def forward_model(self, x_t: Tensor, t: Tensor, beta_t: Tensor, alpha_bar_t: Tensor,
                      alpha_bar_t_minus_one: Tensor) -> NoiseVariance:
        """(noise, variance) with the interpolated learned variance (diffusion_models/iddpm.py:150-164)."""
        noise, v = self.model(x_t, t).chunk(2, dim=1)
        beta_tilde_t = (1 - alpha_bar_t_minus_one) / (1 - alpha_bar_t) * beta_t
        variance = torch.exp(v * torch.log(beta_t) + (1 - v) * torch.log(beta_tilde_t.clamp(1e-12)))
        return NoiseVariance(noise, variance)

    def training_step(self, x_0: Tensor, *, t: Optional[Tensor] = None, noise: Optional[Tensor] = None) -> Optional[Tensor]:
        r"""Hybrid loss :math:`L_\text{simple} + \gamma L_\text{vlb}` (or :math:`L_\text{vlb}` alone for
        ``loss_type="vlb"``), diffusion_models/iddpm.py:62-116.  The whole loss tail (learned-variance
        interpolation, discrete NLL at t = 1, KL elsewhere, MSE) and its gradient run as one fused kernel."""
        if self.loss_type == "hybrid":
            w_simple, w_vlb = 1.0, float(self.gamma)
        elif self.loss_type == "vlb":
            w_simple, w_vlb = 0.0, 1.0
        else:
            w_simple = None  # the reference falls off the end of training_step for any other loss_type (after the forward)
        with torch.cuda.device(x_0.device if x_0.is_cuda else self.beta.device):
            x_0, t, x_t, _, _ = self._noised(x_0, t, noise)
            model_out = self.model(x_t, t)
            if w_simple is None:
                return None
            return _HybridLoss.apply(model_out, x_t, x_0, t, self.beta, self.alpha, self.alpha_bar, w_simple, w_vlb)


class _HybridLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model_out, x_t, x_0, t, beta, alpha, alpha_bar, w_simple, w_vlb):
        model_out = model_out.contiguous()
        d_out = torch.empty_like(model_out)
        loss = ops.iddpm_loss(model_out, x_t.contiguous(), x_0, t, beta, alpha, alpha_bar, w_simple, w_vlb, d_out)
        ctx.save_for_backward(d_out)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (d_out,) = ctx.saved_tensors
        return (d_out * g,) + (None,) * 8
