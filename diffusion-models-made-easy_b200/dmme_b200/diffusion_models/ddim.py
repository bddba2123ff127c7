"""Deterministic DDIM sampling over a sub-sequence tau (drop-in for ``dmme.diffusion_models.DDIM``,
src/dmme/diffusion_models/ddim.py:16-99).  The update is the reference's formula *as written*
(equations/ddim/ddim.py:52-57): x <- sqrt(abar_prev) * ((x - sqrt(1 - abar_i) eps) / sqrt(abar_prev)),
there is no eta and no noise term."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn, Tensor

from .. import ops
from ..equations import ddim as eq_ddim
from .ddpm import DDPM


class DDIM(DDPM):
    r"""Denoising Diffusion Implicit Models

    Args:
        model: model passed to :code:`DDPM`
        timesteps: total timesteps :math:`T`
        sub_timesteps: sub-sequence length
        tau_schedule: tau schedule to use, "linear" or "quadratic"
    """

    tau: Tensor

    def __init__(self, model: nn.Module, timesteps: int = 1000, sub_timesteps: int = 50,
                 tau_schedule: str = "quadratic") -> None:
        super().__init__(model, timesteps)
        self.sub_timesteps = sub_timesteps
        kind = tau_schedule.lower()
        if kind == "linear":
            tau = eq_ddim.linear_tau(timesteps, sub_timesteps)
        elif kind == "quadratic":
            tau = eq_ddim.quadratic_tau(timesteps, sub_timesteps)
        else:
            raise NotImplementedError
        self.register_buffer("tau", tau, persistent=False)

    def sampling_step(self, x_tau_i: Tensor, i: Tensor, noise: Optional[Tensor] = None) -> Tensor:
        r"""Mean of :math:`p_\theta(x_{\tau_{i-1}}|x_{\tau_i})`; ``i`` has shape (1,)."""
        i = self._check_step_index(i, self.tau.shape[0])
        with torch.cuda.device(self.beta.device):
            x = x_tau_i.detach().to(self.beta.device).float().contiguous().clone()
            t_model = ops.gather_i64(self.tau, i, torch.empty(1, dtype=torch.int64, device=x.device))
            return self._denoise_(x, t_model, i, None, 0)

    _sampler_kind = 2  # dmme_b200._lib.SAMPLER_DDIM

    def _update_(self, x: Tensor, model_out: Tensor, noise: Optional[Tensor], i: Tensor, seed: int) -> Tensor:
        return ops.ddim_step_(x, model_out, self.alpha_bar, self.tau, i)

    def _sampler_epilogue(self, x: Tensor, noise: Optional[Tensor], i: Tensor, seed: int):
        return ops.sampler_epilogue(self._sampler_kind, x, i, self.alpha_bar, None, None, self.tau)

    def _counter_start(self) -> int:
        return self.sub_timesteps

    def _num_steps(self) -> int:
        return self.sub_timesteps

    def _graph_step(self, x: Tensor, counter: Tensor, seed: int) -> None:
        t_model = self.__dict__.get("_t_model")
        if t_model is None or t_model.device != x.device:
            t_model = torch.empty(1, dtype=torch.int64, device=x.device)  # stable address across graph replays
            self.__dict__["_t_model"] = t_model
        ops.gather_i64(self.tau, counter, t_model)
        self._denoise_(x, t_model, counter, None, seed)
        ops.add_i64_(counter, -1)
