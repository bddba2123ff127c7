"""Sub-sequence schedules tau (src/dmme/equations/ddim/ddim.py:9-34); int64, tau_0 = 0."""
import torch
from torch import Tensor


def _steps(sub_timesteps: int) -> Tensor:
    return torch.arange(0, sub_timesteps + 1)


def linear_tau(timesteps: int, sub_timesteps: int) -> Tensor:
    return torch.round((timesteps / sub_timesteps) * _steps(sub_timesteps)).long()


def quadratic_tau(timesteps: int, sub_timesteps: int) -> Tensor:
    return torch.round((timesteps / (sub_timesteps ** 2)) * _steps(sub_timesteps) ** 2).long()
