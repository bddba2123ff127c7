"""Classifier guidance (SURVEY par. 8f-4; semantics of src/dmme/guidance/classifier.py:8-63).

The reference module cannot be imported (it needs ``dmme.ddpm`` / ``dmme.ddim``, which do not exist: SURVEY quirk 8), so
there is nothing to run against: **parity is unpinned by the reference itself**.  What is reproduced is the arithmetic *as
written* in that file, including its calling convention -- the pre-refactor one its own test uses
(tests/test_guidance.py:75-89): the sampler owns only the schedule, ``model`` and ``classifier`` are call arguments, and
``t`` has shape (B,).

* ``classifier_grad`` (classifier.py:9-23): gradient w.r.t. ``x_t`` of ``log_softmax(classifier(x_t, t))[:, y].sum()``.
  ``log_probs[:, y]`` with ``y`` of shape (B,) selects a (B, B) block -- every image against every label -- so image i
  receives ``sum_j grad log p(y_j | x_i)``; kept as written.
* DDPM (classifier.py:26-36): ancestral step, then ``x += scale * grad`` evaluated at the NEW x.
* DDIM (classifier.py:39-63): ``eps_hat = eps - sqrt(1 - abar_t) * scale * grad``, then the textbook eta = 0 update
  ``sqrt(abar_{t-1}) (x - sqrt(1 - abar_t) eps_hat) / sqrt(abar_t) + sqrt(1 - abar_{t-1}) eps_hat``.

The gradient "through the classifier" is taken by autograd: when the classifier (or the noise model) is a ``dmme_b200``
UNet, its autograd node runs the explicit CUDA backward kernels (data-gradient convs on tcgen05, GroupNorm / attention
backward; ``models/_train.py``) and returns the input gradient; any other ``nn.Module`` goes through torch.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import Tensor

from ..equations import ddim as eq_ddim
from ..equations import ddpm as eq_ddpm


def _col(table: Tensor, t: Tensor) -> Tensor:
    return table.to(t.device)[t].reshape(-1, 1, 1, 1)


class ClassifierMixin:
    def classifier_grad(self, classifier, y: Tensor, x_t: Tensor, t: Tensor) -> Tensor:
        """``d/dx_t  log_softmax(classifier(x_t, t))[:, y].sum()`` (classifier.py:9-23)."""
        x_in = x_t.detach().requires_grad_(True)
        t_in = t.float()
        with torch.enable_grad():
            logits = classifier(x_in, t_in)
            log_probs = F.log_softmax(logits.float(), dim=1)
            log_probs_of_y = log_probs[:, y]
            (grad,) = torch.autograd.grad(log_probs_of_y.sum(), x_in)
        return grad.detach()


class _Schedule:
    """The schedule tables the pre-refactor ``DDPM(timesteps)`` owned: linear betas, padded at t = 0."""

    def __init__(self, timesteps: int, start: float = 0.0001, end: float = 0.02) -> None:
        self.timesteps = timesteps
        beta, alpha, alpha_bar = eq_ddpm.schedule_tables(eq_ddpm.linear_schedule(timesteps, start, end))
        self.beta, self.alpha, self.alpha_bar = beta.flatten(), alpha.flatten(), alpha_bar.flatten()


class ClassifierGuidedDDPM(_Schedule, ClassifierMixin):
    def __init__(self, timesteps: int = 1000, guidance_scale: float = 10.0) -> None:
        super().__init__(timesteps)
        self.scale = guidance_scale

    @torch.no_grad()
    def reverse_process(self, model, x_t: Tensor, t: Tensor, noise: Tensor) -> Tensor:
        """x_{t-1} = 1/sqrt(alpha_t) (x_t - beta_t / sqrt(1 - abar_t) eps) + sqrt(beta_t) noise (equations/ddpm/ddpm.py:44-72)."""
        beta_t, alpha_t, alpha_bar_t = _col(self.beta, t), _col(self.alpha, t), _col(self.alpha_bar, t)
        eps = model(x_t, t)
        mean = 1 / torch.sqrt(alpha_t) * (x_t - beta_t / torch.sqrt(1 - alpha_bar_t) * eps)
        return mean + torch.sqrt(beta_t) * noise

    def sample(self, model, classifier, y: Tensor, x_t: Tensor, t: Tensor, noise: Tensor) -> Tensor:
        x = self.reverse_process(model, x_t, t, noise)
        return x + self.scale * self.classifier_grad(classifier, y, x, t)


class ClassifierGuidedDDIM(_Schedule, ClassifierMixin):
    def __init__(self, timesteps: int, tau_schedule: str = "quadratic", guidance_scale: float = 10.0) -> None:
        super().__init__(timesteps)
        self.scale = guidance_scale
        kind = tau_schedule.lower()
        if kind not in ("linear", "quadratic"):
            raise NotImplementedError
        self.tau_schedule = kind
        self.tau = (eq_ddim.linear_tau if kind == "linear" else eq_ddim.quadratic_tau)(timesteps, timesteps)

    def reverse_process(self, model, classifier, y: Tensor, x_t: Tensor, t: Tensor) -> Tensor:
        abar_prev, abar_t = _col(self.alpha_bar, t - 1), _col(self.alpha_bar, t)
        grad = self.classifier_grad(classifier, y, x_t, t)
        with torch.no_grad():
            eps = model(x_t, t) - torch.sqrt(1 - abar_t) * self.scale * grad
            return torch.sqrt(abar_prev) * (x_t - torch.sqrt(1 - abar_t) * eps) / torch.sqrt(abar_t) + torch.sqrt(1 - abar_prev) * eps

    def sample(self, model, classifier, y: Tensor, x_t: Tensor, t: Tensor) -> Tensor:
        return self.reverse_process(model, classifier, y, x_t, t)
