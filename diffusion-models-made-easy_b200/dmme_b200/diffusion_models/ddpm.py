"""DDPM sampling (and, later, training) on the CUDA hot path.

Drop-in for ``dmme.diffusion_models.DDPM`` (src/dmme/diffusion_models/ddpm.py:15-144): same
constructor, same non-persistent buffers ``beta / alpha / alpha_bar`` of shape (T+1, 1, 1, 1), same
``sampling_step / generate / forward`` entry points.  ``generate`` captures ONE denoising step
(timestep embedding -> UNet -> fused sampler update -> step-counter decrement) in a CUDA graph whose
kernels read ``t`` and the schedule scalars from device memory, and replays it T times: the
reference's 1000-iteration host loop (ddpm.py:130-131) runs with no host work between steps.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
from torch import nn, Tensor

from .. import ops
from ..common.noise import gaussian, uniform_int
from ..equations import ddpm as eq_ddpm


class DDPM(nn.Module):
    r"""Training and Sampling for DDPM

    Args:
        model: model predicting noise from data, :math:`\epsilon_\theta(x_t, t)`
        timesteps: total timesteps :math:`T`
        start: linear variance schedule start value
        end: linear variance schedule end value
    """

    beta: Tensor
    alpha: Tensor
    alpha_bar: Tensor

    def __init__(self, model: nn.Module, timesteps: int = 1000, start: float = 0.0001, end: float = 0.02) -> None:
        super().__init__()
        self.model = model
        self.timesteps = timesteps
        beta, alpha, alpha_bar = eq_ddpm.schedule_tables(eq_ddpm.linear_schedule(timesteps, start, end))
        self._register_tables(beta, alpha, alpha_bar)

    def _register_tables(self, beta: Tensor, alpha: Tensor, alpha_bar: Tensor) -> None:
        for name, tab in (("beta", beta), ("alpha", alpha), ("alpha_bar", alpha_bar)):
            self.register_buffer(name, tab.reshape(-1, 1, 1, 1).contiguous(), persistent=False)

    # ------------------------------------------------------------------------------------------
    def forward(self, x: Tensor, t: Tensor) -> Tensor:
        """Applies the internal model: :math:`\\epsilon_\\theta(x, t)`."""
        return self.model(x, t)

    def _check_step_index(self, t: Tensor, table_len: Optional[int] = None) -> Tensor:
        if t.numel() != 1:
            # the reference itself fails here (torch.where broadcast, ddpm.py:110)
            raise ValueError("sampling_step takes a step tensor of shape (1,), as DDPM.generate passes it")
        # eager entry point: validate on the host like the reference's `self.beta[t]` (negative indices wrap, anything past
        # the table raises IndexError); the graph-replayed chain owns its counter and needs no check
        n = int(table_len if table_len is not None else self.beta.shape[0])
        tv = int(t.reshape(-1)[0].item())
        if not -n <= tv < n:
            raise IndexError(f"index {tv} is out of bounds for dimension 0 with size {n}")
        return t.reshape(1).to(device=self.beta.device, dtype=torch.int64).contiguous()

    def _model_out(self, x: Tensor, t: Tensor) -> Tensor:
        """eps (or (eps, v)) of the wrapped model: the dmme_b200 UNet's raw executor output, or -- for any other
        ``nn.Module`` (a guidance wrapper, the reference UNet, ...) as the reference allows -- ``self.model(x, t)``."""
        raw = getattr(self.model, "forward_raw", None)
        if raw is not None:
            return raw(x, t)
        with torch.no_grad():
            return self.model(x, t).detach().float().contiguous()

    def _update_(self, x: Tensor, model_out: Tensor, noise: Optional[Tensor], t: Tensor, seed: int) -> Tensor:
        return ops.ddpm_step_(x, model_out, noise, self.beta, self.alpha, self.alpha_bar, t, seed,
                              getattr(self, "_noise_offset", 0))

    _sampler_kind = 1  # dmme_b200._lib.SAMPLER_DDPM

    def _sampler_epilogue(self, x: Tensor, noise: Optional[Tensor], t: Tensor, seed: int):
        """The same update as ``_update_`` as an epilogue of the UNet's output conv (eps never leaves the registers)."""
        return ops.sampler_epilogue(self._sampler_kind, x, t, self.alpha_bar, self.beta, self.alpha, None, noise, seed,
                                    getattr(self, "_noise_offset", 0))

    def _denoise_(self, x: Tensor, t_model: Tensor, t: Tensor, noise: Optional[Tensor], seed: int) -> Tensor:
        """x_t -> x_{t-1} in place: the network at ``t_model`` followed by this sampler's update at step ``t``."""
        raw = getattr(self.model, "forward_raw", None)
        if raw is not None and x.shape[1] * x.shape[2] * x.shape[3] % 4 == 0:
            spec = self._sampler_epilogue(x, noise, t, seed)  # keeps raw pointers: alive until the launch below returned
            out = raw(x, t_model, sampler=spec)
            if self.model.engine.sampler_applied:
                return x
        else:
            out = self._model_out(x, t_model)
        return self._update_(x, out, noise, t, seed)

    def sampling_step(self, x_t: Tensor, t: Tensor, noise: Optional[Tensor] = None) -> Tensor:
        r"""Denoise by sampling from :math:`p_\theta(x_{t-1}|x_t)`.

        Args:
            x_t: image of shape (N, C, H, W)
            t: the step, tensor of shape (1,)
            noise: optional standard-normal tensor shaped like ``x_t`` (default: ``torch.randn_like``)
        """
        t = self._check_step_index(t)
        with torch.cuda.device(self.beta.device):
            x = x_t.detach().to(self.beta.device).float().contiguous().clone()
            if noise is None:
                noise = torch.randn_like(x)
            return self._denoise_(x, t, t, noise.to(x.device).float().contiguous(), 0)

    # ------------------------------------------------------------------------------------------
    def _counter_start(self) -> int:
        return self.timesteps

    def _num_steps(self) -> int:
        return self.timesteps

    def _graph_step(self, x: Tensor, counter: Tensor, seed: int) -> None:
        """One denoising step on device-resident state; everything launched here is graph-capturable."""
        self._denoise_(x, counter, counter, None, seed)
        ops.add_i64_(counter, -1)

    def _run_steps(self, x: Tensor, steps: int, seed: int, graph: bool,
                   on_step: Optional[Callable[[int, Tensor], None]] = None,
                   pre_step: Optional[Callable[[int, Tensor], None]] = None) -> Tensor:
        dev = x.device
        counter = torch.full((1,), self._counter_start(), dtype=torch.int64, device=dev)
        if not graph:
            for k in range(steps):
                if pre_step is not None:
                    pre_step(k, x)
                self._graph_step(x, counter, seed)
                if on_step is not None:
                    on_step(k, x)
            return x
        # warm-up outside capture: packs weights, sizes the workspace, sets kernel attributes
        x_saved = x.clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._graph_step(x, counter, seed)
        torch.cuda.current_stream(dev).wait_stream(side)
        x.copy_(x_saved)
        counter.fill_(self._counter_start())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._graph_step(x, counter, seed)
        # capture does not execute: state is still (x_T, T)
        for k in range(steps):
            if pre_step is not None:
                pre_step(k, x)
            g.replay()
            if on_step is not None:
                on_step(k, x)
        return x

    @torch.no_grad()
    def generate(self, img_size: Tuple[int, int, int, int], *, x_T: Optional[Tensor] = None, seed: Optional[int] = None,
                 graph: bool = True, on_step: Optional[Callable[[int, Tensor], None]] = None,
                 noise_offset: int = 0) -> Tensor:
        """Generate images of shape (N, C, H, W) by running the full denoising chain.

        Args:
            img_size: (N, C, H, W)
            x_T: optional starting noise (default ``dmme_b200.gaussian(img_size)`` like the reference)
            seed: Philox seed of the per-step noise (default: drawn from torch's CPU generator)
            graph: replay one captured CUDA graph per step (default) or launch eagerly
            on_step: optional callback ``(k, x)`` after each step (k = 0 is t = T)
            noise_offset: element index of this batch inside a larger sharded batch (``parallel.generate_sharded``)
        """
        dev = self.beta.device
        if dev.type != "cuda":
            raise RuntimeError("dmme_b200 samplers run on CUDA only; move the module with .cuda()")
        x = gaussian(tuple(img_size), device=dev) if x_T is None else x_T.detach().to(dev).float().clone()
        x = x.contiguous()
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        self._noise_offset = int(noise_offset)
        try:
            with torch.cuda.device(dev):
                return self._run_steps(x, self._num_steps(), seed, graph, on_step)
        finally:
            self._noise_offset = 0

    @staticmethod
    def history_timesteps(timesteps: int, vis_length: int) -> list:
        """The ``save_t`` list of ``GenerateImage.generate_img`` (callbacks/generate.py:73-76), same integer arithmetic."""
        return [int(timesteps / (vis_length - 1) * i) for i in range(vis_length - 1, 0, -1)]

    @torch.no_grad()
    def generate_history(self, img_size: Tuple[int, int, int, int], vis_length: int = 20, *,
                         x_T: Optional[Tensor] = None, seed: Optional[int] = None, graph: bool = True,
                         uint8: bool = False) -> Tensor:
        """The denoising sequence ``GenerateImage.generate_img`` collects (callbacks/generate.py:64-90): ``denorm(x_t)``
        at every ``t`` in ``save_t`` *before* the step at ``t`` runs, plus ``denorm`` of the final sample, stacked as
        ``(len, N, C, H, W)`` in [0, 1] (or uint8 0..255 with ``uint8=True``).  The chain itself is the graph-replayed
        ``generate`` loop; each snapshot is one fused denorm launch between two replays, written straight into the
        history tensor (the reference clones ``x_t`` and runs three elementwise kernels per snapshot)."""
        from ..common.norm import denorm, denorm_uint8
        dev = self.beta.device
        if dev.type != "cuda":
            raise RuntimeError("dmme_b200 samplers run on CUDA only; move the module with .cuda()")
        steps, start = self._num_steps(), self._counter_start()
        save_t = set(self.history_timesteps(start, vis_length))
        hits = [t for t in range(start, start - steps, -1) if t in save_t]
        hist = torch.empty((len(hits) + 1,) + tuple(img_size), dtype=torch.uint8 if uint8 else torch.float32, device=dev)
        put = denorm_uint8 if uint8 else denorm
        slot = {t: i for i, t in enumerate(hits)}

        def tap(k: int, x: Tensor) -> None:
            t = start - k
            if t in slot:
                put(x, out=hist[slot[t]])

        x = gaussian(tuple(img_size), device=dev) if x_T is None else x_T.detach().to(dev).float().clone()
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        with torch.cuda.device(dev):
            x = self._run_steps(x.contiguous(), steps, seed, graph, None, tap)
            put(x, out=hist[len(hits)])
        return hist

    def _noised(self, x_0: Tensor, t: Optional[Tensor] = None, noise: Optional[Tensor] = None):
        """t ~ U{1..T-1} (randint excludes T), x_t ~ q(x_t | x_0); RNG order as the reference:
        randint, then one normal draw (diffusion_models/ddpm.py:65-75).  ``t`` / ``noise`` inject fixed draws
        (Normal.sample() == noise * std + mean)."""
        if not x_0.is_cuda:
            raise RuntimeError("dmme_b200 training runs on CUDA only; there is no CPU path")
        x_0 = x_0.detach().float().contiguous()
        if t is None:
            t = uniform_int(1, self.timesteps, x_0.size(0), device=x_0.device)
        t = t.to(device=x_0.device, dtype=torch.int64).contiguous()
        alpha_bar_t = self.alpha_bar[t]
        mean, std = torch.sqrt(alpha_bar_t) * x_0, torch.sqrt(1 - alpha_bar_t)
        if noise is None:
            # == torch.normal(mean, std.expand_as(mean)) draw for draw (ATen: out.normal_().mul_(std).add_(mean)) without
            # that op's host-side `std >= 0` check, which synchronises and cannot be captured in a CUDA graph
            x_t = torch.empty_like(mean).normal_().mul_(std).add_(mean)
        else:
            x_t = noise.to(x_0.device) * std + mean
        return x_0, t, x_t.contiguous(), mean, std

    def training_step(self, x_0: Tensor, *, t: Optional[Tensor] = None, noise: Optional[Tensor] = None) -> Tensor:
        r"""Training step except for optimization

        Args:
            x_0: image from dataset

        Returns:
            loss, :math:`L_\text{simple}` (0-dim tensor; ``loss.backward()`` runs the explicit backward kernels)
        """
        with torch.cuda.device(x_0.device if x_0.is_cuda else self.beta.device):
            x_0, t, x_t, mean, std = self._noised(x_0, t, noise)
            noise_in_x_t = self.model(x_t, t)
            noise = ((x_t - mean) / std).contiguous()
            return _SimpleLoss.apply(noise_in_x_t, noise)


class _SimpleLoss(torch.autograd.Function):
    """L_simple = mse_loss(noise, eps) (equations/ddpm/losses.py:5-13): value and gradient from one fused kernel."""

    @staticmethod
    def forward(ctx, eps: Tensor, noise: Tensor) -> Tensor:
        eps = eps.contiguous()
        d_eps = torch.empty_like(eps)
        loss = ops.mse_loss(eps, noise, d_eps)
        ctx.save_for_backward(d_eps)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (d_eps,) = ctx.saved_tensors
        return d_eps * g, None
