"""CPU oracle for the dmme hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, in plain functional torch fp32 on the CPU, the algorithm of the reference
(`urw7rs/diffusion-models-made-easy`, dmme 0.5.2) for the UNet denoiser and the DDPM / DDIM /
IDDPM wrappers.  It consumes a ``state_dict`` with the reference's key layout, so the same
weights drive the oracle, the real reference and the CUDA path.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker (or the timed CPU baseline).  Nothing under
``diffusion-models-made-easy_b200/`` imports it; the product path has no CPU fallback.

Pinning: the reference's own tests hold no golden vectors (shape/NaN checks only, SURVEY.md par. 4),
so the oracle is pinned against the *unmodified reference executed in the build container*:
``oracle/make_golden.py`` imports /root/reference through a stub shim, asserts that this
restatement reproduces it (bit-exact tables, <= 1e-6 relative on UNet outputs and trajectories) and
writes the vectors to ``tests/golden/``.  ``tests/test_oracle.py`` re-checks the restatement against
those committed vectors everywhere and against the live reference where /root/reference exists.

Every function cites the reference file:line it follows (paths relative to the reference checkout).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


# --------------------------------------------------------------------------------------------
# schedules (bit-exact restatements: same torch CPU ops in the same order)
# --------------------------------------------------------------------------------------------
def pad_front(x: Tensor, value: float = 0.0) -> Tensor:
    """src/dmme/common/noise.py:19-23 -- prepend one entry so index t addresses step t."""
    return torch.cat([torch.ones_like(x[0:1]) * value, x], dim=0)


def linear_tables(timesteps: int, start: float = 1e-4, end: float = 0.02) -> Tuple[Tensor, Tensor, Tensor]:
    """beta/alpha/alpha_bar of DDPM.__init__ (src/dmme/diffusion_models/ddpm.py:41-51,
    src/dmme/equations/ddpm/ddpm.py:9-21): beta_0 = 0, alpha_bar = cumprod(1 - beta)."""
    beta = pad_front(torch.linspace(start, end, timesteps))
    alpha = 1 - beta
    alpha_bar = torch.cumprod(alpha, dim=0)
    return beta, alpha, alpha_bar


def cosine_tables(timesteps: int, offset: float = 0.008) -> Tuple[Tensor, Tensor, Tensor]:
    """src/dmme/equations/iddpm/iddpm.py:6-20 and src/dmme/diffusion_models/iddpm.py:46-58."""
    def f(t):
        return torch.cos((t / timesteps + offset) / (1 + offset) * math.pi / 2) ** 2

    t = torch.arange(0, timesteps + 1)
    alpha_bar = f(t) / f(torch.tensor([0], dtype=torch.float32))
    beta = torch.clip(1 - alpha_bar[1:] / alpha_bar[:-1], 0, 0.999)
    beta = pad_front(beta, value=1)
    alpha = 1 - beta
    return beta, alpha, alpha_bar


def tau_table(timesteps: int, sub_timesteps: int, kind: str = "quadratic") -> Tensor:
    """src/dmme/equations/ddim/ddim.py:9-34 (int64, tau_0 = 0)."""
    i = torch.arange(0, sub_timesteps + 1)
    if kind == "linear":
        return torch.round((timesteps / sub_timesteps) * i).long()
    if kind == "quadratic":
        return torch.round((timesteps / (sub_timesteps ** 2)) * i ** 2).long()
    raise NotImplementedError(kind)


# --------------------------------------------------------------------------------------------
# UNet (both flavours), driven by the state_dict key layout
# --------------------------------------------------------------------------------------------
def _has(sd: StateDict, key: str) -> bool:
    return key in sd


def _gn(sd: StateDict, prefix: str, x: Tensor, groups: int) -> Tensor:
    return F.group_norm(x, groups, sd[prefix + ".weight"], sd[prefix + ".bias"], eps=1e-5)


def _conv(sd: StateDict, prefix: str, x: Tensor, stride: int = 1) -> Tensor:
    w = sd[prefix + ".weight"]
    return F.conv2d(x, w, sd[prefix + ".bias"], stride=stride, padding=w.shape[-1] // 2)


def _last_conv_key(sd: StateDict, prefix: str) -> str:
    """conv2 is Sequential(norm, act, [drop], conv): the conv sits at index 3 (p > 0) or 2 (p = 0)
    (src/dmme/models/ddpm.py:32-35; IDDPM slices off the norm but keeps the indices, models/iddpm.py:94)."""
    return prefix + (".3" if _has(sd, prefix + ".3.weight") else ".2")


def _channel_dropout(h: Tensor, mask: Optional[Tensor]) -> Tensor:
    """nn.Dropout2d as a given (N, C) keep-mask already scaled by 1/(1-p) (models/ddpm.py:29)."""
    return h if mask is None else h * mask[:, :, None, None]


def _attention(sd: StateDict, prefix: str, x: Tensor, groups: int, heads: Optional[int]) -> Tensor:
    """Attention.forward (src/dmme/models/ddpm.py:54-75) when heads is None,
    MultiHeadAttention.forward (src/dmme/models/iddpm.py:36-59) otherwise -- including its
    "(b head)" fold / "(head b)" unfold mismatch and the full-width dim**-0.5 scale."""
    b, c, hh, ww = x.shape
    h = _gn(sd, prefix + ".norm", x, groups)
    qkv = _conv(sd, prefix + ".qkv_proj", h)
    scale = c ** -0.5
    if heads is None:
        qkv = qkv.flatten(2).transpose(1, 2)  # b (h w) c
        q, k, v = qkv.chunk(3, dim=2)
        k = k.transpose(1, 2) * scale
        att = F.softmax(torch.bmm(q, k), dim=2)
        out = torch.bmm(att, v)
        out = out.transpose(1, 2).reshape(b, c, hh, ww)
    else:
        ch = 3 * c // heads
        qkv = qkv.reshape(b, heads, ch, hh * ww).permute(0, 1, 3, 2).reshape(b * heads, hh * ww, ch)
        q, k, v = qkv.chunk(3, dim=2)
        k = k.transpose(1, 2) * scale
        att = F.softmax(torch.bmm(q, k), dim=2)
        out = torch.bmm(att, v)  # (b*heads, hw, c/heads), leading index = b*heads + head
        dh = c // heads
        out = out.reshape(heads, b, hh * ww, dh)  # ... re-read as (head b)
        out = out.permute(1, 0, 3, 2).reshape(b, c, hh, ww)
    return _conv(sd, prefix + ".proj", out) + x


def _resblock(sd: StateDict, prefix: str, x: Tensor, emb: Tensor, groups: int, flavour: str,
              heads: Optional[int], mask: Optional[Tensor]) -> Tensor:
    """ResBlock.forward: src/dmme/models/ddpm.py:118-133 (ddpm), src/dmme/models/iddpm.py:106-122 (iddpm)."""
    h = _conv(sd, prefix + ".conv1.2", F.silu(_gn(sd, prefix + ".conv1.0", x, groups)))
    cond = F.linear(emb, sd[prefix + ".condition.0.weight"], sd[prefix + ".condition.0.bias"])[:, :, None, None]
    conv2 = _last_conv_key(sd, prefix + ".conv2")
    if flavour == "ddpm":
        h = h + cond
        h = F.silu(_gn(sd, prefix + ".conv2.0", h, groups))
    else:
        shift, scale = cond.chunk(2, dim=1)
        h = _gn(sd, prefix + ".norm", h, groups) * (scale + 1) + shift
        h = F.silu(h)
    h = _conv(sd, conv2, _channel_dropout(h, mask))
    if _has(sd, prefix + ".residual.weight"):
        h = h + _conv(sd, prefix + ".residual", x)
    else:
        h = h + x
    if _has(sd, prefix + ".attention.norm.weight"):
        h = _attention(sd, prefix + ".attention", h, groups, heads)
    return h


def _indices(sd: StateDict, list_name: str) -> List[int]:
    idx = set()
    for k in sd:
        if k.startswith(list_name + "."):
            idx.add(int(k.split(".")[1]))
    return sorted(idx)


def timestep_embedding(sd: StateDict, t: Tensor) -> Tensor:
    """UNet.condition: sinusoidal -> Linear -> SiLU -> Linear -> SiLU
    (src/dmme/models/ddpm.py:211-217, 338-349); `condition.0.embeddings` is the frequency buffer."""
    e = t.unsqueeze(dim=1) * sd["condition.0.embeddings"]
    e = torch.cat((e.sin(), e.cos()), dim=-1)
    e = F.silu(F.linear(e, sd["condition.1.weight"], sd["condition.1.bias"]))
    return F.silu(F.linear(e, sd["condition.3.weight"], sd["condition.3.bias"]))


def unet_forward(sd: StateDict, x: Tensor, t: Tensor, groups: int = 32, flavour: str = "ddpm",
                 heads: int = 4, dropout_masks: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """UNet.forward: src/dmme/models/ddpm.py:281-316 / src/dmme/models/iddpm.py:230-265.

    ``dropout_masks`` maps a ResBlock prefix (e.g. "down_layers.0") to an (N, C) keep-mask already
    scaled by 1/(1-p); None means eval mode.
    """
    hd = heads if flavour == "iddpm" else None
    masks = dropout_masks or {}
    emb = timestep_embedding(sd, t)
    h = _conv(sd, "input_conv", x)
    skips = [h]
    for i in _indices(sd, "down_layers"):
        p = f"down_layers.{i}"
        if _has(sd, p + ".conv1.0.weight"):
            h = _resblock(sd, p, h, emb, groups, flavour, hd, masks.get(p))
        else:
            h = _conv(sd, p, h, stride=2)  # DownSample, models/ddpm.py:136-147
        skips.append(h)
    for i in _indices(sd, "middle_layers"):
        p = f"middle_layers.{i}"
        h = _resblock(sd, p, h, emb, groups, flavour, hd, masks.get(p))
    for i in _indices(sd, "up_layers"):
        p = f"up_layers.{i}"
        if _has(sd, p + ".conv1.0.weight"):
            h = torch.cat([h, skips.pop()], dim=1)
            h = _resblock(sd, p, h, emb, groups, flavour, hd, masks.get(p))
        else:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")  # UpSample, models/ddpm.py:150-173
            h = _conv(sd, p + ".conv", h)
    h = F.silu(_gn(sd, "output_conv.0", h, groups))
    return _conv(sd, "output_conv.2", h)


# --------------------------------------------------------------------------------------------
# diffusion processes
# --------------------------------------------------------------------------------------------
def _col(table: Tensor, idx: Tensor) -> Tensor:
    """table[(T+1,)] indexed like the reference's (T+1,1,1,1) buffers: result broadcasts over C,H,W."""
    return table[idx].reshape(-1, 1, 1, 1)


def ddpm_mean(x_t: Tensor, eps: Tensor, beta_t: Tensor, alpha_t: Tensor, alpha_bar_t: Tensor) -> Tensor:
    """mean of reverse_process, src/dmme/equations/ddpm/ddpm.py:66-70."""
    return 1 / torch.sqrt(alpha_t) * (x_t - beta_t / torch.sqrt(1 - alpha_bar_t) * eps)


def ddpm_step(x_t: Tensor, t: Tensor, eps: Tensor, z: Tensor, tables: Sequence[Tensor]) -> Tensor:
    """DDPM.sampling_step given the model output and the normal draw
    (src/dmme/diffusion_models/ddpm.py:94-111; Normal.sample == z * std + mean)."""
    beta, alpha, alpha_bar = tables
    b, a, ab = _col(beta, t), _col(alpha, t), _col(alpha_bar, t)
    mean = ddpm_mean(x_t, eps, b, a, ab)
    x = z * torch.sqrt(b) + mean
    return torch.where(t.reshape(-1, 1, 1, 1) == 1, mean, x)


def ddim_step(x: Tensor, i: Tensor, eps: Tensor, alpha_bar: Tensor, tau: Tensor) -> Tensor:
    """DDIM.sampling_step as written (src/dmme/diffusion_models/ddim.py:66-77,
    src/dmme/equations/ddim/ddim.py:52-57, src/dmme/equations/ddpm/ddpm.py:36): the returned mean is
    sqrt(abar_prev) * ((x - sqrt(1 - abar_i) eps) / sqrt(abar_prev))."""
    ab_i = _col(alpha_bar, tau[i])
    ab_p = _col(alpha_bar, tau[i - 1])
    x0 = (x - torch.sqrt(1 - ab_i) * eps) / torch.sqrt(ab_p)
    return torch.sqrt(ab_p) * x0


def interpolate_variance(v: Tensor, beta_t: Tensor, beta_tilde_t: Tensor) -> Tensor:
    """src/dmme/equations/iddpm/losses.py:34-37 (v used raw)."""
    return torch.exp(v * torch.log(beta_t) + (1 - v) * torch.log(beta_tilde_t.clamp(1e-12)))


def iddpm_split(model_out: Tensor, t: Tensor, tables: Sequence[Tensor]) -> Tuple[Tensor, Tensor]:
    """IDDPM.forward_model, src/dmme/diffusion_models/iddpm.py:150-164."""
    beta, alpha, alpha_bar = tables
    eps, v = model_out.chunk(2, dim=1)
    b, ab, abp = _col(beta, t), _col(alpha_bar, t), _col(alpha_bar, t - 1)
    beta_tilde = (1 - abp) / (1 - ab) * b
    return eps, interpolate_variance(v, b, beta_tilde)


def iddpm_step(x_t: Tensor, t: Tensor, model_out: Tensor, z: Tensor, tables: Sequence[Tensor]) -> Tensor:
    """IDDPM.sampling_step, src/dmme/diffusion_models/iddpm.py:118-148."""
    beta, alpha, alpha_bar = tables
    eps, var = iddpm_split(model_out, t, tables)
    mean = ddpm_mean(x_t, eps, _col(beta, t), _col(alpha, t), _col(alpha_bar, t))
    x = z * torch.sqrt(var) + mean
    return torch.where(t.reshape(-1, 1, 1, 1) == 1, mean, x)


def forward_noising(x_0: Tensor, t: Tensor, z: Tensor, alpha_bar: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """q(x_t | x_0) sample plus the (mean, std) the losses need
    (src/dmme/equations/ddpm/ddpm.py:24-41, src/dmme/diffusion_models/ddpm.py:72-75)."""
    ab = _col(alpha_bar, t)
    mean = torch.sqrt(ab) * x_0
    std = torch.sqrt(1 - ab)
    return z * std + mean, mean, std


def ddpm_loss(x_t: Tensor, q_mean: Tensor, q_std: Tensor, eps_hat: Tensor) -> Tensor:
    """L_simple with the noise recovered from x_t (src/dmme/diffusion_models/ddpm.py:79-80)."""
    return F.mse_loss((x_t - q_mean) / q_std, eps_hat)


def _normal_cdf(x: Tensor, mean: Tensor, std: Tensor) -> Tensor:
    return 0.5 * (1 + torch.erf((x - mean) / (std * math.sqrt(2))))


def vlb_loss(eps_hat: Tensor, var: Tensor, x_t: Tensor, t: Tensor, x_0: Tensor, tables: Sequence[Tensor]) -> Tensor:
    """loss_vlb, src/dmme/equations/iddpm/losses.py:40-90, restated mask-free: the reference's masked
    gathers + cat + mean equal where(t == 1, nll, kl).mean()."""
    beta, alpha, alpha_bar = tables
    b, a, ab, abp = _col(beta, t), _col(alpha, t), _col(alpha_bar, t), _col(alpha_bar, t - 1)
    p_mean = ddpm_mean(x_t, eps_hat.detach(), b, a, ab)
    p_std = torch.sqrt(var)
    # t == 1: discrete NLL (losses.py:8-19)
    up = torch.where(x_0 < 1, _normal_cdf(x_0 + 1 / 255, p_mean, p_std), torch.ones_like(x_0))
    lo = torch.where(x_0 > -1, _normal_cdf(x_0 - 1 / 255, p_mean, p_std), torch.zeros_like(x_0))
    nll = -torch.log((up - lo).clamp(1e-12))
    # t != 1: KL(q || p) between the true posterior (losses.py:22-31) and p
    q_mean = torch.sqrt(abp) * b / (1 - ab) * x_0 + torch.sqrt(a) * (1 - abp) / (1 - ab) * x_t
    q_var = (1 - abp) / (1 - ab) * b
    is_first = (t == 1).reshape(-1, 1, 1, 1)
    q_std = torch.sqrt(torch.where(is_first, torch.ones_like(q_var), q_var))  # guarded: q_var = 0 at t = 1
    var_ratio = (q_std / p_std) ** 2
    kl = 0.5 * (var_ratio + ((q_mean - p_mean) / p_std) ** 2 - 1 - torch.log(var_ratio))
    return torch.where(is_first, nll, kl).mean()


# --------------------------------------------------------------------------------------------
# whole-trajectory helpers with injected randomness
# --------------------------------------------------------------------------------------------
@torch.no_grad()
def ddpm_generate(sd: StateDict, x_T: Tensor, noises: Sequence[Tensor], tables: Sequence[Tensor], timesteps: int,
                  groups: int = 32, steps: Optional[int] = None, return_trajectory: bool = False):
    """DDPM.generate (src/dmme/diffusion_models/ddpm.py:113-133): t runs T..1 with t of shape (1,).
    ``steps``: stop after that many steps (a prefix of the chain); ``return_trajectory``: also return every x_{t-1}."""
    x = x_T
    traj = []
    for k, t in enumerate(range(timesteps, 0, -1)):
        if steps is not None and k >= steps:
            break
        tt = torch.tensor([t])
        eps = unet_forward(sd, x, tt, groups)
        x = ddpm_step(x, tt, eps, noises[k], tables)
        if return_trajectory:
            traj.append(x.clone())
    return (x, traj) if return_trajectory else x


@torch.no_grad()
def ddim_generate(sd: StateDict, x_T: Tensor, alpha_bar: Tensor, tau: Tensor, groups: int = 32,
                  return_trajectory: bool = False):
    """DDIM.generate (src/dmme/diffusion_models/ddim.py:79-99): i runs S..1, model evaluated at tau_i."""
    x = x_T
    traj = []
    for i in range(tau.numel() - 1, 0, -1):
        ii = torch.tensor([i])
        eps = unet_forward(sd, x, tau[ii], groups)
        x = ddim_step(x, ii, eps, alpha_bar, tau)
        if return_trajectory:
            traj.append(x.clone())
    return (x, traj) if return_trajectory else x


# ---------------------------------------------------------------------------------------------------------------------
# classifier guidance (src/dmme/guidance/classifier.py) -- the reference module is unimportable (it needs dmme.ddpm / dmme.ddim,
# which do not exist), so this restatement follows the file as written and is NOT pinned against a run of the reference:
# PARITY UNPINNED for this part
# ---------------------------------------------------------------------------------------------------------------------
def classifier_grad(classifier, y: Tensor, x_t: Tensor, t: Tensor) -> Tensor:
    """classifier.py:9-23: grad of log_softmax(classifier(x_t, t))[:, y].sum() w.r.t. x_t (a (B, B) selection, as written)."""
    x = x_t.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        log_probs = F.log_softmax(classifier(x, t.float()), dim=1)
        (g,) = torch.autograd.grad(log_probs[:, y].sum(), x)
    return g


def guided_ddpm_sample(model, classifier, y: Tensor, x_t: Tensor, t: Tensor, noise: Tensor, tables: Sequence[Tensor],
                       scale: float = 10.0) -> Tensor:
    """classifier.py:26-36: ancestral step (per-sample t), then x += scale * classifier_grad at the new x."""
    beta, alpha, alpha_bar = tables
    with torch.no_grad():
        x = ddpm_mean(x_t, model(x_t, t), _col(beta, t), _col(alpha, t), _col(alpha_bar, t)) + torch.sqrt(_col(beta, t)) * noise
    return x + scale * classifier_grad(classifier, y, x, t)


def guided_ddim_sample(model, classifier, y: Tensor, x_t: Tensor, t: Tensor, alpha_bar: Tensor, scale: float = 10.0) -> Tensor:
    """classifier.py:39-63: eps_hat = eps - sqrt(1 - abar_t) scale grad, then the eta = 0 DDIM update written there."""
    ab_p, ab_t = _col(alpha_bar, t - 1), _col(alpha_bar, t)
    g = classifier_grad(classifier, y, x_t, t)
    with torch.no_grad():
        eps = model(x_t, t) - torch.sqrt(1 - ab_t) * scale * g
        return torch.sqrt(ab_p) * (x_t - torch.sqrt(1 - ab_t) * eps) / torch.sqrt(ab_t) + torch.sqrt(1 - ab_p) * eps


# ---------------------------------------------------------------------------------------------------------------------
# optimizer tail (test infrastructure for dmme_b200.optim): Adam + WarmupLR + gradient clipping + EMA as the reference
# composes them
# ---------------------------------------------------------------------------------------------------------------------
class WarmupLR(torch.optim.lr_scheduler._LRScheduler):
    """Restatement of lr_scheduler/warmup.py:4-19: linear ramp over ``warmup`` optimizer steps, evaluated on
    ``optimizer._step_count + 1``."""

    def __init__(self, optimizer, warmup=0.0, last_epoch=-1):
        self.warmup_steps = warmup
        # the reference was written against torch 1.13, whose _LRScheduler counts the optimizer's steps in
        # ``optimizer._step_count``; newer torch dropped that attribute, so the training loop below keeps it
        if not hasattr(optimizer, "_step_count"):
            optimizer._step_count = 0
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        steps = self.optimizer._step_count + 1
        if steps < self.warmup_steps:
            return [g["initial_lr"] * (steps / self.warmup_steps) for g in self.optimizer.param_groups]
        return [g["initial_lr"] for g in self.optimizer.param_groups]


def optimizer_tail_reference(params, grads_per_step, lr, warmup, max_norm, decay):
    """Run the reference's training tail on ``params`` (modified in place): per step clip_grad_norm_ (Lightning
    gradient_clip_val, configs/ddpm/cifar10.yaml:24), Adam.step (lit_modules/ddpm.py:130), WarmupLR.step (:131-133) and
    the EMA update of callbacks/ema.py:169-176.  Returns (ema tensors, optimizer, list of lr used per step, norms)."""
    opt = torch.optim.Adam(params, lr=lr)
    sched = WarmupLR(opt, warmup)
    ema = [p.detach().clone() for p in params]
    lrs, norms = [], []
    for grads in grads_per_step:
        for p, g in zip(params, grads):
            p.grad = g.clone()
        if max_norm:
            norms.append(torch.nn.utils.clip_grad_norm_(params, max_norm).detach().clone())
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        opt._step_count = len(lrs)  # what torch 1.13's step wrapper did
        sched.step()
        with torch.no_grad():
            torch._foreach_mul_(ema, decay)
            torch._foreach_add_(ema, [p.detach() for p in params], alpha=(1.0 - decay))
    return ema, opt, lrs, norms
