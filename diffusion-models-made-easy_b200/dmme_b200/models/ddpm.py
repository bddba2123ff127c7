"""DDPM UNet noise predictor on hand-written sm_100a kernels.

Drop-in for ``dmme.models.ddpm.UNet`` (src/dmme/models/ddpm.py:176-316): same constructor arguments,
same ``forward(x, c)`` contract, identical ``state_dict`` keys / shapes / initialisation order, so
``load_state_dict(reference_unet.state_dict())`` works in both directions and
``torch.manual_seed(s); UNet()`` yields the reference's weights.  The sub-modules defined here only
hold parameters; the arithmetic runs in ``_engine.Engine`` through the C-ABI library.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional, Sequence

import torch
from torch import nn, Tensor

from ._engine import Engine
from ._train import TrainEngine

_PRECISIONS = {"bf16": torch.bfloat16, "fp32": torch.float32}


class SinusoidalPositionEmbeddings(nn.Module):
    """Holds the persistent frequency buffer ``embeddings`` of shape (1, dim // 2)
    (src/dmme/models/ddpm.py:327-336); evaluated inside ``dmme_temb_mlp_fwd``."""

    embeddings: Tensor

    def __init__(self, dim: int) -> None:
        super().__init__()
        half = dim // 2
        step = math.log(10000) / (half - 1)
        self.register_buffer("embeddings", torch.exp(torch.arange(half) * -step).unsqueeze(dim=0))


def _norm_act_conv(c_in: int, c_out: int, num_groups: int, p: float, drop_norm: bool = False) -> nn.Sequential:
    """Parameter container with the reference's Sequential indices: 0 GroupNorm, 1 SiLU,
    [2 Dropout2d iff p > 0], last Conv2d 3x3 (src/dmme/models/ddpm.py:25-35).  ``drop_norm`` keeps the
    indices but omits entry 0, as the ``[1:]`` slice of src/dmme/models/iddpm.py:94 does."""
    entries = [("0", nn.GroupNorm(num_groups, c_in)), ("1", nn.SiLU())]
    if p > 0:
        entries.append(("2", nn.Dropout2d(p)))
    entries.append((str(len(entries)), nn.Conv2d(c_in, c_out, kernel_size=3, stride=1, padding=1)))
    if drop_norm:
        entries = entries[1:]
    return nn.Sequential(OrderedDict(entries))


class Attention(nn.Module):
    """Single-head self-attention parameters (src/dmme/models/ddpm.py:38-52)."""

    def __init__(self, dim: int, num_groups: int) -> None:
        super().__init__()
        self.norm = nn.GroupNorm(num_groups, dim)
        self.scale = dim ** -0.5
        self.qkv_proj = nn.Conv2d(dim, dim * 3, kernel_size=1)
        self.proj = nn.Conv2d(dim, dim, kernel_size=1)


class ResBlock(nn.Module):
    """ResBlock parameters (src/dmme/models/ddpm.py:82-116)."""

    def __init__(self, c_in: int, c_out: int, with_attention: bool = False, emb_dim: int = 512,
                 num_groups: int = 32, p: float = 0.1) -> None:
        super().__init__()
        self.conv1 = _norm_act_conv(c_in, c_out, num_groups, 0.0)
        self.condition = nn.Sequential(nn.Linear(emb_dim, c_out), nn.Identity())
        self.conv2 = _norm_act_conv(c_out, c_out, num_groups, p)
        self.residual = nn.Conv2d(c_in, c_out, kernel_size=1) if c_in != c_out else nn.Identity()
        self.attention = Attention(c_out, num_groups) if with_attention else nn.Identity()
        self.p = p


def DownSample(c_in: int, c_out: int) -> nn.Conv2d:
    """Stride-2 3x3 conv (src/dmme/models/ddpm.py:136-147)."""
    return nn.Conv2d(c_in, c_out, kernel_size=3, stride=2, padding=1)


class UpSample(nn.Module):
    """Nearest x2 then 3x3 conv (src/dmme/models/ddpm.py:150-173); key ``conv.*``."""

    def __init__(self, c_in: int, c_out: int) -> None:
        super().__init__()
        self.conv = nn.Conv2d(c_in, c_out, kernel_size=3, stride=1, padding=1)


def build_topology(unet: nn.Module, make_block, in_channels: int, out_channels: int, pos_dim: int, emb_dim: int,
                   num_groups: int, channels_per_depth: Sequence[int], num_blocks: int,
                   attention_depths: Sequence[int]) -> None:
    """Creates the module lists in the reference's construction order (condition, input_conv, down, up,
    middle, output) so that seeded initialisation matches (src/dmme/models/ddpm.py:202-279).

    The reference's ``down_layers[-1] == len(channels) - 1`` test compares a Module with an int and is
    always False: there is never a leading UpSample."""
    widths = [channels_per_depth[0]]
    for c in channels_per_depth:
        widths += [c] * num_blocks
    n_depth = len(channels_per_depth)
    down_at = {num_blocks * i for i in range(1, n_depth)}

    unet.condition = nn.Sequential(SinusoidalPositionEmbeddings(pos_dim), nn.Linear(pos_dim, emb_dim), nn.SiLU(),
                                   nn.Linear(emb_dim, emb_dim), nn.SiLU())
    unet.input_conv = nn.Conv2d(in_channels, widths[0], kernel_size=3, stride=1, padding=1)

    down, depth = [], 1
    for k in range(1, len(widths)):
        down.append(make_block(widths[k - 1], widths[k], depth in attention_depths))
        if k in down_at:
            down.append(DownSample(widths[k], widths[k]))
            depth += 1

    up, depth = [], n_depth
    rev = widths[::-1]
    for i in range(len(rev) - 1):
        c_in, c_out = rev[i], rev[i + 1]
        attn = depth in attention_depths
        up.append(make_block(2 * c_in, c_out, attn))
        if (len(widths) - 1 - i - 1) in down_at:
            up += [make_block(2 * c_out, c_out, attn), UpSample(c_out, c_out)]
            depth -= 1
    up.append(make_block(2 * widths[0], widths[0], 1 in attention_depths))

    unet.down_layers = nn.ModuleList(down)
    unet.up_layers = nn.ModuleList(up)
    unet.middle_layers = nn.ModuleList([make_block(widths[-1], widths[-1], True), make_block(widths[-1], widths[-1], False)])
    unet.output_conv = _norm_act_conv(widths[0], out_channels, num_groups, 0.0)


class _UNetBase(nn.Module):
    flavour = "ddpm"

    def _finish_init(self, precision: str) -> None:
        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        self.precision = precision
        object.__setattr__(self, "_engine_obj", None)

    @property
    def engine(self) -> Engine:
        eng = self.__dict__.get("_engine_obj")
        if eng is None:
            eng = Engine(self, self.flavour)
            object.__setattr__(self, "_engine_obj", eng)
        return eng

    def _dropout_masks(self, n: int, device) -> Optional[Dict[str, Tensor]]:
        """Dropout2d keep-masks (N, C) scaled by 1/(1-p), one per ResBlock (training mode only)."""
        if not self.training:
            return None
        masks = {}
        for name, blk in self.engine.resblocks():
            if blk.p > 0:
                c = blk.conv2[-1].weight.shape[1]
                keep = torch.rand((n, c), device=device) >= blk.p
                masks[name] = keep.float() / (1.0 - blk.p)
        return masks or None

    @property
    def train_engine(self) -> TrainEngine:
        eng = self.__dict__.get("_train_engine_obj")
        if eng is None:
            eng = TrainEngine(self, self.flavour)
            object.__setattr__(self, "_train_engine_obj", eng)
        return eng

    def forward_raw(self, x: Tensor, c: Tensor, masks: Optional[Dict[str, Tensor]] = None, sampler=None) -> Tensor:
        """Inference forward: returns the executor's own output buffer (overwritten by the next call).  No
        autograd graph is recorded; use ``forward`` (with gradients enabled) for training.
        ``sampler``: optional ``ops.sampler_epilogue``; when ``self.engine.sampler_applied`` is True afterwards, the output
        conv applied that update to ``x_t`` in its epilogue and the returned buffer was not written."""
        if not x.is_cuda:
            raise RuntimeError("dmme_b200.UNet runs on CUDA (sm_100a) only; there is no CPU path")
        with torch.no_grad(), torch.cuda.device(x.device):
            if masks is None:
                masks = self._dropout_masks(x.shape[0], x.device)
            eng = self.engine
            eng.force_generic = self.precision == "fp32"
            return eng.forward(x.contiguous(), c.contiguous(), _PRECISIONS[self.precision], masks, sampler)

    def forward(self, x: Tensor, c: Tensor) -> Tensor:
        r"""Predicts noise from x.

        Args:
            x: image of shape (N, C, H, W), float32
            c: timestep of shape (N,) or (1,), integer

        Returns:
            estimated noise (N, C, H, W) (IDDPM flavour: (N, 2C, H, W)), float32.  With gradients enabled the
            result carries an autograd node whose backward runs the explicit backward kernels (``_train.py``).
        """
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            if not x.is_cuda:
                raise RuntimeError("dmme_b200.UNet runs on CUDA (sm_100a) only; there is no CPU path")
            with torch.cuda.device(x.device):
                return _UNetFunction.apply(self, x, c, *self.parameters())
        return self.forward_raw(x, c).clone()


class _UNetFunction(torch.autograd.Function):
    """Autograd node of one UNet call: forward = TrainEngine.forward, backward = TrainEngine.backward."""

    @staticmethod
    def forward(ctx, unet, x, c, *params):
        eng = unet.train_engine
        eng.force_generic = unet.precision == "fp32"
        masks = getattr(unet, "_injected_masks", None) or unet._dropout_masks(x.shape[0], x.device)
        eng.want_input_grad = bool(x.requires_grad)
        out = eng.forward(x.detach(), c, _PRECISIONS[unet.precision], masks)
        eng.generation = getattr(eng, "generation", 0) + 1
        ctx.unet, ctx.generation = unet, eng.generation
        ctx.x_dtype = x.dtype
        return out.clone()

    @staticmethod
    def backward(ctx, d_out):
        unet = ctx.unet
        eng = unet.train_engine
        if eng.generation != ctx.generation:
            raise RuntimeError("dmme_b200: backward() of a UNet call whose saved activations were overwritten by a later "
                               "forward; run backward before the next training forward of the same module")
        with torch.cuda.device(d_out.device):
            grads = eng.backward(d_out.contiguous())
        out = []
        for p in unet.parameters():
            g = grads.get(id(p)) if p.requires_grad else None
            out.append(g.clone() if g is not None else None)
        gx = eng.input_grad.clone().to(ctx.x_dtype) if eng.want_input_grad and eng.input_grad is not None else None
        return (None, gx, None) + tuple(out)


class UNet(_UNetBase):
    r"""U-Net for predicting noise in images (drop-in for ``dmme.models.ddpm.UNet``).

    Args:
        in_channels (int): input channels of image
        pos_dim (int): dimension of position embedding
        emb_dim (int): dimension of timestep embedding
        num_groups (int): number of groups in GroupNorm
        dropout (float): channel dropout rate in the second conv of each ResBlock
        channels_per_depth (Tuple[int, ...]): channels per depth
        num_blocks (int): number of resblocks to use in each depth
        attention_depths (Tuple[int, ...]): depths to use attention blocks
        precision (str): "bf16" (tcgen05 tensor cores, fp32 accumulate) or "fp32" (parity mode)
    """

    flavour = "ddpm"

    def __init__(self, in_channels=3, pos_dim=128, emb_dim=512, num_groups=32, dropout=0.1,
                 channels_per_depth=(128, 256, 256, 256), num_blocks=2, attention_depths=(2,), precision="bf16"):
        super().__init__()

        def make_block(c_in, c_out, attn):
            return ResBlock(c_in, c_out, attn, emb_dim=emb_dim, num_groups=num_groups, p=dropout)

        build_topology(self, make_block, in_channels, in_channels, pos_dim, emb_dim, num_groups,
                       tuple(channels_per_depth), num_blocks, tuple(attention_depths))
        self._finish_init(precision)
