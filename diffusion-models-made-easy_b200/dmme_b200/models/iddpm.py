"""IDDPM UNet (learned variance: 2C output channels, scale-shift conditioning, 4-head attention).

Drop-in for ``dmme.models.iddpm.UNet`` (src/dmme/models/iddpm.py:125-265) with the same
``state_dict`` layout (``X.norm.*`` present, no ``X.conv2.0``, ``X.condition.0.weight`` of width 2C).
Reference behaviour that is reproduced on purpose: the attention output is regrouped as "(head b)"
although the heads were folded as "(b head)" (models/iddpm.py:38 vs :44-46) and the softmax scale is
``dim ** -0.5`` of the full width (models/iddpm.py:32).
"""
from __future__ import annotations

from torch import nn

from .ddpm import _UNetBase, _norm_act_conv, build_topology


class MultiHeadAttention(nn.Module):
    """Multi-head attention parameters (src/dmme/models/iddpm.py:24-34)."""

    def __init__(self, dim: int, num_groups: int, num_heads: int) -> None:
        super().__init__()
        assert dim % num_heads == 0
        self.num_heads = num_heads
        self.norm = nn.GroupNorm(num_groups, dim)
        self.scale = dim ** -0.5
        self.qkv_proj = nn.Conv2d(dim, 3 * dim, kernel_size=1)
        self.proj = nn.Conv2d(dim, dim, kernel_size=1)


class ResBlock(nn.Module):
    """Scale-shift ResBlock parameters (src/dmme/models/iddpm.py:74-104)."""

    def __init__(self, c_in, c_out, with_attention=False, num_heads=4, emb_dim=512, num_groups=32, p=0.1) -> None:
        super().__init__()
        self.conv1 = _norm_act_conv(c_in, c_out, num_groups, 0.0)
        self.norm = nn.GroupNorm(num_groups, c_out)
        self.condition = nn.Sequential(nn.Linear(emb_dim, c_out * 2), nn.Identity())
        self.conv2 = _norm_act_conv(c_out, c_out, num_groups, p, drop_norm=True)
        self.residual = nn.Conv2d(c_in, c_out, kernel_size=1) if c_in != c_out else nn.Identity()
        self.attention = MultiHeadAttention(c_out, num_groups, num_heads) if with_attention else nn.Identity()
        self.p = p


class UNet(_UNetBase):
    r"""U-Net predicting noise and the variance interpolation coefficient (drop-in for
    ``dmme.models.iddpm.UNet``); arguments as ``dmme_b200.models.ddpm.UNet``."""

    flavour = "iddpm"

    def __init__(self, in_channels=3, pos_dim=128, emb_dim=512, num_groups=32, dropout=0.3,
                 channels_per_depth=(128, 256, 256, 256), num_blocks=2, attention_depths=(2, 3), precision="bf16"):
        super().__init__()

        def make_block(c_in, c_out, attn):
            return ResBlock(c_in, c_out, attn, emb_dim=emb_dim, num_groups=num_groups, p=dropout)

        build_topology(self, make_block, in_channels, 2 * in_channels, pos_dim, emb_dim, num_groups,
                       tuple(channels_per_depth), num_blocks, tuple(attention_depths))
        self._finish_init(precision)
