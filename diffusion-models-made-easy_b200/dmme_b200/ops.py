"""Functional wrappers over the C-ABI kernels.  Activations are NHWC tensors ``[N, H, W, C]`` of dtype
bfloat16 (tensor-core mode) or float32 (fp32 parity mode); image-space tensors are NCHW float32 as in
the reference.  Every function launches on torch's current stream and allocates only its output
(callers that need CUDA-graph-stable addresses pass ``out=``)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib as L
from ._lib import ptr

Tensor = torch.Tensor


def _empty(shape, dtype, device, out: Optional[Tensor]) -> Tensor:
    if out is not None:
        if tuple(out.shape) != tuple(shape) or out.dtype != dtype:
            raise ValueError(f"out has shape {tuple(out.shape)}/{out.dtype}, expected {tuple(shape)}/{dtype}")
        return out
    return torch.empty(shape, dtype=dtype, device=device)


# ---------------------------------------------------------------------------------------------
# weights
# ---------------------------------------------------------------------------------------------
def pack_conv_weight(w: Tensor, w_res: Optional[Tensor] = None, tc: bool = True) -> Tensor:
    """OIHW fp32 (+ optional fused 1x1 residual weight) -> GEMM-B layout of the chosen kernel."""
    L.require_cuda(w, w_res)
    w = w.detach().float().contiguous()
    cout, cin, kh, kw = w.shape
    if kh != kw:
        raise ValueError("square kernels only")
    rc = 0
    if w_res is not None:
        w_res = w_res.detach().float().contiguous()
        rc = w_res.shape[1]
    k = kh * kw * cin + rc
    if tc:
        packed = torch.empty((cout, k), dtype=torch.bfloat16, device=w.device)
    else:
        packed = torch.empty((k, cout), dtype=torch.float32, device=w.device)
    L.check(L.load().dmme_pack_conv_weight(ptr(w), cout, cin, kh, ptr(w_res), rc, ptr(packed),
                                           L.CONV_TC if tc else L.CONV_GENERIC, L.stream_ptr()), "pack_conv_weight")
    return packed


# ---------------------------------------------------------------------------------------------
# layout
# ---------------------------------------------------------------------------------------------
def nchw_to_nhwc(x: Tensor, dtype: torch.dtype, out: Optional[Tensor] = None) -> Tensor:
    L.require_cuda(x)
    n, c, h, w = x.shape
    y = _empty((n, h, w, c), dtype, x.device, out)
    L.check(L.load().dmme_nchw_to_nhwc(ptr(x.float()), ptr(y), n, c, h, w, L.act_code(dtype), L.stream_ptr()), "nchw_to_nhwc")
    return y


def nhwc_to_nchw(x: Tensor, out: Optional[Tensor] = None) -> Tensor:
    L.require_cuda(x)
    n, h, w, c = x.shape
    y = _empty((n, c, h, w), torch.float32, x.device, out)
    L.check(L.load().dmme_nhwc_to_nchw(ptr(x), ptr(y), n, c, h, w, L.act_code(x.dtype), L.stream_ptr()), "nhwc_to_nchw")
    return y


def upsample2x(x: Tensor, out: Optional[Tensor] = None) -> Tensor:
    L.require_cuda(x)
    n, h, w, c = x.shape
    y = _empty((n, 2 * h, 2 * w, c), x.dtype, x.device, out)
    L.check(L.load().dmme_upsample2x_nhwc(ptr(x), ptr(y), n, h, w, c, L.act_code(x.dtype), L.stream_ptr()), "upsample2x")
    return y


# ---------------------------------------------------------------------------------------------
# convolution
# ---------------------------------------------------------------------------------------------
def make_conv_desc(src0: Tensor, src1: Optional[Tensor], cout: int, ksize: int, stride: int = 1,
                   upsample: bool = False, res0: Optional[Tensor] = None, res1: Optional[Tensor] = None,
                   in_nchw: bool = False, out_layout: int = L.OUT_NHWC, act_dtype: Optional[torch.dtype] = None,
                   kernel: int = L.CONV_AUTO) -> L.ConvDesc:
    d = L.ConvDesc()
    if in_nchw:
        n, c0, h, w = src0.shape
    else:
        n, h, w, c0 = src0.shape
    d.src0, d.c0 = ptr(src0), c0
    d.src1, d.c1 = (ptr(src1), src1.shape[3]) if src1 is not None else (None, 0)
    d.res0, d.rc0 = (ptr(res0), res0.shape[3]) if res0 is not None else (None, 0)
    d.res1, d.rc1 = (ptr(res1), res1.shape[3]) if res1 is not None else (None, 0)
    d.n, d.h_in, d.w_in = n, h, w
    d.ksize, d.stride, d.upsample, d.cout = ksize, stride, int(upsample), cout
    d.in_layout = L.IN_NCHW_F32 if in_nchw else L.IN_NHWC
    d.out_layout = out_layout
    if act_dtype is None:
        act_dtype = src0.dtype
    d.act_dtype = L.act_code(act_dtype)
    d.kernel = kernel
    return d


def conv_uses_tc(desc: L.ConvDesc) -> bool:
    return bool(L.load().dmme_conv2d_uses_tc(C.byref(desc)))


def conv_writes_stats(desc: L.ConvDesc) -> bool:
    return bool(L.load().dmme_conv2d_writes_stats(C.byref(desc)))


def conv_out_hw(desc: L.ConvDesc) -> Tuple[int, int]:
    h = desc.h_in * (2 if desc.upsample else 1)
    w = desc.w_in * (2 if desc.upsample else 1)
    pad = desc.ksize // 2
    return (h + 2 * pad - desc.ksize) // desc.stride + 1, (w + 2 * pad - desc.ksize) // desc.stride + 1


def conv2d_launch(desc: L.ConvDesc, weight: Tensor, bias: Optional[Tensor], out: Tensor,
                  temb: Optional[Tensor] = None, addend: Optional[Tensor] = None,
                  out2: Optional[Tensor] = None, out3: Optional[Tensor] = None, stats: Optional[Tensor] = None) -> None:
    """Launch one fused convolution described by ``desc`` (see include/dmme_b200.h)."""
    desc.weight, desc.bias = ptr(weight), ptr(bias)
    if temb is not None:
        if temb.dim() != 2 or temb.dtype != torch.float32:
            raise ValueError("temb must be a 2-D fp32 view")
        desc.temb, desc.temb_rows, desc.temb_ld = ptr(temb), temb.shape[0], temb.stride(0)
    else:
        desc.temb, desc.temb_rows, desc.temb_ld = None, 0, 0
    desc.addend = ptr(addend)
    desc.out, desc.out2, desc.out3 = ptr(out), ptr(out2), ptr(out3)
    desc.stats = ptr(stats)
    L.check(L.load().dmme_conv2d_fwd(C.byref(desc), L.stream_ptr()), "conv2d_fwd")


# ---------------------------------------------------------------------------------------------
# GroupNorm / attention / timestep embedding
# ---------------------------------------------------------------------------------------------
def groupnorm(src0: Tensor, src1: Optional[Tensor], groups: int, gamma: Tensor, beta: Tensor, silu: bool,
              scale: Optional[Tensor] = None, shift: Optional[Tensor] = None, chan_mask: Optional[Tensor] = None,
              eps: float = 1e-5, out: Optional[Tensor] = None, stats0: Optional[Tensor] = None,
              stats1: Optional[Tensor] = None) -> Tensor:
    """GN (+ scale/shift) (+ SiLU) (+ channel mask) over the channel-concat of src0|src1 (NHWC).
    stats0/stats1: int64 micro-group sums written by the producing conv (see ``conv2d_launch(stats=)``)."""
    L.require_cuda(src0, src1, out)
    n, h, w, c0 = src0.shape
    c1 = src1.shape[3] if src1 is not None else 0
    y = _empty((n, h, w, c0 + c1), src0.dtype, src0.device, out)
    ss_rows = ss_ld = 0
    if scale is not None:
        if scale.dim() != 2 or shift is None or shift.dim() != 2 or scale.stride(0) != shift.stride(0):
            raise ValueError("scale/shift must be 2-D fp32 views with equal row stride")
        ss_rows, ss_ld = scale.shape[0], scale.stride(0)
    L.check(L.load().dmme_groupnorm_fwd(ptr(src0), ptr(src1), c0, c1, n, h * w, groups, eps, ptr(gamma), ptr(beta),
                                        ptr(scale), ptr(shift), ss_rows, ss_ld, ptr(chan_mask), int(silu), ptr(y),
                                        L.act_code(src0.dtype), ptr(stats0), ptr(stats1), L.stream_ptr()), "groupnorm_fwd")
    return y


def attention(q: Tensor, k: Tensor, v: Tensor, n: int, heads: int, seq: int, dh: int, scale: float,
              batch_stride: int, row_stride: int, head_stride: int, v_transposed: bool, v_batch_stride: int,
              head_batch_swap: bool, out: Tensor, kernel: int = L.CONV_AUTO) -> Tensor:
    L.check(L.load().dmme_attention_fwd(ptr(q), ptr(k), ptr(v), batch_stride, row_stride, head_stride,
                                        int(v_transposed), v_batch_stride, n, heads, seq, dh, scale,
                                        int(head_batch_swap), ptr(out), L.act_code(out.dtype), kernel, L.stream_ptr()),
            "attention_fwd")
    return out


def temb_mlp(t: Tensor, freq: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor,
             out: Optional[Tensor] = None, scratch: Optional[Tensor] = None) -> Tensor:
    L.require_cuda(t, freq, w1, b1, w2, b2)
    if t.dtype != torch.int64:
        t = t.long()
    rows, emb = t.numel(), w2.shape[0]
    y = _empty((rows, emb), torch.float32, t.device, out)
    scratch = _empty((rows, emb), torch.float32, t.device, scratch)
    L.check(L.load().dmme_temb_mlp_fwd(ptr(t), rows, ptr(freq), freq.numel(), ptr(w1), ptr(b1), ptr(w2), ptr(b2), emb,
                                       ptr(scratch), ptr(y), L.stream_ptr()), "temb_mlp_fwd")
    return y


def temb_proj(emb: Tensor, wcat: Tensor, bcat: Tensor, out: Optional[Tensor] = None) -> Tensor:
    rows, emb_dim = emb.shape
    total = wcat.shape[0]
    y = _empty((rows, total), torch.float32, emb.device, out)
    L.check(L.load().dmme_temb_proj_fwd(ptr(emb), rows, emb_dim, ptr(wcat), ptr(bcat), total, ptr(y), L.stream_ptr()),
            "temb_proj_fwd")
    return y


# ---------------------------------------------------------------------------------------------
# sampler updates (in place on x)
# ---------------------------------------------------------------------------------------------
def ddpm_step_(x: Tensor, eps: Tensor, noise: Optional[Tensor], beta: Tensor, alpha: Tensor, alpha_bar: Tensor,
               t: Tensor, seed: int = 0) -> Tensor:
    L.require_cuda(x, eps, noise, beta, alpha, alpha_bar, t)
    L.check(L.load().dmme_ddpm_step(ptr(x), ptr(eps), ptr(noise), ptr(beta), ptr(alpha), ptr(alpha_bar), ptr(t),
                                    x.numel(), seed, L.stream_ptr()), "ddpm_step")
    return x


def ddim_step_(x: Tensor, eps: Tensor, alpha_bar: Tensor, tau: Tensor, i: Tensor) -> Tensor:
    L.require_cuda(x, eps, alpha_bar, tau, i)
    L.check(L.load().dmme_ddim_step(ptr(x), ptr(eps), ptr(alpha_bar), ptr(tau), ptr(i), x.numel(), L.stream_ptr()),
            "ddim_step")
    return x


def iddpm_step_(x: Tensor, model_out: Tensor, noise: Optional[Tensor], beta: Tensor, alpha: Tensor,
                alpha_bar: Tensor, t: Tensor, seed: int = 0) -> Tensor:
    L.require_cuda(x, model_out, noise, beta, alpha, alpha_bar, t)
    n, c, h, w = x.shape
    L.check(L.load().dmme_iddpm_step(ptr(x), ptr(model_out), ptr(noise), ptr(beta), ptr(alpha), ptr(alpha_bar), ptr(t),
                                     n, c, h * w, seed, L.stream_ptr()), "iddpm_step")
    return x


def gather_i64(table: Tensor, idx: Tensor, out: Tensor) -> Tensor:
    L.check(L.load().dmme_gather_i64(ptr(table), ptr(idx), ptr(out), L.stream_ptr()), "gather_i64")
    return out


def add_i64_(value: Tensor, delta: int) -> Tensor:
    L.check(L.load().dmme_add_i64(ptr(value), delta, L.stream_ptr()), "add_i64")
    return value


def philox_normal(shape, seed: int, stream_id: int, device, out: Optional[Tensor] = None) -> Tensor:
    y = _empty(tuple(shape), torch.float32, device, out)
    L.check(L.load().dmme_philox_normal(ptr(y), y.numel(), seed, stream_id, L.stream_ptr()), "philox_normal")
    return y


def launch_count() -> int:
    return int(L.load().dmme_launch_count())


def reset_launch_count() -> None:
    L.load().dmme_reset_launch_count()
