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


class PackBatch:
    """Device table of ``dmme_pack_item`` entries: every bf16 tensor-core weight pack of a model as ONE launch."""

    def __init__(self, entries, device) -> None:
        # entries: [(w, wres | None, packed, dgrad, ci_off, ci_cnt)]; the tensors are kept alive here (raw pointers in the table)
        per = int(L.load().dmme_pack_block_elems())
        self.tensors = entries
        items = (L.PackItem * len(entries))()
        blocks = 0
        for k, (w, wres, packed, dgrad, off, cnt) in enumerate(entries):
            cout, cin, ks, _ = w.shape
            it = items[k]
            it.w, it.wres, it.packed = ptr(w), ptr(wres), ptr(packed)
            it.cout, it.cin, it.ksize, it.rc = cout, cin, ks, (wres.shape[1] if wres is not None else 0)
            it.dgrad, it.ci_off, it.ci_cnt, it.reserved = int(dgrad), int(off), int(cnt), 0
            it.first_block = blocks
            elems = packed.numel()
            blocks += (elems + per - 1) // per
        self.n, self.blocks = len(entries), blocks
        raw = torch.frombuffer(bytearray(bytes(items)), dtype=torch.uint8)
        self.table = raw.to(device)

    def launch(self) -> None:
        L.check(L.load().dmme_pack_conv_weights_batch(ptr(self.table), self.n, self.blocks, L.stream_ptr()), "pack_conv_weights_batch")


def pack_upsample_phase_weight(w: Tensor) -> Tensor:
    """Weights of ``nearest x2 -> conv3x3`` (UpSample, models/ddpm.py:150-173) collapsed onto the low-resolution grid:
    bf16 ``[4 phases][cout][4 taps x cin]`` for ``make_conv_desc(upsample=3)``.  Phase (a, b) is the parity of the output
    pixel; its 2x2 taps (u, v) sum the 3x3 taps that fall on the same low-resolution pixel: rows {0 | 1+2} for a = 0,
    {0+1 | 2} for a = 1, likewise for columns.  Summed in fp32, rounded to bf16 once."""
    L.require_cuda(w)
    w = w.detach().float()
    cout, cin, kh, kw = w.shape
    if kh != 3 or kw != 3:
        raise ValueError("UpSample convs are 3x3")
    groups = (((0,), (1, 2)), ((0, 1), (2,)))  # [parity][tap] -> 3x3 indices
    out = torch.empty((4, cout, 4, cin), dtype=torch.float32, device=w.device)
    for a in range(2):
        for b in range(2):
            for u in range(2):
                for v in range(2):
                    acc = 0
                    for r in groups[a][u]:
                        for s in groups[b][v]:
                            acc = acc + w[:, :, r, s]
                    out[2 * a + b, :, 2 * u + v, :] = acc
    return out.reshape(4 * cout, 4 * cin).to(torch.bfloat16).contiguous()


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
    d.ksize, d.stride, d.upsample, d.cout = ksize, stride, int(upsample), cout  # upsample: 0 | 1 nearest x2 | 3 sub-pixel
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


def conv_fuses_gn(desc: L.ConvDesc) -> bool:
    """True when the kernel that would run ``desc`` can apply a fused GroupNorm(+SiLU) to its input (``gn_ab=``)."""
    return bool(L.load().dmme_conv2d_fuses_gn(C.byref(desc)))


def groupnorm_coeff(stats0: Tensor, stats1: Optional[Tensor], c0: int, c1: int, n: int, hw: int, groups: int,
                    gamma: Tensor, beta: Tensor, scale: Optional[Tensor] = None, shift: Optional[Tensor] = None,
                    eps: float = 1e-5, out: Optional[Tensor] = None) -> Tensor:
    """(a, b) of ``y = a * x + b`` per (image, channel) -- GroupNorm statistics of the producing convs folded with gamma /
    beta (+ IDDPM scale / shift) -- as fp32 ``[n, c0 + c1, 2]`` for ``conv2d_launch(gn_ab=)``."""
    L.require_cuda(stats0, stats1, gamma, beta, out)
    ab = _empty((n, c0 + c1, 2), torch.float32, stats0.device, out)
    ss_rows = ss_ld = 0
    if scale is not None:
        ss_rows, ss_ld = scale.shape[0], scale.stride(0)
    L.check(L.load().dmme_groupnorm_coeff(ptr(stats0), ptr(stats1), c0, c1, n, hw, groups, eps, ptr(gamma), ptr(beta),
                                          ptr(scale), ptr(shift), ss_rows, ss_ld, ptr(ab), L.stream_ptr()), "groupnorm_coeff")
    return ab


def conv_fuses_sampler(desc: L.ConvDesc) -> bool:
    """True when the kernel that would run ``desc`` can apply a sampler update in its epilogue (``sampler=``)."""
    return bool(L.load().dmme_conv2d_fuses_sampler(C.byref(desc)))


def sampler_epilogue(kind: int, x: Tensor, t: Tensor, alpha_bar: Tensor, beta: Optional[Tensor] = None,
                     alpha: Optional[Tensor] = None, tau: Optional[Tensor] = None, noise: Optional[Tensor] = None,
                     seed: int = 0, noise_offset: int = 0) -> L.SamplerEpilogue:
    """One ``dmme_sampler_epilogue``: the DDPM / DDIM / IDDPM update of ``x`` (in place) applied by the output conv.
    The struct holds raw pointers: the caller keeps the tensors alive until the launch has been issued."""
    L.require_cuda(x, t, alpha_bar, beta, alpha, tau, noise)
    s = L.SamplerEpilogue()
    s.kind, s.x, s.noise = int(kind), ptr(x), ptr(noise)
    s.beta, s.alpha, s.alpha_bar = ptr(beta), ptr(alpha), ptr(alpha_bar)
    s.t_ptr, s.tau = ptr(t), ptr(tau)
    s.table_len, s.tau_len = alpha_bar.numel(), tau.numel() if tau is not None else 0
    s.seed, s.noise_offset = int(seed) & (2 ** 64 - 1), int(noise_offset)
    return s


def conv_epilogue_norm(desc: L.ConvDesc) -> bool:
    """True when the unsplit kernel that would run ``desc`` honours ``conv2d_launch(out_norms=)`` itself (8x8 maps on the
    transposed tcgen05 kernel: the epilogue warps hold whole images)."""
    return bool(L.load().dmme_conv2d_epilogue_norm(C.byref(desc)))


def conv_splitk_workspace(desc: L.ConvDesc) -> int:
    """Bytes of fp32 workspace with which ``conv2d_launch(splitk_ws=)`` runs ``desc`` split-K (0: it would not split).
    Only the split-K path honours ``out_norms=``."""
    return int(L.load().dmme_conv2d_splitk_workspace(C.byref(desc)))


def out_norm(out: Tensor, gamma: Tensor, beta: Tensor, cpg: int, silu: bool, eps: float = 1e-5,
             scale: Optional[Tensor] = None, shift: Optional[Tensor] = None) -> L.OutNorm:
    """One ``dmme_out_norm``: the GroupNorm(+SiLU) a consumer applies to this conv's output, written by the split-K
    finishing pass.  ``gamma`` / ``beta`` (and ``scale`` / ``shift``, 2-D fp32 views) start at this tensor's first channel
    inside the consumer's norm; ``cpg`` is the consumer's group width."""
    o = L.OutNorm()
    o.out, o.gamma, o.beta = ptr(out), ptr(gamma), ptr(beta)
    o.cpg, o.silu, o.eps = int(cpg), int(silu), float(eps)
    if scale is not None:
        if scale.dim() != 2 or shift is None or shift.dim() != 2 or scale.stride(0) != shift.stride(0):
            raise ValueError("scale/shift must be 2-D fp32 views with equal row stride")
        o.scale, o.shift, o.ss_rows, o.ss_ld = ptr(scale), ptr(shift), scale.shape[0], scale.stride(0)
    return o


def conv_out_hw(desc: L.ConvDesc) -> Tuple[int, int]:
    h = desc.h_in * (2 if desc.upsample else 1)
    w = desc.w_in * (2 if desc.upsample else 1)
    pad = desc.ksize // 2
    return (h + 2 * pad - desc.ksize) // desc.stride + 1, (w + 2 * pad - desc.ksize) // desc.stride + 1


def conv2d_launch(desc: L.ConvDesc, weight: Tensor, bias: Optional[Tensor], out: Tensor,
                  temb: Optional[Tensor] = None, addend: Optional[Tensor] = None,
                  out2: Optional[Tensor] = None, out3: Optional[Tensor] = None, stats: Optional[Tensor] = None,
                  gn_ab: Optional[Tensor] = None, gn_silu: bool = True, splitk_ws: Optional[Tensor] = None,
                  out_norms=(), sampler: Optional[L.SamplerEpilogue] = None) -> None:
    """Launch one fused convolution described by ``desc`` (see include/dmme_b200.h)."""
    desc.weight, desc.bias = ptr(weight), ptr(bias)
    desc.sampler = C.pointer(sampler) if sampler is not None else None
    desc.splitk_ws = ptr(splitk_ws)
    desc.splitk_ws_bytes = splitk_ws.numel() * splitk_ws.element_size() if splitk_ws is not None else 0
    for k in range(2):
        desc.out_norm[k] = out_norms[k] if k < len(out_norms) else L.OutNorm()
    desc.gn_ab, desc.gn_silu = ptr(gn_ab), int(gn_silu)
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


CHAIN_COUT = 256      # output channels of every conv of a chain (csrc/conv_chain.cu)
CHAIN_MAX_OPS = 16


def chain_op(src0: Optional[Tensor], src1: Optional[Tensor], weight: Tensor, bias: Optional[Tensor], *,
             c0: Optional[int] = None, res0: Optional[Tensor] = None, res1: Optional[Tensor] = None,
             temb: Optional[Tensor] = None, addend: Optional[Tensor] = None, out: Optional[Tensor] = None,
             stats: Optional[Tensor] = None, out_norms=(), keep: int = -1) -> L.ChainOp:
    """One ``dmme_chain_op`` (see include/dmme_b200.h).  ``src0 = None``: the operand the previous op kept in shared
    memory (``c0`` channels).  The returned struct only holds raw pointers: the caller keeps the tensors alive."""
    o = L.ChainOp()
    o.src0, o.c0 = (ptr(src0), src0.shape[3]) if src0 is not None else (None, int(c0))
    o.src1, o.c1 = (ptr(src1), src1.shape[3]) if src1 is not None else (None, 0)
    o.res0, o.rc0 = (ptr(res0), res0.shape[3]) if res0 is not None else (None, 0)
    o.res1, o.rc1 = (ptr(res1), res1.shape[3]) if res1 is not None else (None, 0)
    o.weight, o.bias = ptr(weight), ptr(bias)
    if temb is not None:
        if temb.dim() != 2 or temb.dtype != torch.float32:
            raise ValueError("temb must be a 2-D fp32 view")
        o.temb, o.temb_rows, o.temb_ld = ptr(temb), temb.shape[0], temb.stride(0)
    o.addend, o.out, o.stats = ptr(addend), ptr(out), ptr(stats)
    for k in range(2):
        o.out_norm[k] = out_norms[k] if k < len(out_norms) else L.OutNorm()
    o.keep = int(keep)
    return o


def conv_chain_supported(n: int, h: int, w: int, cout: int) -> bool:
    return bool(L.load().dmme_conv_chain_supported(n, h, w, cout))


def conv_chain(chain, n: int, h: int, w: int) -> None:
    """Run a list of ``chain_op`` back to back in one persistent launch (csrc/conv_chain.cu)."""
    arr = (L.ChainOp * len(chain))(*chain)
    L.check(L.load().dmme_conv_chain_fwd(arr, len(chain), n, h, w, L.stream_ptr()), "conv_chain_fwd")


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


def attention_block_supported(heads: int, seq: int, c: int, dtype: torch.dtype) -> bool:
    """True when ``attention_block`` (the one-launch attention block) takes this shape."""
    return bool(L.load().dmme_attention_block_supported(heads, seq, c, L.act_code(dtype)))


def attention_block(x: Tensor, gn_ab: Optional[Tensor], wqkv: Tensor, bias_qkv: Tensor, wproj: Tensor, bias_proj: Tensor,
                    scale: float, out: Optional[Tensor] = None, stats: Optional[Tensor] = None, *,
                    stats_in: Optional[Tensor] = None, gamma: Optional[Tensor] = None, beta: Optional[Tensor] = None,
                    groups: int = 0, eps: float = 1e-5) -> Tensor:
    """``x + proj(attention(qkv_proj(GroupNorm(x))))`` (Attention.forward, models/ddpm.py:54-75) in one launch.
    x: NHWC bf16 [n, h, w, c]; the block's norm either as ``gn_ab`` (``groupnorm_coeff``) or, with ``gn_ab=None``, as the
    producer's statistics ``stats_in`` + ``gamma`` / ``beta`` / ``groups`` / ``eps`` (the kernel forms the coefficients);
    wqkv / wproj: ``pack_conv_weight`` of the two 1x1 convs; stats: optional zeroed int64 micro-group sums of the output."""
    L.require_cuda(x, gn_ab, wqkv, bias_qkv, wproj, bias_proj, out, stats, stats_in, gamma, beta)
    n, h, w, c = x.shape
    y = _empty((n, h, w, c), x.dtype, x.device, out)
    L.check(L.load().dmme_attention_block_fwd(ptr(x), ptr(gn_ab), ptr(stats_in), ptr(gamma), ptr(beta), int(groups), float(eps),
                                              ptr(wqkv), ptr(bias_qkv), ptr(wproj), ptr(bias_proj), n, 1, h * w, c,
                                              float(scale), ptr(y), ptr(stats), L.act_code(x.dtype), L.stream_ptr()),
            "attention_block_fwd")
    return y


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
               t: Tensor, seed: int = 0, noise_offset: int = 0) -> Tensor:
    L.require_cuda(x, eps, noise, beta, alpha, alpha_bar, t)
    L.check(L.load().dmme_ddpm_step(ptr(x), ptr(eps), ptr(noise), ptr(beta), ptr(alpha), ptr(alpha_bar), ptr(t),
                                    beta.numel(), x.numel(), seed, noise_offset, L.stream_ptr()), "ddpm_step")
    return x


def ddim_step_(x: Tensor, eps: Tensor, alpha_bar: Tensor, tau: Tensor, i: Tensor) -> Tensor:
    L.require_cuda(x, eps, alpha_bar, tau, i)
    L.check(L.load().dmme_ddim_step(ptr(x), ptr(eps), ptr(alpha_bar), ptr(tau), ptr(i), alpha_bar.numel(), tau.numel(),
                                    x.numel(), L.stream_ptr()),
            "ddim_step")
    return x


def iddpm_step_(x: Tensor, model_out: Tensor, noise: Optional[Tensor], beta: Tensor, alpha: Tensor,
                alpha_bar: Tensor, t: Tensor, seed: int = 0, noise_offset: int = 0) -> Tensor:
    L.require_cuda(x, model_out, noise, beta, alpha, alpha_bar, t)
    n, c, h, w = x.shape
    L.check(L.load().dmme_iddpm_step(ptr(x), ptr(model_out), ptr(noise), ptr(beta), ptr(alpha), ptr(alpha_bar), ptr(t),
                                     beta.numel(), n, c, h * w, seed, noise_offset, L.stream_ptr()), "iddpm_step")
    return x


def gather_i64(table: Tensor, idx: Tensor, out: Tensor) -> Tensor:
    L.check(L.load().dmme_gather_i64(ptr(table), ptr(idx), ptr(out), L.stream_ptr()), "gather_i64")
    return out


def add_i64_(value: Tensor, delta: int) -> Tensor:
    L.check(L.load().dmme_add_i64(ptr(value), delta, L.stream_ptr()), "add_i64")
    return value


def philox_normal(shape, seed: int, stream_id: int, device, out: Optional[Tensor] = None, noise_offset: int = 0) -> Tensor:
    y = _empty(tuple(shape), torch.float32, device, out)
    L.check(L.load().dmme_philox_normal(ptr(y), y.numel(), seed, stream_id, noise_offset, L.stream_ptr()), "philox_normal")
    return y


def launch_count() -> int:
    return int(L.load().dmme_launch_count())


def reset_launch_count() -> None:
    L.load().dmme_reset_launch_count()


# ---------------------------------------------------------------------------------------------
# backward pass
# ---------------------------------------------------------------------------------------------
def pack_conv_weight_dgrad(w: Tensor, ci_off: int, ci_cnt: int, tc: bool) -> Tensor:
    """Weights of the convolution that maps grad_out to the gradient of input channels [ci_off, ci_off+ci_cnt)."""
    L.require_cuda(w)
    w = w.detach().float().contiguous()
    cout, cin, kh, _ = w.shape
    k = kh * kh * cout
    packed = torch.empty((ci_cnt, k) if tc else (k, ci_cnt), dtype=torch.bfloat16 if tc else torch.float32, device=w.device)
    L.check(L.load().dmme_pack_conv_weight_dgrad(ptr(w), cout, cin, kh, ci_off, ci_cnt, ptr(packed),
                                                 L.CONV_TC if tc else L.CONV_GENERIC, L.stream_ptr()), "pack_conv_weight_dgrad")
    return packed


def conv_wgrad_workspace(desc: L.ConvDesc) -> int:
    return int(L.load().dmme_conv2d_wgrad_workspace(C.byref(desc)))


def conv_wgrad_uses_tc(desc: L.ConvDesc) -> bool:
    return bool(L.load().dmme_conv2d_wgrad_uses_tc(C.byref(desc)))


def conv2d_wgrad(desc: L.ConvDesc, grad_out: Tensor, dweight: Tensor, dweight_res: Optional[Tensor],
                 dbias: Optional[Tensor], workspace: Tensor) -> None:
    """Weight / fused-residual-weight / bias gradients of the forward call described by ``desc``."""
    L.require_cuda(grad_out, dweight, dweight_res, dbias, workspace)
    L.check(L.load().dmme_conv2d_wgrad(C.byref(desc), ptr(grad_out), ptr(dweight), ptr(dweight_res), ptr(dbias),
                                       ptr(workspace), workspace.numel() * workspace.element_size(), L.stream_ptr()),
            "conv2d_wgrad")


def groupnorm_bwd(grad_out: Tensor, src0: Tensor, src1: Optional[Tensor], groups: int, gamma: Tensor, beta: Tensor,
                  silu: bool, scale: Optional[Tensor], shift: Optional[Tensor], chan_mask: Optional[Tensor], eps: float,
                  gin0: Optional[Tensor], gin1: Optional[Tensor], add0: Optional[Tensor], add1: Optional[Tensor],
                  dgamma: Tensor, dbeta: Tensor, dscale: Optional[Tensor], dshift: Optional[Tensor], sums: Tensor) -> None:
    L.require_cuda(grad_out, src0, src1, gin0, gin1, add0, add1)
    n, h, w, c0 = src0.shape
    c1 = src1.shape[3] if src1 is not None else 0
    ss_rows = ss_ld = dss_ld = 0
    if scale is not None:
        ss_rows, ss_ld = scale.shape[0], scale.stride(0)
    if dscale is not None:
        dss_ld = dscale.stride(0)
        if dshift is None or dshift.stride(0) != dss_ld:
            raise ValueError("dscale/dshift must be 2-D fp32 views with equal row stride")
    L.check(L.load().dmme_groupnorm_bwd(ptr(grad_out), ptr(src0), ptr(src1), c0, c1, n, h * w, groups, eps, ptr(gamma),
                                        ptr(beta), ptr(scale), ptr(shift), ss_rows, ss_ld, ptr(chan_mask), int(silu),
                                        ptr(gin0), ptr(gin1), ptr(add0), ptr(add1), ptr(dgamma), ptr(dbeta), ptr(dscale),
                                        ptr(dshift), dss_ld, ptr(sums), L.act_code(src0.dtype), L.stream_ptr()),
            "groupnorm_bwd")


def attention_bwd_workspace(n: int, heads: int, seq: int, dh: int) -> int:
    return int(L.load().dmme_attention_bwd_workspace(n, heads, seq, dh))


def attention_bwd(q: Tensor, k: Tensor, v: Tensor, n: int, heads: int, seq: int, dh: int, scale: float, batch_stride: int,
                  row_stride: int, head_stride: int, head_batch_swap: bool, dout: Tensor, dq: Tensor, dk: Tensor,
                  dv: Tensor, workspace: Tensor, p_saved: Optional[Tensor] = None) -> None:
    L.check(L.load().dmme_attention_bwd(ptr(q), ptr(k), ptr(v), batch_stride, row_stride, head_stride, n, heads, seq, dh,
                                        scale, int(head_batch_swap), ptr(dout), ptr(dq), ptr(dk), ptr(dv),
                                        L.act_code(dout.dtype), ptr(p_saved), ptr(workspace),
                                        workspace.numel() * workspace.element_size(), L.stream_ptr()), "attention_bwd")


def attention_bwd_fused_supported(heads: int, seq: int, dh: int, dtype: torch.dtype) -> bool:
    return bool(L.load().dmme_attention_bwd_fused_supported(heads, seq, dh, L.act_code(dtype)))


def attention_bwd_fused(qkv: Tensor, out: Tensor, dout: Tensor, dqkv: Tensor, n: int, heads: int, seq: int, dh: int,
                        scale: float, head_batch_swap: bool) -> Tensor:
    """dqkv = gradient of the packed ``[n][L][heads][q | k | v][dh]`` tensor given the forward output and its gradient
    (csrc/attention_bwd_tc.cu: fused tcgen05 kernel, softmax recomputed, no L x L matrix in global memory)."""
    L.require_cuda(qkv, out, dout, dqkv)
    L.check(L.load().dmme_attention_bwd_fused(ptr(qkv), ptr(out), ptr(dout), ptr(dqkv), n, heads, seq, dh, float(scale),
                                              int(head_batch_swap), L.act_code(qkv.dtype), L.stream_ptr()), "attention_bwd_fused")
    return dqkv


def attention_fwd_train(q: Tensor, k: Tensor, v: Tensor, n: int, heads: int, seq: int, dh: int, scale: float,
                        batch_stride: int, row_stride: int, head_stride: int, head_batch_swap: bool, out: Tensor,
                        p_out: Tensor, o_tmp: Tensor) -> Tensor:
    """Attention core for training: keeps the softmax matrix ``p_out`` (fp32 [n*heads, L, L]) for the backward pass."""
    L.check(L.load().dmme_attention_fwd_train(ptr(q), ptr(k), ptr(v), batch_stride, row_stride, head_stride, n, heads, seq,
                                              dh, scale, int(head_batch_swap), ptr(out), L.act_code(out.dtype), ptr(p_out),
                                              ptr(o_tmp), L.stream_ptr()), "attention_fwd_train")
    return out


def temb_bwd_workspace(rows: int, half: int, emb_dim: int) -> int:
    return int(L.load().dmme_temb_bwd_workspace(rows, half, emb_dim))


def temb_bwd(t: Tensor, freq: Tensor, w1: Tensor, b1: Tensor, w2: Tensor, b2: Tensor, hidden: Tensor, emb: Tensor,
             wcat: Tensor, d_all: Tensor, dw1: Tensor, db1: Tensor, dw2: Tensor, db2: Tensor, dwcat: Tensor, dbcat: Tensor,
             workspace: Tensor, bf16_mma: bool = False) -> None:
    """``bf16_mma``: the two large products run on the tensor cores with bf16-rounded operands (bf16 training mode)."""
    L.require_cuda(t, freq, w1, b1, w2, b2, hidden, emb, wcat, d_all, dw1, db1, dw2, db2, dwcat, dbcat, workspace)
    rows, emb_dim = emb.shape
    L.check(L.load().dmme_temb_bwd(ptr(t), rows, ptr(freq), freq.numel(), ptr(w1), ptr(b1), ptr(w2), ptr(b2), emb_dim,
                                   ptr(hidden), ptr(emb), ptr(wcat), wcat.shape[0], ptr(d_all), ptr(dw1), ptr(db1), ptr(dw2),
                                   ptr(db2), ptr(dwcat), ptr(dbcat), ptr(workspace),
                                   workspace.numel() * workspace.element_size(), int(bf16_mma), L.stream_ptr()), "temb_bwd")


def gemm_strided(a: Tensor, a_str, b: Tensor, b_str, c: Tensor, c_str, m: int, n: int, k: int, outer: int = 1,
                 heads: int = 1, alpha: float = 1.0, accumulate: bool = False) -> Tensor:
    """c[b](i,j) = alpha * sum_k a[b](i,k) b[b](k,j); *_str = (outer-batch, head, row, column) element strides."""
    L.check(L.load().dmme_gemm_strided(ptr(a), L.act_code(a.dtype), *a_str, ptr(b), L.act_code(b.dtype), *b_str, ptr(c),
                                       L.act_code(c.dtype), *c_str, m, n, k, outer, heads, alpha, int(accumulate),
                                       L.stream_ptr()), "gemm_strided")
    return c


def add(a: Tensor, b: Tensor, out: Optional[Tensor] = None) -> Tensor:
    L.require_cuda(a, b, out)
    y = _empty(tuple(a.shape), a.dtype, a.device, out)
    L.check(L.load().dmme_add(ptr(y), ptr(a), ptr(b), a.numel(), L.act_code(a.dtype), L.stream_ptr()), "add")
    return y


def pixel_sum(g: Tensor, out: Tensor) -> Tensor:
    """out[n][c] = sum over pixels of g[n][y][x][c]; ``out`` is a 2-D fp32 view (row stride honoured)."""
    n, h, w, c = g.shape
    L.check(L.load().dmme_pixel_sum(ptr(g), n, h * w, c, ptr(out), out.stride(0), L.act_code(g.dtype), L.stream_ptr()),
            "pixel_sum")
    return out


def pool2x_sum(g: Tensor, out: Optional[Tensor] = None) -> Tensor:
    L.require_cuda(g, out)
    n, h2, w2, c = g.shape
    y = _empty((n, h2 // 2, w2 // 2, c), g.dtype, g.device, out)
    L.check(L.load().dmme_pool2x_sum_nhwc(ptr(g), ptr(y), n, h2 // 2, w2 // 2, c, L.act_code(g.dtype), L.stream_ptr()),
            "pool2x_sum")
    return y


def dilate2x(g: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """Zero-dilation x2 of an NHWC tensor (values at even coordinates)."""
    L.require_cuda(g, out)
    n, h, w, c = g.shape
    y = _empty((n, 2 * h, 2 * w, c), g.dtype, g.device, out)
    L.check(L.load().dmme_dilate2x_nhwc(ptr(g), ptr(y), n, h, w, c, L.act_code(g.dtype), L.stream_ptr()), "dilate2x")
    return y


def colsum(x: Tensor, out: Tensor, accumulate: bool = False) -> Tensor:
    """out[c] (+)= sum_r x[r][c] for a 2-D fp32 view."""
    L.check(L.load().dmme_colsum_f32(ptr(x), x.shape[0], x.shape[1], x.stride(0), ptr(out), int(accumulate),
                                     L.stream_ptr()), "colsum")
    return out


def mse_loss(eps: Tensor, noise: Tensor, d_eps: Optional[Tensor], grad_scale: float = 1.0) -> Tensor:
    """L_simple = mean((eps - noise)^2) as a 1-element fp32 tensor; fills ``d_eps`` with its gradient w.r.t. eps."""
    L.require_cuda(eps, noise, d_eps)
    loss = torch.empty(1, dtype=torch.float32, device=eps.device)
    partial = torch.empty(1024, dtype=torch.float32, device=eps.device)
    L.check(L.load().dmme_mse_loss(ptr(eps), ptr(noise), eps.numel(), grad_scale, ptr(d_eps), ptr(loss), ptr(partial),
                                   L.stream_ptr()), "mse_loss")
    return loss


def iddpm_loss(model_out: Tensor, x_t: Tensor, x_0: Tensor, t: Tensor, beta: Tensor, alpha: Tensor, alpha_bar: Tensor,
               w_simple: float, w_vlb: float, d_out: Optional[Tensor], grad_scale: float = 1.0) -> Tensor:
    """[w_simple L_simple + w_vlb L_vlb, L_simple, L_vlb] as a 3-element fp32 tensor; fills ``d_out`` with the
    gradient of the first entry w.r.t. ``model_out``."""
    L.require_cuda(model_out, x_t, x_0, t, beta, alpha, alpha_bar, d_out)
    n, c, h, w = x_t.shape
    loss = torch.empty(3, dtype=torch.float32, device=x_t.device)
    partial = torch.empty(2048, dtype=torch.float32, device=x_t.device)
    L.check(L.load().dmme_iddpm_loss(ptr(model_out), ptr(x_t), ptr(x_0), ptr(t), ptr(beta), ptr(alpha), ptr(alpha_bar), n, c,
                                     h * w, w_simple, w_vlb, grad_scale, ptr(d_out), ptr(loss), ptr(partial),
                                     L.stream_ptr()), "iddpm_loss")
    return loss
