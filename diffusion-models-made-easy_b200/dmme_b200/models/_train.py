"""Training executor: the forward pass of ``_engine.Engine`` with every intermediate kept, plus the explicit
backward pass (the reference gets it from autograd: ``loss.backward()`` after ``DDPM.training_step``
src/dmme/diffusion_models/ddpm.py:53-81 / ``IDDPM.training_step`` src/dmme/diffusion_models/iddpm.py:62-116).

The forward records one closure per fused op on a tape; ``backward`` replays the tape in reverse.  Every closure
launches C-ABI kernels only:

    conv        ->  wgrad (+bias, +fused 1x1 residual weights)  | dgrad = forward conv kernels on grad_out with
                    flipped/transposed weights (stride 2: zero-dilated gather; nearest x2: 2x2 sum pool afterwards)
                    | pixel sums for the timestep-embedding gradient | pass-through to the fused addend
    GroupNorm   ->  dmme_groupnorm_bwd (SiLU, dropout mask and IDDPM scale/shift folded in; skip / residual
                    gradient accumulation fused through its addend inputs)
    attention   ->  dmme_attention_bwd on the strided q/k/v views of the NHWC qkv tensor
    conditioning->  dmme_temb_bwd (batched ResBlock projections + the two-layer MLP)

Gradient contributions to an activation that feeds several consumers (skip connections, identity residuals) are
collected per tensor and summed either inside the next kernel that can take an addend or by ``dmme_add``.
"""
from __future__ import annotations

import os

from typing import Callable, Dict, List, Optional, Tuple

import torch
from torch import nn

from .. import _lib as L
from .. import ops
from ._engine import Engine

Tensor = torch.Tensor


class TrainEngine(Engine):
    def __init__(self, unet: nn.Module, flavour: str) -> None:
        super().__init__(unet, flavour)
        self.tape: List[Callable[[], None]] = []
        self.pending: Dict[int, List[Tensor]] = {}
        self.param_grads: Dict[int, Tensor] = {}
        self._k = 0
        self._d_all: Optional[Tensor] = None
        self._temb_ctx = None
        # parameter gradients live in one flat fp32 arena, handed out in backward order: a finished prefix of the
        # arena is a contiguous bucket that parallel.GradBucketer all-reduces in place while backward continues
        self._garena: Optional[Tensor] = None
        self._gcur = 0
        self.grad_sync = None          # None: single process; else dict(group=..., bucket_bytes=...)
        # multi-head attention backward on the fused tcgen05 kernel where it applies (DMME_FUSED_ATTN_BWD=0: strided products)
        self.fused_attn_bwd = os.environ.get("DMME_FUSED_ATTN_BWD", "1") != "0"
        # split-K plans for the training convs too (DMME_SPLITK_TRAIN=0: unsplit kernels, A/B)
        self.splitk_train = os.environ.get("DMME_SPLITK_TRAIN", "1") != "0"
        self.want_input_grad = False   # set per call when the image tensor requires grad
        self.input_grad: Optional[Tensor] = None
        self.last_buckets = []

    # -- buffers -------------------------------------------------------------------------------
    def _buf(self, tag: str, shape, dtype, dev) -> Tensor:
        self._k += 1
        return self.ws.get(f"train{self._k}.{tag}", shape, dtype, dev)

    def _like(self, tag: str, t: Tensor) -> Tensor:
        return self._buf(tag, tuple(t.shape), t.dtype, t.device)

    def _arena_take(self, shape, dev) -> Tensor:
        numel = 1
        for v in shape:
            numel *= int(v)
        if self._garena is None or self._garena.device != dev:
            need = sum(p.numel() + 4 for p in self.unet.parameters()) + 64
            self._garena = torch.zeros(need, dtype=torch.float32, device=dev)
        if self._gcur + numel > self._garena.numel():
            raise RuntimeError("dmme_b200: gradient arena overflow (parameters were added after the first backward?)")
        g = self._garena[self._gcur:self._gcur + numel].view(tuple(shape))
        self._gcur += (numel + 3) // 4 * 4
        return g

    def _pgrad(self, p: Optional[Tensor]) -> Optional[Tensor]:
        if p is None:
            return None
        g = self._arena_take(tuple(p.shape), p.device)
        self.param_grads[id(p)] = g
        return g

    # -- gradient bookkeeping ------------------------------------------------------------------
    def _contribute(self, t: Tensor, g: Tensor) -> None:
        self.pending.setdefault(t.data_ptr(), []).append(g)

    def _take(self, t: Tensor, keep: int = 1) -> List[Tensor]:
        """Removes and returns the pending contributions of ``t`` reduced to at most ``keep`` tensors."""
        lst = self.pending.pop(t.data_ptr(), [])
        while len(lst) > max(keep, 1):
            a, b = lst.pop(), lst.pop()
            lst.append(ops.add(a, b, out=self._like("gsum", a)))
        return lst

    def _grad_of(self, t: Tensor) -> Optional[Tensor]:
        lst = self._take(t, 1)
        return lst[0] if lst else None

    # -- convolution ---------------------------------------------------------------------------
    def _dgrad_weight(self, conv: nn.Conv2d, off: int, cnt: int, tc: bool) -> Tensor:
        key = ("wd", id(conv), off, cnt, tc)
        if tc and conv.weight.dtype == torch.float32:
            self._pack_specs[key] = (conv.weight, None, 1, off, cnt)
        return self._cached(key, self._ver(conv.weight), lambda: ops.pack_conv_weight_dgrad(conv.weight, off, cnt, tc))

    def _launch(self, d, w, b, out, temb=None, addend=None, stats=None) -> None:
        """conv2d_launch with the split-K workspace where the C side's cost model splits (the 4x4 / 8x8 levels at training
        batch sizes: a 3x3 conv over 2048 pixels is 16 tiles for 148 SMs); same plan as the sampling executor's."""
        ws = None
        if self.splitk_train and not self.force_generic and out is not None and out.dtype == torch.bfloat16:
            ws_bytes = ops.conv_splitk_workspace(d)
            if ws_bytes:
                if self._splitk_ws is None or self._splitk_ws.device != out.device or self._splitk_ws.numel() * 4 < ws_bytes:
                    self._splitk_ws = torch.empty(max(ws_bytes // 4, 1 << 22), dtype=torch.float32, device=out.device)
                ws = self._splitk_ws
        ops.conv2d_launch(d, w, b, out, temb, addend, stats=stats, splitk_ws=ws)

    def conv(self, name: str, src0: Tensor, src1: Optional[Tensor], conv: nn.Conv2d, *, stride: int = 1,
             upsample: bool = False, res: Optional[nn.Conv2d] = None, res0: Optional[Tensor] = None,
             res1: Optional[Tensor] = None, temb: Optional[Tensor] = None, addend: Optional[Tensor] = None,
             in_nchw: bool = False, out_layout: int = L.OUT_NHWC, act_dtype: Optional[torch.dtype] = None,
             temb_cols: Optional[Tuple[int, int]] = None):
        if src1 is not None or out_layout == L.OUT_QKV:
            raise NotImplementedError("training path: convs take one main source and write NHWC / NCHW outputs")
        cout, ks = conv.weight.shape[0], conv.weight.shape[2]
        act_dtype = act_dtype or src0.dtype
        kernel = L.CONV_GENERIC if self.force_generic else L.CONV_AUTO
        src_lo = None  # low-resolution source when the x2 tensor is materialised
        if upsample and not self.force_generic and act_dtype == torch.bfloat16 and src0.shape[3] % 64 == 0 and cout % 64 == 0:
            n, h, w, c = src0.shape
            src_lo = src0
            src0 = ops.upsample2x(src0, out=self._buf(name + ".up", (n, 2 * h, 2 * w, c), src0.dtype, src0.device))
            upsample = False
        d = ops.make_conv_desc(src0, None, cout, ks, stride, upsample, res0, res1, in_nchw, out_layout, act_dtype, kernel)
        tc = ops.conv_uses_tc(d)
        w = self.packed_weight(conv, res, tc)
        b = self.fused_bias(conv, res)
        ho, wo = ops.conv_out_hw(d)
        dev = src0.device
        if out_layout == L.OUT_NCHW_F32:
            out = self._buf(name, (d.n, cout, ho, wo), torch.float32, dev)
            ops.conv2d_launch(d, w, b, out, temb, addend)
        else:
            out = self._buf(name, (d.n, ho, wo, cout), act_dtype, dev)
            stats = self._stats_for(out, d.n, cout) if ops.conv_writes_stats(d) else None
            self._launch(d, w, b, out, temb, addend, stats=stats)

        def backward() -> None:
            g = self._grad_of(out)
            if g is None:
                return
            # stride-2 conv on the tensor-core path: dilate grad_out once, then both gradients are stride-1 problems
            g_w, d_w, g_dil = g, d, None
            if stride == 2 and tc and out_layout == L.OUT_NHWC and not in_nchw:
                g_dil = ops.dilate2x(g, out=self._buf(name + ".gdil", (d.n, 2 * ho, 2 * wo, cout), g.dtype, dev))
                d_w = ops.make_conv_desc(src0, None, cout, ks, 1, False, None, None, False, L.OUT_NHWC, act_dtype, kernel)
                g_w = g_dil
            # weight / bias gradients (the forward descriptor still holds the source pointers and geometry)
            dw, db = self._pgrad(conv.weight), self._pgrad(conv.bias)
            dwr = self._pgrad(res.weight) if res is not None else None
            if res is not None and res.bias is not None:
                self.param_grads[id(res.bias)] = db  # out = conv + res: both biases see the same gradient
            wsz = ops.conv_wgrad_workspace(d_w)
            wsb = self.ws.get("train.wgrad_ws", (max(wsz, 4) // 4,), torch.float32, dev)
            ops.conv2d_wgrad(d_w, g_w, dw, dwr, db, wsb)
            # timestep-embedding gradient: column block of d_all
            if temb is not None and temb_cols is not None:
                o, width = temb_cols
                if temb.shape[0] == d.n:
                    ops.pixel_sum(g, self._d_all[:, o:o + width])
                else:
                    tmp = ops.pixel_sum(g, self._buf("temb_px", (d.n, width), torch.float32, dev))
                    ops.colsum(tmp, self._d_all[0, o:o + width])
            if addend is not None:
                self._contribute(addend, g)
            if in_nchw:
                if self.want_input_grad:
                    # d loss / d image (classifier-guidance style callers): NHWC grad -> NCHW fp32, cout' = image channels
                    cin = conv.weight.shape[1]
                    xd = ops.make_conv_desc(g, None, cin, ks, 1, False, None, None, False, L.OUT_NCHW_F32, act_dtype, kernel)
                    xtc = ops.conv_uses_tc(xd)
                    self.input_grad = self._buf(name + ".gx", (d.n, cin, d.h_in, d.w_in), torch.float32, dev)
                    ops.conv2d_launch(xd, self._dgrad_weight(conv, 0, cin, xtc), None, self.input_grad)
                return
            g_nchw = out_layout == L.OUT_NCHW_F32
            # data gradient of the main source
            cin = conv.weight.shape[1]
            if g_dil is not None:
                dd = ops.make_conv_desc(g_dil, None, cin, ks, 1, False, None, None, False, L.OUT_NHWC, act_dtype, kernel)
            else:
                dd = ops.make_conv_desc(g, None, cin, ks, 1, 2 if stride == 2 else False, None, None, g_nchw, L.OUT_NHWC,
                                        act_dtype, kernel)
            dtc = ops.conv_uses_tc(dd)
            hi, wi = ops.conv_out_hw(dd)
            target = src0 if src_lo is None and not upsample else None
            fuse = self._take(target, 1) if target is not None else []
            gin = self._buf(name + ".gin", (d.n, hi, wi, cin), act_dtype, dev)
            self._launch(dd, self._dgrad_weight(conv, 0, cin, dtc), None, gin, None, fuse[0] if fuse else None)
            if target is not None:
                self._contribute(src0, gin)
            else:  # nearest x2 in front of the conv (models/ddpm.py:161): sum each 2x2 block
                lo = src_lo if src_lo is not None else src0
                self._contribute(lo, ops.pool2x_sum(gin, out=self._like(name + ".gpool", lo)))
            # data gradient through the fused 1x1 residual conv
            off = 0
            for r in (res0, res1):
                if r is None:
                    continue
                cnt = r.shape[3]
                rd = ops.make_conv_desc(g, None, cnt, 1, 1, False, None, None, False, L.OUT_NHWC, act_dtype, kernel)
                rtc = ops.conv_uses_tc(rd)
                fuse = self._take(r, 1)
                gr = self._like(name + ".gres", r)
                self._launch(rd, self._dgrad_weight(res, off, cnt, rtc), None, gr, None, fuse[0] if fuse else None)
                self._contribute(r, gr)
                off += cnt

        self.tape.append(backward)
        return out

    # -- GroupNorm -----------------------------------------------------------------------------
    def gn(self, name: str, norm: nn.GroupNorm, src0: Tensor, src1: Optional[Tensor], silu: bool,
           scale: Optional[Tensor] = None, shift: Optional[Tensor] = None, mask: Optional[Tensor] = None,
           ss_cols: Optional[Tuple[int, int]] = None) -> Tensor:
        n, h, w, c0 = src0.shape
        c = c0 + (src1.shape[3] if src1 is not None else 0)
        dev = src0.device
        out = self._buf(name, (n, h, w, c), src0.dtype, dev)
        st0 = self._stats.get(src0.data_ptr())
        st1 = self._stats.get(src1.data_ptr()) if src1 is not None else None
        ops.groupnorm(src0, src1, norm.num_groups, norm.weight.detach(), norm.bias.detach(), silu, scale, shift, mask,
                      norm.eps, out, st0, st1)

        def backward() -> None:
            g = self._grad_of(out)
            if g is None:
                return
            gin0 = self._like(name + ".gin0", src0)
            gin1 = self._like(name + ".gin1", src1) if src1 is not None else None
            add0 = self._take(src0, 1)
            add1 = self._take(src1, 1) if src1 is not None else []
            dgamma, dbeta = self._pgrad(norm.weight), self._pgrad(norm.bias)
            dscale = dshift = None
            per_image = scale is not None and scale.shape[0] == n
            sc_in, sh_in, tmp_ss = scale, shift, None
            if scale is not None and ss_cols is not None:
                o, cc = ss_cols  # cond = [shift | scale] (models/iddpm.py:116-119)
                if per_image:
                    dshift, dscale = self._d_all[:, o:o + cc], self._d_all[:, o + cc:o + 2 * cc]
                else:
                    # (1,)-shaped timestep: one conditioning row broadcast over the batch; per-image gradients go to a
                    # scratch [n][2C] and are summed over the images afterwards
                    tmp_ss = self._buf(name + ".dss", (n, 2 * cc), torch.float32, dev)
                    dshift, dscale = tmp_ss[:, :cc], tmp_ss[:, cc:]
            sums = self._buf(name + ".sums", (n, c, 2), torch.float32, dev)
            ops.groupnorm_bwd(g, src0, src1, norm.num_groups, norm.weight.detach(), norm.bias.detach(), silu, sc_in, sh_in,
                              mask, norm.eps, gin0, gin1, add0[0] if add0 else None, add1[0] if add1 else None, dgamma,
                              dbeta, dscale, dshift, sums)
            if tmp_ss is not None:
                o, cc = ss_cols
                ops.colsum(tmp_ss, self._d_all[0, o:o + 2 * cc])
            self._contribute(src0, gin0)
            if src1 is not None:
                self._contribute(src1, gin1)

        self.tape.append(backward)
        return out

    # -- attention -----------------------------------------------------------------------------
    def attention_block(self, name: str, att: nn.Module, x: Tensor) -> Tensor:
        n, h, w, c = x.shape
        seq = h * w
        dev = x.device
        a = self.gn(name + ".attn_norm", att.norm, x, None, silu=False)
        heads = getattr(att, "num_heads", None)
        qkv = self.conv(name + ".qkv", a, None, att.qkv_proj)  # NHWC [n, h, w, 3c]
        ao = self._buf(name + ".attn_out", (n, h, w, c), x.dtype, dev)
        flat = qkv.view(-1)
        if heads is None:  # channels [q | k | v] (models/ddpm.py:56-57)
            nh, dh, hs, swap = 1, c, 0, False
            q, k, v = flat, flat[c:], flat[2 * c:]
        else:              # channels [head][q | k | v][dh] (models/iddpm.py:38-39)
            nh, dh, swap = heads, c // heads, True
            hs = 3 * dh
            q, k, v = flat, flat[dh:], flat[2 * dh:]
        fused = (heads is not None and not self.force_generic and self.fused_attn_bwd
                 and ops.attention_bwd_fused_supported(nh, seq, dh, x.dtype))
        if fused:
            # fused tcgen05 backward (csrc/attention_bwd_tc.cu) recomputes the softmax from q and k: the forward is the
            # inference kernel, nothing but its output is kept
            ops.attention(q, k, v, n, nh, seq, dh, att.scale, seq * 3 * c, 3 * c, hs, False, 0, swap, ao)
            p_saved = None
        else:
            # strided products with the softmax matrix kept for the backward pass
            p_saved = self._buf(name + ".attn_p", (n * nh, seq, seq), torch.float32, dev)
            o_tmp = self.ws.get("train.attn_otmp", (n * nh * seq * dh,), torch.float32, dev)
            ops.attention_fwd_train(q, k, v, n, nh, seq, dh, att.scale, seq * 3 * c, 3 * c, hs, swap, ao, p_saved, o_tmp)

        def backward() -> None:
            g = self._grad_of(ao)
            if g is None:
                return
            dqkv = self._like(name + ".dqkv", qkv)
            if fused:
                ops.attention_bwd_fused(qkv, ao, g.contiguous(), dqkv, n, nh, seq, dh, att.scale, swap)
                self._contribute(qkv, dqkv)
                return
            dflat = dqkv.view(-1)
            if heads is None:
                dq, dk, dv = dflat, dflat[c:], dflat[2 * c:]
            else:
                dq, dk, dv = dflat, dflat[dh:], dflat[2 * dh:]
            wsz = ops.attention_bwd_workspace(n, nh, seq, dh)
            wsb = self.ws.get("train.attn_ws", (wsz // 4,), torch.float32, dev)
            ops.attention_bwd(q, k, v, n, nh, seq, dh, att.scale, seq * 3 * c, 3 * c, hs, swap, g, dq, dk, dv, wsb, p_saved)
            self._contribute(qkv, dqkv)

        self.tape.append(backward)
        return self.conv(name + ".attn", ao, None, att.proj, addend=x)

    # -- blocks --------------------------------------------------------------------------------
    def resblock(self, name: str, blk: nn.Module, x0: Tensor, x1: Optional[Tensor], temb_all: Tensor,
                 offs: Dict[int, Tuple[int, int]], masks: Optional[Dict[str, Tensor]]) -> Tensor:
        o, width = offs[id(blk)]
        cond = temb_all[:, o:o + width]
        mask = masks.get(name) if masks else None
        a1 = self.gn(name + ".a1", blk.conv1[0], x0, x1, silu=True)
        conv2 = blk.conv2[-1]
        if self.flavour == "ddpm":
            h1 = self.conv(name + ".h1", a1, None, blk.conv1[2], temb=cond, temb_cols=(o, width))
            a2 = self.gn(name + ".a2", blk.conv2[0], h1, None, silu=True, mask=mask)
        else:
            h1 = self.conv(name + ".h1", a1, None, blk.conv1[2])
            cout = width // 2
            a2 = self.gn(name + ".a2", blk.norm, h1, None, silu=True, shift=cond[:, :cout], scale=cond[:, cout:], mask=mask,
                         ss_cols=(o, cout))
        if isinstance(blk.residual, nn.Identity):
            h2 = self.conv(name + ".h2", a2, None, conv2, addend=x0)
        else:
            h2 = self.conv(name + ".h2", a2, None, conv2, res=blk.residual, res0=x0, res1=x1)
        if not isinstance(blk.attention, nn.Identity):
            h2 = self.attention_block(name, blk.attention, h2)
        return h2

    # -- whole network -------------------------------------------------------------------------
    def forward(self, x: Tensor, c: Tensor, act_dtype: torch.dtype, masks: Optional[Dict[str, Tensor]] = None) -> Tensor:
        u = self.unet
        L.require_cuda(x, c)
        x = x.float().contiguous()
        c = c.long().contiguous()
        if c.dim() != 1 or c.numel() not in (1, x.shape[0]):
            raise ValueError(f"timestep tensor must have shape (1,) or (N,), got {tuple(c.shape)}")
        dev = x.device
        self._act_dtype = act_dtype
        self.tape.clear()
        self.pending.clear()
        self.param_grads.clear()
        self._k = 0
        self._begin_stats(dev)
        cond = u.condition
        rows, emb_dim = c.numel(), cond[3].weight.shape[0]
        hidden = self._buf("temb.hidden", (rows, emb_dim), torch.float32, dev)
        emb = ops.temb_mlp(c, cond[0].embeddings, cond[1].weight.detach(), cond[1].bias.detach(), cond[3].weight.detach(),
                           cond[3].bias.detach(), out=self._buf("temb.emb", (rows, emb_dim), torch.float32, dev), scratch=hidden)
        wcat, bcat, offs = self.temb_tables()
        temb_all = ops.temb_proj(emb, wcat, bcat, out=self._buf("temb.all", (rows, wcat.shape[0]), torch.float32, dev))
        self._d_all = self._buf("temb.d_all", (rows, wcat.shape[0]), torch.float32, dev)
        self._temb_ctx = (c, hidden, emb, wcat)

        h = self.conv("input_conv", x, None, u.input_conv, in_nchw=True, act_dtype=act_dtype)
        skips = [h]
        for i, m in enumerate(u.down_layers):
            name = f"down_layers.{i}"
            if hasattr(m, "conv1"):
                h = self.resblock(name, m, h, None, temb_all, offs, masks)
            else:
                h = self.conv(name, h, None, m, stride=2)
            skips.append(h)
        for i, m in enumerate(u.middle_layers):
            h = self.resblock(f"middle_layers.{i}", m, h, None, temb_all, offs, masks)
        for i, m in enumerate(u.up_layers):
            name = f"up_layers.{i}"
            if hasattr(m, "conv1"):
                h = self.resblock(name, m, h, skips.pop(), temb_all, offs, masks)
            else:
                h = self.conv(name, h, None, m.conv, upsample=True)
        a = self.gn("out_norm", u.output_conv[0], h, None, silu=True)
        self._out = self.conv("output_conv", a, None, u.output_conv[2], out_layout=L.OUT_NCHW_F32)
        return self._out

    def backward(self, d_out: Tensor) -> Dict[int, Tensor]:
        """d_out: gradient of the loss w.r.t. the network output (NCHW fp32).  Returns {id(param): fp32 gradient}."""
        u = self.unet
        dev = d_out.device
        self._contribute(self._out, d_out.float().contiguous())
        self._gcur = 0
        bucketer = None
        if self.grad_sync is not None:
            from ..parallel import GradBucketer
            self._arena_take((0,), dev)  # make sure the arena exists
            bucketer = GradBucketer(self._garena, self.grad_sync.get("group"), self.grad_sync.get("bucket_bytes", 32 << 20))
        for fn in reversed(self.tape):
            fn()
            if bucketer is not None:
                bucketer.mark(self._gcur)
        # conditioning: batched ResBlock projections, then the two-layer MLP
        c, hidden, emb, wcat = self._temb_ctx
        cond = u.condition
        dw1, db1 = self._pgrad(cond[1].weight), self._pgrad(cond[1].bias)
        dw2, db2 = self._pgrad(cond[3].weight), self._pgrad(cond[3].bias)
        dwcat = self._arena_take(tuple(wcat.shape), dev)
        dbcat = self._arena_take((wcat.shape[0],), dev)
        half = cond[0].embeddings.numel()
        wsz = ops.temb_bwd_workspace(c.numel(), half, emb.shape[1])
        wsb = self._buf("temb.bwd_ws", (wsz // 4,), torch.float32, dev)
        ops.temb_bwd(c, cond[0].embeddings, cond[1].weight.detach(), cond[1].bias.detach(), cond[3].weight.detach(),
                     cond[3].bias.detach(), hidden, emb, wcat, self._d_all, dw1, db1, dw2, db2, dwcat, dbcat, wsb,
                     bf16_mma=self._act_dtype == torch.bfloat16 and not self.force_generic)
        _, _, offs = self.temb_tables()
        for _, blk in self.resblocks():
            o, width = offs[id(blk)]
            lin = blk.condition[0]
            self.param_grads[id(lin.weight)] = dwcat[o:o + width]
            self.param_grads[id(lin.bias)] = dbcat[o:o + width]
        if bucketer is not None:
            bucketer.finish(self._gcur)
            self.last_buckets = bucketer.buckets
        self.tape.clear()
        self.pending.clear()
        return self.param_grads
