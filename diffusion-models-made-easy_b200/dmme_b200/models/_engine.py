"""Forward executor shared by both UNet flavours.

The ``nn.Module`` tree (see ``ddpm.py`` / ``iddpm.py``) only *holds* parameters under the reference's
``state_dict`` names.  This executor walks that tree and issues the fused CUDA kernels:

    ResBlock (models/ddpm.py:118-133)      ->  GN+SiLU | conv3x3(+temb) | GN+SiLU(+mask) | conv3x3 (+fused 1x1 residual | +x)
    Attention (models/ddpm.py:54-75)       ->  GN | 1x1 qkv (Q, K, V^T) | attention core | 1x1 proj (+x)
    torch.cat skip (models/ddpm.py:310)    ->  never materialised: two source pointers
    UNet.condition + 22 block Linears      ->  temb_mlp + one batched temb_proj

Packed weights (bf16 [cout][K] for the tcgen05 kernel, fp32 [K][cout] for the generic one) are cached per
conv site and re-packed when a parameter's ``_version`` changes (optimizer step, load_state_dict).
Activation buffers are cached per (site, shape) so a forward pass allocates nothing after the first call
and can be captured in a CUDA graph.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from .. import _lib as L
from .. import ops

Tensor = torch.Tensor


class Workspace:
    """Named, shape-keyed device buffers (stable addresses across calls)."""

    def __init__(self) -> None:
        self._bufs: Dict[Tuple, Tensor] = {}

    def get(self, name: str, shape, dtype, device) -> Tensor:
        key = (name, tuple(shape), dtype, str(device))
        t = self._bufs.get(key)
        if t is None:
            t = torch.empty(tuple(shape), dtype=dtype, device=device)
            self._bufs[key] = t
        return t

    def clear(self) -> None:
        self._bufs.clear()


class Engine:
    def __init__(self, unet: nn.Module, flavour: str) -> None:
        self.unet = unet
        self.flavour = flavour  # "ddpm" | "iddpm"
        self.ws = Workspace()
        self._packed: Dict[Tuple, Tuple[Tuple, Tensor]] = {}
        # (weight, residual weight, dgrad, ci_off, ci_cnt) of every tensor-core weight pack, for pack_batch(); the keys whose
        # cache entries a batched launch refreshes in place
        self._pack_specs: Dict[Tuple, Tuple] = {}
        self._batched_packs: set = set()
        self._blocks: Optional[List[Tuple[str, nn.Module]]] = None
        self.force_generic = False  # debugging / fp32 mode: never take the tcgen05 path
        # True while a training step is captured into a CUDA graph: weight-derived buffers are rebuilt on every call, so
        # the pack kernels become part of the graph and replays see the optimizer's latest weights
        self.always_repack = False
        self._side: Optional[torch.cuda.Stream] = None  # side stream of the timestep-embedding branch
        # GroupNorm + SiLU applied inside the halo conv kernel where it runs (DMME_FUSE_GN=0: always the stand-alone pass)
        self.fuse_gn = os.environ.get("DMME_FUSE_GN", "1") != "0"
        # GroupNorm statistics written by conv epilogues: one zeroed int64 arena per forward pass
        self._arena: Optional[Tensor] = None
        self._arena_cursor = 0
        self._arena_need = 0
        self._stats: Dict[int, Tensor] = {}
        # split-K convs (4x4 / 8x8 levels, small batches): shared fp32 partial-tile workspace, the GroupNorm(+SiLU) outputs
        # their finishing pass wrote for the consumers {(raw tensor ptr, id(norm module)): normalised tensor}, and the
        # static producer -> consumers plan of the topology
        self._splitk_ws: Optional[Tensor] = None
        self._normed: Dict[Tuple[int, int], Tensor] = {}
        self._consumers: Optional[Dict[str, List[Tuple]]] = None
        self.fuse_out_norm = os.environ.get("DMME_FUSE_OUT_NORM", "1") != "0"
        # 8x8 maps: the unsplit transposed conv can finish its consumers' norms in its own epilogue (csrc/conv_tc.cu, NORM).
        # Measured no faster than conv + stand-alone GroupNorm (3.38 vs 3.34 ms per step at batch 256, equal at 128: that
        # epilogue is store-request bound and the norm doubles its stores), so it is opt-in: DMME_EPI_NORM=1
        self.epi_norm = os.environ.get("DMME_EPI_NORM", "0") == "1"
        # sampler update applied by the output conv's epilogue (eps stays in registers); sampler_applied reports whether the
        # last forward did it (the callers run the stand-alone update kernel otherwise)
        self.fuse_sampler = os.environ.get("DMME_FUSE_SAMPLER", "1") != "0"
        # single-head 256-token x 256-channel attention blocks as one launch (DMME_FUSE_ATTN=0: norm | qkv | core | proj launches)
        self.fuse_attn = os.environ.get("DMME_FUSE_ATTN", "1") != "0"
        self.attn16_min_batch = int(os.environ.get("DMME_ATTN16_MIN_BATCH", "128"))
        # the block kernels can form their GroupNorm coefficients themselves from the producer's statistics (six launches fewer
        # per step, same bits); measured within noise of the coefficient launch it replaces (3.33 vs 3.25 - 3.31 ms at batch
        # 256, +0.01 ms at the smaller batches: the dependent loads sit in front of the first tile's norm), so opt-in
        self.attn_coeff_in_kernel = os.environ.get("DMME_ATTN_COEFF_IN_KERNEL", "0") == "1"
        # consecutive ResBlocks of the 8x8 / 4x4 levels in one persistent launch (csrc/conv_chain.cu); DMME_CHAIN=0: per-conv
        # launches as at the higher resolutions
        self.use_chain = os.environ.get("DMME_CHAIN", "1") != "0"
        self.chain_max_hw = int(os.environ.get("DMME_CHAIN_MAX_HW", "8"))  # A/B: 4 = only the 4x4 level
        self._layers: Optional[List[Tuple[str, nn.Module, str, str]]] = None
        self.sampler_applied = False

    # -- caches ------------------------------------------------------------------------------
    def _cached(self, key: Tuple, versions: Tuple, build):
        hit = self._packed.get(key)
        if hit is not None and hit[0] == versions and not self.always_repack:
            return hit[1]
        if hit is not None and self.always_repack and key in self._batched_packs:
            return hit[1]  # refreshed in place by the batched pack launch at the head of the captured step
        val = build()
        self._packed[key] = (versions, val)
        return val

    def pack_batch(self) -> Optional["ops.PackBatch"]:
        """One-launch refresh of every bf16 tensor-core weight pack made so far (``packed_weight`` / ``_dgrad_weight`` with
        ``tc``), writing into the cached buffers; while ``always_repack`` is set those entries are then served from the cache.
        The graph-captured training step calls it at the head of the captured region."""
        entries, keys = [], []
        for key, (w, wres, dgrad, off, cnt) in self._pack_specs.items():
            hit = self._packed.get(key)
            if hit is None or hit[1].dtype != torch.bfloat16:
                continue
            entries.append((w.detach(), wres.detach() if wres is not None else None, hit[1], dgrad, off, cnt))
            keys.append(key)
        if not entries:
            return None
        self._batched_packs = set(keys)
        return ops.PackBatch(entries, entries[0][2].device)

    @staticmethod
    def _ver(*params: Optional[Tensor]) -> Tuple:
        # _dmme_gen: bumped by optim.FusedAdamEMA, whose kernels update the weights without touching torch's counter
        return tuple((p.data_ptr(), p._version, getattr(p, "_dmme_gen", 0)) if p is not None else None for p in params)

    def packed_weight(self, conv: nn.Conv2d, res: Optional[nn.Conv2d], tc: bool) -> Tensor:
        wr = res.weight if res is not None else None
        key = ("w", id(conv), tc)
        if tc and conv.weight.dtype == torch.float32 and (wr is None or wr.dtype == torch.float32):
            self._pack_specs[key] = (conv.weight, wr, 0, 0, 0)
        return self._cached(key, self._ver(conv.weight, wr), lambda: ops.pack_conv_weight(conv.weight, wr, tc))

    def packed_weight_identity(self, conv: nn.Conv2d) -> Tensor:
        """[W | I]: the conv weight with an identity 1x1 residual block appended, so that ``conv(a) + x`` runs as extra
        K chunks of the same tensor-core GEMM (bf16 x times 1.0 accumulates exactly in fp32) instead of an epilogue
        read of x -- measured 36 vs 44 us for the 256-channel attention projection at batch 256."""
        def build():
            c = conv.weight.shape[0]
            eye = torch.eye(c, device=conv.weight.device, dtype=torch.float32).view(c, c, 1, 1)
            return ops.pack_conv_weight(conv.weight, eye, True)
        return self._cached(("w", id(conv), "eye"), self._ver(conv.weight), build)

    def fused_bias(self, conv: nn.Conv2d, res: Optional[nn.Conv2d]) -> Tensor:
        if res is None:
            return conv.bias.detach()
        return self._cached(("b", id(conv)), self._ver(conv.bias, res.bias),
                            lambda: (conv.bias.detach() + res.bias.detach()).contiguous())

    def resblocks(self) -> List[Tuple[str, nn.Module]]:
        if self._blocks is None:
            u = self.unet
            out = []
            for name, lst in (("down_layers", u.down_layers), ("middle_layers", u.middle_layers), ("up_layers", u.up_layers)):
                for i, m in enumerate(lst):
                    if hasattr(m, "conv1"):
                        out.append((f"{name}.{i}", m))
            self._blocks = out
        return self._blocks

    def temb_tables(self) -> Tuple[Tensor, Tensor, Dict[int, Tuple[int, int]]]:
        """Concatenated [total][emb] weight / [total] bias of every ResBlock.condition Linear + column ranges."""
        blocks = self.resblocks()
        lins = [b.condition[0] for _, b in blocks]
        vers = self._ver(*[l.weight for l in lins], *[l.bias for l in lins])

        def build():
            w = torch.cat([l.weight.detach().float() for l in lins], dim=0).contiguous()
            b = torch.cat([l.bias.detach().float() for l in lins], dim=0).contiguous()
            return w, b

        w, b = self._cached(("temb",), vers, build)
        offs, o = {}, 0
        for (_, blk), l in zip(blocks, lins):
            offs[id(blk)] = (o, l.weight.shape[0])
            o += l.weight.shape[0]
        return w, b, offs

    # -- who normalises what ---------------------------------------------------------------------
    def consumers(self) -> Dict[str, List[Tuple]]:
        """producer tensor name -> [(GroupNorm module, first channel inside that norm, channels of that norm, silu)]: the
        norms that read a block / down-sampling output -- the next ResBlock's conv1 norm (part 0 of a concat on the up
        path), the attention norm of its own block, and the up-path ResBlock that later pops it as a skip (part 1)."""
        if self._consumers is not None:
            return self._consumers
        u = self.unet
        plan: Dict[str, List[Tuple]] = {}

        def out_name(name, m):
            return name + ".attn" if not isinstance(m.attention, nn.Identity) else name

        seq = []  # (name, module, is_resblock, section)
        for sec, lst in (("down", u.down_layers), ("mid", u.middle_layers), ("up", u.up_layers)):
            pre = {"down": "down_layers", "mid": "middle_layers", "up": "up_layers"}[sec]
            for i, m in enumerate(lst):
                seq.append((f"{pre}.{i}", m, hasattr(m, "conv1"), sec))
        # channels of every produced tensor
        chan = {"input_conv": u.input_conv.weight.shape[0]}
        skips = ["input_conv"]
        prev = "input_conv"
        for name, m, is_rb, sec in seq:
            if is_rb:
                cout = m.conv1[2].weight.shape[0]
                norm = m.conv1[0]
                if sec == "up":
                    sk = skips.pop()
                    total = chan[prev] + chan[sk]
                    plan.setdefault(prev, []).append((norm, 0, total, True))
                    plan.setdefault(sk, []).append((norm, chan[prev], total, True))
                else:
                    plan.setdefault(prev, []).append((norm, 0, chan[prev], True))
                if not isinstance(m.attention, nn.Identity):
                    plan.setdefault(name + ".pre", []).append((m.attention.norm, 0, cout, False))
                    chan[name + ".pre"] = cout
                prev = out_name(name, m)
                chan[prev] = cout
            else:
                conv = m.conv if hasattr(m, "conv") else m
                prev_c = conv.weight.shape[0]
                prev = name
                chan[prev] = prev_c
            if sec == "down":
                skips.append(prev)
        plan.setdefault(prev, []).append((u.output_conv[0], 0, chan[prev], True))
        # the chain consumer first: when a tensor has more than two readers the skip reader falls back to its own pass
        self._consumers = plan
        return plan

    # -- GroupNorm statistics arena ------------------------------------------------------------
    def _begin_stats(self, device) -> None:
        if self._arena is None or self._arena.device != device or self._arena.numel() < self._arena_need:
            self._arena = torch.zeros(max(self._arena_need, 1 << 18), dtype=torch.int64, device=device)
        else:
            self._arena.zero_()
        self._arena_cursor = 0
        self._arena_need = 0
        self._stats.clear()
        self._normed.clear()

    def _stats_for(self, out: Tensor, n: int, cout: int) -> Optional[Tensor]:
        size = n * (cout // 4) * 2
        self._arena_need += size
        if self._arena_cursor + size > self._arena.numel():
            self._stats.pop(out.data_ptr(), None)
            return None  # arena grows on the next forward; this tensor's consumer reduces its own statistics
        st = self._arena[self._arena_cursor:self._arena_cursor + size]
        self._arena_cursor += size
        self._stats[out.data_ptr()] = st
        return st

    # -- kernels -----------------------------------------------------------------------------
    def conv(self, name: str, src0: Tensor, src1: Optional[Tensor], conv: nn.Conv2d, *, stride: int = 1,
             upsample: bool = False, res: Optional[nn.Conv2d] = None, res0: Optional[Tensor] = None,
             res1: Optional[Tensor] = None, temb: Optional[Tensor] = None, addend: Optional[Tensor] = None,
             in_nchw: bool = False, out_layout: int = L.OUT_NHWC, act_dtype: Optional[torch.dtype] = None,
             addend_in_gemm: bool = False, gn_ab: Optional[Tensor] = None, gn_silu: bool = True,
             consumers: Optional[List[Tuple]] = None, sampler=None):
        """``consumers``: [(norm module, first channel, channels of that norm, silu[, scale, shift])] reading this conv's
        output; when the conv runs split-K its finishing pass writes their GroupNorm(+SiLU) too (``self._normed``)."""
        cout, ks = conv.weight.shape[0], conv.weight.shape[2]
        act_dtype = act_dtype or src0.dtype
        kernel = L.CONV_GENERIC if self.force_generic else L.CONV_AUTO
        identity = False
        if addend_in_gemm and addend is not None and res is None and not self.force_generic and \
                act_dtype == torch.bfloat16 and cout % 64 == 0 and src0.shape[3] % 64 == 0:
            probe = ops.make_conv_desc(src0, src1, cout, ks, stride, upsample, addend, None, in_nchw, out_layout, act_dtype, kernel)
            if ops.conv_uses_tc(probe):
                identity, res0, addend = True, addend, None
        if upsample and not self.force_generic and act_dtype == torch.bfloat16 and src1 is None and res is None \
                and temb is None and addend is None and out_layout == L.OUT_NHWC:
            # nearest x2 + conv3x3 as four 2x2 phase convs on the low-resolution tensor (2.25x fewer FLOPs, no x2 tensor)
            d = ops.make_conv_desc(src0, None, cout, ks, stride, 3, None, None, in_nchw, out_layout, act_dtype, kernel)
            if ops.conv_uses_tc(d):
                w = self._cached(("wup", id(conv)), self._ver(conv.weight), lambda: ops.pack_upsample_phase_weight(conv.weight))
                ho, wo = ops.conv_out_hw(d)
                out = self.ws.get(name, (d.n, ho, wo, cout), act_dtype, src0.device)
                stats = self._stats_for(out, d.n, cout) if ops.conv_writes_stats(d) else None
                if stats is None:
                    self._stats.pop(out.data_ptr(), None)
                ops.conv2d_launch(d, w, conv.bias.detach(), out, stats=stats)
                return out
        if upsample and not self.force_generic and act_dtype == torch.bfloat16 and src0.shape[3] % 64 == 0 and cout % 64 == 0:
            # tensor-core path has no upsampling gather: materialise the x2 tensor once (memory-bound copy)
            n, h, w, c = src0.shape
            src0 = ops.upsample2x(src0, out=self.ws.get(name + ".up", (n, 2 * h, 2 * w, c), src0.dtype, src0.device))
            upsample = False
        d = ops.make_conv_desc(src0, src1, cout, ks, stride, upsample, res0, res1, in_nchw, out_layout, act_dtype, kernel)
        tc = ops.conv_uses_tc(d)
        w = self.packed_weight_identity(conv) if identity else self.packed_weight(conv, res, tc)
        b = self.fused_bias(conv, res)
        ho, wo = ops.conv_out_hw(d)
        dev = src0.device
        if out_layout == L.OUT_NCHW_F32:
            out = self.ws.get(name, (d.n, cout, ho, wo), torch.float32, dev)
            if sampler is not None and self.fuse_sampler and temb is None and addend is None and ops.conv_fuses_sampler(d):
                d.out = None  # eps is consumed in the epilogue and never written
                ops.conv2d_launch(d, w, b, None, sampler=sampler, gn_ab=gn_ab, gn_silu=gn_silu)
                self.sampler_applied = True
                return out
            ops.conv2d_launch(d, w, b, out, temb, addend, gn_ab=gn_ab, gn_silu=gn_silu)
            return out
        if out_layout == L.OUT_QKV:
            c = cout // 3
            q = self.ws.get(name + ".q", (d.n, ho * wo, c), act_dtype, dev)
            k = self.ws.get(name + ".k", (d.n, ho * wo, c), act_dtype, dev)
            vt = self.ws.get(name + ".vt", (d.n, c, ho * wo), act_dtype, dev)
            ops.conv2d_launch(d, w, b, q, temb, addend, k, vt)
            return q, k, vt
        out = self.ws.get(name, (d.n, ho, wo, cout), act_dtype, dev)
        stats = self._stats_for(out, d.n, cout) if ops.conv_writes_stats(d) else None
        if stats is None:
            self._stats.pop(out.data_ptr(), None)  # the buffer may be a reused scratch with stale statistics
        ws_bytes = ops.conv_splitk_workspace(d) if (gn_ab is None and act_dtype == torch.bfloat16 and not self.force_generic) else 0
        # 8x8 maps on the unsplit transposed kernel: the conv's own epilogue finishes the consumers' norms (opt-in, see __init__)
        epi_norm = (not ws_bytes and self.epi_norm and gn_ab is None and act_dtype == torch.bfloat16 and not self.force_generic
                    and self.fuse_out_norm and bool(consumers) and stats is not None and ops.conv_epilogue_norm(d))
        if ws_bytes or epi_norm:
            if ws_bytes and (self._splitk_ws is None or self._splitk_ws.device != dev or self._splitk_ws.numel() * 4 < ws_bytes):
                self._splitk_ws = torch.empty(max(ws_bytes // 4, 1 << 22), dtype=torch.float32, device=dev)
            norms = []
            for k, spec in enumerate((consumers or [])[:2] if self.fuse_out_norm else []):
                norm, off, total, silu = spec[:4]
                scale, shift = (spec[4], spec[5]) if len(spec) > 4 else (None, None)
                cpg = total // norm.num_groups
                if cpg < 1 or 32 % cpg or off % cpg or cout % cpg:
                    continue
                y = self.ws.get(f"{name}.normed{k}", (d.n, ho, wo, cout), act_dtype, dev)
                norms.append(ops.out_norm(y, norm.weight.detach()[off:off + cout], norm.bias.detach()[off:off + cout], cpg,
                                          silu, norm.eps, scale, shift))
                self._normed[(out.data_ptr(), id(norm))] = y
            ops.conv2d_launch(d, w, b, out, temb, addend, stats=stats, splitk_ws=self._splitk_ws if ws_bytes else None,
                              out_norms=norms)
            return out
        ops.conv2d_launch(d, w, b, out, temb, addend, stats=stats, gn_ab=gn_ab, gn_silu=gn_silu)
        return out

    def norm_conv(self, name: str, norm: nn.GroupNorm, x0: Tensor, x1: Optional[Tensor], conv: nn.Conv2d, *,
                  scale: Optional[Tensor] = None, shift: Optional[Tensor] = None, mask: Optional[Tensor] = None,
                  gn_name: str = "scratch.a", **conv_kw) -> Tensor:
        """``conv(silu(norm(cat(x0, x1))))`` (norm_act_drop_conv, models/ddpm.py:25-35).  When the conv takes the halo
        kernel and the producers left their statistics, the norm + SiLU is applied to the halo tile inside the conv kernel
        (one small coefficient launch instead of a pass over the tensor); otherwise GroupNorm runs as its own kernel."""
        c0 = x0.shape[3]
        c1 = x1.shape[3] if x1 is not None else 0
        cpg = (c0 + c1) // norm.num_groups
        if mask is None and self._normed:
            # parts already normalised by the split-K finishing pass of their producer; a part that is not gets its own
            # stand-alone pass (groups never straddle the two parts of a concat)
            n0 = self._normed.get((x0.data_ptr(), id(norm)))
            n1 = self._normed.get((x1.data_ptr(), id(norm))) if x1 is not None else None
            if n0 is not None or n1 is not None:
                if n0 is None and c0 % cpg == 0:
                    n0 = self.gn_part(gn_name + ".p0", norm, x0, 0, cpg, scale, shift)
                if x1 is not None and n1 is None and n0 is not None and c1 % cpg == 0:
                    n1 = self.gn_part(gn_name + ".p1", norm, x1, c0, cpg, scale, shift)
                if n0 is not None and (x1 is None or n1 is not None):
                    return self.conv(name, n0, n1, conv, **conv_kw)
        st0 = self._stats.get(x0.data_ptr())
        st1 = self._stats.get(x1.data_ptr()) if x1 is not None else None
        if (self.fuse_gn and mask is None and not self.force_generic and x0.dtype == torch.bfloat16 and st0 is not None
                and (x1 is None or st1 is not None) and cpg % 4 == 0 and c0 % cpg == 0):
            cout, ks = conv.weight.shape[0], conv.weight.shape[2]
            probe = ops.make_conv_desc(x0, x1, cout, ks, 1, False, conv_kw.get("res0"), conv_kw.get("res1"), False,
                                       conv_kw.get("out_layout", L.OUT_NHWC), x0.dtype, L.CONV_AUTO)
            if ops.conv_fuses_gn(probe):
                n, h, w, _ = x0.shape
                ab = ops.groupnorm_coeff(st0, st1, c0, c1, n, h * w, norm.num_groups, norm.weight.detach(),
                                         norm.bias.detach(), scale, shift, norm.eps,
                                         out=self.ws.get(gn_name + ".ab", (n, c0 + c1, 2), torch.float32, x0.device))
                return self.conv(name, x0, x1, conv, gn_ab=ab, gn_silu=True, **conv_kw)
        a = self.gn(gn_name, norm, x0, x1, silu=True, scale=scale, shift=shift, mask=mask)
        return self.conv(name, a, None, conv, **conv_kw)

    def gn_part(self, name: str, norm: nn.GroupNorm, src: Tensor, off: int, cpg: int, scale: Optional[Tensor] = None,
                shift: Optional[Tensor] = None) -> Tensor:
        """GroupNorm + SiLU of ONE part of a concat (channels [off, off + c) of ``norm``): its groups lie inside the part."""
        n, h, w, c = src.shape
        out = self.ws.get(name, (n, h, w, c), src.dtype, src.device)
        sc = scale[:, off:off + c] if scale is not None else None
        sh = shift[:, off:off + c] if shift is not None else None
        return ops.groupnorm(src, None, c // cpg, norm.weight.detach()[off:off + c], norm.bias.detach()[off:off + c], True,
                             sc, sh, None, norm.eps, out, self._stats.get(src.data_ptr()), None)

    def gn(self, name: str, norm: nn.GroupNorm, src0: Tensor, src1: Optional[Tensor], silu: bool,
           scale: Optional[Tensor] = None, shift: Optional[Tensor] = None, mask: Optional[Tensor] = None) -> Tensor:
        n, h, w, c0 = src0.shape
        c = c0 + (src1.shape[3] if src1 is not None else 0)
        out = self.ws.get(name, (n, h, w, c), src0.dtype, src0.device)
        st0 = self._stats.get(src0.data_ptr())
        st1 = self._stats.get(src1.data_ptr()) if src1 is not None else None
        return ops.groupnorm(src0, src1, norm.num_groups, norm.weight.detach(), norm.bias.detach(), silu, scale, shift,
                             mask, norm.eps, out, st0, st1)

    # -- blocks ------------------------------------------------------------------------------
    def attention_fused(self, att: nn.Module, seq: int, c: int, dtype: torch.dtype, n: int = 1 << 30) -> bool:
        """True when the block runs as the one-launch kernel (csrc/attention_block.cu) given its producer's statistics."""
        if seq == 16 and n < self.attn16_min_batch:
            # the 16-token kernel takes eight images per CTA and ~20 us per CTA whatever the batch (weight streaming and
            # phase latencies): below ~128 images the four small launches it replaces are faster (batch 64: 1.488 vs 1.502 ms
            # per step, batch 32: 1.135 vs 1.150)
            return False
        return (self.fuse_attn and not self.force_generic and getattr(att, "num_heads", None) is None
                and c % att.norm.num_groups == 0 and (c // att.norm.num_groups) % 4 == 0
                and ops.attention_block_supported(1, seq, c, dtype))

    def attention_block(self, name: str, att: nn.Module, x: Tensor) -> Tensor:
        n, h, w, c = x.shape
        seq = h * w
        heads = getattr(att, "num_heads", None)
        st = self._stats.get(x.data_ptr())
        if st is not None and self.attention_fused(att, seq, c, x.dtype, n):
            # the whole block in one launch (csrc/attention_block.cu): norm, qkv, softmax(q k^T) v, proj and + x
            norm = att.norm
            out = self.ws.get(name + ".attn", (n, h, w, c), x.dtype, x.device)
            stats = self._stats_for(out, n, c)
            if stats is None:
                self._stats.pop(out.data_ptr(), None)
            wq, wp = self.packed_weight(att.qkv_proj, None, True), self.packed_weight(att.proj, None, True)
            if self.attn_coeff_in_kernel:
                return ops.attention_block(x, None, wq, att.qkv_proj.bias.detach(), wp, att.proj.bias.detach(), att.scale, out,
                                           stats, stats_in=st, gamma=norm.weight.detach(), beta=norm.bias.detach(),
                                           groups=norm.num_groups, eps=norm.eps)
            ab = ops.groupnorm_coeff(st, None, c, 0, n, seq, norm.num_groups, norm.weight.detach(), norm.bias.detach(), None, None,
                                     norm.eps, out=self.ws.get("scratch.attn_ab", (n, c, 2), torch.float32, x.device))
            return ops.attention_block(x, ab, wq, att.qkv_proj.bias.detach(), wp, att.proj.bias.detach(), att.scale, out, stats)
        a = self._normed.get((x.data_ptr(), id(att.norm)))
        if a is None:
            a = self.gn("scratch.attn_norm", att.norm, x, None, silu=False)
        ao = self.ws.get("scratch.attn_out", (n, h, w, c), x.dtype, x.device)
        if heads is None:
            # single head; scale on K in the reference (models/ddpm.py:58) == scale on the scores
            q, k, vt = self.conv("scratch.qkv", a, None, att.qkv_proj, out_layout=L.OUT_QKV)
            ops.attention(q, k, vt, n, 1, seq, c, att.scale, seq * c, c, 0, True, c * seq, False, ao)
        else:
            # channels are [head][q | k | v][dh] (models/iddpm.py:38-39)
            qkv = self.conv("scratch.qkv", a, None, att.qkv_proj)
            dh = c // heads
            flat = qkv.view(-1)
            ops.attention(flat, flat[dh:], flat[2 * dh:], n, heads, seq, dh, att.scale, seq * 3 * c, 3 * c, 3 * dh,
                          False, 0, True, ao)
        return self.conv(name + ".attn", ao, None, att.proj, addend=x, addend_in_gemm=True)

    def resblock(self, name: str, blk: nn.Module, x0: Tensor, x1: Optional[Tensor], temb_all: Tensor,
                 offs: Dict[int, Tuple[int, int]], masks: Optional[Dict[str, Tensor]]) -> Tensor:
        o, width = offs[id(blk)]
        cond = temb_all[:, o:o + width]
        mask = masks.get(name) if masks else None
        conv2 = blk.conv2[-1]
        has_attn = not isinstance(blk.attention, nn.Identity)
        out_name = name + (".pre" if has_attn else "")
        res_kw = dict(addend=x0) if isinstance(blk.residual, nn.Identity) else dict(res=blk.residual, res0=x0, res1=x1)
        res_kw["consumers"] = self.consumers().get(out_name)
        if has_attn and res_kw["consumers"] and x0.dtype == torch.bfloat16 and \
                self.attention_fused(blk.attention, x0.shape[1] * x0.shape[2], conv2.weight.shape[0], x0.dtype, x0.shape[0]):
            # the one-launch attention block applies its own norm: the producer's finishing pass need not write it
            res_kw["consumers"] = [c for c in res_kw["consumers"] if c[0] is not blk.attention.norm]
        if self.flavour == "ddpm":
            c_mid = blk.conv1[2].weight.shape[0]
            h1 = self.norm_conv("scratch.h1", blk.conv1[0], x0, x1, blk.conv1[2], gn_name="scratch.a1", temb=cond,
                                consumers=None if mask is not None else [(blk.conv2[0], 0, c_mid, True)])
            h2 = self.norm_conv(out_name, blk.conv2[0], h1, None, conv2, mask=mask, gn_name="scratch.a2", **res_kw)
        else:
            cout = width // 2
            shift, scale = cond[:, :cout], cond[:, cout:]
            h1 = self.norm_conv("scratch.h1", blk.conv1[0], x0, x1, blk.conv1[2], gn_name="scratch.a1",
                                consumers=None if mask is not None else [(blk.norm, 0, cout, True, scale, shift)])
            h2 = self.norm_conv(out_name, blk.norm, h1, None, conv2, shift=shift, scale=scale, mask=mask,
                                gn_name="scratch.a2", **res_kw)
        if has_attn:
            h2 = self.attention_block(name, blk.attention, h2)
        return h2

    # -- chains of low-resolution ResBlocks --------------------------------------------------------
    def layer_seq(self) -> List[Tuple[str, nn.Module, str, str]]:
        """[(state_dict prefix, module, kind, section)] of UNet.forward's walk (models/ddpm.py:295-313); kind = "res"
        (ResBlock) | "down" (stride-2 conv) | "up" (UpSample)."""
        if self._layers is None:
            u = self.unet
            seq = []
            for sec, pre, lst in (("down", "down_layers", u.down_layers), ("mid", "middle_layers", u.middle_layers),
                                  ("up", "up_layers", u.up_layers)):
                for i, m in enumerate(lst):
                    kind = "res" if hasattr(m, "conv1") else ("down" if sec == "down" else "up")
                    seq.append((f"{pre}.{i}", m, kind, sec))
            self._layers = seq
        return self._layers

    def chain_run(self, seq, i: int, h: Tensor, skips: List[Tensor], masks) -> int:
        """How many consecutive ResBlocks starting at seq[i] the chain kernel takes (0: none).  A run ends after a block
        with attention, at a resolution change, and wherever a GroupNorm's groups do not fit the epilogue's shuffles."""
        if not self.use_chain or self.force_generic or masks or h.dtype != torch.bfloat16:
            return 0
        n, hh, ww, c_in = h.shape
        cout = ops.CHAIN_COUT
        if hh > self.chain_max_hw or not ops.conv_chain_supported(n, hh, ww, cout):
            return 0
        run, pops, pushed = 0, 0, False
        while i + run < len(seq) and run < ops.CHAIN_MAX_OPS // 2:
            name, blk, kind, sec = seq[i + run]
            if kind != "res":
                break
            conv1, conv2 = blk.conv1[2], blk.conv2[-1]
            c1 = 0
            if sec == "up":
                # the skip this block pops must exist before the launch (pushed by a block outside this run)
                if pushed or pops + 1 > len(skips):
                    break
                pops += 1
                c1 = skips[-pops].shape[3]
            norm1, norm2 = blk.conv1[0], (blk.conv2[0] if self.flavour == "ddpm" else blk.norm)
            ok = (c_in == cout and c1 % 64 == 0 and conv1.weight.shape[0] == cout and conv2.weight.shape[0] == cout
                  and conv1.weight.shape[1] == c_in + c1 and conv1.weight.shape[2] == 3 and conv2.weight.shape[2] == 3
                  and (c_in + c1) % norm1.num_groups == 0 and cout % norm2.num_groups == 0)
            if ok:
                cpg1, cpg2 = (c_in + c1) // norm1.num_groups, cout // norm2.num_groups
                ok = 32 % cpg1 == 0 and c_in % cpg1 == 0 and 32 % cpg2 == 0
            if ok and isinstance(blk.residual, nn.Identity):
                ok = c1 == 0
            elif ok:
                ok = blk.residual.weight.shape[1] == c_in + c1
            if not ok:
                break
            pushed = pushed or sec == "down"
            run += 1
            if not isinstance(blk.attention, nn.Identity):
                break
        return run

    def norm_inputs(self, name: str, norm: nn.GroupNorm, x0: Tensor, x1: Optional[Tensor]) -> Tuple[Tensor, Optional[Tensor]]:
        """The GroupNorm + SiLU'd operand(s) of a conv1 (no scale / shift): what a producer's fused pass already wrote, or
        a stand-alone pass now."""
        c0 = x0.shape[3]
        n0 = self._normed.get((x0.data_ptr(), id(norm)))
        if x1 is None:
            return (n0 if n0 is not None else self.gn(name + ".n0", norm, x0, None, silu=True)), None
        c1 = x1.shape[3]
        cpg = (c0 + c1) // norm.num_groups
        n1 = self._normed.get((x1.data_ptr(), id(norm)))
        if c0 % cpg == 0 and c1 % cpg == 0:
            if n0 is None:
                n0 = self.gn_part(name + ".n0", norm, x0, 0, cpg)
            if n1 is None:
                n1 = self.gn_part(name + ".n1", norm, x1, c0, cpg)
            return n0, n1
        return self.gn(name + ".n01", norm, x0, x1, silu=True), None

    def chain_blocks(self, blocks, h: Tensor, skips: List[Tensor], temb_all: Tensor, offs: Dict[int, Tuple[int, int]]) -> Tensor:
        """``blocks`` consecutive ResBlocks (same resolution, 256 channels) as ONE launch of the chain kernel: per block
        op A = conv1 (+temb) whose epilogue leaves conv2's normalised operand in shared memory, op B = conv2 (+residual)
        whose epilogue stores the raw block output and the GroupNorm(+SiLU) its consumers apply -- the next block's, kept
        in shared memory, and e.g. the up-path skip reader's or the attention norm's, to global memory."""
        n, hh, ww, _ = h.shape
        dev, dt = h.device, h.dtype
        cout = ops.CHAIN_COUT
        chain = []
        x_raw, resident = h, False
        for bi, (name, blk, _, sec) in enumerate(blocks):
            o, width = offs[id(blk)]
            cond = temb_all[:, o:o + width]
            conv1, conv2 = blk.conv1[2], blk.conv2[-1]
            has_attn = not isinstance(blk.attention, nn.Identity)
            out_name = name + (".pre" if has_attn else "")
            x1 = skips.pop() if sec == "up" else None
            norm1 = blk.conv1[0]
            if resident:
                s0 = None
                s1 = None
                if x1 is not None:
                    s1 = self._normed.get((x1.data_ptr(), id(norm1)))
                    if s1 is None:
                        s1 = self.gn_part(name + ".n1", norm1, x1, cout, (cout + x1.shape[3]) // norm1.num_groups)
            else:
                s0, s1 = self.norm_inputs(name, norm1, x_raw, x1)
            if self.flavour == "ddpm":
                norm2, temb, scale, shift = blk.conv2[0], cond, None, None
            else:
                norm2, temb = blk.norm, None
                shift, scale = cond[:, :cout], cond[:, cout:]
            on_a = ops.out_norm(None, norm2.weight.detach(), norm2.bias.detach(), cout // norm2.num_groups, True, norm2.eps,
                                scale, shift)
            chain.append(ops.chain_op(s0, s1, self.packed_weight(conv1, None, True), conv1.bias.detach(), c0=cout, temb=temb,
                                      out_norms=[on_a], keep=0))
            out = self.ws.get(out_name, (n, hh, ww, cout), dt, dev)
            stats = self._stats_for(out, n, cout)
            if stats is None:
                self._stats.pop(out.data_ptr(), None)
            if isinstance(blk.residual, nn.Identity):
                # h + x (models/ddpm.py:131) inside the GEMM: [W | I] weights, the raw input as 1x1 chunks (bf16 x 1.0
                # accumulates exactly in fp32) -- the chunk loads overlap the MMAs, an epilogue addend read would not
                kw = dict(res0=x_raw)
                w2, b2 = self.packed_weight_identity(conv2), conv2.bias.detach()
            else:
                kw = dict(res0=x_raw, res1=x1)
                w2, b2 = self.packed_weight(conv2, blk.residual, True), self.fused_bias(conv2, blk.residual)
            nxt_norm = blocks[bi + 1][1].conv1[0] if bi + 1 < len(blocks) else None
            norms, keep = [], -1
            for k, spec in enumerate((self.consumers().get(out_name) or [])[:2]):
                norm, off, total, silu = spec[:4]
                cpg = total // norm.num_groups
                if cpg < 1 or 32 % cpg or off % cpg or cout % cpg:
                    continue
                gam, bet = norm.weight.detach()[off:off + cout], norm.bias.detach()[off:off + cout]
                if norm is nxt_norm and off == 0:
                    keep = len(norms)
                    norms.append(ops.out_norm(None, gam, bet, cpg, silu, norm.eps))
                else:
                    y = self.ws.get(f"{out_name}.normed{k}", (n, hh, ww, cout), dt, dev)
                    norms.append(ops.out_norm(y, gam, bet, cpg, silu, norm.eps))
                    self._normed[(out.data_ptr(), id(norm))] = y
            if nxt_norm is not None and keep < 0:
                raise RuntimeError(f"dmme_b200: chain planning lost the consumer of {out_name}")
            chain.append(ops.chain_op(None, None, w2, b2, c0=cout, out=out, stats=stats, out_norms=norms, keep=keep, **kw))
            x_raw, resident = out, keep >= 0
            if sec == "down":
                skips.append(out)
        ops.conv_chain(chain, n, hh, ww)
        last_name, last = blocks[-1][0], blocks[-1][1]
        if not isinstance(last.attention, nn.Identity):
            x_raw = self.attention_block(last_name, last.attention, x_raw)
            if blocks[-1][3] == "down":
                skips[-1] = x_raw
        return x_raw

    # -- whole network -------------------------------------------------------------------------
    def forward(self, x: Tensor, c: Tensor, act_dtype: torch.dtype, masks: Optional[Dict[str, Tensor]] = None,
                sampler=None) -> Tensor:
        """``sampler``: an ``ops.sampler_epilogue`` to apply in the output conv's epilogue when the kernel supports it
        (``self.sampler_applied`` tells; the returned eps buffer is then NOT written)."""
        u = self.unet
        self.sampler_applied = False
        L.require_cuda(x, c)
        if x.dtype != torch.float32:
            x = x.float()
        if c.dtype != torch.int64:
            c = c.long()
        if c.dim() != 1 or c.numel() not in (1, x.shape[0]):
            raise ValueError(f"timestep tensor must have shape (1,) or (N,), got {tuple(c.shape)}")
        dev = x.device
        self._begin_stats(dev)
        cond = u.condition
        wcat, bcat, offs = self.temb_tables()
        # the timestep embedding (three small launches) only depends on t: it runs on a side stream -- a parallel branch of
        # the step's CUDA graph -- while the input conv and the first GroupNorm run, and joins before the first use
        main = torch.cuda.current_stream(dev)
        if self._side is None or self._side.device != dev:
            self._side = torch.cuda.Stream(device=dev)
        side = self._side
        side.wait_stream(main)
        with torch.cuda.stream(side):
            emb = ops.temb_mlp(c, cond[0].embeddings, cond[1].weight.detach(), cond[1].bias.detach(), cond[3].weight.detach(),
                               cond[3].bias.detach(), out=self.ws.get("temb.emb", (c.numel(), cond[3].weight.shape[0]), torch.float32, dev),
                               scratch=self.ws.get("temb.hidden", (c.numel(), cond[3].weight.shape[0]), torch.float32, dev))
            temb_all = ops.temb_proj(emb, wcat, bcat, out=self.ws.get("temb.all", (c.numel(), wcat.shape[0]), torch.float32, dev))

        plan = self.consumers()
        h = self.conv("input_conv", x, None, u.input_conv, in_nchw=True, act_dtype=act_dtype)
        main.wait_stream(side)
        skips = [h]
        seq = self.layer_seq()
        i = 0
        while i < len(seq):
            name, m, kind, sec = seq[i]
            if kind == "res":
                run = self.chain_run(seq, i, h, skips, masks)
                if run:
                    h = self.chain_blocks(seq[i:i + run], h, skips, temb_all, offs)
                    i += run
                    continue
                h = self.resblock(name, m, h, skips.pop() if sec == "up" else None, temb_all, offs, masks)
            elif kind == "down":
                h = self.conv(name, h, None, m, stride=2, consumers=plan.get(name))
            else:
                h = self.conv(name, h, None, m.conv, upsample=True)
            if sec == "down":
                skips.append(h)
            i += 1
        # GroupNorm + SiLU inside the output conv on 32x32 maps (csrc/conv_out_tc.cu), its own pass otherwise
        return self.norm_conv("output_conv", u.output_conv[0], h, None, u.output_conv[2], gn_name="scratch.out_norm",
                              out_layout=L.OUT_NCHW_F32, sampler=sampler)
