"""Backward-pass parity on the GPU.

Kernel level: every backward C-ABI entry point against torch autograd (fp32, CPU) of the same op.
Path level: DDPM / IDDPM ``training_step`` loss and parameter gradients against autograd through the CPU oracle
(oracle/dmme_oracle.py) with the same weights, timesteps and noise.  Tolerances: fp32 kernels 1e-4 rel-L2
(accumulation order; 2e-5 typical), whole-network fp32 gradients 1e-3, bf16 mode 5e-2 on gradients (bf16 storage of
activations AND activation gradients; the forward bar of 1e-2 on the output is checked in test_unet_gpu.py)."""
import math

import pytest
import torch
import torch.nn.functional as F

import dmme_oracle as O
from helpers import rel_l2, to_nchw, to_nhwc

pytestmark = pytest.mark.gpu
DEV = "cuda"
TINY = dict(in_channels=3, pos_dim=4, emb_dim=8, num_groups=2, channels_per_depth=(4, 8, 16, 32), num_blocks=3)


def _ops():
    from dmme_b200 import ops, _lib
    return ops, _lib


def _rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


# ---------------------------------------------------------------------------------------------
# strided product
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,n,k,outer,heads", [(70, 33, 50, 1, 1), (64, 64, 16, 3, 2), (5, 130, 257, 2, 1)])
def test_gemm_strided(m, n, k, outer, heads):
    ops, L = _ops()
    a = _rand(outer, heads, m, k, seed=1)
    b = _rand(outer, heads, n, k, seed=2)  # stored transposed: B(k, j) = b[j][k]
    c0 = _rand(outer, heads, m, n, seed=3)
    want = 0.5 * a @ b.transpose(-1, -2) + c0
    cd = c0.clone().to(DEV)
    ops.gemm_strided(a.to(DEV), (heads * m * k, m * k, k, 1), b.to(DEV), (heads * n * k, n * k, 1, k), cd,
                     (heads * m * n, m * n, n, 1), m, n, k, outer, heads, 0.5, True)
    assert rel_l2(cd.cpu(), want) < 1e-5


# ---------------------------------------------------------------------------------------------
# convolution: weight gradient and data gradient
# ---------------------------------------------------------------------------------------------
CONV_BWD = [
    dict(n=2, h=8, cin=5, cout=7, ks=3, stride=1),
    dict(n=3, h=16, cin=16, cout=32, ks=3, stride=2),
    dict(n=2, h=8, cin=8, cout=4, ks=3, stride=1, upsample=True),
    dict(n=2, h=8, cin=12, cout=6, ks=1, stride=1),
    dict(n=2, h=8, cin=6, cin1=10, cout=9, ks=3, stride=1, res=True),
    dict(n=1, h=32, cin=64, cout=64, ks=3, stride=1),
]


def _conv_case(cfg, dtype=torch.float32):
    n, h, cin, cout, ks, stride = (cfg[k] for k in ("n", "h", "cin", "cout", "ks", "stride"))
    cin1 = cfg.get("cin1", 0)
    x = _rand(n, cin, h, h, seed=1).requires_grad_()
    x1 = _rand(n, cin1, h, h, seed=2).requires_grad_() if cin1 else None
    ctot = cin + cin1
    w = (_rand(cout, ctot, ks, ks, seed=3) / math.sqrt(ks * ks * ctot)).requires_grad_()
    b = _rand(cout, seed=4).requires_grad_()
    wres = (_rand(cout, ctot, 1, 1, seed=5) / math.sqrt(ctot)).requires_grad_() if cfg.get("res") else None
    return x, x1, w, b, wres


@pytest.mark.parametrize("cfg", CONV_BWD)
def test_conv_wgrad_generic_fp32(cfg):
    ops, L = _ops()
    x, x1, w, b, wres = _conv_case(cfg)
    up = cfg.get("upsample", False)
    xin = x if x1 is None else torch.cat([x, x1], 1)
    # a conv with a fused residual reads its main input a and the residual sources (x, x1) separately
    a = _rand(*xin.shape, seed=9).requires_grad_() if wres is not None else None
    main = a if a is not None else xin
    if up:
        main = F.interpolate(main, scale_factor=2.0, mode="nearest")
    y = F.conv2d(main, w, b, stride=cfg["stride"], padding=cfg["ks"] // 2)
    if wres is not None:
        y = y + F.conv2d(xin, wres)
    g = _rand(*y.shape, seed=7)
    y.backward(g)

    dt = torch.float32
    if wres is not None:
        s0, s1 = to_nhwc(a.detach(), dt).to(DEV), None
        r0, r1 = to_nhwc(x.detach(), dt).to(DEV), to_nhwc(x1.detach(), dt).to(DEV)
    else:
        s0 = to_nhwc(x.detach(), dt).to(DEV)
        s1 = to_nhwc(x1.detach(), dt).to(DEV) if x1 is not None else None
        r0 = r1 = None
    d = ops.make_conv_desc(s0, s1, cfg["cout"], cfg["ks"], cfg["stride"], up, r0, r1, False, L.OUT_NHWC, dt, L.CONV_GENERIC)
    dw = torch.empty_like(w, device=DEV)
    db = torch.empty(cfg["cout"], device=DEV)
    dwr = torch.empty(cfg["cout"], wres.shape[1], device=DEV) if wres is not None else None
    ws = torch.empty(ops.conv_wgrad_workspace(d) // 4, device=DEV)
    ops.conv2d_wgrad(d, to_nhwc(g, dt).to(DEV), dw, dwr, db, ws)
    assert rel_l2(dw.cpu(), w.grad) < 1e-4
    assert rel_l2(db.cpu(), b.grad) < 1e-4
    if wres is not None:
        assert rel_l2(dwr.cpu(), wres.grad.flatten(1)) < 1e-4


WGRAD_TC = [
    dict(n=4, h=32, cin=128, cout=128, ks=3),
    dict(n=8, h=16, cin=256, cin1=128, cout=256, ks=3, res=True),
    dict(n=16, h=8, cin=64, cout=128, ks=3),                  # 64-channel blocks
    dict(n=6, h=16, cin=256, cout=384, ks=1),                 # 1x1 (qkv projection)
    dict(n=33, h=4, cin=256, cout=256, ks=3),                 # ragged last chunk: out-of-range images are zero-filled
    dict(n=64, h=32, cin=128, cout=128, ks=3),                # many pixel slices
]


@pytest.mark.parametrize("cfg", WGRAD_TC)
def test_conv_wgrad_tc(cfg):
    """tcgen05 weight gradient (MN-major operands straight from the NHWC tensors) against fp32 autograd on the same
    bf16-rounded operands: only the fp32 accumulation order differs."""
    ops, L = _ops()
    n, h, cout, ks = cfg["n"], cfg["h"], cfg["cout"], cfg["ks"]
    cin, cin1 = cfg["cin"], cfg.get("cin1", 0)
    c0 = cin - cin1
    bf = lambda t: t.to(torch.bfloat16).float()
    a = bf(_rand(n, cin, h, h, seed=1))
    w = (_rand(cout, cin, ks, ks, seed=3) / math.sqrt(ks * ks * cin)).requires_grad_()
    b = _rand(cout, seed=4).requires_grad_()
    y = F.conv2d(a, w, b, padding=ks // 2)
    wres = xr = None
    if cfg.get("res"):
        xr = bf(_rand(n, cin, h, h, seed=5))
        wres = (_rand(cout, cin, 1, 1, seed=6) / math.sqrt(cin)).requires_grad_()
        y = y + F.conv2d(xr, wres)
    g = bf(_rand(*y.shape, seed=7))
    y.backward(g)
    dt = torch.bfloat16
    s0 = to_nhwc(a[:, :c0], dt).to(DEV)
    s1 = to_nhwc(a[:, c0:], dt).to(DEV) if cin1 else None
    r0 = to_nhwc(xr[:, :c0], dt).to(DEV) if xr is not None else None
    r1 = to_nhwc(xr[:, c0:], dt).to(DEV) if xr is not None and cin1 else None
    d = ops.make_conv_desc(s0, s1, cout, ks, 1, False, r0, r1, False, L.OUT_NHWC, dt, L.CONV_AUTO)
    assert ops.conv_wgrad_uses_tc(d)
    dw = torch.empty_like(w, device=DEV)
    db = torch.empty(cout, device=DEV)
    dwr = torch.empty(cout, cin, device=DEV) if wres is not None else None
    ws = torch.empty(ops.conv_wgrad_workspace(d) // 4, device=DEV)
    gd = to_nhwc(g, dt).to(DEV)
    ops.conv2d_wgrad(d, gd, dw, dwr, db, ws)
    assert rel_l2(dw.cpu(), w.grad) < 1e-4, rel_l2(dw.cpu(), w.grad)
    assert rel_l2(db.cpu(), b.grad) < 1e-4
    if wres is not None:
        assert rel_l2(dwr.cpu(), wres.grad.flatten(1)) < 1e-4
    # and the CUDA-core kernel agrees on the same inputs
    d.kernel = L.CONV_GENERIC
    assert not ops.conv_wgrad_uses_tc(d)
    dw2 = torch.empty_like(dw)
    ws2 = torch.empty(ops.conv_wgrad_workspace(d) // 4, device=DEV)
    ops.conv2d_wgrad(d, gd, dw2, torch.empty_like(dwr) if dwr is not None else None, torch.empty_like(db), ws2)
    assert rel_l2(dw2.cpu(), dw.cpu()) < 1e-4


def test_conv_wgrad_nchw_image_ends():
    """input conv reads the NCHW fp32 image; output conv's grad_out is the NCHW fp32 image gradient."""
    ops, L = _ops()
    x = _rand(2, 3, 16, 16, seed=1)
    w = (_rand(8, 3, 3, 3, seed=2) / 5).requires_grad_()
    b = _rand(8, seed=3).requires_grad_()
    y = F.conv2d(x, w, b, padding=1)
    g = _rand(*y.shape, seed=4)
    y.backward(g)
    xd = x.to(DEV)  # the descriptor holds a raw pointer: keep the tensor alive
    d = ops.make_conv_desc(xd, None, 8, 3, 1, False, None, None, True, L.OUT_NHWC, torch.float32, L.CONV_GENERIC)
    dw, db = torch.empty(8, 3, 3, 3, device=DEV), torch.empty(8, device=DEV)
    ws = torch.empty(ops.conv_wgrad_workspace(d) // 4, device=DEV)
    ops.conv2d_wgrad(d, to_nhwc(g, torch.float32).to(DEV), dw, None, db, ws)
    assert rel_l2(dw.cpu(), w.grad) < 1e-4 and rel_l2(db.cpu(), b.grad) < 1e-4

    a = _rand(2, 8, 16, 16, seed=5)
    w2 = (_rand(3, 8, 3, 3, seed=6) / 8).requires_grad_()
    b2 = _rand(3, seed=7).requires_grad_()
    y2 = F.conv2d(a, w2, b2, padding=1)
    g2 = _rand(*y2.shape, seed=8)
    y2.backward(g2)
    ad = to_nhwc(a, torch.float32).to(DEV)
    d2 = ops.make_conv_desc(ad, None, 3, 3, 1, False, None, None, False, L.OUT_NCHW_F32, torch.float32, L.CONV_GENERIC)
    dw2, db2 = torch.empty(3, 8, 3, 3, device=DEV), torch.empty(3, device=DEV)
    ws = torch.empty(ops.conv_wgrad_workspace(d2) // 4, device=DEV)
    ops.conv2d_wgrad(d2, g2.to(DEV).contiguous(), dw2, None, db2, ws)
    assert rel_l2(dw2.cpu(), w2.grad) < 1e-4 and rel_l2(db2.cpu(), b2.grad) < 1e-4


@pytest.mark.parametrize("cfg", [c for c in CONV_BWD if not c.get("res") and not c.get("upsample")])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_conv_dgrad(cfg, mode):
    """dgrad = the forward conv kernels on grad_out with dmme_pack_conv_weight_dgrad weights."""
    ops, L = _ops()
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    x, x1, w, b, _ = _conv_case(cfg)
    xin = x if x1 is None else torch.cat([x, x1], 1)
    wq = w.detach().to(dt).float().requires_grad_() if mode == "bf16" else w
    y = F.conv2d(xin, wq, b, stride=cfg["stride"], padding=cfg["ks"] // 2)
    g = _rand(*y.shape, seed=7).to(dt).float()
    y.backward(g)
    kernel = L.CONV_GENERIC if mode == "fp32" else L.CONV_AUTO
    gd = to_nhwc(g, dt).to(DEV)
    off = 0
    for src in (x, x1):
        if src is None:
            continue
        cnt = src.shape[1]
        dd = ops.make_conv_desc(gd, None, cnt, cfg["ks"], 1, 2 if cfg["stride"] == 2 else False, None, None, False,
                                L.OUT_NHWC, dt, kernel)
        tc = ops.conv_uses_tc(dd)
        wp = ops.pack_conv_weight_dgrad(w.detach().to(DEV), off, cnt, tc)
        out = torch.empty((cfg["n"], cfg["h"], cfg["h"], cnt), dtype=dt, device=DEV)
        ops.conv2d_launch(dd, wp, None, out)
        assert rel_l2(to_nchw(out.cpu()), src.grad) < (1e-4 if mode == "fp32" else 6e-3), (tc, cnt)
        off += cnt


def test_pool2x_pixel_sum_add():
    ops, L = _ops()
    g = _rand(2, 8, 6, 5, seed=1)
    want = F.avg_pool2d(to_nchw(g), 2) * 4
    assert rel_l2(to_nchw(ops.pool2x_sum(g.to(DEV)).cpu()), want) < 1e-6
    out = torch.zeros(2, 11, device=DEV)
    ops.pixel_sum(g.to(DEV), out[:, 3:8])
    assert rel_l2(out[:, 3:8].cpu(), g.sum(dim=(1, 2))) < 1e-6 and float(out[:, :3].abs().sum()) == 0
    a, b = _rand(3, 4, 4, 6, seed=2).bfloat16(), _rand(3, 4, 4, 6, seed=3).bfloat16()
    assert torch.equal(ops.add(a.to(DEV), b.to(DEV)).cpu(), (a.float() + b.float()).bfloat16())


# ---------------------------------------------------------------------------------------------
# GroupNorm (+ scale/shift) (+ SiLU) (+ mask) backward
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [
    dict(n=2, c0=8, c1=0, h=8, groups=2, silu=True),
    dict(n=3, c0=32, c1=32, h=4, groups=32, silu=True, mask=True, adds=True),
    dict(n=2, c0=16, c1=0, h=8, groups=4, silu=True, ss=True, mask=True),
    dict(n=2, c0=12, c1=4, h=6, groups=8, silu=False),
])
def test_groupnorm_bwd(cfg):
    ops, L = _ops()
    n, c0, c1, h, groups = (cfg[k] for k in ("n", "c0", "c1", "h", "groups"))
    C = c0 + c1
    x0 = _rand(n, c0, h, h, seed=1).requires_grad_()
    x1 = _rand(n, c1, h, h, seed=2).requires_grad_() if c1 else None
    gamma = (1 + 0.3 * _rand(C, seed=3)).requires_grad_()
    beta = (0.2 * _rand(C, seed=4)).requires_grad_()
    scale = (0.5 * _rand(n, C, seed=5)).requires_grad_() if cfg.get("ss") else None
    shift = (0.5 * _rand(n, C, seed=6)).requires_grad_() if cfg.get("ss") else None
    mask = ((torch.rand(n, C, generator=torch.Generator().manual_seed(7)) > 0.3).float() / 0.7) if cfg.get("mask") else None
    xin = x0 if x1 is None else torch.cat([x0, x1], 1)
    y = F.group_norm(xin, groups, gamma, beta, 1e-5)
    if scale is not None:
        y = y * (1 + scale[:, :, None, None]) + shift[:, :, None, None]
    if cfg["silu"]:
        y = F.silu(y)
    if mask is not None:
        y = y * mask[:, :, None, None]
    g = _rand(*y.shape, seed=8)
    y.backward(g)
    dt = torch.float32
    add0 = _rand(n, h, h, c0, seed=9) if cfg.get("adds") else None
    add1 = _rand(n, h, h, c1, seed=10) if cfg.get("adds") and c1 else None
    s0 = to_nhwc(x0.detach(), dt).to(DEV)
    s1 = to_nhwc(x1.detach(), dt).to(DEV) if c1 else None
    gin0 = torch.empty_like(s0)
    gin1 = torch.empty_like(s1) if c1 else None
    dgamma, dbeta = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    dss = torch.zeros(n, 2 * C + 3, device=DEV) if scale is not None else None
    sums = torch.empty(n, C, 2, device=DEV)
    ops.groupnorm_bwd(to_nhwc(g, dt).to(DEV), s0, s1, groups, gamma.detach().to(DEV), beta.detach().to(DEV), cfg["silu"],
                      scale.detach().to(DEV) if scale is not None else None,
                      shift.detach().to(DEV) if shift is not None else None,
                      mask.to(DEV) if mask is not None else None, 1e-5, gin0, gin1,
                      add0.to(DEV) if add0 is not None else None, add1.to(DEV) if add1 is not None else None, dgamma, dbeta,
                      dss[:, C + 3:] if dss is not None else None, dss[:, :C] if dss is not None else None, sums)
    want0 = x0.grad + (to_nchw(add0) if add0 is not None else 0)
    assert rel_l2(to_nchw(gin0.cpu()), want0) < 1e-4
    if c1:
        want1 = x1.grad + (to_nchw(add1) if add1 is not None else 0)
        assert rel_l2(to_nchw(gin1.cpu()), want1) < 1e-4
    assert rel_l2(dgamma.cpu(), gamma.grad) < 1e-4 and rel_l2(dbeta.cpu(), beta.grad) < 1e-4
    if scale is not None:
        assert rel_l2(dss[:, C + 3:].cpu(), scale.grad) < 1e-4 and rel_l2(dss[:, :C].cpu(), shift.grad) < 1e-4


@pytest.mark.parametrize("cfg", [
    dict(n=3, c0=128, c1=0, h=16, groups=32, silu=True, mask=True),
    dict(n=2, c0=128, c1=128, h=32, groups=32, silu=True, adds=True),
    dict(n=2, c0=256, c1=0, h=8, groups=32, silu=True, ss=True),
    dict(n=5, c0=64, c1=0, h=4, groups=32, silu=False),
])
def test_groupnorm_bwd_bf16_slab(cfg):
    """register-resident slab kernel (bf16 storage) against fp32 autograd on the same bf16-rounded tensors; the only
    differences are fp32 summation order and the bf16 rounding of the stored gradient (2^-9 relative)."""
    ops, L = _ops()
    n, c0, c1, h, groups = (cfg[k] for k in ("n", "c0", "c1", "h", "groups"))
    C = c0 + c1
    bf = lambda t: t.to(torch.bfloat16).float()
    x0 = bf(_rand(n, c0, h, h, seed=1)).requires_grad_()
    x1 = bf(_rand(n, c1, h, h, seed=2)).requires_grad_() if c1 else None
    gamma = (1 + 0.3 * _rand(C, seed=3)).requires_grad_()
    beta = (0.2 * _rand(C, seed=4)).requires_grad_()
    scale = (0.5 * _rand(n, C, seed=5)).requires_grad_() if cfg.get("ss") else None
    shift = (0.5 * _rand(n, C, seed=6)).requires_grad_() if cfg.get("ss") else None
    mask = ((torch.rand(n, C, generator=torch.Generator().manual_seed(7)) > 0.3).float() / 0.7) if cfg.get("mask") else None
    xin = x0 if x1 is None else torch.cat([x0, x1], 1)
    y = F.group_norm(xin, groups, gamma, beta, 1e-5)
    if scale is not None:
        y = y * (1 + scale[:, :, None, None]) + shift[:, :, None, None]
    if cfg["silu"]:
        y = F.silu(y)
    if mask is not None:
        y = y * mask[:, :, None, None]
    g = bf(_rand(*y.shape, seed=8))
    y.backward(g)
    dt = torch.bfloat16
    add0 = bf(_rand(n, h, h, c0, seed=9)) if cfg.get("adds") else None
    add1 = bf(_rand(n, h, h, c1, seed=10)) if cfg.get("adds") and c1 else None
    s0 = to_nhwc(x0.detach(), dt).to(DEV)
    s1 = to_nhwc(x1.detach(), dt).to(DEV) if c1 else None
    gin0 = torch.empty_like(s0)
    gin1 = torch.empty_like(s1) if c1 else None
    dgamma, dbeta = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    dss = torch.zeros(n, 2 * C, device=DEV) if scale is not None else None
    sums = torch.empty(n, C, 2, device=DEV)
    ops.groupnorm_bwd(to_nhwc(g, dt).to(DEV), s0, s1, groups, gamma.detach().to(DEV), beta.detach().to(DEV), cfg["silu"],
                      scale.detach().to(DEV) if scale is not None else None,
                      shift.detach().to(DEV) if shift is not None else None,
                      mask.to(DEV) if mask is not None else None, 1e-5, gin0, gin1,
                      add0.to(dt).to(DEV) if add0 is not None else None, add1.to(dt).to(DEV) if add1 is not None else None,
                      dgamma, dbeta, dss[:, C:] if dss is not None else None, dss[:, :C] if dss is not None else None, sums)
    want0 = x0.grad + (to_nchw(add0) if add0 is not None else 0)
    assert rel_l2(to_nchw(gin0.cpu()), want0) < 4e-3
    if c1:
        want1 = x1.grad + (to_nchw(add1) if add1 is not None else 0)
        assert rel_l2(to_nchw(gin1.cpu()), want1) < 4e-3
    assert rel_l2(dgamma.cpu(), gamma.grad) < 1e-4 and rel_l2(dbeta.cpu(), beta.grad) < 1e-4
    if scale is not None:
        assert rel_l2(dss[:, C:].cpu(), scale.grad) < 1e-4 and rel_l2(dss[:, :C].cpu(), shift.grad) < 1e-4


# ---------------------------------------------------------------------------------------------
# attention backward
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,c,heads,L_,swap", [(2, 16, 1, 64, False), (3, 32, 4, 16, True), (2, 64, 4, 64, True)])
def test_attention_bwd(n, c, heads, L_, swap):
    ops, L = _ops()
    dh = c // heads
    qkv = _rand(n, L_, 3 * c, seed=1).requires_grad_()
    scale = c ** -0.5
    if heads == 1:
        q, k, v = qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:]
        out = torch.softmax(q @ (k * scale).transpose(1, 2), dim=2) @ v
    else:  # models/iddpm.py:36-47: channels [head][q|k|v][dh], "(b head)" folded, "(head b)" unfolded
        t = qkv.reshape(n, L_, heads, 3, dh).permute(3, 0, 2, 1, 4).reshape(3, n * heads, L_, dh)
        o = torch.softmax(t[0] @ (t[1] * scale).transpose(1, 2), dim=2) @ t[2]
        out = o.reshape(heads, n, L_, dh).permute(1, 2, 0, 3).reshape(n, L_, c)
    g = _rand(*out.shape, seed=2)
    out.backward(g)
    qd = qkv.detach().to(DEV)
    flat = qd.view(-1)
    dqkv = torch.empty_like(qd)
    dflat = dqkv.view(-1)
    step = c if heads == 1 else dh
    hs = 0 if heads == 1 else 3 * dh
    ws = torch.empty(ops.attention_bwd_workspace(n, heads, L_, dh) // 4, device=DEV)
    ops.attention_bwd(flat, flat[step:], flat[2 * step:], n, heads, L_, dh, scale, L_ * 3 * c, 3 * c, hs, swap, g.to(DEV),
                      dflat, dflat[step:], dflat[2 * step:], ws)
    assert rel_l2(dqkv.cpu(), qkv.grad) < 1e-4
    # training-mode forward (keeps the softmax matrix) and the backward that reuses it
    od = torch.empty(n, L_, c, device=DEV)
    psave = torch.empty(n * heads, L_, L_, device=DEV)
    otmp = torch.empty(n * heads * L_ * dh, device=DEV)
    ops.attention_fwd_train(flat, flat[step:], flat[2 * step:], n, heads, L_, dh, scale, L_ * 3 * c, 3 * c, hs, swap, od, psave, otmp)
    assert rel_l2(od.cpu(), out.detach()) < 1e-5
    dqkv2 = torch.empty_like(qd)
    d2 = dqkv2.view(-1)
    ops.attention_bwd(flat, flat[step:], flat[2 * step:], n, heads, L_, dh, scale, L_ * 3 * c, 3 * c, hs, swap, g.to(DEV),
                      d2, d2[step:], d2[2 * step:], ws, psave)
    assert rel_l2(dqkv2.cpu(), qkv.grad) < 1e-4


@pytest.mark.parametrize("n,heads,L_,swap,dh", [(3, 4, 256, True, 64), (2, 1, 256, False, 64), (5, 4, 64, True, 64), (4, 2, 64, False, 64),
                                                 (1, 4, 64, True, 64), (128, 4, 256, True, 64),
                                                 # 32-channel heads (the 128-channel 16x16 sites of the IDDPM UNet, models/iddpm.py)
                                                 (3, 4, 256, True, 32), (2, 2, 256, False, 32), (5, 4, 64, True, 32), (1, 1, 256, False, 32),
                                                 (128, 4, 256, True, 32)])
def test_attention_bwd_fused(n, heads, L_, swap, dh):
    """fused tcgen05 attention backward (csrc/attention_bwd_tc.cu) against autograd through fp32 attention on the SAME
    bf16-rounded q, k, v and output gradient: what differs is the bf16 rounding of P and dS inside the kernel (2^-9 relative
    per element) and of the stored gradients -- rel-L2 <= 1.5e-2 per gradient"""
    ops, L = _ops()
    c = heads * dh
    nn_ = min(n, 6)  # the reference is computed for the first images only (the batch-128 case checks the grid / regrouping)
    qkv = (_rand(n, L_, 3 * c, seed=11) * 0.7).to(torch.bfloat16).float().requires_grad_()
    scale = c ** -0.5
    t = qkv.reshape(n, L_, heads, 3, dh).permute(3, 0, 2, 1, 4).reshape(3, n * heads, L_, dh)
    o = torch.softmax(t[0] @ (t[1] * scale).transpose(1, 2), dim=2) @ t[2]
    if swap:
        out = o.reshape(heads, n, L_, dh).permute(1, 2, 0, 3).reshape(n, L_, c)
    else:
        out = o.reshape(n, heads, L_, dh).permute(0, 2, 1, 3).reshape(n, L_, c)
    g = _rand(*out.shape, seed=12).to(torch.bfloat16).float()
    out.backward(g)
    qd = qkv.detach().to(torch.bfloat16).to(DEV)
    od = torch.empty(n, L_, c, dtype=torch.bfloat16, device=DEV)
    flat = qd.view(-1)
    ops.attention(flat, flat[dh:], flat[2 * dh:], n, heads, L_, dh, scale, L_ * 3 * c, 3 * c, 3 * dh, False, 0, swap, od)
    assert rel_l2(od.float().cpu()[:nn_], out.detach()[:nn_]) < 8e-3
    dqkv = torch.full_like(qd, float("nan"))
    assert ops.attention_bwd_fused_supported(heads, L_, dh, torch.bfloat16)
    ops.attention_bwd_fused(qd, od, g.to(torch.bfloat16).to(DEV), dqkv, n, heads, L_, dh, scale, swap)
    torch.cuda.synchronize()
    got = dqkv.float().cpu().reshape(n, L_, heads, 3, dh)
    want = qkv.grad.reshape(n, L_, heads, 3, dh)
    assert torch.isfinite(got).all()
    for j, name in enumerate(("dq", "dk", "dv")):
        e = rel_l2(got[..., j, :], want[..., j, :])
        assert e < 1.5e-2, f"{name}: rel-L2 {e}"


@pytest.mark.parametrize("n,c,heads,L_", [(3, 256, 4, 256), (2, 128, 4, 256), (5, 256, 4, 64), (1, 64, 1, 256), (2, 128, 4, 64)])
def test_attention_fwd_train_bf16_keeps_softmax(n, c, heads, L_):
    """bf16 training forward of the multi-head layout (tcgen05 kernels at 256 tokens and at 64 tokens with 64-channel heads,
    mma.sync elsewhere): the kept fp32 softmax matrix and the output against fp32 attention on the same bf16 operands"""
    ops, L = _ops()
    dh = c // heads
    g = torch.Generator().manual_seed(21)
    qkv = torch.randn(n, L_, 3 * c, generator=g).to(torch.bfloat16)
    t = qkv.float().reshape(n, L_, heads, 3, dh).permute(3, 0, 2, 1, 4).reshape(3, n * heads, L_, dh)
    scale = c ** -0.5
    prob = torch.softmax(t[0] @ (t[1] * scale).transpose(1, 2), dim=2)
    want = (prob @ t[2]).reshape(heads, n, L_, dh).permute(1, 2, 0, 3).reshape(n, L_, c)
    dev = qkv.to(DEV).contiguous()
    flat = dev.view(-1)
    od = torch.full((n, L_, c), float("nan"), dtype=torch.bfloat16, device=DEV)
    psave = torch.full((n * heads, L_, L_), float("nan"), device=DEV)
    otmp = torch.empty(n * heads * L_ * dh, device=DEV)
    ops.attention_fwd_train(flat, flat[dh:], flat[2 * dh:], n, heads, L_, dh, scale, L_ * 3 * c, 3 * c, 3 * dh, True, od, psave, otmp)
    torch.cuda.synchronize()
    assert rel_l2(psave.cpu(), prob) < 1e-4, rel_l2(psave.cpu(), prob)
    assert rel_l2(od.float().cpu(), want) < 6e-3, rel_l2(od.float().cpu(), want)


# ---------------------------------------------------------------------------------------------
# conditioning MLP backward
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,pos,emb,total,mma", [(5, 128, 512, 300, False), (3, 4, 8, 20, False), (128, 128, 512, 3000, True),
                                                    (7, 128, 512, 300, True)])
def test_temb_bwd(rows, pos, emb, total, mma):
    ops, L = _ops()
    half = pos // 2
    freq = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1))).unsqueeze(0)
    t = torch.randint(1, 1000, (rows,), generator=torch.Generator().manual_seed(1))
    w1 = (_rand(emb, pos, seed=2) / math.sqrt(pos)).requires_grad_()
    b1 = (0.1 * _rand(emb, seed=3)).requires_grad_()
    w2 = (_rand(emb, emb, seed=4) / math.sqrt(emb)).requires_grad_()
    b2 = (0.1 * _rand(emb, seed=5)).requires_grad_()
    wc = (_rand(total, emb, seed=6) / math.sqrt(emb)).requires_grad_()
    bc = (0.1 * _rand(total, seed=7)).requires_grad_()
    e = t[:, None] * freq
    s = torch.cat([e.sin(), e.cos()], dim=1)
    hid = F.silu(F.linear(s, w1, b1))
    em = F.silu(F.linear(hid, w2, b2))
    allp = F.linear(em, wc, bc)
    g = _rand(rows, total, seed=8)
    allp.backward(g)
    D = lambda x: x.detach().to(DEV).contiguous()
    outs = [torch.empty_like(D(p)) for p in (w1, b1, w2, b2, wc, bc)]
    ws = torch.empty(ops.temb_bwd_workspace(rows, half, emb) // 4, device=DEV)
    ops.temb_bwd(t.to(DEV), freq.to(DEV), D(w1), D(b1), D(w2), D(b2), D(hid), D(em), D(wc), g.to(DEV), *outs, ws, bf16_mma=mma)
    # bf16_mma (bf16 training mode): the two products over the batched projection round their operands to bf16 (2^-9 per
    # element, fp32 accumulation)
    for got, p in zip(outs, (w1, b1, w2, b2, wc, bc)):
        assert rel_l2(got.cpu(), p.grad) < (6e-3 if mma else 1e-4)


# ---------------------------------------------------------------------------------------------
# losses
# ---------------------------------------------------------------------------------------------
def test_mse_loss_value_and_gradient():
    ops, L = _ops()
    eps = _rand(4, 3, 32, 32, seed=1).requires_grad_()
    noise = _rand(4, 3, 32, 32, seed=2)
    want = F.mse_loss(noise, eps)
    want.backward()
    d = torch.empty(4, 3, 32, 32, device=DEV)
    got = ops.mse_loss(eps.detach().to(DEV), noise.to(DEV), d)
    assert abs(float(got[0]) - float(want.detach())) < 1e-6 * abs(float(want.detach()))
    assert rel_l2(d.cpu(), eps.grad) < 1e-6


def _oracle_iddpm_loss(dtype, tabs, t, x0, z, out_init, loss_type):
    tb = [x.to(dtype) for x in tabs]
    out = out_init.detach().to(dtype).clone().requires_grad_()
    x_t, qm, qs = O.forward_noising(x0.to(dtype), t, z.to(dtype), tb[2])
    eps, var = O.iddpm_split(out, t, tb)
    vlb = O.vlb_loss(eps, var, x_t, t, x0.to(dtype), tb)
    simple = O.ddpm_loss(x_t, qm, qs, eps)
    total = simple + 0.001 * vlb if loss_type == "hybrid" else vlb
    total.backward()
    return float(total.detach()), float(simple.detach()), float(vlb.detach()), out.grad.double(), x_t


@pytest.mark.parametrize("schedule", ["cosine", "linear"])
@pytest.mark.parametrize("loss_type", ["hybrid", "vlb"])
def test_iddpm_loss_value_and_gradient(schedule, loss_type):
    """The t = 1 discrete-NLL term is a difference of two normal CDFs at a standard deviation down to 1e-6: in fp32 it
    cancels catastrophically, so the reference's own fp32 arithmetic (the oracle) is far from its fp64 evaluation on
    part of the elements (measured here: 1e-3 on the loss, 0.7-0.9 rel-L2 on the gradient).  Parity is therefore
    asserted (a) on the loss within 3x the fp32 oracle's own distance from fp64 and (b) on the gradient over the
    elements where fp32 and fp64 oracle agree to 1e-3 (the well-conditioned ones, required to be > 90% of all)."""
    ops, L = _ops()
    T = 100
    tabs = O.cosine_tables(T) if schedule == "cosine" else O.linear_tables(T)
    n = 6
    t = torch.tensor([1, 1, 2, 50, 99, 17])
    x0 = _rand(n, 3, 16, 16, seed=1).clamp(-1, 1)
    x0[0, 0, 0, :4] = torch.tensor([1.0, -1.0, 0.999, -0.999])  # the edge bins of the discrete NLL
    z = _rand(n, 3, 16, 16, seed=2)
    out0 = 0.5 * _rand(n, 6, 16, 16, seed=3)
    w32, s32, v32, g32, x_t = _oracle_iddpm_loss(torch.float32, tabs, t, x0, z, out0, loss_type)
    w64, s64, v64, g64, _ = _oracle_iddpm_loss(torch.float64, tabs, t, x0, z, out0, loss_type)
    ws, wv = (1.0, 0.001) if loss_type == "hybrid" else (0.0, 1.0)
    d = torch.empty(n, 6, 16, 16, device=DEV)
    got = ops.iddpm_loss(out0.to(DEV), x_t.to(DEV), x0.to(DEV), t.to(DEV), *[tb.to(DEV) for tb in tabs], ws, wv, d)
    assert abs(float(got[1]) - s64) < 2e-5 * abs(s64)
    assert abs(float(got[2]) - v64) < 3 * abs(v32 - v64) + 1e-5 * abs(v64), (float(got[2]), v32, v64)
    assert abs(float(got[0]) - w64) < 3 * abs(w32 - w64) + 1e-5 * abs(w64), (float(got[0]), w32, w64)
    good = (g32 - g64).abs() <= 1e-3 * g64.abs() + 1e-12
    assert float(good.float().mean()) > 0.9
    dd = d.cpu().double()
    assert float(((dd - g64) * good).norm() / (g64 * good).norm()) < 2e-3


# ---------------------------------------------------------------------------------------------
# whole training step against autograd through the oracle
# ---------------------------------------------------------------------------------------------
def _oracle_training(flavour, sd, x0, t, z, tables, groups, masks=None, loss_type="hybrid", gamma=0.001):
    params = {k: v.clone().requires_grad_(v.is_floating_point() and "embeddings" not in k) for k, v in sd.items()}
    x_t, qm, qs = O.forward_noising(x0, t, z, tables[2])
    out = O.unet_forward(params, x_t, t, groups=groups, flavour=flavour, dropout_masks=masks)
    if flavour == "ddpm":
        loss = O.ddpm_loss(x_t, qm, qs, out)
    else:
        eps, var = O.iddpm_split(out, t, tables)
        vlb = O.vlb_loss(eps, var, x_t, t, x0, tables)
        loss = vlb if loss_type == "vlb" else O.ddpm_loss(x_t, qm, qs, eps) + gamma * vlb
    loss.backward()
    return float(loss.detach()), {k: p.grad for k, p in params.items() if p.requires_grad}


def _model(flavour, precision, dropout, seed=0, **kw):
    from dmme_b200 import DDPM, IDDPM
    from dmme_b200.models import ddpm, iddpm
    torch.manual_seed(seed)
    m = (ddpm.UNet if flavour == "ddpm" else iddpm.UNet)(precision=precision, dropout=dropout, **kw)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    return m, sd


def _check_grads(model, want, tol, worst_tol=None):
    total_num = total_den = 0.0
    worst = ("", 0.0)
    for k, p in model.named_parameters():
        assert p.grad is not None, f"no gradient for {k}"
        w = want[k]
        num, den = float((p.grad.cpu().double() - w.double()).norm() ** 2), float(w.double().norm() ** 2)
        total_num += num
        total_den += den
        r = math.sqrt(num / max(den, 1e-30))
        if r > worst[1]:
            worst = (k, r)
    overall = math.sqrt(total_num / total_den)
    assert overall < tol, (overall, worst)
    if worst_tol is not None:
        assert worst[1] < worst_tol, worst
    return overall, worst


@pytest.mark.parametrize("flavour", ["ddpm", "iddpm"])
def test_tiny_training_step_fp32(flavour):
    """the reference's own training test (tests/test_ddpm.py:7-23, tests/test_iddpm.py:17-34), with numbers."""
    from dmme_b200 import DDPM, IDDPM
    m, sd = _model(flavour, "fp32", 0.0, **TINY)
    T = 100
    dm = (DDPM(m, timesteps=T) if flavour == "ddpm" else IDDPM(m, timesteps=T)).to(DEV)
    g = torch.Generator().manual_seed(5)
    x0 = torch.rand(4, 3, 32, 32, generator=g) * 2 - 1
    z = torch.randn(4, 3, 32, 32, generator=g)
    t = torch.tensor([1, 7, 50, 99])
    tables = O.linear_tables(T) if flavour == "ddpm" else O.cosine_tables(T)
    want_loss, want = _oracle_training(flavour, sd, x0, t, z, tables, 2)
    loss = dm.training_step(x0.to(DEV), t=t.to(DEV), noise=z.to(DEV))
    assert loss.dim() == 0 and not torch.isnan(loss)
    loss.backward()
    assert abs(float(loss) - want_loss) < 1e-4 * abs(want_loss), (float(loss), want_loss)
    _check_grads(m, want, 1e-3, 2e-2)


def test_tiny_training_step_dropout_masks_fp32():
    from dmme_b200 import DDPM
    m, sd = _model("ddpm", "fp32", 0.1, **TINY)
    dm = DDPM(m, timesteps=100).to(DEV)
    g = torch.Generator().manual_seed(6)
    x0 = torch.rand(3, 3, 32, 32, generator=g) * 2 - 1
    z = torch.randn(3, 3, 32, 32, generator=g)
    t = torch.tensor([3, 40, 80])
    masks = {}
    for name, blk in m.engine.resblocks():
        c = blk.conv2[-1].weight.shape[1]
        masks[name] = (torch.rand(3, c, generator=g) > 0.1).float() / 0.9
    want_loss, want = _oracle_training("ddpm", sd, x0, t, z, O.linear_tables(100), 2, masks=masks)
    m._injected_masks = {k: v.to(DEV) for k, v in masks.items()}
    loss = dm.training_step(x0.to(DEV), t=t.to(DEV), noise=z.to(DEV))
    loss.backward()
    assert abs(float(loss) - want_loss) < 1e-4 * abs(want_loss)
    _check_grads(m, want, 1e-3, 2e-2)


@pytest.mark.parametrize("flavour,precision", [("ddpm", "fp32"), ("ddpm", "bf16"), ("iddpm", "bf16")])
def test_mid_training_step(flavour, precision):
    """64/128-channel UNet: tensor-core forward + dgrad kernels in bf16 mode, attention at 16x16 and 8x8."""
    from dmme_b200 import DDPM, IDDPM
    kw = dict(pos_dim=32, emb_dim=64, channels_per_depth=(64, 128), num_blocks=1, attention_depths=(2,))
    m, sd = _model(flavour, precision, 0.0, **kw)
    T = 1000
    dm = (DDPM(m, timesteps=T) if flavour == "ddpm" else IDDPM(m, timesteps=T)).to(DEV)
    g = torch.Generator().manual_seed(7)
    x0 = torch.rand(4, 3, 32, 32, generator=g) * 2 - 1
    z = torch.randn(4, 3, 32, 32, generator=g)
    t = torch.tensor([1, 200, 500, 999])
    tables = O.linear_tables(T) if flavour == "ddpm" else O.cosine_tables(T)
    want_loss, want = _oracle_training(flavour, sd, x0, t, z, tables, 32)
    loss = dm.training_step(x0.to(DEV), t=t.to(DEV), noise=z.to(DEV))
    loss.backward()
    if precision == "fp32":
        assert abs(float(loss) - want_loss) < 1e-4 * abs(want_loss)
        _check_grads(m, want, 1e-3)
    else:
        assert abs(float(loss) - want_loss) < 2e-2 * abs(want_loss), (float(loss), want_loss)
        _check_grads(m, want, 5e-2)


def test_default_iddpm_training_step_bf16():
    """BASELINE config #4 at full size: the default 36.2 M-parameter IDDPM UNet (256 / 512-channel weight gradients, 11
    multi-head attention sites with their kept softmax matrices, cosine schedule, hybrid loss), batch 8, dropout 0,
    bf16 tensor-core path -- loss and every parameter gradient against autograd through the fp32 CPU oracle."""
    from dmme_b200 import IDDPM
    m, sd = _model("iddpm", "bf16", 0.0)
    T = 1000
    dm = IDDPM(m, timesteps=T).to(DEV)
    g = torch.Generator().manual_seed(3)
    x0 = torch.rand(8, 3, 32, 32, generator=g) * 2 - 1
    z = torch.randn(8, 3, 32, 32, generator=g)
    t = torch.tensor([2, 17, 250, 500, 640, 800, 999, 3])
    want_loss, want = _oracle_training("iddpm", sd, x0, t, z, O.cosine_tables(T), 32)
    loss = dm.training_step(x0.to(DEV), t=t.to(DEV), noise=z.to(DEV))
    loss.backward()
    print(f"default iddpm training step: loss {float(loss):.6f} (oracle {want_loss:.6f})")
    assert abs(float(loss) - want_loss) < 2e-2 * abs(want_loss), (float(loss), want_loss)
    _check_grads(m, want, 5e-2)


def test_training_step_draws_like_the_reference_and_optimizer_step():
    """training_step(x_0) without injection: t from randint(1, T), loss finite, Adam step changes the weights and the
    packed bf16 weights are rebuilt (the next forward differs)."""
    from dmme_b200 import DDPM
    m, _ = _model("ddpm", "bf16", 0.1, pos_dim=32, emb_dim=64, channels_per_depth=(64, 128), num_blocks=1)
    dm = DDPM(m, timesteps=1000).to(DEV).train()
    opt = torch.optim.Adam(dm.parameters(), lr=1e-3)
    x0 = (torch.rand(8, 3, 32, 32) * 2 - 1).to(DEV)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        loss = dm.training_step(x0)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(math.isfinite(v) for v in losses)
    assert losses[-1] < losses[0]


@pytest.mark.parametrize("flavour", ["ddpm", "iddpm"])
def test_input_gradient_and_broadcast_timestep_fp32(flavour):
    """d out / d x (what classifier-guidance style callers differentiate through, guidance/classifier.py:9-23) and a
    (1,)-shaped timestep in a training-mode forward (one conditioning row broadcast over the batch)."""
    from dmme_b200.models import ddpm, iddpm
    torch.manual_seed(0)
    m = (ddpm.UNet if flavour == "ddpm" else iddpm.UNet)(precision="fp32", dropout=0.0, **TINY)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.to(DEV)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 3, 32, 32, generator=g)
    t = torch.tensor([42])
    w = torch.randn(3, 3 if flavour == "ddpm" else 6, 32, 32, generator=g)
    params = {k: v.clone().requires_grad_(v.is_floating_point() and "embeddings" not in k) for k, v in sd.items()}
    xr = x.clone().requires_grad_()
    want = (O.unet_forward(params, xr, t, groups=2, flavour=flavour) * w).sum()
    want.backward()
    xd = x.to(DEV).requires_grad_()
    got = (m(xd, t.to(DEV)) * w.to(DEV)).sum()
    got.backward()
    assert abs(float(got) - float(want)) < 1e-3 * abs(float(want))
    assert rel_l2(xd.grad.cpu(), xr.grad) < 1e-3
    _check_grads(m, {k: p.grad for k, p in params.items() if p.requires_grad}, 1e-3, 2e-2)
