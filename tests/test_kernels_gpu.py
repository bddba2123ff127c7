"""Kernel-level parity: every C-ABI entry point against a plain torch fp32 computation of the same op
on the CPU (the oracle's building blocks).  Tolerances: fp32 kernels 1e-5 rel-L2 (accumulation order),
bf16-operand kernels 4e-3 rel-L2 against an fp32 computation on the SAME bf16-rounded operands
(only the fp32 accumulation order and the bf16 output rounding differ), sampler updates bit-exact."""
import math

import pytest
import torch
import torch.nn.functional as F

from helpers import bf16_round, rel_l2, to_nchw, to_nhwc

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from dmme_b200 import ops, _lib
    return ops, _lib


def run_conv(x, w, b, *, x1=None, stride=1, upsample=False, res=None, wres=None, bres=None, temb=None,
             addend=None, dtype=torch.float32, kernel=None, in_nchw=False, out_layout=None):
    """x, x1, res: NCHW fp32 CPU tensors.  Returns NCHW fp32 CPU output of the CUDA conv."""
    ops, L = _ops()
    kernel = L.CONV_AUTO if kernel is None else kernel
    out_layout = L.OUT_NHWC if out_layout is None else out_layout
    if in_nchw:
        s0 = x.to(DEV).contiguous()
        s1 = None
    else:
        s0 = to_nhwc(x, dtype).to(DEV)
        s1 = to_nhwc(x1, dtype).to(DEV) if x1 is not None else None
    r0 = to_nhwc(res, dtype).to(DEV) if res is not None else None
    cout = w.shape[0]
    d = ops.make_conv_desc(s0, s1, cout, w.shape[2], stride, upsample, r0, None, in_nchw, out_layout, dtype, kernel)
    tc = ops.conv_uses_tc(d)
    if kernel == L.CONV_TC:
        assert tc, "shape must be eligible for the tcgen05 path"
    wp = ops.pack_conv_weight(w.to(DEV), wres.to(DEV) if wres is not None else None, tc)
    bias = b if bres is None else b + bres
    ho, wo = ops.conv_out_hw(d)
    n = x.shape[0]
    ad = to_nhwc(addend, dtype).to(DEV) if addend is not None else None
    tb = temb.to(DEV).contiguous() if temb is not None else None
    if out_layout == L.OUT_NCHW_F32:
        out = torch.empty((n, cout, ho, wo), dtype=torch.float32, device=DEV)
        ops.conv2d_launch(d, wp, bias.to(DEV), out, tb, ad)
        torch.cuda.synchronize()
        return out.cpu(), tc
    if out_layout == L.OUT_QKV:
        c = cout // 3
        q = torch.empty((n, ho * wo, c), dtype=dtype, device=DEV)
        k = torch.empty_like(q)
        vt = torch.empty((n, c, ho * wo), dtype=dtype, device=DEV)
        ops.conv2d_launch(d, wp, bias.to(DEV), q, tb, ad, k, vt)
        torch.cuda.synchronize()
        return (q.float().cpu(), k.float().cpu(), vt.float().cpu()), tc
    out = torch.empty((n, ho, wo, cout), dtype=dtype, device=DEV)
    ops.conv2d_launch(d, wp, bias.to(DEV), out, tb, ad)
    torch.cuda.synchronize()
    return to_nchw(out.cpu()), tc


def ref_conv(x, w, b, *, x1=None, stride=1, upsample=False, res=None, wres=None, bres=None, temb=None, addend=None):
    xin = x if x1 is None else torch.cat([x, x1], dim=1)
    if upsample:
        xin = F.interpolate(xin, scale_factor=2.0, mode="nearest")
    y = F.conv2d(xin, w, b, stride=stride, padding=w.shape[2] // 2)
    if res is not None:
        y = y + F.conv2d(res, wres, bres)
    if temb is not None:
        rows = temb if temb.shape[0] == x.shape[0] else temb.expand(x.shape[0], -1)
        y = y + rows[:, :, None, None]
    if addend is not None:
        y = y + addend
    return y


# ---------------------------------------------------------------------------------------------
# generic (FFMA) convolution: arbitrary shapes, fp32 storage
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [
    dict(n=2, cin=3, cout=4, h=32, w=32, k=3),
    dict(n=3, cin=8, cout=16, h=16, w=16, k=3, stride=2),
    dict(n=2, cin=16, cout=8, h=8, w=8, k=3, upsample=True),
    dict(n=2, cin=5, cout=7, h=6, w=10, k=1),
    dict(n=1, cin=32, cout=32, h=4, w=4, k=3, cin1=32),
    dict(n=2, cin=16, cout=24, h=8, w=8, k=3, res=40, temb="rows", addend=False),
    dict(n=2, cin=16, cout=16, h=8, w=8, k=3, temb="bcast", addend=True),
    dict(n=1, cin=70, cout=130, h=5, w=7, k=3),
])
def test_conv_generic_fp32(cfg):
    _, L = _ops()
    g = torch.Generator().manual_seed(11)
    n, cin, cout, h, w, k = (cfg[s] for s in ("n", "cin", "cout", "h", "w", "k"))
    stride, up = cfg.get("stride", 1), cfg.get("upsample", False)
    x = torch.randn(n, cin, h, w, generator=g)
    x1 = torch.randn(n, cfg["cin1"], h, w, generator=g) if cfg.get("cin1") else None
    ctot = cin + (cfg.get("cin1") or 0)
    wt = torch.randn(cout, ctot, k, k, generator=g) / math.sqrt(ctot * k * k)
    b = torch.randn(cout, generator=g)
    ho, wo = (h * (2 if up else 1)) // stride, (w * (2 if up else 1)) // stride
    kw = {}
    if cfg.get("res"):
        kw.update(res=torch.randn(n, cfg["res"], ho, wo, generator=g),
                  wres=torch.randn(cout, cfg["res"], 1, 1, generator=g) / math.sqrt(cfg["res"]),
                  bres=torch.randn(cout, generator=g))
    if cfg.get("temb") == "rows":
        kw["temb"] = torch.randn(n, cout, generator=g)
    elif cfg.get("temb") == "bcast":
        kw["temb"] = torch.randn(1, cout, generator=g)
    if cfg.get("addend"):
        kw["addend"] = torch.randn(n, cout, ho, wo, generator=g)
    got, tc = run_conv(x, wt, b, x1=x1, stride=stride, upsample=up, kernel=L.CONV_GENERIC, **kw)
    assert not tc
    want = ref_conv(x, wt, b, x1=x1, stride=stride, upsample=up, **kw)
    assert got.shape == want.shape
    assert rel_l2(got, want) < 1e-5


def test_conv_generic_nchw_in_out():
    _, L = _ops()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 32, 32, generator=g)
    w = torch.randn(16, 3, 3, 3, generator=g) / 5
    b = torch.randn(16, generator=g)
    got, _ = run_conv(x, w, b, kernel=L.CONV_GENERIC, in_nchw=True)
    assert rel_l2(got, F.conv2d(x, w, b, padding=1)) < 1e-5
    w2 = torch.randn(3, 16, 3, 3, generator=g) / 12
    b2 = torch.randn(3, generator=g)
    y = torch.randn(2, 16, 8, 8, generator=g)
    got2, _ = run_conv(y, w2, b2, kernel=L.CONV_GENERIC, out_layout=L.OUT_NCHW_F32)
    assert rel_l2(got2, F.conv2d(y, w2, b2, padding=1)) < 1e-5


def test_conv_generic_bf16_storage():
    _, L = _ops()
    g = torch.Generator().manual_seed(6)
    x = bf16_round(torch.randn(2, 24, 8, 8, generator=g))
    w = torch.randn(40, 24, 3, 3, generator=g) / 15
    b = torch.randn(40, generator=g)
    got, _ = run_conv(x, w, b, dtype=torch.bfloat16, kernel=L.CONV_GENERIC)
    assert rel_l2(got, F.conv2d(x, w, b, padding=1)) < 4e-3


@pytest.mark.parametrize("n,h,cout", [(3, 32, 128), (2, 16, 64), (5, 8, 256)])
def test_conv_in_image_to_nhwc_bf16(n, h, cout):
    """input_conv fast path: NCHW fp32 image (3 channels) -> NHWC bf16"""
    _, L = _ops()
    g = torch.Generator().manual_seed(31)
    x = torch.randn(n, 3, h, h, generator=g)
    w = torch.randn(cout, 3, 3, 3, generator=g) / 5
    b = torch.randn(cout, generator=g)
    got, tc = run_conv(x, w, b, dtype=torch.bfloat16, in_nchw=True)
    assert not tc
    assert rel_l2(got, F.conv2d(x, w, b, padding=1)) < 4e-3


@pytest.mark.parametrize("n,cin", [(1, 3), (37, 3), (300, 3), (2, 6), (131, 6)])
def test_conv_in_tc_32x32(n, cin):
    """input_conv on tcgen05 (operand rows [hi | lo] built from the fp32 image, bf16 weights): against the fp32 conv on
    bf16-rounded weights, and against the FFMA kernel it replaces (fp32 weights) within the weight rounding"""
    ops, L = _ops()
    lib = L.load()
    g = torch.Generator().manual_seed(36)
    # cin = 6: the data gradient of the IDDPM output conv (a 6-channel fp32 NCHW image-space gradient into 128 channels)
    x = torch.randn(n, cin, 32, 32, generator=g) * 1.7
    w = torch.randn(128, cin, 3, 3, generator=g) / 5
    b = torch.randn(128, generator=g)
    lib.dmme_set_conv_in_tc_mode(2)  # the default takes it from 64 images up
    try:
        got, tc = run_conv(x, w, b, dtype=torch.bfloat16, in_nchw=True)
    finally:
        lib.dmme_set_conv_in_tc_mode(1)
    assert not tc  # fp32 [K][cout] weights, like the FFMA kernel
    want = F.conv2d(x, bf16_round(w), b, padding=1)
    assert rel_l2(got, want) < 2.5e-3  # bf16 rounding of the stored output
    assert (got - bf16_round(want)).abs().max() <= 2 * bf16_round(want).abs().max() * 2 ** -8
    lib.dmme_set_conv_in_tc_mode(0)
    try:
        ffma, _ = run_conv(x, w, b, dtype=torch.bfloat16, in_nchw=True)
    finally:
        lib.dmme_set_conv_in_tc_mode(1)
    assert rel_l2(got, ffma) < 4e-3


@pytest.mark.parametrize("n,tc_mode", [(3, 2), (3, 0), (200, 1)])
def test_conv_in_writes_groupnorm_stats(n, tc_mode):
    ops, L = _ops()
    L.load().dmme_set_conv_in_tc_mode(tc_mode)
    try:
        _conv_in_stats_case(ops, L, n)
    finally:
        L.load().dmme_set_conv_in_tc_mode(1)


def _conv_in_stats_case(ops, L, n):
    g = torch.Generator().manual_seed(33)
    h, cout = 32, 128
    x = torch.randn(n, 3, h, h, generator=g)
    w = torch.randn(cout, 3, 3, 3, generator=g) / 5
    b = torch.randn(cout, generator=g)
    d = ops.make_conv_desc(x.to(DEV), None, cout, 3, 1, False, None, None, True, L.OUT_NHWC, torch.bfloat16, L.CONV_AUTO)
    assert ops.conv_writes_stats(d) and not ops.conv_uses_tc(d)
    out = torch.empty((n, h, h, cout), dtype=torch.bfloat16, device=DEV)
    st = torch.zeros(n * (cout // 4) * 2, dtype=torch.int64, device=DEV)
    ops.conv2d_launch(d, ops.pack_conv_weight(w.to(DEV), None, False), b.to(DEV), out, stats=st)
    torch.cuda.synchronize()
    y = to_nchw(out.cpu()).double()
    sums = st.cpu().view(n, cout // 4, 2).double() / 2 ** 20
    assert torch.allclose(sums[..., 0], y.reshape(n, cout // 4, -1).sum(-1), rtol=1e-5, atol=1e-2)
    assert torch.allclose(sums[..., 1], (y ** 2).reshape(n, cout // 4, -1).sum(-1), rtol=1e-5, atol=1e-2)


@pytest.mark.parametrize("n,h,cin,cout", [(3, 32, 128, 3), (2, 16, 128, 6), (5, 8, 64, 3), (1, 32, 256, 6)])
def test_conv_out_nhwc_bf16_to_image(n, h, cin, cout):
    """output_conv FFMA path (fp32 weights): NHWC bf16 -> eps NCHW fp32 with 3 (DDPM) or 6 (IDDPM) channels"""
    _, L = _ops()
    lib = L.load()
    lib.dmme_set_conv_out_tc_mode(0)
    try:
        g = torch.Generator().manual_seed(32)
        x = bf16_round(torch.randn(n, cin, h, h, generator=g))
        w = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)
        b = torch.randn(cout, generator=g)
        got, tc = run_conv(x, w, b, dtype=torch.bfloat16, out_layout=L.OUT_NCHW_F32)
        assert not tc
        assert rel_l2(got, F.conv2d(x, w, b, padding=1)) < 1e-5
    finally:
        lib.dmme_set_conv_out_tc_mode(1)


@pytest.mark.parametrize("n,h,cin,cout", [(3, 32, 128, 3), (2, 16, 128, 6), (5, 8, 64, 3), (1, 32, 256, 6), (1, 8, 128, 1),
                                          (260, 32, 128, 3), (131, 16, 192, 6), (77, 32, 128, 8)])
def test_conv_out_tc(n, h, cin, cout):
    """output_conv on tcgen05 (positions on the M side, bf16 weights): against fp32 conv on the same bf16 operands"""
    _, L = _ops()
    g = torch.Generator().manual_seed(33)
    x = bf16_round(torch.randn(n, cin, h, h, generator=g))
    w = bf16_round(torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin))
    b = torch.randn(cout, generator=g)
    got, tc = run_conv(x, w, b, dtype=torch.bfloat16, out_layout=L.OUT_NCHW_F32)
    assert tc
    assert rel_l2(got, F.conv2d(x, w, b, padding=1)) < 1e-5


@pytest.mark.parametrize("n,cin,cout", [(3, 128, 3), (150, 128, 6), (1, 256, 8)])
def test_conv_out_tc_per_tap_kernel_32x32(n, cin, cout):
    """the per-tap tcgen05 output conv stays covered on 32x32 maps (mode 2), where the row-tile kernel is the default"""
    _, L = _ops()
    lib = L.load()
    lib.dmme_set_conv_out_tc_mode(2)
    try:
        g = torch.Generator().manual_seed(34)
        x = bf16_round(torch.randn(n, cin, 32, 32, generator=g))
        w = bf16_round(torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin))
        b = torch.randn(cout, generator=g)
        got, tc = run_conv(x, w, b, dtype=torch.bfloat16, out_layout=L.OUT_NCHW_F32)
        assert tc
        assert rel_l2(got, F.conv2d(x, w, b, padding=1)) < 1e-5
    finally:
        lib.dmme_set_conv_out_tc_mode(1)


@pytest.mark.parametrize("n,cin,cout,silu", [(3, 128, 3, True), (149, 128, 3, True), (300, 128, 6, True), (5, 256, 6, False),
                                             (2, 128, 1, True)])
def test_conv_out_fused_groupnorm(n, cin, cout, silu):
    """models/ddpm.py:277 (GroupNorm -> SiLU -> Conv) in ONE launch on 32x32 maps: the norm applied to the operand tile
    inside the output conv == gn_apply followed by the same conv, bit for bit; both against the fp32 reference"""
    ops, L = _ops()
    g = torch.Generator().manual_seed(35)
    h = 32
    # producer: a conv that writes the tensor and its statistics
    x = bf16_round(torch.randn(n, 64, h, h, generator=g))
    w = bf16_round(torch.randn(cin, 64, 3, 3, generator=g) / 8)
    b = torch.randn(cin, generator=g) * 2
    d = ops.make_conv_desc(to_nhwc(x, torch.bfloat16).to(DEV), None, cin, 3, 1, False, None, None, False, L.OUT_NHWC,
                           torch.bfloat16, L.CONV_AUTO)
    src = torch.empty((n, h, h, cin), dtype=torch.bfloat16, device=DEV)
    st = torch.zeros(n * (cin // 4) * 2, dtype=torch.int64, device=DEV)
    ops.conv2d_launch(d, ops.pack_conv_weight(w.to(DEV), None, True), b.to(DEV), src, stats=st)
    gamma, beta = torch.randn(cin, generator=g).to(DEV), torch.randn(cin, generator=g).to(DEV)
    wt = bf16_round(torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin))
    bias = torch.randn(cout, generator=g)
    wp = ops.pack_conv_weight(wt.to(DEV), None, True)

    a = ops.groupnorm(src, None, 32, gamma, beta, silu, stats0=st)
    d = ops.make_conv_desc(a, None, cout, 3, 1, False, None, None, False, L.OUT_NCHW_F32, torch.bfloat16, L.CONV_AUTO)
    want = torch.empty((n, cout, h, h), dtype=torch.float32, device=DEV)
    ops.conv2d_launch(d, wp, bias.to(DEV), want)

    d = ops.make_conv_desc(src, None, cout, 3, 1, False, None, None, False, L.OUT_NCHW_F32, torch.bfloat16, L.CONV_AUTO)
    assert ops.conv_fuses_gn(d)
    ab = ops.groupnorm_coeff(st, None, cin, 0, n, h * h, 32, gamma, beta)
    got = torch.full((n, cout, h, h), float("nan"), dtype=torch.float32, device=DEV)
    ops.conv2d_launch(d, wp, bias.to(DEV), got, gn_ab=ab, gn_silu=silu)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    ref = F.conv2d(to_nchw(a.cpu()).float(), wt, bias, padding=1)
    assert rel_l2(got.cpu(), ref) < 1e-5
    # smaller maps keep the norm as its own pass
    d16 = ops.make_conv_desc(src[:, :16, :16].contiguous(), None, cout, 3, 1, False, None, None, False, L.OUT_NCHW_F32,
                             torch.bfloat16, L.CONV_AUTO)
    assert not ops.conv_fuses_gn(d16)


# ---------------------------------------------------------------------------------------------
# tcgen05 convolution
# ---------------------------------------------------------------------------------------------
TC_CASES = [
    dict(n=2, cin=64, cout=64, h=32, w=32, k=3),                      # BN=64, smallest case
    dict(n=4, cin=128, cout=128, h=32, w=32, k=3, temb="bcast"),       # the dominant shape
    dict(n=2, cin=128, cout=256, h=16, w=16, k=3, temb="rows"),
    dict(n=3, cin=256, cout=256, h=16, w=16, k=3, addend=True),
    dict(n=5, cin=256, cout=256, h=8, w=8, k=3),                       # 2 images per tile, ragged batch
    dict(n=11, cin=256, cout=256, h=4, w=4, k=3),                      # 8 images per tile, ragged batch
    dict(n=2, cin=128, cout=128, h=32, w=32, k=3, stride=2),           # DownSample
    dict(n=3, cin=256, cout=256, h=8, w=8, k=3, stride=2),
    dict(n=2, cin=256, cout=128, h=32, w=32, k=3, cin1=128, res=True),  # concat + fused 1x1 residual
    dict(n=2, cin=512, cout=256, h=16, w=16, k=3, cin1=256, res=True),
    dict(n=2, cin=256, cout=256, h=16, w=16, k=1),                     # attention proj
    dict(n=2, cin=128, cout=128, h=16, w=16, k=1, addend=True),
    dict(n=40, cin=256, cout=256, h=16, w=16, k=3),                    # enough tiles to select BN=256
]


@pytest.mark.parametrize("cfg", TC_CASES)
def test_conv_tc(cfg):
    _, L = _ops()
    g = torch.Generator().manual_seed(21)
    n, cin, cout, h, w, k = (cfg[s] for s in ("n", "cin", "cout", "h", "w", "k"))
    stride = cfg.get("stride", 1)
    c1 = cfg.get("cin1", 0)
    c0 = cin - c1
    xa = bf16_round(torch.randn(n, cin, h, w, generator=g))
    x, x1 = (xa[:, :c0].contiguous(), xa[:, c0:].contiguous()) if c1 else (xa, None)
    wt = bf16_round(torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k))
    b = torch.randn(cout, generator=g)
    ho, wo = h // stride, w // stride
    kw = {}
    if cfg.get("res"):
        # the fused residual conv reads the same concat input
        kw.update(res=xa, wres=bf16_round(torch.randn(cout, cin, 1, 1, generator=g) / math.sqrt(cin)),
                  bres=torch.randn(cout, generator=g))
    if cfg.get("temb") == "rows":
        kw["temb"] = torch.randn(n, cout, generator=g)
    elif cfg.get("temb") == "bcast":
        kw["temb"] = torch.randn(1, cout, generator=g)
    if cfg.get("addend"):
        kw["addend"] = bf16_round(torch.randn(n, cout, ho, wo, generator=g))
    got, tc = run_conv(x, wt, b, x1=x1, stride=stride, dtype=torch.bfloat16, kernel=L.CONV_TC, **kw)
    assert tc
    want = ref_conv(x, wt, b, x1=x1, stride=stride, **kw)
    err = rel_l2(got, want)
    assert err < 4e-3, f"rel-L2 {err}"


@pytest.fixture
def tct_everywhere():
    """force the transposed tcgen05 kernel (thread = channel epilogue) wherever it is supported"""
    _, L = _ops()
    lib = L.load()
    old = lib.dmme_get_conv_tct_mode()
    lib.dmme_set_conv_tct_mode(2)
    yield
    lib.dmme_set_conv_tct_mode(old)


TCT_CASES = [c for c in TC_CASES if c["cout"] % 128 == 0] + [
    dict(n=70, cin=256, cout=256, h=16, w=16, k=1, addend=True),       # weight-stationary 1x1, 256-pixel tiles, 2 n tiles
    dict(n=33, cin=512, cout=128, h=16, w=16, k=1, temb="rows"),       # K = 512 slab
    dict(n=130, cin=256, cout=256, h=8, w=8, k=3, temb="rows"),        # 4 images per 256-pixel tile
    dict(n=200, cin=128, cout=256, h=4, w=4, k=3, addend=True, temb="rows"),  # 2 images per 32-pixel chunk
    dict(n=66, cin=128, cout=128, h=32, w=32, k=3, stride=2, temb="bcast"),
    # LSUN-256 shapes (configs/ddpm/lsun_bedroom.yaml:78-90, batch 2): 256 / 128 / 64 pixel wide maps
    dict(n=2, cin=128, cout=128, h=256, w=256, k=3, temb="bcast"),
    dict(n=2, cin=256, cout=128, h=256, w=256, k=3, cin1=128, res=True),
    dict(n=2, cin=128, cout=128, h=256, w=256, k=3, stride=2),
    dict(n=2, cin=128, cout=128, h=128, w=128, k=3, addend=True),
    dict(n=2, cin=128, cout=256, h=64, w=64, k=3, temb="rows"),
    dict(n=1, cin=128, cout=128, h=256, w=256, k=1, addend=True),
]


@pytest.mark.parametrize("cfg", TCT_CASES)
def test_conv_tct(cfg, tct_everywhere):
    """transposed kernel against fp32 conv on the same bf16 operands, plus its GroupNorm statistics"""
    ops, L = _ops()
    g = torch.Generator().manual_seed(29)
    n, cin, cout, h, w, k = (cfg[s] for s in ("n", "cin", "cout", "h", "w", "k"))
    stride = cfg.get("stride", 1)
    c1 = cfg.get("cin1", 0)
    c0 = cin - c1
    xa = bf16_round(torch.randn(n, cin, h, w, generator=g))
    wt = bf16_round(torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k))
    b = torch.randn(cout, generator=g)
    ho, wo = h // stride, w // stride
    s0 = to_nhwc(xa[:, :c0], torch.bfloat16).to(DEV)
    s1 = to_nhwc(xa[:, c0:], torch.bfloat16).to(DEV) if c1 else None
    want = F.conv2d(xa, wt, b, stride=stride, padding=k // 2)
    wres, r0, r1, bias = None, None, None, b
    if cfg.get("res"):
        wres = bf16_round(torch.randn(cout, cin, 1, 1, generator=g) / math.sqrt(cin))
        bres = torch.randn(cout, generator=g)
        want = want + F.conv2d(xa, wres, bres)
        bias = b + bres
        r0, r1 = s0, s1
    temb = None
    if cfg.get("temb"):
        temb = torch.randn(n if cfg["temb"] == "rows" else 1, cout, generator=g)
        want = want + (temb if temb.shape[0] == n else temb.expand(n, -1))[:, :, None, None]
        temb = temb.to(DEV)
    addend = None
    if cfg.get("addend"):
        ad = bf16_round(torch.randn(n, cout, ho, wo, generator=g))
        want = want + ad
        addend = to_nhwc(ad, torch.bfloat16).to(DEV)
    d = ops.make_conv_desc(s0, s1, cout, k, stride, False, r0, r1, False, L.OUT_NHWC, torch.bfloat16, L.CONV_TC)
    assert ops.conv_uses_tc(d)
    wp = ops.pack_conv_weight(wt.to(DEV), wres.to(DEV) if wres is not None else None, True)
    out = torch.full((n, ho, wo, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
    st = torch.zeros(n * (cout // 4) * 2, dtype=torch.int64, device=DEV)
    ops.conv2d_launch(d, wp, bias.to(DEV), out, temb, addend, stats=st)
    torch.cuda.synchronize()
    got = to_nchw(out.cpu())
    err = rel_l2(got, want)
    assert err < 4e-3, f"rel-L2 {err}"
    sums = st.cpu().view(n, cout // 4, 2).double() / 2 ** 20
    assert torch.allclose(sums[..., 0], got.double().reshape(n, cout // 4, -1).sum(-1), rtol=1e-5, atol=2e-2)
    assert torch.allclose(sums[..., 1], (got.double() ** 2).reshape(n, cout // 4, -1).sum(-1), rtol=1e-5, atol=2e-2)


@pytest.fixture
def pair_everywhere():
    """force the cta_group::2 (two-SM MMA) variant of the transposed kernel wherever it is supported"""
    _, L = _ops()
    lib = L.load()
    lib.dmme_set_conv_pair_mode(2)
    yield
    lib.dmme_set_conv_pair_mode(0)


@pytest.mark.parametrize("cfg", [c for c in TCT_CASES if c["cout"] % 256 == 0] + [
    dict(n=1, cin=64, cout=256, h=16, w=16, k=3),                                # a single pair
    dict(n=75, cin=256, cout=512, h=8, w=8, k=3, temb="bcast", addend=True),     # two channel pairs, ragged last tile
    dict(n=300, cin=256, cout=256, h=16, w=16, k=3, temb="rows"),                # several units per pair
])
def test_conv_tct_pair(cfg, pair_everywhere):
    if not _ops()[1].load().dmme_has_experimental():
        pytest.skip("cta_group::2 variant is compiled only with -DDMME_EXPERIMENTAL (measured no faster, DESIGN.md)")
    test_conv_tct.__wrapped__(cfg, None) if hasattr(test_conv_tct, "__wrapped__") else test_conv_tct(cfg, None)


@pytest.mark.parametrize("n,c,h", [(3, 128, 16), (50, 256, 16), (7, 256, 4), (300, 128, 4)])
def test_conv_tct_qkv_layout(n, c, h, tct_everywhere):
    _, L = _ops()
    g = torch.Generator().manual_seed(9)
    x = bf16_round(torch.randn(n, c, h, h, generator=g))
    w = bf16_round(torch.randn(3 * c, c, 1, 1, generator=g) / math.sqrt(c))
    b = torch.randn(3 * c, generator=g)
    (q, k, vt), tc = run_conv(x, w, b, dtype=torch.bfloat16, kernel=L.CONV_TC, out_layout=L.OUT_QKV)
    assert tc
    y = F.conv2d(x, w, b).flatten(2)  # n, 3c, L
    assert rel_l2(q, y[:, :c].transpose(1, 2)) < 4e-3
    assert rel_l2(k, y[:, c:2 * c].transpose(1, 2)) < 4e-3
    assert rel_l2(vt, y[:, 2 * c:]) < 4e-3


HALO_CASES = [
    dict(n=1, cin=64, cout=128, h=16, w=16),
    dict(n=4, cin=128, cout=128, h=32, w=32, temb="bcast"),
    dict(n=3, cin=256, cout=256, h=16, w=16, addend=True, temb="rows"),
    dict(n=2, cin=256, cout=128, h=32, w=32, cin1=128, res=True),
    dict(n=37, cin=512, cout=256, h=16, w=16, cin1=256, res=True),   # > 148 work units: persistent loop, both TMEM stages
    dict(n=150, cin=128, cout=128, h=32, w=32),                      # several units per CTA at 32x32
    dict(n=5, cin=256, cout=256, h=8, w=8, temb="rows"),             # 8x8: tiles span several images
    dict(n=70, cin=256, cout=256, h=8, w=8, addend=True),
    dict(n=3, cin=256, cout=256, h=4, w=4, temb="bcast"),            # 4x4: 42 padded rows = 7 images per tile
    dict(n=256, cin=512, cout=256, h=4, w=4, cin1=256, res=True),
    dict(n=64, cin=128, cout=128, h=8, w=8),
]


def _halo_params():
    out = [pytest.param(c, "halo", id=f"halo-{i}") for i, c in enumerate(HALO_CASES)]
    # the tilings bench.py times (rows per tile are chosen by wave fill FROM THE BATCH): per-GPU batches 256 / 128 / 64 / 32
    # of BASELINE config #2 (256 images split over 1 / 2 / 4 / 8 GPUs) at the 16x16 and 32x32 levels
    for nb in (256, 128, 64, 32):
        out.append(pytest.param(dict(n=nb, cin=256, cout=256, h=16, w=16, temb="bcast"), "halo", id=f"halo-timed-16x16-c256-n{nb}"))
        out.append(pytest.param(dict(n=nb, cin=256, cout=128, h=16, w=16, addend=False), "halo", id=f"halo-timed-16x16-c128-n{nb}"))
        out.append(pytest.param(dict(n=nb, cin=128, cout=128, h=32, w=32, temb="bcast", addend=True), "halo", id=f"halo-timed-32x32-n{nb}"))
    return out


@pytest.mark.parametrize("cfg,which", _halo_params())
def test_conv_halo(cfg, which):
    """halo-reuse persistent tcgen05 kernel against fp32 conv on the same
    bf16 operands, plus their GroupNorm stats"""
    ops, L = _ops()
    g = torch.Generator().manual_seed(23)
    n, cin, cout, h, w = (cfg[s] for s in ("n", "cin", "cout", "h", "w"))
    c1 = cfg.get("cin1", 0)
    c0 = cin - c1
    xa = bf16_round(torch.randn(n, cin, h, w, generator=g))
    wt = bf16_round(torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9))
    b = torch.randn(cout, generator=g)
    s0 = to_nhwc(xa[:, :c0], torch.bfloat16).to(DEV)
    s1 = to_nhwc(xa[:, c0:], torch.bfloat16).to(DEV) if c1 else None
    want = F.conv2d(xa, wt, b, padding=1)
    wres = None
    r0 = r1 = None
    bias = b
    if cfg.get("res"):
        wres = bf16_round(torch.randn(cout, cin, 1, 1, generator=g) / math.sqrt(cin))
        bres = torch.randn(cout, generator=g)
        want = want + F.conv2d(xa, wres, bres)
        bias = b + bres
        r0, r1 = s0, s1
    temb = None
    if cfg.get("temb"):
        temb = torch.randn(n if cfg["temb"] == "rows" else 1, cout, generator=g)
        want = want + (temb if temb.shape[0] == n else temb.expand(n, -1))[:, :, None, None]
        temb = temb.to(DEV)
    addend = None
    if cfg.get("addend"):
        ad = bf16_round(torch.randn(n, cout, h, w, generator=g))
        want = want + ad
        addend = to_nhwc(ad, torch.bfloat16).to(DEV)
    d = ops.make_conv_desc(s0, s1, cout, 3, 1, False, r0, r1, False, L.OUT_NHWC, torch.bfloat16,
                           L.CONV_HALO)
    assert ops.conv_uses_tc(d)
    wp = ops.pack_conv_weight(wt.to(DEV), wres.to(DEV) if wres is not None else None, True)
    out = torch.empty((n, h, w, cout), dtype=torch.bfloat16, device=DEV)
    st = torch.zeros(n * (cout // 4) * 2, dtype=torch.int64, device=DEV)
    ops.conv2d_launch(d, wp, bias.to(DEV), out, temb, addend, stats=st)
    torch.cuda.synchronize()
    got = to_nchw(out.cpu())
    err = rel_l2(got, want)
    assert err < 4e-3, f"rel-L2 {err}"
    sums = st.cpu().view(n, cout // 4, 2).double() / 2 ** 20
    assert torch.allclose(sums[..., 0], got.double().reshape(n, cout // 4, -1).sum(-1), rtol=1e-5, atol=2e-2)
    assert torch.allclose(sums[..., 1], (got.double() ** 2).reshape(n, cout // 4, -1).sum(-1), rtol=1e-5, atol=2e-2)


def test_conv_tc_res_two_sources():
    """fused residual conv whose operand is itself a two-source concat (the up-path ResBlocks)."""
    ops, L = _ops()
    g = torch.Generator().manual_seed(3)
    n, c0, c1, cout, h = 2, 128, 128, 128, 16
    a2 = bf16_round(torch.randn(n, cout, h, h, generator=g))
    xa = bf16_round(torch.randn(n, c0 + c1, h, h, generator=g))
    w2 = bf16_round(torch.randn(cout, cout, 3, 3, generator=g) / math.sqrt(9 * cout))
    wr = bf16_round(torch.randn(cout, c0 + c1, 1, 1, generator=g) / math.sqrt(c0 + c1))
    b = torch.randn(cout, generator=g)
    s0 = to_nhwc(a2, torch.bfloat16).to(DEV)
    r0 = to_nhwc(xa[:, :c0], torch.bfloat16).to(DEV)
    r1 = to_nhwc(xa[:, c0:], torch.bfloat16).to(DEV)
    d = ops.make_conv_desc(s0, None, cout, 3, 1, False, r0, r1, False, L.OUT_NHWC, torch.bfloat16, L.CONV_TC)
    wp = ops.pack_conv_weight(w2.to(DEV), wr.to(DEV), True)
    out = torch.empty((n, h, h, cout), dtype=torch.bfloat16, device=DEV)
    ops.conv2d_launch(d, wp, b.to(DEV), out)
    torch.cuda.synchronize()
    want = F.conv2d(a2, w2, b, padding=1) + F.conv2d(xa, wr)
    assert rel_l2(to_nchw(out.cpu()), want) < 4e-3


@pytest.mark.parametrize("kernel", ["tc", "generic"])
@pytest.mark.parametrize("c,h", [(128, 16), (256, 16), (256, 4)])
def test_conv_qkv_layout(kernel, c, h):
    _, L = _ops()
    g = torch.Generator().manual_seed(8)
    n = 3
    x = bf16_round(torch.randn(n, c, h, h, generator=g))
    w = bf16_round(torch.randn(3 * c, c, 1, 1, generator=g) / math.sqrt(c))
    b = torch.randn(3 * c, generator=g)
    (q, k, vt), tc = run_conv(x, w, b, dtype=torch.bfloat16, kernel=L.CONV_TC if kernel == "tc" else L.CONV_GENERIC,
                              out_layout=L.OUT_QKV)
    assert tc == (kernel == "tc")
    y = F.conv2d(x, w, b).flatten(2)  # n, 3c, L
    assert rel_l2(q, y[:, :c].transpose(1, 2)) < 4e-3
    assert rel_l2(k, y[:, c:2 * c].transpose(1, 2)) < 4e-3
    assert rel_l2(vt, y[:, 2 * c:]) < 4e-3


# ---------------------------------------------------------------------------------------------
# GroupNorm (+SiLU)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [
    dict(n=3, c=128, hw=(32, 32), groups=32, dtype="bf16", silu=True),
    dict(n=2, c=256, hw=(16, 16), groups=32, dtype="bf16", silu=True, c1=128),
    dict(n=2, c=512, hw=(8, 8), groups=32, dtype="bf16", silu=False),
    dict(n=5, c=256, hw=(4, 4), groups=32, dtype="bf16", silu=True, ss=True, mask=True),
    dict(n=2, c=8, hw=(16, 16), groups=2, dtype="fp32", silu=True),
    dict(n=2, c=12, hw=(5, 7), groups=3, dtype="fp32", silu=False, c1=4, ss=True, mask=True),
    dict(n=2, c=64, hw=(64, 64), groups=32, dtype="bf16", silu=True),
])
def test_groupnorm(cfg):
    ops, _ = _ops()
    g = torch.Generator().manual_seed(4)
    n, c, (h, w), groups = cfg["n"], cfg["c"], cfg["hw"], cfg["groups"]
    dtype = torch.bfloat16 if cfg["dtype"] == "bf16" else torch.float32
    x = torch.randn(n, c, h, w, generator=g) * 2 + 0.5
    if dtype == torch.bfloat16:
        x = bf16_round(x)
    gamma, beta = torch.randn(c, generator=g), torch.randn(c, generator=g)
    c1 = cfg.get("c1", 0)
    s0 = to_nhwc(x[:, :c - c1], dtype).to(DEV)
    s1 = to_nhwc(x[:, c - c1:], dtype).to(DEV) if c1 else None
    scale = shift = mask = None
    want = F.group_norm(x, groups, gamma, beta, eps=1e-5)
    if cfg.get("ss"):
        both = torch.randn(n, 2 * c, generator=g)
        shift, scale = both[:, :c], both[:, c:]
        want = want * (scale[:, :, None, None] + 1) + shift[:, :, None, None]
        both = both.to(DEV)
        shift, scale = both[:, :c], both[:, c:]
    if cfg["silu"]:
        want = F.silu(want)
    if cfg.get("mask"):
        mask = (torch.rand(n, c, generator=g) > 0.3).float() / 0.7
        want = want * mask[:, :, None, None]
        mask = mask.to(DEV)
    got = ops.groupnorm(s0, s1, groups, gamma.to(DEV), beta.to(DEV), cfg["silu"], scale, shift, mask)
    torch.cuda.synchronize()
    tol = 4e-3 if dtype == torch.bfloat16 else 1e-5
    assert rel_l2(to_nchw(got.cpu()), want) < tol


@pytest.mark.parametrize("n,c0,c1,h,groups", [(3, 128, 0, 32, 32), (2, 256, 128, 16, 32), (5, 256, 256, 8, 32),
                                                (11, 256, 0, 4, 32), (2, 64, 0, 16, 16)])
def test_conv_stats_feed_streaming_groupnorm(n, c0, c1, h, groups):
    """conv epilogue statistics (int64 fixed point) -> single-pass GroupNorm+SiLU == GN of the stored tensor"""
    ops, L = _ops()
    g = torch.Generator().manual_seed(41)
    outs, stats = [], []
    for c in (c0, c1):
        if c == 0:
            outs.append(None); stats.append(None)
            continue
        x = bf16_round(torch.randn(n, 64, h, h, generator=g))
        w = bf16_round(torch.randn(c, 64, 3, 3, generator=g) / 8)
        b = torch.randn(c, generator=g) * 2
        s0 = to_nhwc(x, torch.bfloat16).to(DEV)
        d = ops.make_conv_desc(s0, None, c, 3, 1, False, None, None, False, L.OUT_NHWC, torch.bfloat16, L.CONV_TC)
        wp = ops.pack_conv_weight(w.to(DEV), None, True)
        out = torch.empty((n, h, h, c), dtype=torch.bfloat16, device=DEV)
        st = torch.zeros(n * (c // 4) * 2, dtype=torch.int64, device=DEV)
        ops.conv2d_launch(d, wp, b.to(DEV), out, stats=st)
        outs.append(out); stats.append(st)
    torch.cuda.synchronize()
    y0 = to_nchw(outs[0].cpu())
    # the statistics describe the stored (bf16) tensor
    sums = stats[0].cpu().view(n, c0 // 4, 2).double() / 2 ** 20
    ref1 = y0.double().reshape(n, c0 // 4, 4 * h * h).sum(-1)
    ref2 = (y0.double() ** 2).reshape(n, c0 // 4, 4 * h * h).sum(-1)
    assert torch.allclose(sums[..., 0], ref1, rtol=1e-5, atol=1e-2)
    assert torch.allclose(sums[..., 1], ref2, rtol=1e-5, atol=1e-2)
    full = y0 if c1 == 0 else torch.cat([y0, to_nchw(outs[1].cpu())], dim=1)
    c = c0 + c1
    gamma, beta = torch.randn(c, generator=g), torch.randn(c, generator=g)
    want = F.silu(F.group_norm(full, groups, gamma, beta, eps=1e-5))
    got = ops.groupnorm(outs[0], outs[1], groups, gamma.to(DEV), beta.to(DEV), True, stats0=stats[0], stats1=stats[1])
    ref_kernel = ops.groupnorm(outs[0], outs[1], groups, gamma.to(DEV), beta.to(DEV), True)
    torch.cuda.synchronize()
    assert rel_l2(to_nchw(got.cpu()), want) < 4e-3
    assert rel_l2(to_nchw(got.cpu()), to_nchw(ref_kernel.cpu())) < 4e-3


# ---------------------------------------------------------------------------------------------
# attention core
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,c,L_,dtype", [(2, 16, 64, "fp32"), (3, 128, 256, "bf16"), (2, 256, 16, "bf16"), (1, 8, 1024, "fp32")])
def test_attention_single_head(n, c, L_, dtype):
    ops, _ = _ops()
    g = torch.Generator().manual_seed(9)
    dt = torch.bfloat16 if dtype == "bf16" else torch.float32
    q, k, v = (torch.randn(n, L_, c, generator=g) for _ in range(3))
    if dt == torch.bfloat16:
        q, k, v = bf16_round(q), bf16_round(k), bf16_round(v)
    scale = c ** -0.5
    want = torch.bmm(F.softmax(torch.bmm(q, k.transpose(1, 2) * scale), dim=2), v)
    out = torch.empty(n, L_, c, dtype=dt, device=DEV)
    vt = v.transpose(1, 2).contiguous().to(dt).to(DEV)
    ops.attention(q.to(dt).to(DEV), k.to(dt).to(DEV), vt, n, 1, L_, c, scale, L_ * c, c, 0, True, c * L_, False, out)
    torch.cuda.synchronize()
    assert rel_l2(out.float().cpu(), want) < (4e-3 if dt == torch.bfloat16 else 1e-5)


@pytest.mark.parametrize("n,c", [(3, 128), (2, 256), (5, 64), (300, 128)])
def test_attention_tc(n, c):
    """fused tcgen05 attention (L = 256) against fp32 softmax attention on the same bf16 operands"""
    ops, L = _ops()
    g = torch.Generator().manual_seed(19)
    seq = 256
    q, k, v = (bf16_round(torch.randn(n, seq, c, generator=g)) for _ in range(3))
    q = q * 2.0  # spread the logits so the softmax is far from uniform
    scale = c ** -0.5
    want = torch.bmm(F.softmax(torch.bmm(q, k.transpose(1, 2) * scale), dim=2), v)
    out = torch.empty(n, seq, c, dtype=torch.bfloat16, device=DEV)
    vt = v.transpose(1, 2).contiguous().to(torch.bfloat16).to(DEV)
    args = (q.to(torch.bfloat16).to(DEV), k.to(torch.bfloat16).to(DEV), vt, n, 1, seq, c, scale, seq * c, c, 0, True,
            c * seq, False)
    ops.attention(*args, out, kernel=L.CONV_TC)
    torch.cuda.synchronize()
    err = rel_l2(out.float().cpu(), want)
    assert err < 6e-3, err  # P is rounded to bf16 before the PV product
    gen = torch.empty_like(out)
    ops.attention(*args, gen, kernel=L.CONV_GENERIC)
    torch.cuda.synchronize()
    assert rel_l2(out.float().cpu(), gen.float().cpu()) < 6e-3


def _stats_of(x_nhwc):
    """int64 micro-group sums (2^-20 fixed point) of a stored NHWC tensor, as a producing conv's epilogue leaves them"""
    n, h, w, c = x_nhwc.shape
    v = x_nhwc.double().reshape(n, h * w, c // 4, 4)
    s1, s2 = v.sum((1, 3)), (v * v).sum((1, 3))
    return (torch.stack([s1, s2], dim=-1) * 2 ** 20).round().to(torch.int64).reshape(-1)


@pytest.mark.parametrize("n,hgt,c", [(1, 16, 256), (2, 16, 256), (5, 16, 256), (75, 16, 256), (150, 16, 256), (256, 16, 256),
                                     (1, 4, 256), (5, 4, 256), (8, 4, 256), (37, 4, 256), (256, 4, 256),
                                     (1, 16, 128), (3, 16, 128), (80, 16, 128), (256, 16, 128)])
def test_attention_block_fused(n, hgt, c):
    """The one-launch attention block (norm | qkv | softmax(q k^T) v | proj | + x, models/ddpm.py:54-75) against the fp32
    computation of the module on the same bf16-rounded input and weights.  16x16: n = 75 / 150 / 256: some / all clusters
    take two, three or four images (74 clusters of two CTAs); c = 128: the V^T and projection products run with duplicated
    weight rows.  4x4: eight images per CTA, n = 1 / 5 / 37: a ragged last group."""
    ops, L = _ops()
    g = torch.Generator().manual_seed(7 + n + hgt + c)
    groups = 32
    seq = hgt * hgt
    assert ops.attention_block_supported(1, seq, c, torch.bfloat16)
    # per-image and per-channel offsets / gains so that the norm matters
    x = torch.randn(n, c, hgt, hgt, generator=g) * (0.5 + torch.rand(n, c, 1, 1, generator=g)) + torch.randn(n, c, 1, 1, generator=g)
    x = bf16_round(x)
    gamma, beta = 1 + 0.3 * torch.randn(c, generator=g), 0.3 * torch.randn(c, generator=g)
    wqkv = bf16_round(torch.randn(3 * c, c, 1, 1, generator=g) * (2.0 / math.sqrt(c)))
    bqkv = torch.randn(3 * c, generator=g) * 0.5
    wproj = bf16_round(torch.randn(c, c, 1, 1, generator=g) / math.sqrt(c))
    bproj = torch.randn(c, generator=g) * 0.5
    scale = c ** -0.5

    hn = F.group_norm(x, groups, gamma, beta, eps=1e-5)
    qkv = F.conv2d(hn, wqkv, bqkv)
    q, k, v = (t.reshape(n, c, seq).transpose(1, 2) for t in qkv.chunk(3, dim=1))
    att = torch.bmm(F.softmax(torch.bmm(q, k.transpose(1, 2) * scale), dim=2), v)  # [n, L, c]
    att = att.transpose(1, 2).reshape(n, c, hgt, hgt)
    want = x + F.conv2d(att, wproj, bproj)

    xd = to_nhwc(x, torch.bfloat16).to(DEV)
    st = _stats_of(xd.cpu().float()).to(DEV)
    ab = ops.groupnorm_coeff(st, None, c, 0, n, seq, groups, gamma.to(DEV), beta.to(DEV), None, None, 1e-5)
    out_stats = torch.zeros(n * (c // 4) * 2, dtype=torch.int64, device=DEV)
    out = ops.attention_block(xd, ab, ops.pack_conv_weight(wqkv.to(DEV), None, True), bqkv.to(DEV),
                              ops.pack_conv_weight(wproj.to(DEV), None, True), bproj.to(DEV), scale, stats=out_stats)
    # the same with the coefficients formed inside the kernel from the producer's statistics: identical bits
    out2 = ops.attention_block(xd, None, ops.pack_conv_weight(wqkv.to(DEV), None, True), bqkv.to(DEV),
                               ops.pack_conv_weight(wproj.to(DEV), None, True), bproj.to(DEV), scale,
                               stats_in=st, gamma=gamma.to(DEV), beta=beta.to(DEV), groups=groups, eps=1e-5)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    got = to_nchw(out.cpu())
    err = rel_l2(got, want)
    # the attention branch alone (the residual x dominates the output norm)
    err_branch = rel_l2(got - x, want - x)
    assert err < 6e-3, err
    assert err_branch < 1.5e-2, err_branch  # h, q, k, v, P and O are each rounded to bf16 once
    # every image separately: a wrong image -> cluster assignment would hide in the aggregate
    per = ((got - want).flatten(1).norm(dim=1) / want.flatten(1).norm(dim=1)).max()
    assert per < 8e-3, per
    # statistics of the stored output
    sums = out_stats.cpu().view(n, c // 4, 2).double() / 2 ** 20
    yo = out.cpu().double().reshape(n, seq, c // 4, 4)
    assert torch.allclose(sums[..., 0], yo.sum((1, 3)), rtol=1e-5, atol=1e-2)
    assert torch.allclose(sums[..., 1], (yo * yo).sum((1, 3)), rtol=1e-5, atol=1e-2)


@pytest.mark.parametrize("n,c,heads,L_", [(3, 32, 4, 64), (2, 128, 4, 256), (4, 16, 4, 16)])
def test_attention_multi_head_iddpm_regrouping(n, c, heads, L_):
    """models/iddpm.py:36-47 including the (b head) -> (head b) output regrouping."""
    ops, _ = _ops()
    g = torch.Generator().manual_seed(10)
    qkv = torch.randn(n, L_, 3 * c, generator=g)  # NHWC view of the qkv conv output, channels [head][q|k|v][dh]
    dh = c // heads
    t = qkv.reshape(n, L_, heads, 3 * dh).permute(0, 2, 1, 3).reshape(n * heads, L_, 3 * dh)
    q, k, v = t.chunk(3, dim=2)
    scale = c ** -0.5
    o = torch.bmm(F.softmax(torch.bmm(q, k.transpose(1, 2) * scale), dim=2), v)
    want = o.reshape(heads, n, L_, dh).permute(1, 2, 0, 3).reshape(n, L_, c)
    dev = qkv.to(DEV).contiguous()
    flat = dev.view(-1)
    out = torch.empty(n, L_, c, dtype=torch.float32, device=DEV)
    ops.attention(flat, flat[dh:], flat[2 * dh:], n, heads, L_, dh, scale, L_ * 3 * c, 3 * c, 3 * dh, False, 0, True, out)
    torch.cuda.synchronize()
    assert rel_l2(out.cpu(), want) < 1e-5


# ---------------------------------------------------------------------------------------------
# timestep embedding
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,pos,emb", [(1, 128, 512), (7, 128, 512), (3, 4, 8)])
def test_temb(rows, pos, emb):
    ops, _ = _ops()
    g = torch.Generator().manual_seed(12)
    half = pos // 2
    freq = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1))).unsqueeze(0)
    t = torch.randint(0, 1001, (rows,), generator=g)
    w1, b1 = torch.randn(emb, pos, generator=g) / math.sqrt(pos), torch.randn(emb, generator=g)
    w2, b2 = torch.randn(emb, emb, generator=g) / math.sqrt(emb), torch.randn(emb, generator=g)
    e = t.unsqueeze(1) * freq
    e = torch.cat((e.sin(), e.cos()), dim=-1)
    want = F.silu(F.linear(F.silu(F.linear(e, w1, b1)), w2, b2))
    got = ops.temb_mlp(t.to(DEV), freq.to(DEV), w1.to(DEV), b1.to(DEV), w2.to(DEV), b2.to(DEV))
    torch.cuda.synchronize()
    assert rel_l2(got.cpu(), want) < 2e-5
    wc, bc = torch.randn(300, emb, generator=g) / math.sqrt(emb), torch.randn(300, generator=g)
    got2 = ops.temb_proj(got, wc.to(DEV), bc.to(DEV))
    torch.cuda.synchronize()
    assert rel_l2(got2.cpu(), F.linear(got.cpu(), wc, bc)) < 1e-5


# ---------------------------------------------------------------------------------------------
# sampler updates: bit-exact against the oracle given the same eps and z
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("t", [1000, 500, 2, 1])
def test_ddpm_step_bit_exact(t):
    import dmme_oracle as O
    ops, _ = _ops()
    g = torch.Generator().manual_seed(13)
    tabs = O.linear_tables(1000)
    x, eps, z = (torch.randn(5, 3, 32, 32, generator=g) for _ in range(3))
    tt = torch.tensor([t])
    want = O.ddpm_step(x, tt, eps, z, tabs)
    xd = x.to(DEV).clone()
    ops.ddpm_step_(xd, eps.to(DEV), z.to(DEV), *(tb.to(DEV) for tb in tabs), tt.to(DEV))
    torch.cuda.synchronize()
    assert torch.equal(xd.cpu(), want)


@pytest.mark.parametrize("i", [50, 25, 2, 1])
def test_ddim_step_bit_exact(i):
    import dmme_oracle as O
    ops, _ = _ops()
    g = torch.Generator().manual_seed(14)
    _, _, ab = O.linear_tables(1000)
    tau = O.tau_table(1000, 50)
    x, eps = (torch.randn(4, 3, 32, 32, generator=g) for _ in range(2))
    ii = torch.tensor([i])
    want = O.ddim_step(x, ii, eps, ab, tau)
    xd = x.to(DEV).clone()
    ops.ddim_step_(xd, eps.to(DEV), ab.to(DEV), tau.to(DEV), ii.to(DEV))
    torch.cuda.synchronize()
    assert torch.equal(xd.cpu(), want)


@pytest.mark.parametrize("t", [1000, 400, 1])
def test_iddpm_step(t):
    import dmme_oracle as O
    ops, _ = _ops()
    g = torch.Generator().manual_seed(15)
    tabs = O.cosine_tables(1000)
    x, z = (torch.randn(4, 3, 32, 32, generator=g) for _ in range(2))
    mo = torch.randn(4, 6, 32, 32, generator=g)
    tt = torch.tensor([t])
    want = O.iddpm_step(x, tt, mo, z, tabs)
    xd = x.to(DEV).clone()
    ops.iddpm_step_(xd, mo.to(DEV), z.to(DEV), *(tb.to(DEV) for tb in tabs), tt.to(DEV))
    torch.cuda.synchronize()
    assert rel_l2(xd.cpu(), want) < 1e-6  # expf/logf differ from the CPU's by an ulp


def test_philox_normal_moments_and_step_counter():
    ops, _ = _ops()
    z = ops.philox_normal((1 << 20,), 1234, 7, DEV)
    z2 = ops.philox_normal((1 << 20,), 1234, 8, DEV)
    torch.cuda.synchronize()
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1) < 5e-3
    assert abs(float((z * z2).mean())) < 5e-3
    assert abs(float((z ** 4).mean()) - 3) < 5e-2
    c = torch.full((1,), 10, dtype=torch.int64, device=DEV)
    ops.add_i64_(c, -1)
    tau = torch.arange(0, 40, 2, device=DEV)
    out = torch.zeros(1, dtype=torch.int64, device=DEV)
    ops.gather_i64(tau, c, out)
    torch.cuda.synchronize()
    assert int(c) == 9 and int(out) == 18


# ---------------------------------------------------------------------------------------------
# image-space tail: denorm / snapshot history (callbacks/generate.py, common/norm.py)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(3, 3, 32, 32), (1, 1, 5, 7), (64, 3, 32, 32)])
def test_denorm_bit_exact(shape):
    import dmme_b200
    g = torch.Generator().manual_seed(5)
    x = torch.randn(shape, generator=g) * 1.5
    x.view(-1)[:4] = torch.tensor([-1.0, 1.0, -3.0, 3.0])
    want = torch.clip((x + 1) / 2, 0, 1)  # common/norm.py:9-11
    got = dmme_b200.denorm(x.to(DEV)).cpu()
    assert torch.equal(got, want)
    u8 = dmme_b200.denorm_uint8(x.to(DEV)).cpu()
    assert u8.dtype == torch.uint8 and torch.equal(u8, (want * 255).round().to(torch.uint8))


@pytest.mark.parametrize("graph", [True, False])
def test_generate_history_matches_generate_taps(graph):
    """generate_history == denorm of the chain's states at the reference's save_t (callbacks/generate.py:73-83)"""
    import dmme_b200
    from dmme_b200.models.ddpm import UNet
    torch.manual_seed(0)
    model = UNet(pos_dim=32, emb_dim=64, channels_per_depth=(64, 128), num_blocks=1, attention_depths=(2,)).eval()
    ddpm = dmme_b200.DDPM(model, timesteps=20).to(DEV)
    g = torch.Generator().manual_seed(2)
    x_T = torch.randn(2, 3, 32, 32, generator=g)
    save_t = dmme_b200.DDPM.history_timesteps(20, 5)
    assert save_t == [int(20 / 4 * i) for i in range(4, 0, -1)] == [20, 15, 10, 5]
    states = {20: x_T.clone()}

    def after(k, x):  # state after step k = x_{T-k-1}
        states[20 - k - 1] = x.detach().cpu().clone()

    final = ddpm.generate((2, 3, 32, 32), x_T=x_T, seed=77, graph=graph, on_step=after).cpu()
    hist = ddpm.generate_history((2, 3, 32, 32), vis_length=5, x_T=x_T, seed=77, graph=graph).cpu()
    assert hist.shape == (5, 2, 3, 32, 32)
    for i, t in enumerate(save_t):
        assert torch.equal(hist[i], torch.clip((states[t] + 1) / 2, 0, 1)), t
    assert torch.equal(hist[4], torch.clip((final + 1) / 2, 0, 1))
    h8 = ddpm.generate_history((2, 3, 32, 32), vis_length=5, x_T=x_T, seed=77, graph=graph, uint8=True).cpu()
    assert torch.equal(h8, (hist * 255).round().to(torch.uint8))


@pytest.mark.parametrize("n,c,cout,h", [(3, 128, 128, 16), (40, 256, 256, 8), (2, 64, 128, 32), (70, 128, 256, 16), (5, 256, 256, 8)])
def test_conv_upsample_subpixel(n, c, cout, h):
    """UpSample (nearest x2 -> conv3x3, models/ddpm.py:150-173) as four 2x2 phase convs on the low-resolution tensor:
    against F.interpolate + conv2d on the same bf16 input (phase weights are summed in fp32 then rounded to bf16, so the
    reference here uses fp32 weights and the tolerance is the bf16 weight rounding), plus the GroupNorm statistics"""
    ops, L = _ops()
    g = torch.Generator().manual_seed(51)
    x = bf16_round(torch.randn(n, c, h, h, generator=g))
    w = torch.randn(cout, c, 3, 3, generator=g) / math.sqrt(9 * c)
    b = torch.randn(cout, generator=g)
    s0 = to_nhwc(x, torch.bfloat16).to(DEV)
    d = ops.make_conv_desc(s0, None, cout, 3, 1, 3, None, None, False, L.OUT_NHWC, torch.bfloat16, L.CONV_AUTO)
    assert ops.conv_uses_tc(d) and ops.conv_out_hw(d) == (2 * h, 2 * h)
    wp = ops.pack_upsample_phase_weight(w.to(DEV))
    out = torch.full((n, 2 * h, 2 * h, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
    st = torch.zeros(n * (cout // 4) * 2, dtype=torch.int64, device=DEV)
    ops.conv2d_launch(d, wp, b.to(DEV), out, stats=st)
    torch.cuda.synchronize()
    got = to_nchw(out.cpu())
    want = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, b, padding=1)
    err = rel_l2(got, want)
    assert err < 6e-3, f"rel-L2 {err}"
    sums = st.cpu().view(n, cout // 4, 2).double() / 2 ** 20
    assert torch.allclose(sums[..., 0], got.double().reshape(n, cout // 4, -1).sum(-1), rtol=1e-5, atol=2e-2)
    assert torch.allclose(sums[..., 1], (got.double() ** 2).reshape(n, cout // 4, -1).sum(-1), rtol=1e-5, atol=2e-2)


@pytest.mark.parametrize("n,c,heads,L_,swap", [(3, 256, 4, 256, True), (2, 128, 4, 256, True), (5, 256, 4, 64, True),
                                               (1, 256, 4, 64, False), (7, 128, 2, 64, True), (9, 64, 2, 256, False),
                                               (2, 256, 4, 16, True), (130, 128, 4, 64, False), (1, 64, 1, 256, False),
                                               (2, 256, 2, 128, True)])
def test_attention_mma_multi_head(n, c, heads, L_, swap):
    """multi-head kernels (tcgen05 at 256 tokens and at 64 tokens with 64-channel heads, mma.sync elsewhere; bf16 storage)
    against fp32 attention on the same bf16 operands, IDDPM channel layout
    [head][q|k|v][dh] and the (b head) -> (head b) regrouping (models/iddpm.py:36-47)"""
    ops, L = _ops()
    g = torch.Generator().manual_seed(12)
    qkv = bf16_round(torch.randn(n, L_, 3 * c, generator=g))
    dh = c // heads
    t = qkv.reshape(n, L_, heads, 3 * dh).permute(0, 2, 1, 3).reshape(n * heads, L_, 3 * dh)
    q, k, v = t.chunk(3, dim=2)
    scale = c ** -0.5
    o = torch.bmm(F.softmax(torch.bmm(q, k.transpose(1, 2) * scale), dim=2), v)
    if swap:
        want = o.reshape(heads, n, L_, dh).permute(1, 2, 0, 3).reshape(n, L_, c)
    else:
        want = o.reshape(n, heads, L_, dh).permute(0, 2, 1, 3).reshape(n, L_, c)
    dev = qkv.to(torch.bfloat16).to(DEV).contiguous()
    flat = dev.view(-1)
    guarded = torch.full((n + 2, L_, c), float("nan"), dtype=torch.bfloat16, device=DEV)  # one canary image on either side
    out = guarded[1:-1]
    ops.attention(flat, flat[dh:], flat[2 * dh:], n, heads, L_, dh, scale, L_ * 3 * c, 3 * c, 3 * dh, False, 0, swap, out)
    ref = torch.empty_like(out)
    lib = L.load()
    lib.dmme_set_attn_mma_mode(0)
    try:
        ops.attention(flat, flat[dh:], flat[2 * dh:], n, heads, L_, dh, scale, L_ * 3 * c, 3 * c, 3 * dh, False, 0, swap, ref)
    finally:
        lib.dmme_set_attn_mma_mode(1)
    torch.cuda.synchronize()
    assert rel_l2(ref.float().cpu(), want) < 4e-3          # CUDA-core kernel: only the bf16 output rounding
    assert rel_l2(out.float().cpu(), want) < 6e-3, rel_l2(out.float().cpu(), want)
    assert torch.isnan(guarded[0]).all() and torch.isnan(guarded[-1]).all()  # nothing written outside the output


@pytest.mark.parametrize("cfg", [
    dict(n=4, c0=128, c1=0, cout=128, h=32, temb="bcast"),
    dict(n=37, c0=128, c1=128, cout=128, h=32, res=True),            # concat source, fused 1x1 residual, several units per CTA
    dict(n=3, c0=256, c1=0, cout=256, h=16, addend=True, temb="rows"),
    dict(n=70, c0=256, c1=256, cout=256, h=16, silu=False),          # tiles spanning two images, plain norm
    dict(n=150, c0=128, c1=0, cout=128, h=32),
    # the tilings bench.py times: 13 / 11 rows per tile at 16x16 only appear at batch 256; per-GPU shards 128 / 64 / 32
    dict(n=256, c0=256, c1=0, cout=256, h=16, temb="bcast"),
    dict(n=256, c0=256, c1=0, cout=128, h=16),
    dict(n=128, c0=256, c1=256, cout=256, h=16),
    dict(n=64, c0=256, c1=0, cout=256, h=16, addend=True),
    dict(n=32, c0=128, c1=128, cout=128, h=32, res=True),
    dict(n=32, c0=256, c1=0, cout=256, h=16, temb="bcast"),
    # 8x8 maps: row tiles of two whole images in the shared-padding layout (AUTO's choice from 160 images per GPU up)
    dict(n=256, c0=256, c1=0, cout=256, h=8, temb="bcast"),
    dict(n=256, c0=256, c1=256, cout=256, h=8),                      # concat input of the up path
    dict(n=5, c0=256, c1=0, cout=256, h=8, addend=True, temb="rows"),  # odd batch: the last tile holds one image
    dict(n=200, c0=256, c1=256, cout=256, h=8, res=True),            # fused 1x1 residual over the concat (eight single-tap chunks)
    dict(n=161, c0=128, c1=128, cout=128, h=8, silu=False),
])
def test_conv_halo_fused_groupnorm(cfg):
    """GroupNorm(+SiLU) applied to the halo tile inside the conv kernel == gn_apply followed by the same conv, bit for bit
    (same coefficient function, same fma / SiLU, zero padding after the activation)"""
    ops, L = _ops()
    g = torch.Generator().manual_seed(61)
    n, c0, c1, cout, h = (cfg[k] for k in ("n", "c0", "c1", "cout", "h"))
    C = c0 + c1
    silu = cfg.get("silu", True)
    # producers: convs that write the tensors and their statistics
    srcs, stats = [], []
    for c in (c0, c1):
        if c == 0:
            srcs.append(None); stats.append(None)
            continue
        x = bf16_round(torch.randn(n, 64, h, h, generator=g))
        w = bf16_round(torch.randn(c, 64, 3, 3, generator=g) / 8)
        b = torch.randn(c, generator=g) * 2
        s0 = to_nhwc(x, torch.bfloat16).to(DEV)
        d = ops.make_conv_desc(s0, None, c, 3, 1, False, None, None, False, L.OUT_NHWC, torch.bfloat16, L.CONV_AUTO)
        out = torch.empty((n, h, h, c), dtype=torch.bfloat16, device=DEV)
        st = torch.zeros(n * (c // 4) * 2, dtype=torch.int64, device=DEV)
        ops.conv2d_launch(d, ops.pack_conv_weight(w.to(DEV), None, True), b.to(DEV), out, stats=st)
        srcs.append(out); stats.append(st)
    gamma, beta = torch.randn(C, generator=g).to(DEV), torch.randn(C, generator=g).to(DEV)
    wt = bf16_round(torch.randn(cout, C, 3, 3, generator=g) / math.sqrt(9 * C))
    bias = torch.randn(cout, generator=g)
    wres = None
    if cfg.get("res"):
        wres = bf16_round(torch.randn(cout, C, 1, 1, generator=g) / math.sqrt(C))
    temb = None
    if cfg.get("temb"):
        temb = torch.randn(n if cfg["temb"] == "rows" else 1, cout, generator=g).to(DEV)
    addend = bf16_round(torch.randn(n, h, h, cout, generator=g)).to(torch.bfloat16).to(DEV) if cfg.get("addend") else None
    wp = ops.pack_conv_weight(wt.to(DEV), wres.to(DEV) if wres is not None else None, True)
    r0, r1 = (srcs[0], srcs[1]) if cfg.get("res") else (None, None)

    # two-kernel path
    a = ops.groupnorm(srcs[0], srcs[1], 32, gamma, beta, silu, stats0=stats[0], stats1=stats[1])
    d = ops.make_conv_desc(a, None, cout, 3, 1, False, r0, r1, False, L.OUT_NHWC, torch.bfloat16, L.CONV_HALO)
    want = torch.empty((n, h, h, cout), dtype=torch.bfloat16, device=DEV)
    st_want = torch.zeros(n * (cout // 4) * 2, dtype=torch.int64, device=DEV)
    ops.conv2d_launch(d, wp, bias.to(DEV), want, temb, addend, stats=st_want)

    # fused path: the conv reads the raw tensors
    d = ops.make_conv_desc(srcs[0], srcs[1], cout, 3, 1, False, r0, r1, False, L.OUT_NHWC, torch.bfloat16, L.CONV_HALO)
    assert ops.conv_fuses_gn(d)
    ab = ops.groupnorm_coeff(stats[0], stats[1], c0, c1, n, h * h, 32, gamma, beta)
    got = torch.full((n, h, h, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
    st_got = torch.zeros_like(st_want)
    ops.conv2d_launch(d, wp, bias.to(DEV), got, temb, addend, stats=st_got, gn_ab=ab, gn_silu=silu)
    torch.cuda.synchronize()
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    assert torch.equal(st_got, st_want)


@pytest.mark.parametrize("cfg", [
    dict(n=150, c0=128, c1=0, cout=128, h=32, gn=True),             # 4 full rounds of 7-row tiles + one round of short tiles
    dict(n=256, c0=128, c1=128, cout=128, h=32, res=True, gn=True),  # the step's dominant launch: 8 rounds + a 2-row tail
    dict(n=128, c0=256, c1=0, cout=256, h=16, gn=True),             # two channel tiles per row tile
    dict(n=64, c0=128, c1=0, cout=128, h=32),
    dict(n=37, c0=256, c1=256, cout=256, h=16),
])
def test_conv_halo_tail_tiles(cfg):
    """tail tiles (the rows of the partial last wave as one round of short tiles, csrc/conv_halo.cu) == equal tiles: raw
    outputs bit for bit (same K order per output); GroupNorm statistics to the fixed-point rounding of the per-tile sums"""
    ops, L = _ops()
    lib = L.load()
    g = torch.Generator().manual_seed(71)
    n, c0, c1, cout, h = (cfg[k] for k in ("n", "c0", "c1", "cout", "h"))
    C = c0 + c1
    s0 = to_nhwc(bf16_round(torch.randn(n, c0, h, h, generator=g)), torch.bfloat16).to(DEV)
    s1 = to_nhwc(bf16_round(torch.randn(n, c1, h, h, generator=g)), torch.bfloat16).to(DEV) if c1 else None
    wt = bf16_round(torch.randn(cout, C, 3, 3, generator=g) / math.sqrt(9 * C))
    wres = bf16_round(torch.randn(cout, C, 1, 1, generator=g) / math.sqrt(C)) if cfg.get("res") else None
    wp = ops.pack_conv_weight(wt.to(DEV), wres.to(DEV) if wres is not None else None, True)
    bias = torch.randn(cout, generator=g).to(DEV)
    ab = torch.randn(n, C, 2, generator=g).to(DEV) if cfg.get("gn") else None
    r0, r1 = (s0, s1) if cfg.get("res") else (None, None)
    d = ops.make_conv_desc(s0, s1, cout, 3, 1, False, r0, r1, False, L.OUT_NHWC, torch.bfloat16, L.CONV_HALO)
    outs = []
    mode0 = lib.dmme_get_conv_halo_mode()
    try:
        for mode in (mode0 | 32, mode0 & ~32):  # bit 5: equal tiles only
            lib.dmme_set_conv_halo_mode(mode)
            out = torch.full((n, h, h, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
            st = torch.zeros(n * (cout // 4) * 2, dtype=torch.int64, device=DEV)
            ops.conv2d_launch(d, wp, bias, out, stats=st, gn_ab=ab)
            torch.cuda.synchronize()
            outs.append((out, st))
    finally:
        lib.dmme_set_conv_halo_mode(mode0)
    assert torch.isfinite(outs[0][0].float()).all()
    assert torch.equal(outs[0][0].view(torch.int16), outs[1][0].view(torch.int16))
    a, b = (o[1].double() / 2 ** 20 for o in outs)
    assert torch.allclose(a, b, rtol=1e-5, atol=2e-2)


@pytest.mark.parametrize("cfg", [
    dict(n=5, c0=128, c1=0, cout=128, h=32),                       # odd number of row tiles: the pair's last tile is past the batch
    dict(n=37, c0=128, c1=128, cout=128, h=32, res=True, gn=True),
    dict(n=9, c0=256, c1=0, cout=256, h=16, gn=True),              # two channel tiles per row-tile pair
    dict(n=70, c0=256, c1=256, cout=256, h=16),
    dict(n=300, c0=128, c1=0, cout=128, h=32, gn=True),            # several items per CTA pair
])
def test_conv_halo_weight_multicast(cfg):
    """clusters of two CTAs sharing the weight stream (TMA multicast) == independent CTAs, bit for bit (outputs and
    GroupNorm statistics; same MMA order per output)"""
    if not _ops()[1].load().dmme_has_experimental():
        pytest.skip("weight-multicast variant is compiled only with -DDMME_EXPERIMENTAL (measured slower, DESIGN.md)")
    ops, L = _ops()
    lib = L.load()
    g = torch.Generator().manual_seed(67)
    n, c0, c1, cout, h = (cfg[k] for k in ("n", "c0", "c1", "cout", "h"))
    C = c0 + c1
    s0 = to_nhwc(bf16_round(torch.randn(n, c0, h, h, generator=g)), torch.bfloat16).to(DEV)
    s1 = to_nhwc(bf16_round(torch.randn(n, c1, h, h, generator=g)), torch.bfloat16).to(DEV) if c1 else None
    wt = bf16_round(torch.randn(cout, C, 3, 3, generator=g) / math.sqrt(9 * C))
    wres = bf16_round(torch.randn(cout, C, 1, 1, generator=g) / math.sqrt(C)) if cfg.get("res") else None
    wp = ops.pack_conv_weight(wt.to(DEV), wres.to(DEV) if wres is not None else None, True)
    bias = torch.randn(cout, generator=g).to(DEV)
    ab = torch.randn(n, C, 2, generator=g).to(DEV) if cfg.get("gn") else None
    r0, r1 = (s0, s1) if cfg.get("res") else (None, None)
    d = ops.make_conv_desc(s0, s1, cout, 3, 1, False, r0, r1, False, L.OUT_NHWC, torch.bfloat16, L.CONV_HALO)
    outs = []
    mode0 = lib.dmme_get_conv_halo_mode()
    lib.dmme_set_conv_halo_mode(mode0 | 32)  # equal row tiles in both runs (pairs never take tail tiles): same per-tile sums
    try:
        for mode in (0, 2):
            lib.dmme_set_conv_halo_multicast(mode)
            out = torch.full((n, h, h, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
            st = torch.zeros(n * (cout // 4) * 2, dtype=torch.int64, device=DEV)
            ops.conv2d_launch(d, wp, bias, out, stats=st, gn_ab=ab)
            torch.cuda.synchronize()
            outs.append((out, st))
    finally:
        lib.dmme_set_conv_halo_multicast(0)
        lib.dmme_set_conv_halo_mode(mode0)
    assert torch.isfinite(outs[0][0].float()).all()
    assert torch.equal(outs[0][0].view(torch.int16), outs[1][0].view(torch.int16))
    assert torch.equal(outs[0][1], outs[1][1])


# ---------------------------------------------------------------------------------------------------------------------
# split-K convolution + finishing pass with the consumers' GroupNorm(+SiLU) (4x4 / 8x8 levels, small per-GPU batches)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture
def splitk_everywhere():
    _, L = _ops()
    lib = L.load()
    lib.dmme_set_conv_splitk_mode(2)
    yield
    lib.dmme_set_conv_splitk_mode(1)


SPLITK_CASES = [
    dict(n=32, cin=256, cout=256, h=4, temb="bcast", norms=[(8, True)]),                       # 4x4 at the 8-GPU shard batch
    dict(n=32, cin=256, cout=256, h=8, addend=True, norms=[(8, True), (16, True)]),            # 8x8, chain + skip consumer
    dict(n=256, cin=512, cin1=256, cout=256, h=4, res=True, temb="bcast", norms=[(16, True)]),  # up path: concat + fused residual
    dict(n=4, cin=256, cout=256, h=4, temb="rows", norms=[(8, False)]),                        # fewer pixels than one tile; plain norm
    dict(n=7, cin=128, cout=128, h=8, norms=[(4, True)], scale_shift=True),                    # ragged batch, IDDPM scale / shift
    dict(n=32, cin=256, cout=256, h=16, stride=2, norms=[(8, True), (16, True)]),              # stride-2 down-sampling conv 16 -> 8
    dict(n=64, cin=256, cout=256, h=8, norms=[]),                                              # no consumer norm: raw + stats only
    dict(n=128, cin=256, cout=256, h=4, temb="bcast", addend=True, norms=[(8, True)]),
]


@pytest.mark.parametrize("cfg", SPLITK_CASES)
def test_conv_splitk_with_output_norms(cfg, splitk_everywhere):
    """split-K tcgen05 conv (K slices as work units, fp32 partial tiles) + finishing pass == fp32 conv on the same bf16
    operands; raw output, its GroupNorm statistics, and the fused GroupNorm(+SiLU) outputs for up to two consumers"""
    _conv_out_norms_case(cfg, split=True)


EPI_NORM_CASES = [
    dict(n=256, cin=256, cout=256, h=8, temb="bcast", norms=[(8, True)]),                          # the timed batch: 256-pixel tiles
    dict(n=256, cin=512, cin1=256, cout=256, h=8, res=True, temb="bcast", norms=[(16, True), (8, False)]),  # up path, two consumers
    dict(n=128, cin=256, cout=256, h=8, addend=True, norms=[(8, True), (16, True)]),               # 128-pixel tiles
    dict(n=203, cin=128, cout=128, h=8, norms=[(4, True)], scale_shift=True, temb="rows"),         # ragged batch, IDDPM scale / shift
    dict(n=128, cin=256, cout=256, h=16, stride=2, norms=[(8, True), (16, True)]),                 # stride-2 down-sampling conv 16 -> 8
    dict(n=230, cin=256, cout=128, h=8, norms=[(32, True)]),                                       # one group per warp
]


@pytest.mark.parametrize("cfg", EPI_NORM_CASES)
def test_conv_tct_epilogue_norm(cfg):
    """8x8 maps on the UNSPLIT transposed tcgen05 conv: the conv's own epilogue (whole images per warp, a GroupNorm group =
    neighbouring lanes) writes the raw output, its statistics and the consumers' GroupNorm(+SiLU)"""
    _conv_out_norms_case(cfg, split=False)


def _conv_out_norms_case(cfg, split):
    ops, L = _ops()
    g = torch.Generator().manual_seed(91)
    n, cin, cout, h = (cfg[s] for s in ("n", "cin", "cout", "h"))
    stride = cfg.get("stride", 1)
    c1 = cfg.get("cin1", 0)
    c0 = cin - c1
    ho = h // stride
    xa = bf16_round(torch.randn(n, cin, h, h, generator=g))
    wt = bf16_round(torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9))
    b = torch.randn(cout, generator=g)
    s0 = to_nhwc(xa[:, :c0], torch.bfloat16).to(DEV)
    s1 = to_nhwc(xa[:, c0:], torch.bfloat16).to(DEV) if c1 else None
    want = F.conv2d(xa, wt, b, stride=stride, padding=1)
    wres, r0, r1, bias = None, None, None, b
    if cfg.get("res"):
        wres = bf16_round(torch.randn(cout, cin, 1, 1, generator=g) / math.sqrt(cin))
        bres = torch.randn(cout, generator=g)
        want = want + F.conv2d(xa, wres, bres)
        bias = b + bres
        r0, r1 = s0, s1
    temb = None
    if cfg.get("temb"):
        temb = torch.randn(n if cfg["temb"] == "rows" else 1, cout, generator=g)
        want = want + (temb if temb.shape[0] == n else temb.expand(n, -1))[:, :, None, None]
        temb = temb.to(DEV)
    addend = None
    if cfg.get("addend"):
        ad = bf16_round(torch.randn(n, cout, ho, ho, generator=g))
        want = want + ad
        addend = to_nhwc(ad, torch.bfloat16).to(DEV)
    d = ops.make_conv_desc(s0, s1, cout, 3, stride, False, r0, r1, False, L.OUT_NHWC, torch.bfloat16, L.CONV_AUTO)
    if split:
        ws_bytes = ops.conv_splitk_workspace(d)
        assert ws_bytes > 0, "the forced split-K mode must accept this shape"
        ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=DEV)
    else:
        assert ops.conv_epilogue_norm(d), "the unsplit transposed kernel must take this shape with its epilogue norm"
        ws = None
    wp = ops.pack_conv_weight(wt.to(DEV), wres.to(DEV) if wres is not None else None, True)
    out = torch.empty((n, ho, ho, cout), dtype=torch.bfloat16, device=DEV)
    st = torch.zeros(n * (cout // 4) * 2, dtype=torch.int64, device=DEV)
    norms, specs = [], []
    for cpg, silu in cfg["norms"]:
        gamma, beta = torch.randn(cout, generator=g), torch.randn(cout, generator=g)
        sc = sh = None
        if cfg.get("scale_shift"):
            sc, sh = torch.randn(n, cout, generator=g) * 0.3, torch.randn(n, cout, generator=g)
        y = torch.full((n, ho, ho, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
        # dmme_out_norm holds raw pointers: the device tensors must outlive the launch
        dev_t = [t.to(DEV) if t is not None else None for t in (gamma, beta, sc, sh)]
        norms.append(ops.out_norm(y, dev_t[0], dev_t[1], cpg, silu, 1e-5, dev_t[2], dev_t[3]))
        specs.append((y, gamma, beta, cpg, silu, sc, sh, dev_t))
    ops.conv2d_launch(d, wp, bias.to(DEV), out, temb, addend, stats=st, splitk_ws=ws, out_norms=norms)
    torch.cuda.synchronize()
    got = to_nchw(out.cpu())
    err = rel_l2(got, want)
    assert err < 4e-3, f"raw output rel-L2 {err}"
    sums = st.cpu().view(n, cout // 4, 2).double() / 2 ** 20
    assert torch.allclose(sums[..., 0], got.double().reshape(n, cout // 4, -1).sum(-1), rtol=1e-5, atol=2e-2)
    assert torch.allclose(sums[..., 1], (got.double() ** 2).reshape(n, cout // 4, -1).sum(-1), rtol=1e-5, atol=2e-2)
    for y, gamma, beta, cpg, silu, sc, sh, _ in specs:
        ref = F.group_norm(got, cout // cpg, gamma, beta, 1e-5)  # the norm of the tensor AS STORED (bf16-rounded)
        if sc is not None:
            ref = ref * (1 + sc[:, :, None, None]) + sh[:, :, None, None]
        if silu:
            ref = F.silu(ref)
        e = rel_l2(to_nchw(y.cpu()), ref)
        assert e < 6e-3, f"fused output norm (cpg {cpg}, silu {silu}) rel-L2 {e}"


@pytest.mark.parametrize("n,cin,cpg,addend", [(256, 256, 8, False), (37, 512, 16, True), (3, 256, 4, True)])
def test_splitk_finish_small_bit_equal(n, cin, cpg, addend, splitk_everywhere):
    """4x4 maps: the warp-per-slab finishing kernel writes the same bits as the block-per-slab kernel (same summation
    order): raw output, statistics, and both consumers' GroupNorm outputs"""
    ops, L = _ops()
    lib = L.load()
    g = torch.Generator().manual_seed(97)
    cout, h = 256, 4
    x = torch.randn(n, h, h, cin, generator=g).to(torch.bfloat16).to(DEV)
    wp = ops.pack_conv_weight((torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)).to(DEV), None, True)
    bias, gamma, beta = (torch.randn(cout, generator=g).to(DEV) for _ in range(3))
    temb = torch.randn(n, cout, generator=g).to(DEV)
    ad = torch.randn(n, h, h, cout, generator=g).to(torch.bfloat16).to(DEV) if addend else None
    res = []
    for mode in (0, 2):
        lib.dmme_set_splitk_finish_small(mode)
        lib.dmme_set_conv_splitk_cluster(0)  # the two finishing kernels, not the in-cluster reduction
        try:
            d = ops.make_conv_desc(x, None, cout, 3, 1, False, None, None, False, L.OUT_NHWC, torch.bfloat16, L.CONV_AUTO)
            ws = torch.empty(ops.conv_splitk_workspace(d) // 4, dtype=torch.float32, device=DEV)
            out = torch.empty((n, h, h, cout), dtype=torch.bfloat16, device=DEV)
            y0, y1 = torch.empty_like(out), torch.empty_like(out)
            st = torch.zeros(n * (cout // 4) * 2, dtype=torch.int64, device=DEV)
            ops.conv2d_launch(d, wp, bias, out, temb, ad, stats=st, splitk_ws=ws,
                              out_norms=[ops.out_norm(y0, gamma, beta, cpg, True), ops.out_norm(y1, beta, gamma, 32, False)])
            torch.cuda.synchronize()
            res.append((out, st, y0, y1))
        finally:
            lib.dmme_set_splitk_finish_small(1)
            lib.dmme_set_conv_splitk_cluster(1)
    for a, b in zip(*res):
        assert torch.equal(a, b)


@pytest.mark.parametrize("n,cin,cin1,cpg,addend,res", [(256, 256, 0, 8, False, False), (256, 512, 256, 16, True, True),
                                                      (37, 512, 256, 16, True, False), (3, 256, 0, 4, True, False),
                                                      (32, 256, 0, 8, False, True), (128, 512, 256, 8, False, False)])
@pytest.mark.parametrize("bend", [1, 2])
def test_conv_splitk_cluster_bit_equal(n, cin, cin1, cpg, addend, res, bend, splitk_everywhere):
    """4x4 maps: the K slices reduced inside a thread-block cluster through distributed shared memory and finished by the
    same launch write the same bits as split-K GEMM + finishing pass (same slice order): raw output, statistics, both
    consumers' GroupNorm outputs.  Concat inputs, fused 1x1 residual, ragged batches."""
    ops, L = _ops()
    lib = L.load()
    g = torch.Generator().manual_seed(101 + n)
    cout, h = 256, 4
    c0 = cin - cin1
    xa = torch.randn(n, h, h, cin, generator=g).to(torch.bfloat16)
    s0 = xa[..., :c0].contiguous().to(DEV)
    s1 = xa[..., c0:].contiguous().to(DEV) if cin1 else None
    w = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)).to(DEV)
    wres = (torch.randn(cout, cin, 1, 1, generator=g) / math.sqrt(cin)).to(DEV) if res else None
    wp = ops.pack_conv_weight(w, wres, True)
    bias, gamma, beta = (torch.randn(cout, generator=g).to(DEV) for _ in range(3))
    temb = torch.randn(n, cout, generator=g).to(DEV)
    ad = torch.randn(n, h, h, cout, generator=g).to(torch.bfloat16).to(DEV) if addend else None
    outs = []
    # bend = 2: the plan restricted to the shapes the cluster reduction takes (also CTAs that own a single image); the
    # reference run needs the same plan without the cluster: switch value 3
    for cluster in ((1, 0) if bend == 1 else (2, 3)):
        lib.dmme_set_conv_splitk_cluster(cluster)
        lib.dmme_set_splitk_finish_small(2)
        try:
            d = ops.make_conv_desc(s0, s1, cout, 3, 1, False, s0 if res else None, s1 if res else None, False, L.OUT_NHWC,
                                   torch.bfloat16, L.CONV_AUTO)
            ws = torch.empty(ops.conv_splitk_workspace(d) // 4, dtype=torch.float32, device=DEV)
            out = torch.full((n, h, h, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
            y0, y1 = torch.full_like(out, float("nan")), torch.full_like(out, float("nan"))
            st = torch.zeros(n * (cout // 4) * 2, dtype=torch.int64, device=DEV)
            ops.conv2d_launch(d, wp, bias, out, temb, ad, stats=st, splitk_ws=ws,
                              out_norms=[ops.out_norm(y0, gamma, beta, cpg, True), ops.out_norm(y1, beta, gamma, 32, False)])
            torch.cuda.synchronize()
            outs.append((out, st, y0, y1))
        finally:
            lib.dmme_set_conv_splitk_cluster(1)
            lib.dmme_set_splitk_finish_small(1)
    assert not torch.isnan(outs[0][0].float()).any()
    for a, b in zip(*outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("n,cin,cin1,cpg,addend,res,shape", [(64, 256, 0, 8, False, False, (256, 4)), (64, 512, 256, 16, True, True, (256, 4)),
                                                            (37, 256, 0, 8, True, False, (256, 4)), (32, 256, 0, 4, False, False, (256, 2)),
                                                            (50, 512, 256, 8, False, True, (128, 2))])
def test_conv_splitk_cluster_8x8(n, cin, cin1, cpg, addend, res, shape):
    """8x8 maps: the in-cluster reduction (one image per CTA, or more) against split-K GEMM + finishing pass with the same
    plan: the raw output is bit-equal (same slice order), statistics and the consumers' norms agree to fp32 summation order"""
    ops, L = _ops()
    lib = L.load()
    import ctypes as C
    lib.dmme_debug_force_splitk_4x4.argtypes = [C.c_int, C.c_int]
    lib.dmme_debug_force_splitk_4x4.restype = None
    g = torch.Generator().manual_seed(211 + n)
    cout, h = 256, 8
    c0 = cin - cin1
    xa = torch.randn(n, h, h, cin, generator=g).to(torch.bfloat16)
    s0 = xa[..., :c0].contiguous().to(DEV)
    s1 = xa[..., c0:].contiguous().to(DEV) if cin1 else None
    w = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)).to(DEV)
    wres = (torch.randn(cout, cin, 1, 1, generator=g) / math.sqrt(cin)).to(DEV) if res else None
    wp = ops.pack_conv_weight(w, wres, True)
    bias, gamma, beta = (torch.randn(cout, generator=g).to(DEV) for _ in range(3))
    temb = torch.randn(n, cout, generator=g).to(DEV)
    ad = torch.randn(n, h, h, cout, generator=g).to(torch.bfloat16).to(DEV) if addend else None
    outs = []
    lib.dmme_debug_force_splitk_4x4(*shape)
    try:
        for cluster in (1, 0):
            lib.dmme_set_conv_splitk_cluster(cluster)
            d = ops.make_conv_desc(s0, s1, cout, 3, 1, False, s0 if res else None, s1 if res else None, False, L.OUT_NHWC,
                                   torch.bfloat16, L.CONV_AUTO)
            ws = torch.empty(ops.conv_splitk_workspace(d) // 4, dtype=torch.float32, device=DEV)
            out = torch.full((n, h, h, cout), float("nan"), dtype=torch.bfloat16, device=DEV)
            y0, y1 = torch.full_like(out, float("nan")), torch.full_like(out, float("nan"))
            st = torch.zeros(n * (cout // 4) * 2, dtype=torch.int64, device=DEV)
            ops.conv2d_launch(d, wp, bias, out, temb, ad, stats=st, splitk_ws=ws,
                              out_norms=[ops.out_norm(y0, gamma, beta, cpg, True), ops.out_norm(y1, beta, gamma, 32, False)])
            torch.cuda.synchronize()
            outs.append((out, st, y0, y1))
    finally:
        lib.dmme_set_conv_splitk_cluster(1)
        lib.dmme_debug_force_splitk_4x4(0, 0)
    (oa, sa, ya0, ya1), (ob, sb, yb0, yb1) = outs
    assert not torch.isnan(oa.float()).any() and not torch.isnan(ya0.float()).any() and not torch.isnan(ya1.float()).any()
    assert torch.equal(oa, ob)
    assert torch.allclose(sa.double() / 2 ** 20, sb.double() / 2 ** 20, rtol=1e-5, atol=1e-2)
    assert rel_l2(ya0.float().cpu(), yb0.float().cpu()) < 2e-3
    assert rel_l2(ya1.float().cpu(), yb1.float().cpu()) < 2e-3


def test_conv_splitk_default_plan_matches_unsplit():
    """the cost model's own choice at the strong-scaling shard batches: wherever it splits, the result equals the unsplit
    kernels' (same operands, fp32 accumulation in a different order) and out_norm is refused without the workspace"""
    ops, L = _ops()
    lib = L.load()
    g = torch.Generator().manual_seed(93)
    split_seen = 0
    for n, c, h in ((32, 256, 4), (32, 256, 8), (64, 256, 4), (128, 256, 8), (256, 256, 4), (256, 256, 8)):
        x = torch.randn(n, h, h, c, generator=g).to(torch.bfloat16).to(DEV)
        w = torch.randn(c, c, 3, 3, generator=g) / math.sqrt(9 * c)
        wp = ops.pack_conv_weight(w.to(DEV), None, True)
        bias = torch.randn(c, generator=g).to(DEV)
        d = ops.make_conv_desc(x, None, c, 3, 1, False, None, None, False, L.OUT_NHWC, torch.bfloat16, L.CONV_AUTO)
        ws_bytes = ops.conv_splitk_workspace(d)
        a = torch.empty((n, h, h, c), dtype=torch.bfloat16, device=DEV)
        ops.conv2d_launch(d, wp, bias, a)  # no workspace: the unsplit kernels
        if not ws_bytes:
            continue
        split_seen += 1
        ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=DEV)
        b = torch.empty_like(a)
        d2 = ops.make_conv_desc(x, None, c, 3, 1, False, None, None, False, L.OUT_NHWC, torch.bfloat16, L.CONV_AUTO)
        ops.conv2d_launch(d2, wp, bias, b, splitk_ws=ws)
        torch.cuda.synchronize()
        assert rel_l2(b.float().cpu(), a.float().cpu()) < 3e-3
    assert split_seen >= 3
    # out_norm without a kernel that honours it (a 16x16 map: neither split-K nor the 8x8 epilogue norm) is refused
    x16 = torch.randn(8, 16, 16, c, generator=g).to(torch.bfloat16).to(DEV)
    a16 = torch.empty((8, 16, 16, c), dtype=torch.bfloat16, device=DEV)
    y = torch.empty_like(a16)
    d3 = ops.make_conv_desc(x16, None, c, 3, 1, False, None, None, False, L.OUT_NHWC, torch.bfloat16, L.CONV_AUTO)
    assert not ops.conv_splitk_workspace(d3) and not ops.conv_epilogue_norm(d3)
    with pytest.raises(RuntimeError, match="split-K"):
        ops.conv2d_launch(d3, wp, bias, a16, out_norms=[ops.out_norm(y, bias, bias, 8, True)])


def test_pack_weights_batch_bit_equal():
    """dmme_pack_conv_weights_batch (one launch for every tensor-core weight pack of a model, the head of the graph-captured
    training step) == the per-tensor pack kernels, bit for bit: 3x3 and 1x1, fused residual rows, data-gradient slices,
    shapes whose 2048-element blocks cut rows mid-way"""
    ops, L = _ops()
    g = torch.Generator().manual_seed(5)
    entries, want = [], []
    for cout, cin, ks, rc, dgrad, off, cnt in [(128, 128, 3, 0, 0, 0, 0), (256, 512, 3, 512, 0, 0, 0), (128, 256, 3, 256, 0, 0, 0),
                                               (256, 256, 3, 0, 1, 0, 256), (256, 512, 3, 0, 1, 256, 256), (128, 384, 3, 0, 1, 128, 256),
                                               (6, 128, 3, 0, 0, 0, 0), (128, 3, 3, 0, 1, 0, 3), (768, 256, 1, 0, 0, 0, 0),
                                               (256, 256, 1, 0, 1, 0, 256), (192, 64, 3, 64, 0, 0, 0), (64, 192, 3, 0, 1, 64, 128)]:
        w = torch.randn(cout, cin, ks, ks, generator=g).to(DEV)
        wres = torch.randn(cout, rc, 1, 1, generator=g).to(DEV) if rc else None
        ref = ops.pack_conv_weight_dgrad(w, off, cnt, True) if dgrad else ops.pack_conv_weight(w, wres, True)
        entries.append((w, wres, torch.full_like(ref, float("nan")), dgrad, off, cnt))
        want.append(ref)
    ops.PackBatch(entries, torch.device(DEV)).launch()
    torch.cuda.synchronize()
    for (w, wres, packed, dgrad, off, cnt), ref in zip(entries, want):
        assert torch.equal(packed.view(torch.int16), ref.view(torch.int16)), (tuple(w.shape), wres is not None, dgrad, off, cnt)
