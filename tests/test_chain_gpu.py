"""The chain kernel (csrc/conv_chain.cu: consecutive low-resolution ResBlock convs in one persistent launch) against a
plain torch fp32 computation of the same op sequence on the SAME bf16-rounded operands: only the fp32 accumulation
order and the bf16 rounding of a value that lies on a rounding boundary can differ (raw outputs: rel-L2 <= 4e-3; a
normalised tensor additionally carries the statistics of those outputs: <= 6e-3)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from dmme_b200 import ops, _lib
    return ops, _lib


def nhwc(t):  # NCHW fp32 -> NHWC bf16 on the device
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(DEV)


def nchw(t):  # NHWC device tensor -> NCHW fp32 cpu
    return t.float().permute(0, 3, 1, 2).contiguous().cpu()


def bf(t):
    return t.to(torch.bfloat16).float()


def ref_norm(raw, gamma, beta, cpg, silu, eps, scale=None, shift=None):
    """GroupNorm over groups of cpg channels of the (rounded) raw tensor, NCHW fp32; then optional (1 + scale), shift."""
    n, c, h, w = raw.shape
    g = raw.view(n, c // cpg, cpg * h * w)
    mean = g.mean(dim=2, keepdim=True)
    var = g.var(dim=2, unbiased=False, keepdim=True)
    y = ((g - mean) * torch.rsqrt(var + eps)).view(n, c, h, w) * gamma.view(1, c, 1, 1) + beta.view(1, c, 1, 1)
    if scale is not None:
        y = y * (1 + scale.view(n, c, 1, 1)) + shift.view(n, c, 1, 1)
    if silu:
        y = F.silu(y)
    return bf(y)


def make_case(n, hw, seed, iddpm=False):
    """Two ResBlocks like the up path: block 0 = identity residual, block 1 = concat input (256 + 128) with a 1x1 residual
    conv; returns the tensors and the reference results."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, k=1.0: torch.randn(*s, generator=g) * k
    C = 256
    t = {}
    t["x_raw"] = bf(rn(n, C, hw, hw))
    t["skip_raw"] = bf(rn(n, 128, hw, hw))
    t["x_norm"] = bf(F.silu(rn(n, C, hw, hw)))          # what the producer's norm left for block 0
    t["skip_norm"] = bf(F.silu(rn(n, 128, hw, hw)))      # normalised skip half of block 1's concat
    t["temb"] = rn(n, 2 * C, k=0.3)
    for name, cin, rc in (("a0", C, 0), ("b0", C, 0), ("a1", C + 128, 0), ("b1", C, C + 128)):
        t["w_" + name] = rn(C, cin, 3, 3, k=(9 * cin) ** -0.5)
        t["bias_" + name] = rn(C, k=0.1)
        if rc:
            t["wr_" + name] = rn(C, rc, 1, 1, k=rc ** -0.5)
            t["br_" + name] = rn(C, k=0.1)
    for name, c in (("n_b0", C), ("n_a1", C + 128), ("n_b1", C), ("n_out", C), ("n_skip", 2 * C)):
        t["g_" + name] = 1 + rn(c, k=0.2)
        t["be_" + name] = rn(c, k=0.2)
    return t


def reference(t, n, hw, iddpm):
    C = 256
    eps = 1e-5
    conv = lambda a, w: F.conv2d(a, bf(w), padding=w.shape[2] // 2)
    sc0 = (t["temb"][:, :C], t["temb"][:, C:]) if iddpm else (None, None)
    r = {}
    # block 0: conv1 (+temb in the ddpm flavour) -> norm (scale/shift in the iddpm flavour) -> conv2 + x
    h = conv(t["x_norm"], t["w_a0"]) + t["bias_a0"].view(1, C, 1, 1)
    if not iddpm:
        h = h + t["temb"][:, :C].view(n, C, 1, 1)
    h = bf(h)
    a = ref_norm(h, t["g_n_b0"], t["be_n_b0"], 8, True, eps, *(sc0 if iddpm else (None, None)))
    o0 = bf(conv(a, t["w_b0"]) + t["bias_b0"].view(1, C, 1, 1) + t["x_raw"])
    r["out0"] = o0
    # consumers of block 0's output: block 1's concat norm (part 0 of 384 channels: 12 per group -> not a divisor of 32, so
    # the test's concat norm uses 24 groups of 16) and a skip reader (512-channel norm, 16 per group, no SiLU)
    r["o0_for_a1"] = ref_norm(o0, t["g_n_a1"][:C], t["be_n_a1"][:C], 16, True, eps)
    r["o0_for_skip"] = ref_norm(o0, t["g_n_skip"][C:], t["be_n_skip"][C:], 16, False, eps)
    # block 1: conv1 over cat(norm(o0), skip_norm)
    h = conv(torch.cat([r["o0_for_a1"], t["skip_norm"]], 1), t["w_a1"]) + t["bias_a1"].view(1, C, 1, 1)
    if not iddpm:
        h = h + t["temb"][:, C:].view(n, C, 1, 1)
    h = bf(h)
    a = ref_norm(h, t["g_n_b1"], t["be_n_b1"], 8, True, eps)
    res = F.conv2d(torch.cat([o0, t["skip_raw"]], 1), bf(t["wr_b1"]))
    o1 = bf(conv(a, t["w_b1"]) + res + (t["bias_b1"] + t["br_b1"]).view(1, C, 1, 1))
    r["out1"] = o1
    r["o1_norm"] = ref_norm(o1, t["g_n_out"], t["be_n_out"], 8, True, eps)
    return r


@pytest.mark.parametrize("hw,n,ipc", [(8, 5, 0), (8, 5, 2), (8, 2, 2), (4, 9, 0), (4, 9, 2), (4, 11, 5), (8, 160, 0), (4, 300, 0)])
@pytest.mark.parametrize("iddpm", [False, True])
def test_conv_chain_two_blocks(hw, n, ipc, iddpm):
    ops, L = _ops()
    if not L.load().dmme_has_experimental():
        pytest.skip("chain kernel is compiled only with -DDMME_EXPERIMENTAL (measured slower, DESIGN.md)")
    if iddpm and n > 20:
        pytest.skip("flavour covered at the small sizes")
    t = make_case(n, hw, seed=hw * 1000 + n)
    want = reference(t, n, hw, iddpm)
    C = 256
    d = {k: v.to(DEV) for k, v in t.items() if k[:2] in ("g_", "be", "bi", "br", "te")}
    x_raw, skip_raw = nhwc(t["x_raw"]), nhwc(t["skip_raw"])
    x_norm, skip_norm = nhwc(t["x_norm"]), nhwc(t["skip_norm"])
    keepalive = []  # chain_op only records raw pointers

    def pack(w, wr=None):
        keepalive.append(ops.pack_conv_weight(t[w].to(DEV), t[wr].to(DEV) if wr else None, True))
        return keepalive[-1]

    bias_b1 = (d["bias_b1"] + d["br_b1"]).contiguous()
    new = lambda c=C: torch.full((n, hw, hw, c), float("nan"), dtype=torch.bfloat16, device=DEV)
    out0, out1, o0_a1, o0_skip, o1_norm = new(), new(), new(), new(), new()
    stats0 = torch.zeros(n * (C // 4) * 2, dtype=torch.int64, device=DEV)
    temb = d["temb"]
    ss = (temb[:, :C], temb[:, C:]) if iddpm else (None, None)
    chain = [
        ops.chain_op(x_norm, None, pack("w_a0"), d["bias_a0"], temb=None if iddpm else temb[:, :C],
                     out_norms=[ops.out_norm(None, d["g_n_b0"], d["be_n_b0"], 8, True, 1e-5, *ss)], keep=0),
        ops.chain_op(None, None, pack("w_b0"), d["bias_b0"], c0=C, addend=x_raw, out=out0, stats=stats0,
                     out_norms=[ops.out_norm(o0_a1, d["g_n_a1"][:C], d["be_n_a1"][:C], 16, True),
                                ops.out_norm(o0_skip, d["g_n_skip"][C:], d["be_n_skip"][C:], 16, False)], keep=0),
        ops.chain_op(None, skip_norm, pack("w_a1"), d["bias_a1"], c0=C, temb=None if iddpm else temb[:, C:],
                     out_norms=[ops.out_norm(None, d["g_n_b1"], d["be_n_b1"], 8, True)], keep=0),
        ops.chain_op(None, None, pack("w_b1", "wr_b1"), bias_b1, c0=C, res0=out0, res1=skip_raw,
                     out=out1, out_norms=[ops.out_norm(o1_norm, d["g_n_out"], d["be_n_out"], 8, True)], keep=-1),
    ]
    L.load().dmme_set_conv_chain_ipc(ipc)
    try:
        ops.conv_chain(chain, n, hw, hw)
        torch.cuda.synchronize()
    finally:
        L.load().dmme_set_conv_chain_ipc(0)
    assert rel_l2(nchw(out0), want["out0"]) < 4e-3
    assert rel_l2(nchw(o0_a1), want["o0_for_a1"]) < 6e-3
    assert rel_l2(nchw(o0_skip), want["o0_for_skip"]) < 6e-3
    assert rel_l2(nchw(out1), want["out1"]) < 6e-3
    assert rel_l2(nchw(o1_norm), want["o1_norm"]) < 8e-3
    # micro-group statistics of the stored raw output (what a later stand-alone GroupNorm would consume)
    st = stats0.view(n, C // 4, 2).double().cpu() / float(1 << 20)
    o = nchw(out0).double().view(n, C // 4, 4 * hw * hw)
    assert torch.allclose(st[:, :, 0], o.sum(2), atol=2e-2, rtol=1e-4)
    assert torch.allclose(st[:, :, 1], (o * o).sum(2), atol=2e-2, rtol=1e-4)
