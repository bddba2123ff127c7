"""Whole-path parity on the GPU: UNet forward, sampler steps and trajectories against the CPU oracle
(oracle/dmme_oracle.py, itself pinned to the unmodified reference by tests/test_oracle.py).

Tolerances are BASELINE.json's: per-call UNet output rel-L2 <= 1e-2 in bf16 mode and <= 1e-4 in fp32
mode against the reference fp32; DDIM eta=0 trajectories from a fixed x_T within the same tolerance."""
import pytest
import torch

import dmme_oracle as O
from helpers import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda"

TINY = dict(in_channels=3, pos_dim=4, emb_dim=8, num_groups=2, channels_per_depth=(4, 8, 16, 32), num_blocks=3)
BF16_TOL = 1e-2
FP32_TOL = 1e-4


@pytest.fixture(autouse=True)
def _inference_mode():
    """these tests cover the inference executor; with gradients enabled UNet.forward records the training tape
    (covered by tests/test_backward_gpu.py)."""
    with torch.no_grad():
        yield


def _unet(flavour, seed=0, **kw):
    from dmme_b200.models import ddpm, iddpm
    torch.manual_seed(seed)
    m = (ddpm.UNet if flavour == "ddpm" else iddpm.UNet)(**kw).eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    return m.to(DEV), sd


@pytest.mark.parametrize("flavour", ["ddpm", "iddpm"])
@pytest.mark.parametrize("tshape", ["one", "per_sample"])
def test_tiny_unet_fp32(flavour, tshape):
    """the reference's own test fixture (tests/test_ddpm.py:8-15): odd channel counts, 2 groups."""
    m, sd = _unet(flavour, precision="fp32", **TINY)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 3, 32, 32, generator=g)
    t = torch.tensor([37]) if tshape == "one" else torch.tensor([1, 50, 99])
    want = O.unet_forward(sd, x, t, groups=2, flavour=flavour)
    got = m(x.to(DEV), t.to(DEV)).cpu()
    assert got.shape == want.shape
    assert rel_l2(got, want) < FP32_TOL


def test_tiny_iddpm_unet_64x64_fp32():
    """tests/test_iddpm.py:30 feeds 64x64 images: attention at L = 1024 and L = 256."""
    m, sd = _unet("iddpm", precision="fp32", **TINY)
    x = torch.randn(4, 3, 64, 64, generator=torch.Generator().manual_seed(2))
    t = torch.tensor([1, 1, 1, 1])
    want = O.unet_forward(sd, x, t, groups=2, flavour="iddpm")
    assert rel_l2(m(x.to(DEV), t.to(DEV)).cpu(), want) < FP32_TOL


def test_tiny_unet_dropout_masks_fp32():
    """training-mode channel dropout with injected masks (torch's Philox stream cannot be matched)."""
    m, sd = _unet("ddpm", precision="fp32", **TINY)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 32, 32, generator=g)
    t = torch.tensor([5, 60])
    masks = {}
    for name, blk in m.engine.resblocks():
        c = blk.conv2[-1].weight.shape[1]
        masks[name] = (torch.rand(2, c, generator=g) > 0.1).float() / 0.9
    want = O.unet_forward(sd, x, t, groups=2, dropout_masks=masks)
    got = m.forward_raw(x.to(DEV), t.to(DEV), {k: v.to(DEV) for k, v in masks.items()}).cpu()
    assert rel_l2(got, want) < FP32_TOL


def _c1_inputs(n=16):
    torch.manual_seed(1234)
    return torch.randn(256, 3, 32, 32)[:n].contiguous(), torch.tensor([500])


@pytest.mark.parametrize("t", [1, 500, 1000])
def test_default_ddpm_unet_bf16(t):
    """BASELINE config #1 inputs (batch 16, seed-0 weights, x from seed 1234) through the tcgen05 path."""
    m, sd = _unet("ddpm")
    x, _ = _c1_inputs(16)
    tt = torch.tensor([t])
    want = O.unet_forward(sd, x, tt)
    got = m(x.to(DEV), tt.to(DEV)).cpu()
    err = rel_l2(got, want)
    print(f"default ddpm unet bf16 t={t}: rel-L2 {err:.3e}")
    assert err < BF16_TOL


@pytest.mark.parametrize("flavour,size", [("ddpm", 64), ("iddpm", 16)])
def test_default_unet_bf16_other_resolutions(flavour, size):
    """the default UNets at 64x64 (SURVEY C5 scale: 64x64 maps take the transposed tcgen05 kernel, the output conv the FFMA
    kernel) and 16x16 (2x2 bottleneck): same tolerance as at 32x32"""
    m, sd = _unet(flavour)
    g = torch.Generator().manual_seed(77)
    x = torch.randn(2, 3, size, size, generator=g)
    tt = torch.tensor([300])
    want = O.unet_forward(sd, x, tt) if flavour == "ddpm" else O.unet_forward(sd, x, tt, flavour="iddpm")
    got = m(x.to(DEV), tt.to(DEV)).cpu()
    err = rel_l2(got, want)
    print(f"default {flavour} unet bf16 {size}x{size}: rel-L2 {err:.3e}")
    assert err < BF16_TOL


def test_default_ddpm_unet_fp32_mode():
    m, sd = _unet("ddpm", precision="fp32")
    x, tt = _c1_inputs(4)
    want = O.unet_forward(sd, x, tt)
    err = rel_l2(m(x.to(DEV), tt.to(DEV)).cpu(), want)
    print(f"default ddpm unet fp32 mode: rel-L2 {err:.3e}")
    assert err < FP32_TOL


def test_default_iddpm_unet_bf16_per_sample_t():
    """IDDPM flavour, t of shape (N,): the attention regrouping couples samples, so the whole batch is compared."""
    m, sd = _unet("iddpm")
    x, _ = _c1_inputs(8)
    t = torch.tensor([1, 17, 250, 500, 640, 800, 999, 3])
    want = O.unet_forward(sd, x, t, flavour="iddpm")
    err = rel_l2(m(x.to(DEV), t.to(DEV)).cpu(), want)
    print(f"default iddpm unet bf16: rel-L2 {err:.3e}")
    assert err < BF16_TOL


def test_ddpm_sampling_step_injected_noise():
    from dmme_b200 import DDPM
    m, sd = _unet("ddpm", precision="fp32", **TINY)
    d = DDPM(m, timesteps=100).to(DEV)
    tabs = O.linear_tables(100)
    g = torch.Generator().manual_seed(4)
    x, z = torch.randn(3, 3, 32, 32, generator=g), torch.randn(3, 3, 32, 32, generator=g)
    for t in (100, 1):
        tt = torch.tensor([t])
        want = O.ddpm_step(x, tt, O.unet_forward(sd, x, tt, groups=2), z, tabs)
        got = d.sampling_step(x.to(DEV), tt.to(DEV), noise=z.to(DEV)).cpu()
        assert rel_l2(got, want) < FP32_TOL
    with pytest.raises(ValueError):
        d.sampling_step(x.to(DEV), torch.tensor([3, 4, 5]).to(DEV))


def test_iddpm_sampling_step_injected_noise():
    from dmme_b200 import IDDPM
    m, sd = _unet("iddpm", precision="fp32", **TINY)
    d = IDDPM(m, timesteps=100).to(DEV)
    tabs = O.cosine_tables(100)
    g = torch.Generator().manual_seed(5)
    x, z = torch.randn(3, 3, 32, 32, generator=g), torch.randn(3, 3, 32, 32, generator=g)
    tt = torch.tensor([57])
    want = O.iddpm_step(x, tt, O.unet_forward(sd, x, tt, groups=2, flavour="iddpm"), z, tabs)
    got = d.sampling_step(x.to(DEV), tt.to(DEV), noise=z.to(DEV)).cpu()
    assert rel_l2(got, want) < FP32_TOL


def test_ddim_trajectory_tiny_fp32():
    from dmme_b200 import DDIM
    m, sd = _unet("ddpm", precision="fp32", **TINY)
    d = DDIM(m, timesteps=100, sub_timesteps=5).to(DEV)
    x_T = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(6))
    _, _, ab = O.linear_tables(100)
    want = O.ddim_generate(sd, x_T, ab, O.tau_table(100, 5), groups=2)
    for graph in (False, True):
        got = d.generate(x_T.shape, x_T=x_T.to(DEV), graph=graph).cpu()
        assert rel_l2(got, want) < FP32_TOL


def test_ddim_trajectory_default_bf16():
    """BASELINE config #3: 50-step deterministic DDIM (quadratic tau) from a fixed x_T, whole trajectory."""
    from dmme_b200 import DDIM
    m, sd = _unet("ddpm")
    d = DDIM(m).to(DEV)
    x_T, _ = _c1_inputs(4)
    _, _, ab = O.linear_tables(1000)
    want, traj = O.ddim_generate(sd, x_T, ab, O.tau_table(1000, 50), return_trajectory=True)
    seen = []
    got = d.generate(x_T.shape, x_T=x_T.to(DEV), on_step=lambda k, x: seen.append(x.cpu().clone())).cpu()
    errs = [rel_l2(a, b) for a, b in zip(seen, traj)]
    print("ddim trajectory rel-L2: first %.3e  max %.3e  last %.3e" % (errs[0], max(errs), errs[-1]))
    assert len(seen) == 50 and max(errs) < BF16_TOL
    assert rel_l2(got, want) < BF16_TOL


def test_ddpm_generate_graph_equals_eager_and_is_seeded():
    from dmme_b200 import DDPM
    m, _ = _unet("ddpm", **TINY)
    d = DDPM(m, timesteps=20).to(DEV)
    x_T = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(7)).to(DEV)
    a = d.generate(x_T.shape, x_T=x_T, seed=99, graph=True)
    b = d.generate(x_T.shape, x_T=x_T, seed=99, graph=False)
    c = d.generate(x_T.shape, x_T=x_T, seed=100, graph=True)
    assert torch.equal(a, b)
    assert not torch.equal(a, c)
    assert torch.isfinite(a).all()
    out = d.generate((2, 3, 32, 32))  # reference smoke test: tests/test_ddpm.py:45-60
    assert out.shape == (2, 3, 32, 32)


# ---------------------------------------------------------------------------------------------------------------------
# the configurations bench.py times (BASELINE config #2: 256 images split over 1 / 2 / 4 / 8 GPUs)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("batch", [256, 128, 64, 32])
def test_default_ddpm_unet_bf16_timed_batches(batch):
    """per-GPU batches of the strong-scaling bench: shard r of n owns images [r * 256 / n, (r + 1) * 256 / n) of the
    seed-1234 x_T (SURVEY par. 8d).  The conv kernels pick their tilings from the batch (rows per halo tile by wave fill,
    pixel-tile width by unit count), so every timed batch gets its own whole-UNet parity check -- here on the LAST shard."""
    m, sd = _unet("ddpm")
    x_all, _ = _c1_inputs(256)
    world = 256 // batch
    x = x_all[(world - 1) * batch:].contiguous()
    tt = torch.tensor([1000])
    want = O.unet_forward(sd, x, tt)
    got = m(x.to(DEV), tt.to(DEV)).cpu()
    err = rel_l2(got, want)
    worst = max(rel_l2(got[i], want[i]) for i in range(batch))
    print(f"default ddpm unet bf16 batch {batch}: rel-L2 {err:.3e}, worst image {worst:.3e}")
    assert err < BF16_TOL and worst < BF16_TOL


def _philox_noises(ops, shape, seed, ts, noise_offset=0):
    """the normals ddpm_step_kernel draws in-kernel at step t (stream id = t), reproduced for the oracle"""
    return [ops.philox_normal(shape, seed, t, DEV, noise_offset=noise_offset).cpu() for t in ts]


def test_ddpm_ancestral_chain_default_bf16():
    """BASELINE config #2, multi-step: the first 20 steps (t = 1000 .. 981) of the graph-replayed ancestral chain with
    in-kernel Philox noise against the oracle's DDPM.generate fed the same noise; rel-L2 <= 1e-2 after EVERY step."""
    from dmme_b200 import DDPM, ops
    m, sd = _unet("ddpm")
    d = DDPM(m).to(DEV)
    x_T, _ = _c1_inputs(4)
    seed, steps = 4242, 20
    ts = list(range(1000, 1000 - steps, -1))
    noises = _philox_noises(ops, x_T.shape, seed, ts)
    _, traj = O.ddpm_generate(sd, x_T, noises, O.linear_tables(1000), 1000, steps=steps, return_trajectory=True)
    seen = []
    with torch.cuda.device(0):
        d._run_steps(x_T.to(DEV).clone(), steps, seed, True, lambda k, x: seen.append(x.cpu().clone()))
    errs = [rel_l2(a, b) for a, b in zip(seen, traj)]
    print("ddpm ancestral chain (T=1000, first 20 steps) rel-L2: first %.3e  max %.3e  last %.3e" % (errs[0], max(errs), errs[-1]))
    assert len(seen) == steps and max(errs) < BF16_TOL


def test_ddpm_ancestral_chain_full_short_schedule_bf16():
    """a COMPLETE ancestral chain (T = 20: every step down to t = 1, whose noise is drawn and discarded as in
    diffusion_models/ddpm.py:107-110) through DDPM.generate's CUDA-graph path, against the oracle on the same noise."""
    from dmme_b200 import DDPM, ops
    m, sd = _unet("ddpm")
    d = DDPM(m, timesteps=20).to(DEV)
    x_T, _ = _c1_inputs(4)
    seed = 777
    ts = list(range(20, 0, -1))
    noises = _philox_noises(ops, x_T.shape, seed, ts)
    want, traj = O.ddpm_generate(sd, x_T, noises, O.linear_tables(20), 20, return_trajectory=True)
    seen = []
    got = d.generate(x_T.shape, x_T=x_T.to(DEV), seed=seed, on_step=lambda k, x: seen.append(x.cpu().clone())).cpu()
    errs = [rel_l2(a, b) for a, b in zip(seen, traj)]
    print("ddpm ancestral chain (T=20, complete) rel-L2: first %.3e  max %.3e  last %.3e" % (errs[0], max(errs), errs[-1]))
    assert len(seen) == 20 and max(errs) < BF16_TOL
    assert rel_l2(got, want) < BF16_TOL


def test_ddpm_graph_step_timed_batch_matches_oracle_subset():
    """the exact object bench.py replays -- DDPM._graph_step captured at batch 256 -- for three steps; DDPM's UNet has no
    cross-sample coupling (SURVEY par. 8e), so the oracle runs on a 16-image subset (first / last images and the images
    around the halo kernel's tile boundaries) with the Philox noise of those images."""
    from dmme_b200 import DDPM, ops
    m, sd = _unet("ddpm")
    d = DDPM(m).to(DEV)
    x_T, _ = _c1_inputs(256)
    seed, steps = 99, 3
    pick = torch.tensor([0, 1, 2, 3, 36, 37, 73, 74, 127, 128, 129, 200, 252, 253, 254, 255])
    ts = list(range(1000, 1000 - steps, -1))
    noises = [z[pick] for z in _philox_noises(ops, x_T.shape, seed, ts)]
    _, traj = O.ddpm_generate(sd, x_T[pick], noises, O.linear_tables(1000), 1000, steps=steps, return_trajectory=True)
    seen = []
    with torch.cuda.device(0):
        d._run_steps(x_T.to(DEV).clone(), steps, seed, True, lambda k, x: seen.append(x[pick.to(DEV)].cpu().clone()))
    errs = [rel_l2(a, b) for a, b in zip(seen, traj)]
    print("ddpm graph step at batch 256, 16-image subset, rel-L2 per step:", ["%.3e" % e for e in errs])
    assert max(errs) < BF16_TOL


@pytest.mark.parametrize("kind", ["ddpm_philox", "ddpm_injected", "ddpm_last_step", "ddim", "iddpm_philox", "iddpm_injected"])
def test_sampler_update_in_output_conv_epilogue_is_bit_exact(kind):
    """the sampler update applied by the output conv's epilogue (eps / v still in registers, x_t updated in place, noise
    drawn in the epilogue) == output conv writing eps + the stand-alone dmme_*_step kernel, bit for bit"""
    from dmme_b200 import DDIM, DDPM, IDDPM
    flavour = "iddpm" if kind.startswith("iddpm") else "ddpm"
    m, _ = _unet(flavour)
    d = {"ddpm": DDPM, "ddim": DDIM, "iddpm": IDDPM}[kind.split("_")[0]](m).to(DEV)
    g = torch.Generator().manual_seed(31)
    x0 = torch.randn(5, 3, 32, 32, generator=g).to(DEV)
    z = torch.randn(5, 3, 32, 32, generator=g).to(DEV) if kind.endswith("injected") else None
    if kind == "ddim":
        i = torch.tensor([37], device=DEV)
        t_model = d.tau[i].clone()
        t = i
    else:
        t = torch.tensor([1 if kind == "ddpm_last_step" else 640], device=DEV)
        t_model = t
    outs = []
    with torch.cuda.device(0):
        for fuse in (True, False):
            m.engine.fuse_sampler = fuse
            x = x0.clone()
            d._denoise_(x, t_model, t, z, 1234)
            assert m.engine.sampler_applied == fuse
            outs.append(x.clone())
    m.engine.fuse_sampler = True
    assert torch.isfinite(outs[0]).all()
    assert not torch.equal(outs[0], x0)
    assert torch.equal(outs[0], outs[1])


LSUN = dict(dropout=0.0, channels_per_depth=(128, 128, 256, 256, 512, 512), attention_depths=(5,))


def test_lsun_unet_bf16_256x256():
    """SURVEY par. 8f-3: the LSUN-256 UNet of configs/ddpm/lsun_bedroom.yaml:78-90 (6 depths, 128..512 channels, 256x256
    maps down to 8x8, attention at depth 5, dropout 0 => `conv2.2` keys) at that config's batch 2, against the oracle"""
    m, sd = _unet("ddpm", **LSUN)
    assert "down_layers.0.conv2.2.weight" in sd and sum(v.numel() for k, v in sd.items() if "embeddings" not in k) > 90e6
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 256, 256, generator=g)
    tt = torch.tensor([777])
    want = O.unet_forward(sd, x, tt)
    got = m(x.to(DEV), tt.to(DEV)).cpu()
    err = rel_l2(got, want)
    print(f"lsun unet bf16 256x256 batch 2: rel-L2 {err:.3e}")
    assert err < BF16_TOL


def test_lsun_ddpm_sampling_steps_256x256():
    """three graph-replayed ancestral steps of the LSUN-256 model (T = 1000 .. 998) against the oracle on the same noise"""
    from dmme_b200 import DDPM, ops
    m, sd = _unet("ddpm", **LSUN)
    d = DDPM(m).to(DEV)
    x_T = torch.randn(1, 3, 256, 256, generator=torch.Generator().manual_seed(6))
    seed, steps = 11, 3
    noises = _philox_noises(ops, x_T.shape, seed, [1000, 999, 998])
    _, traj = O.ddpm_generate(sd, x_T, noises, O.linear_tables(1000), 1000, steps=steps, return_trajectory=True)
    seen = []
    with torch.cuda.device(0):
        d._run_steps(x_T.to(DEV).clone(), steps, seed, True, lambda k, x: seen.append(x.cpu().clone()))
    errs = [rel_l2(a, b) for a, b in zip(seen, traj)]
    print("lsun ddpm steps rel-L2:", ["%.3e" % e for e in errs])
    assert max(errs) < BF16_TOL
