"""The oracle against the committed golden vectors (generated from the unmodified reference by
oracle/make_golden.py) and, where the reference checkout exists, against the live reference.
Also pins the drop-in contract: dmme_b200's modules rebuild the reference's seeded weights exactly."""
import hashlib
import os

import pytest
import torch

import dmme_oracle as O
import ref_shim
from helpers import rel_l2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TINY = dict(in_channels=3, pos_dim=4, emb_dim=8, num_groups=2, channels_per_depth=(4, 8, 16, 32), num_blocks=3)


def gold(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def sd_digest(sd) -> str:
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def seeded(cls, seed=0, **kw):
    torch.manual_seed(seed)
    return cls(**kw)


def test_schedule_tables_bit_exact():
    g = gold("schedules.pt")
    for a, b in zip(O.linear_tables(1000), g["linear_1000"]):
        assert torch.equal(a, b)
    for a, b in zip(O.linear_tables(4000, 0.000025, 0.005), g["linear_4000_iddpm_yaml"]):
        assert torch.equal(a, b)
    for a, b in zip(O.cosine_tables(1000), g["cosine_1000"]):
        assert torch.equal(a, b)
    assert torch.equal(O.tau_table(1000, 50, "quadratic"), g["tau_quadratic_1000_50"])
    assert torch.equal(O.tau_table(1000, 50, "linear"), g["tau_linear_1000_50"])
    # reference facts the build relies on
    beta, alpha, alpha_bar = g["linear_1000"]
    assert beta[0] == 0 and alpha[0] == 1 and alpha_bar[0] == 1
    cb = g["cosine_1000"][0]
    assert cb[0] == 1 and int((cb[1:] == 0.999).sum()) == 1
    assert g["tau_quadratic_1000_50"][:2].tolist() == [0, 0]


def test_product_schedule_tables_bit_exact():
    """the buffers the CUDA samplers index are the reference's, bit for bit"""
    from dmme_b200 import DDIM, DDPM, IDDPM
    g = gold("schedules.pt")
    null = torch.nn.Identity()
    d = DDPM(null, 1000)
    for got, want in zip((d.beta, d.alpha, d.alpha_bar), g["linear_1000"]):
        assert torch.equal(got.flatten(), want)
    d = IDDPM(null, 4000, schedule="linear", start=0.000025, end=0.005)
    for got, want in zip((d.beta, d.alpha, d.alpha_bar), g["linear_4000_iddpm_yaml"]):
        assert torch.equal(got.flatten(), want)
    d = IDDPM(null, 1000)
    for got, want in zip((d.beta, d.alpha, d.alpha_bar), g["cosine_1000"]):
        assert torch.equal(got.flatten(), want)
    assert torch.equal(DDIM(null, 1000, 50, "quadratic").tau, g["tau_quadratic_1000_50"])
    assert torch.equal(DDIM(null, 1000, 50, "linear").tau, g["tau_linear_1000_50"])


def test_timestep_draws_bit_exact():
    import dmme_b200
    torch.manual_seed(3)
    t = dmme_b200.uniform_int(1, 1000, 128)
    assert torch.equal(t, gold("schedules.pt")["randint_seed3_1_1000_128"])
    assert int(t.max()) < 1000 and int(t.min()) >= 1


@pytest.mark.parametrize("flavour", ["ddpm", "iddpm"])
def test_tiny_unet_against_golden(flavour):
    from dmme_b200.models import ddpm, iddpm
    g = gold("tiny_unet.pt")[flavour]
    m = seeded(ddpm.UNet if flavour == "ddpm" else iddpm.UNet, **TINY)
    sd = m.state_dict()
    assert sd_digest(sd) == g["digest"], "dmme_b200 did not rebuild the reference's seeded weights"
    for name in ("t_one", "t_per_sample"):
        with torch.no_grad():
            y = O.unet_forward(sd, g["x"], g[name], groups=2, flavour=flavour)
        assert rel_l2(y, g["out_" + name]) < 1e-6


def test_default_unet_against_golden_c1():
    from dmme_b200.models.ddpm import UNet
    g = gold("default_ddpm_c1.pt")
    sd = seeded(UNet).state_dict()
    assert sd_digest(sd) == g["digest"]
    torch.manual_seed(g["x_seed"])
    x = torch.randn(256, 3, 32, 32)[:16]
    with torch.no_grad():
        y = O.unet_forward(sd, x, torch.tensor([500]))
    assert rel_l2(y, g["out_t500"]) < 1e-6


def test_sampler_steps_and_losses_against_golden():
    from dmme_b200.models import ddpm, iddpm
    g = gold("steps_and_losses.pt")
    sd = seeded(ddpm.UNet, **TINY).eval().state_dict()
    with torch.no_grad():
        for t in (100, 1):
            e = g[f"ddpm_t{t}"]
            tt = torch.tensor([t])
            y = O.ddpm_step(e["x"], tt, O.unet_forward(sd, e["x"], tt, groups=2), e["z"], O.linear_tables(100))
            assert rel_l2(y, e["out"]) < 1e-6
        sdi = seeded(iddpm.UNet, **TINY).eval().state_dict()
        e = g["iddpm_t57"]
        tt = torch.tensor([57])
        y = O.iddpm_step(e["x"], tt, O.unet_forward(sdi, e["x"], tt, groups=2, flavour="iddpm"), e["z"], O.cosine_tables(100))
        assert rel_l2(y, e["out"]) < 1e-6
        # losses (dropout 0)
        sd0 = seeded(ddpm.UNet, dropout=0.0, **TINY).state_dict()
        e = g["ddpm_loss"]
        x_t, qm, qs = O.forward_noising(e["x0"], e["t"], e["z"], O.linear_tables(100)[2])
        loss = O.ddpm_loss(x_t, qm, qs, O.unet_forward(sd0, x_t, e["t"], groups=2))
        assert abs(float(loss) - float(e["loss"])) < 1e-6 * max(1.0, abs(float(e["loss"])))
        sdi0 = seeded(iddpm.UNet, dropout=0.0, **TINY).state_dict()
        for T, name in ((100, "iddpm_hybrid_T100"), (2, "iddpm_hybrid_T2")):
            e = g[name]
            tabs = O.cosine_tables(T)
            x_t, qm, qs = O.forward_noising(e["x0"], e["t"], e["z"], tabs[2])
            eps, var = O.iddpm_split(O.unet_forward(sdi0, x_t, e["t"], groups=2, flavour="iddpm"), e["t"], tabs)
            loss = O.ddpm_loss(x_t, qm, qs, eps) + 0.001 * O.vlb_loss(eps, var, x_t, e["t"], e["x0"], tabs)
            assert abs(float(loss) - float(e["loss"])) < 2e-6 * max(1.0, abs(float(e["loss"])))


def test_ddim_trajectory_against_golden_last_steps():
    """cheap slice of config #3: re-run only the last two steps from the stored snapshot (i = 2 -> 1 -> done)"""
    from dmme_b200.models.ddpm import UNet
    g = gold("ddim_trajectory.pt")
    sd = seeded(UNet).state_dict()
    assert sd_digest(sd) == g["digest"]
    _, _, ab = O.linear_tables(1000)
    tau = O.tau_table(1000, 50)
    with torch.no_grad():
        x = g["snapshots"][2]
        ii = torch.tensor([1])
        y = O.ddim_step(x, ii, O.unet_forward(sd, x, tau[ii]), ab, tau)
    assert rel_l2(y, g["snapshots"][1]) < 1e-6
    # tau_1 = tau_0 = 0: the last step is an identity on x that still evaluates the UNet at t = 0
    assert rel_l2(g["snapshots"][1], g["snapshots"][2]) < 1e-6


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present (GPU box)")
def test_oracle_against_live_reference():
    ref_shim.load()
    from dmme.diffusion_models import DDPM as RefDDPM
    from dmme.models import iddpm as ref_iddpm
    from dmme.models.ddpm import UNet as RefUNet
    from dmme_b200.models import ddpm, iddpm
    for mine, theirs, flavour in ((ddpm.UNet, RefUNet, "ddpm"), (iddpm.UNet, ref_iddpm.UNet, "iddpm")):
        a = seeded(mine, **TINY).state_dict()
        m = seeded(theirs, **TINY).eval()
        b = m.state_dict()
        assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
        x = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(9))
        t = torch.tensor([3, 77])
        with torch.no_grad():
            assert rel_l2(O.unet_forward(a, x, t, groups=2, flavour=flavour), m(x, t)) < 1e-6
    # cross-loading: reference weights into the dmme_b200 module and back
    m = seeded(RefUNet, 5, **TINY)
    mine = ddpm.UNet(**TINY)
    mine.load_state_dict(m.state_dict())
    m.load_state_dict(mine.state_dict())
    d = RefDDPM(m, 100)
    assert torch.equal(d.beta.flatten(), O.linear_tables(100)[0])
