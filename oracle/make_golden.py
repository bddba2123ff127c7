"""Generate tests/golden/*.pt from the UNMODIFIED reference -- run in the build container only.

    python oracle/make_golden.py

For every fixture the script first asserts that the oracle restatement (oracle/dmme_oracle.py)
reproduces the live reference, then stores the reference's output.  Weights are never stored: they
are re-created from a seed (torch's CPU generator is deterministic), and their SHA-256 is stored so a
consumer can prove it rebuilt the same state_dict.  TEST INFRASTRUCTURE ONLY.
"""
import hashlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import dmme_oracle as O  # noqa: E402
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
TINY = dict(in_channels=3, pos_dim=4, emb_dim=8, num_groups=2, channels_per_depth=(4, 8, 16, 32), num_blocks=3)


def sd_digest(sd) -> str:
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main():
    dmme = ref_shim.load()
    from dmme.diffusion_models import DDIM, DDPM, IDDPM
    from dmme.models import iddpm as ref_iddpm
    from dmme.models.ddpm import UNet as RefUNet

    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)

    # ---- 1. schedule tables and tau: bit-exact ----------------------------------------------
    class _Null(torch.nn.Module):
        pass

    sched = {}
    d = DDPM(_Null(), 1000)
    sched["linear_1000"] = [d.beta.flatten().clone(), d.alpha.flatten().clone(), d.alpha_bar.flatten().clone()]
    d = IDDPM(_Null(), 4000, schedule="linear", start=0.000025, end=0.005)  # configs/iddpm/cifar10.yaml:78-81
    sched["linear_4000_iddpm_yaml"] = [d.beta.flatten().clone(), d.alpha.flatten().clone(), d.alpha_bar.flatten().clone()]
    d = IDDPM(_Null(), 1000)
    sched["cosine_1000"] = [d.beta.flatten().clone(), d.alpha.flatten().clone(), d.alpha_bar.flatten().clone()]
    sched["tau_quadratic_1000_50"] = DDIM(_Null(), 1000, 50, "quadratic").tau.clone()
    sched["tau_linear_1000_50"] = DDIM(_Null(), 1000, 50, "linear").tau.clone()
    for a, b in zip(sched["linear_1000"], O.linear_tables(1000)):
        assert torch.equal(a, b)
    for a, b in zip(sched["linear_4000_iddpm_yaml"], O.linear_tables(4000, 0.000025, 0.005)):
        assert torch.equal(a, b)
    for a, b in zip(sched["cosine_1000"], O.cosine_tables(1000)):
        assert torch.equal(a, b)
    assert torch.equal(sched["tau_quadratic_1000_50"], O.tau_table(1000, 50, "quadratic"))
    assert torch.equal(sched["tau_linear_1000_50"], O.tau_table(1000, 50, "linear"))
    torch.manual_seed(3)
    sched["randint_seed3_1_1000_128"] = dmme.uniform_int(1, 1000, 128)  # never draws t = T (quirk 2)
    torch.save(sched, os.path.join(OUT, "schedules.pt"))

    # ---- 2. tiny fixture model (tests/test_ddpm.py:8-15), both flavours ------------------------
    tiny = {}
    for flavour, cls in (("ddpm", RefUNet), ("iddpm", ref_iddpm.UNet)):
        torch.manual_seed(0)
        m = cls(**TINY).eval()
        sd = m.state_dict()
        g = torch.Generator().manual_seed(1)
        x = torch.randn(3, 3, 32, 32, generator=g)
        entry = {"digest": sd_digest(sd), "x": x}
        for name, t in (("t_one", torch.tensor([37])), ("t_per_sample", torch.tensor([1, 50, 99]))):
            y = m(x, t)
            assert rel(O.unet_forward(sd, x, t, groups=2, flavour=flavour), y) < 1e-6
            entry[name] = t
            entry["out_" + name] = y.clone()
        tiny[flavour] = entry
    torch.save(tiny, os.path.join(OUT, "tiny_unet.pt"))

    # ---- 3. default DDPM UNet at BASELINE config #1 inputs ------------------------------------
    torch.manual_seed(0)
    m = RefUNet().eval()
    sd = m.state_dict()
    torch.manual_seed(1234)
    x256 = torch.randn(256, 3, 32, 32)
    c1 = {"digest": sd_digest(sd), "x_seed": 1234}
    for t in (1, 500, 1000):
        tt = torch.tensor([t])
        y = m(x256[:16], tt)
        assert rel(O.unet_forward(sd, x256[:16], tt), y) < 1e-6
        c1[f"out_t{t}"] = y.clone()
    torch.save(c1, os.path.join(OUT, "default_ddpm_c1.pt"))

    # ---- 4. DDIM 50-step trajectory (config #3), 4 images --------------------------------------
    ddim = DDIM(m)  # quadratic tau, T = 1000, S = 50
    x_T = x256[:4].clone()
    snaps = {}
    x = x_T
    all_i = torch.arange(0, 51).unsqueeze(dim=1)
    for i in range(50, 0, -1):
        x = ddim.sampling_step(x, all_i[i])
        if i in (50, 25, 2, 1):
            snaps[i] = x.clone()
    want, traj = O.ddim_generate(sd, x_T, ddim.alpha_bar.flatten(), ddim.tau, return_trajectory=True)
    assert rel(want, x) < 1e-6 and rel(traj[0], snaps[50]) < 1e-6
    torch.save({"digest": c1["digest"], "n": 4, "snapshots": snaps}, os.path.join(OUT, "ddim_trajectory.pt"))

    # ---- 5. sampler steps with the reference's own RNG stream --------------------------------
    steps = {}
    torch.manual_seed(0)
    mt = RefUNet(**TINY).eval()
    ddpm = DDPM(mt, timesteps=100)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 3, 32, 32, generator=g)
    for t in (100, 1):
        tt = torch.tensor([t])
        torch.manual_seed(77)
        y = ddpm.sampling_step(x, tt)
        torch.manual_seed(77)
        z = torch.randn(x.shape)  # Normal.sample == randn * std + mean, same stream position
        mine = O.ddpm_step(x, tt, O.unet_forward(mt.state_dict(), x, tt, groups=2), z, O.linear_tables(100))
        assert rel(mine, y) < 1e-6, rel(mine, y)
        steps[f"ddpm_t{t}"] = {"x": x, "z": z, "out": y.clone()}
    torch.manual_seed(0)
    mi = ref_iddpm.UNet(**TINY).eval()
    iddpm = IDDPM(mi, timesteps=100)
    tt = torch.tensor([57])
    torch.manual_seed(78)
    y = iddpm.sampling_step(x, tt)
    torch.manual_seed(78)
    z = torch.randn(x.shape)
    mine = O.iddpm_step(x, tt, O.unet_forward(mi.state_dict(), x, tt, groups=2, flavour="iddpm"), z, O.cosine_tables(100))
    assert rel(mine, y) < 1e-6, rel(mine, y)
    steps["iddpm_t57"] = {"x": x, "z": z, "out": y.clone()}

    # ---- 6. training losses (dropout 0), reference RNG order: randint then normal --------------
    torch.manual_seed(0)
    m0 = RefUNet(dropout=0.0, **TINY).train()
    ddpm0 = DDPM(m0, timesteps=100)
    x0 = torch.randn(3, 3, 32, 32, generator=torch.Generator().manual_seed(5))
    torch.manual_seed(79)
    loss = ddpm0.training_step(x0)
    torch.manual_seed(79)
    t = torch.randint(1, 100, (3,))
    z = torch.randn(x0.shape)
    x_t, qm, qs = O.forward_noising(x0, t, z, O.linear_tables(100)[2])
    mine = O.ddpm_loss(x_t, qm, qs, O.unet_forward(m0.state_dict(), x_t, t, groups=2))
    assert abs(float(mine) - float(loss)) < 1e-6 * max(1.0, abs(float(loss))), (float(mine), float(loss))
    steps["ddpm_loss"] = {"x0": x0, "t": t, "z": z, "loss": loss.clone()}

    torch.manual_seed(0)
    mi0 = ref_iddpm.UNet(dropout=0.0, **TINY).train()
    for T, name in ((100, "hybrid_T100"), (2, "hybrid_T2")):  # T = 2 forces t == 1 rows (tests/test_iddpm.py:28)
        idd = IDDPM(mi0, timesteps=T)
        tabs = O.cosine_tables(T)
        x0 = torch.randn(4, 3, 32, 32, generator=torch.Generator().manual_seed(6)).clamp(-1, 1)
        torch.manual_seed(80)
        loss = idd.training_step(x0)
        torch.manual_seed(80)
        t = torch.randint(1, T, (4,))
        z = torch.randn(x0.shape)
        x_t, qm, qs = O.forward_noising(x0, t, z, tabs[2])
        out = O.unet_forward(mi0.state_dict(), x_t, t, groups=2, flavour="iddpm")
        eps, var = O.iddpm_split(out, t, tabs)
        mine = O.ddpm_loss(x_t, qm, qs, eps) + 0.001 * O.vlb_loss(eps, var, x_t, t, x0, tabs)
        assert abs(float(mine) - float(loss)) < 2e-6 * max(1.0, abs(float(loss))), (name, float(mine), float(loss))
        steps["iddpm_" + name] = {"x0": x0, "t": t, "z": z, "loss": loss.clone()}
    torch.save(steps, os.path.join(OUT, "steps_and_losses.pt"))

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
