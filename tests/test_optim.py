"""Fused optimizer tail: dmme_b200.optim.FusedAdamEMA against the reference's composition of
clip_grad_norm_ + torch.optim.Adam + WarmupLR + EMA (restated in oracle/dmme_oracle.py)."""
import warnings

import pytest
import torch

import dmme_oracle as O


def test_warmup_lr_matches_reference_scheduler():
    """the learning rate used by the k-th optimizer step, bit for bit"""
    from dmme_b200.optim import warmup_lr
    p = [torch.nn.Parameter(torch.zeros(3))]
    grads = [[torch.ones(3)] for _ in range(12)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, _, lrs, _ = O.optimizer_tail_reference(p, grads, lr=2e-4, warmup=5, max_norm=0, decay=0.9)
    for k, want in enumerate(lrs, start=1):
        assert warmup_lr(2e-4, k, 5) == want, (k, want)
    assert warmup_lr(1e-4, 1, 0) == 1e-4


SHAPES = [(128, 3, 3, 3), (128,), (256, 128, 3, 3), (1,), (4097,), (512, 512), (768, 256, 1, 1), (5, 7)]


@pytest.mark.gpu
@pytest.mark.parametrize("grad_scale,max_norm", [(1.0, 1.0), (1e-4, 1.0), (1.0, None)])
def test_fused_adam_ema_matches_torch(grad_scale, max_norm):
    from dmme_b200.optim import FusedAdamEMA
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(7)
    init = [torch.randn(s, generator=g) for s in SHAPES]
    steps = 4
    grads = [[torch.randn(s, generator=g) * grad_scale for s in SHAPES] for _ in range(steps)]
    # reference on the GPU with torch's own kernels
    ref_p = [torch.nn.Parameter(t.clone().to(dev)) for t in init]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref_ema, ref_opt, _, ref_norms = O.optimizer_tail_reference(ref_p, [[x.to(dev) for x in gs] for gs in grads], lr=2e-4,
                                                                   warmup=3, max_norm=max_norm or 0, decay=0.999)
    ours = [torch.nn.Parameter(t.clone().to(dev)) for t in init]
    opt = FusedAdamEMA(ours, lr=2e-4, warmup=3, max_grad_norm=max_norm, ema_decay=0.999)
    for k, gs in enumerate(grads):
        for p, x in zip(ours, gs):
            p.grad = x.clone().to(dev)
        opt.step()
        if max_norm:
            assert torch.allclose(opt.grad_norm.cpu(), ref_norms[k].cpu().reshape(1), rtol=1e-5)
    torch.cuda.synchronize()

    def close(a, b, what):
        err = float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
        assert err < 2e-6, f"{what}: rel-L2 {err}"

    for i, (p, q) in enumerate(zip(ours, ref_p)):
        close(p.detach(), q.detach(), f"weight {i}")
        st = ref_opt.state[q]
        close(opt.exp_avg[i], st["exp_avg"], f"exp_avg {i}")
        close(opt.exp_avg_sq[i], st["exp_avg_sq"], f"exp_avg_sq {i}")
        close(opt.ema[i], ref_ema[i], f"ema {i}")
        assert getattr(p, "_dmme_gen", 0) == steps


@pytest.mark.gpu
def test_fused_adam_ema_skips_missing_grads_and_swaps():
    from dmme_b200.optim import FusedAdamEMA
    dev = torch.device("cuda")
    a = torch.nn.Parameter(torch.ones(10, device=dev))
    b = torch.nn.Parameter(torch.ones(5000, device=dev))
    opt = FusedAdamEMA([a, b], lr=0.1, max_grad_norm=None, ema_decay=0.5)
    b.grad = torch.ones_like(b)
    opt.step()
    torch.cuda.synchronize()
    assert torch.equal(a.detach().cpu(), torch.ones(10))           # no gradient: untouched
    assert torch.allclose(b.detach().cpu(), torch.full((5000,), 0.9), atol=1e-6)  # first Adam step moves by lr
    assert torch.allclose(opt.ema[1].cpu(), torch.full((5000,), 0.95), atol=1e-6)
    opt.swap_ema()
    assert torch.allclose(b.detach().cpu(), torch.full((5000,), 0.95), atol=1e-6)
    opt.swap_ema()
    assert torch.allclose(b.detach().cpu(), torch.full((5000,), 0.9), atol=1e-6)
    sd = opt.state_dict()
    opt2 = FusedAdamEMA([a, b], lr=0.1, max_grad_norm=None, ema_decay=0.5)
    opt2.load_state_dict(sd)
    assert opt2.step_count == 1 and torch.equal(opt2._m, opt._m)


@pytest.mark.gpu
def test_graphed_training_step_matches_eager():
    """forward + loss + backward replayed from a CUDA graph == the eager step on the same draws; with the fused
    optimizer the two training loops stay together"""
    from dmme_b200 import DDPM
    from dmme_b200.models.ddpm import UNet
    from dmme_b200.optim import FusedAdamEMA
    from dmme_b200.training import GraphedTrainingStep
    dev = torch.device("cuda")

    def build():
        torch.manual_seed(0)
        unet = UNet(pos_dim=32, emb_dim=64, channels_per_depth=(64, 128), num_blocks=1, attention_depths=(2,), dropout=0.0)
        dm = DDPM(unet, timesteps=50).to(dev).train()
        return dm, FusedAdamEMA([p for p in dm.parameters() if p.requires_grad], lr=1e-3, warmup=2, ema_decay=0.99)

    x = torch.rand(8, 3, 32, 32, device=dev) * 2 - 1
    # graph replays draw from the CUDA generator at other Philox offsets than eager calls do: fix the draws
    t_fixed = torch.randint(1, 50, (8,), device=dev)
    noise_fixed = torch.randn(8, 3, 32, 32, device=dev)
    eager, opt_e = build()
    graphed, opt_g = build()
    for dm in (eager, graphed):
        inner = dm.training_step
        dm.training_step = (lambda f: (lambda x_0: f(x_0, t=t_fixed, noise=noise_fixed)))(inner)
    step = GraphedTrainingStep(graphed, opt_g, warmup=1)
    le, lg = [], []
    for _ in range(3):
        for p in eager.parameters():
            p.grad = None
        loss = eager.training_step(x)
        loss.backward()
        opt_e.step()
        le.append(float(loss.detach()))
        lg.append(float(step(x).detach()))
    assert le == pytest.approx(lg, rel=1e-5), (le, lg)
    assert le[2] != le[0]  # the optimizer moved the weights
    for p, q in zip(eager.parameters(), graphed.parameters()):
        assert torch.allclose(p.detach(), q.detach(), rtol=1e-4, atol=1e-6)
