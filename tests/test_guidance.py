"""Classifier guidance (SURVEY par. 8f-4, src/dmme/guidance/classifier.py:8-63).  The reference module cannot be imported
(its own tests/test_guidance.py fails at collection), so these tests pin dmme_b200.guidance to the oracle restatement of that
file -- PARITY UNPINNED against a run of the reference.  CPU: the reference test's own toy model / classifier
(tests/test_guidance.py:41-72).  GPU: the gradient through a dmme_b200 UNet classifier (explicit CUDA backward kernels)."""
import pytest
import torch
from torch import nn

import dmme_oracle as O
from helpers import rel_l2

NUM_CLASSES, BATCH, T = 10, 8, 10


class Model(nn.Module):
    """tests/test_guidance.py:41-58 of the reference (einops Rearrange layers written as reshapes)"""

    def __init__(self) -> None:
        super().__init__()
        self.conv1 = nn.Conv2d(3, 4, 3, 1, 1)
        self.act = nn.SiLU()
        self.conv2 = nn.Conv2d(4, 3, 3, 1, 1)
        self.linear = nn.Linear(1, 4)

    def forward(self, x, t):
        x = self.conv1(x)
        x = x + self.linear(t.float().reshape(-1, 1)).reshape(-1, 4, 1, 1)
        return self.conv2(x)


class Classifier(Model):
    def __init__(self) -> None:
        super().__init__()
        self.fc = nn.Sequential(nn.Flatten(), nn.Linear(32 * 32 * 3, NUM_CLASSES))

    def forward(self, x, t):
        return self.fc(super().forward(x, t))


def _inputs(seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randint(0, NUM_CLASSES, (BATCH,), generator=g), torch.randn(BATCH, 3, 32, 32, generator=g),
            torch.randint(1, T, (BATCH,), generator=g), torch.randn(BATCH, 3, 32, 32, generator=g))


def test_classifier_guided_ddpm_and_ddim_match_the_restatement_cpu():
    from dmme_b200.guidance import ClassifierGuidedDDIM, ClassifierGuidedDDPM
    torch.manual_seed(1)
    model, classifier = Model(), Classifier()
    y, x_t, t, noise = _inputs()
    tabs = O.linear_tables(T)
    got = ClassifierGuidedDDPM(timesteps=T).sample(model, classifier, y, x_t.clone(), t, noise)
    assert got.size() == x_t.size()  # the reference's own (shape-only) assertion, tests/test_guidance.py:75-81
    assert torch.allclose(got, O.guided_ddpm_sample(model, classifier, y, x_t, t, noise, tabs), rtol=1e-5, atol=1e-6)
    got = ClassifierGuidedDDIM(timesteps=T).sample(model, classifier, y, x_t.clone(), t)
    assert got.size() == x_t.size()
    assert torch.allclose(got, O.guided_ddim_sample(model, classifier, y, x_t, t, tabs[2]), rtol=1e-5, atol=1e-6)
    with pytest.raises(NotImplementedError):
        ClassifierGuidedDDIM(timesteps=T, tau_schedule="cubic")


class _UNetClassifier(nn.Module):
    """a classifier whose feature extractor is a dmme_b200 UNet: its input gradient comes from the CUDA backward kernels"""

    def __init__(self, unet, head):
        super().__init__()
        self.unet, self.head = unet, head

    def forward(self, x, t):
        return self.head(self.unet(x, t.long()).flatten(1))


class _OracleClassifier:
    def __init__(self, sd, head, groups):
        self.sd, self.head, self.groups = sd, head, groups

    def __call__(self, x, t):
        return self.head(O.unet_forward(self.sd, x, t.long(), groups=self.groups).flatten(1))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["ddpm", "ddim"])
def test_classifier_guidance_through_cuda_backward_kernels(kind):
    from dmme_b200.guidance import ClassifierGuidedDDIM, ClassifierGuidedDDPM
    from dmme_b200.models.ddpm import UNet
    tiny = dict(in_channels=3, pos_dim=4, emb_dim=8, num_groups=2, channels_per_depth=(4, 8, 16, 32), num_blocks=3)
    torch.manual_seed(2)
    eps_net = UNet(precision="fp32", dropout=0.0, **tiny).eval()
    cls_net = UNet(precision="fp32", dropout=0.0, **tiny)
    head = nn.Linear(3 * 32 * 32, NUM_CLASSES)
    sd_eps = {k: v.clone() for k, v in eps_net.state_dict().items()}
    sd_cls = {k: v.clone() for k, v in cls_net.state_dict().items()}
    head_cpu = nn.Linear(3 * 32 * 32, NUM_CLASSES)
    head_cpu.load_state_dict(head.state_dict())
    y, x_t, t, noise = _inputs(3)
    tabs = O.linear_tables(T)
    model_cpu = lambda x, tt: O.unet_forward(sd_eps, x, tt, groups=2)  # noqa: E731
    cls_cpu = _OracleClassifier(sd_cls, head_cpu, 2)
    dev = "cuda"
    model_gpu = eps_net.to(dev)
    cls_gpu = _UNetClassifier(cls_net.to(dev), head.to(dev))
    if kind == "ddpm":
        want = O.guided_ddpm_sample(model_cpu, cls_cpu, y, x_t, t, noise, tabs)
        got = ClassifierGuidedDDPM(timesteps=T).sample(model_gpu, cls_gpu, y.to(dev), x_t.to(dev), t.to(dev), noise.to(dev))
    else:
        want = O.guided_ddim_sample(model_cpu, cls_cpu, y, x_t, t, tabs[2])
        got = ClassifierGuidedDDIM(timesteps=T).sample(model_gpu, cls_gpu, y.to(dev), x_t.to(dev), t.to(dev))
    # the guidance term itself, not just the (dominant) sampler update
    g_want = O.classifier_grad(cls_cpu, y, x_t, t)
    g_got = ClassifierGuidedDDPM(timesteps=T).classifier_grad(cls_gpu, y.to(dev), x_t.to(dev), t.to(dev))
    assert rel_l2(g_got, g_want) < 1e-3, rel_l2(g_got, g_want)
    assert rel_l2(got, want) < 1e-4, rel_l2(got, want)
