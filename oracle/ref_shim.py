"""Import the UNMODIFIED reference (dmme 0.5.2) from its read-only checkout -- TEST INFRASTRUCTURE ONLY.

pytorch_lightning / torchmetrics are not installed in this image; the reference's top-level
``dmme/__init__.py`` imports them for its Lightning glue, which is outside the hot path.  Inert
stand-ins for exactly those names are registered in ``sys.modules`` so that ``dmme.models``,
``dmme.equations`` and ``dmme.diffusion_models`` import and run unmodified.  Nothing is copied into
the repository's history: the reference is used where it lies -- its read-only checkout (default
/root/reference, override with DMME_REFERENCE_ROOT) or, where that does not exist (the GPU box), the
git-ignored offline install ``baseline/_ref`` made by ``baseline/install_reference.sh``
(``pip install --no-index --no-deps --target baseline/_ref``).  Use ``available()`` to skip.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("DMME_REFERENCE_ROOT", "/root/reference")
INSTALLED_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def source_dir():
    """Directory to put on sys.path so that ``import dmme`` finds the unmodified reference, or None."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "dmme")):
        return os.path.join(REFERENCE_ROOT, "src")
    if os.path.isdir(os.path.join(INSTALLED_ROOT, "dmme")):
        return INSTALLED_ROOT
    return None


def available() -> bool:
    return source_dir() is not None


def _mod(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Inert:
    def __init__(self, *a, **k):
        pass


class _LightningModule(torch.nn.Module):
    pass


def load():
    """Returns the imported reference package ``dmme``."""
    if "dmme" in sys.modules and getattr(sys.modules["dmme"], "__version__", None):
        return sys.modules["dmme"]
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT} or {INSTALLED_ROOT}")
    if "pytorch_lightning" not in sys.modules:
        pl = _mod("pytorch_lightning", LightningModule=_LightningModule, LightningDataModule=_Inert,
                  Callback=_Inert, Trainer=_Inert)
        pl.loggers = _mod("pytorch_lightning.loggers", WandbLogger=_Inert, TensorBoardLogger=_Inert)
        pl.utilities = _mod("pytorch_lightning.utilities")
        pl.utilities.exceptions = _mod("pytorch_lightning.utilities.exceptions", MisconfigurationException=Exception)
        pl.cli = _mod("pytorch_lightning.cli", LightningCLI=_Inert)
    if "torchmetrics" not in sys.modules:
        tm = _mod("torchmetrics")
        tm.image = _mod("torchmetrics.image")
        tm.image.fid = _mod("torchmetrics.image.fid", FrechetInceptionDistance=_Inert)
        tm.image.inception = _mod("torchmetrics.image.inception", InceptionScore=_Inert)
    src = source_dir()
    if src not in sys.path:
        sys.path.insert(0, src)
    # DDIM's last step builds Normal(mean, std=0); the reference only runs with validation off
    torch.distributions.Distribution.set_default_validate_args(False)
    import dmme  # noqa: E402

    return dmme
