"""dmme_b200 -- B200-native (sm_100a) implementation of the hot path of diffusion-models-made-easy:
the UNet denoiser inside the DDPM / DDIM / IDDPM loops, behind the reference's Python interface."""
__version__ = "0.1.0"

from .common.noise import gaussian, gaussian_like, uniform_int, pad
from .common.norm import norm, denorm, denorm_uint8

from . import equations
from . import models
from . import diffusion_models
from . import optim
from . import training
from . import guidance
from .diffusion_models import DDPM, DDIM, IDDPM

__all__ = ["DDPM", "DDIM", "IDDPM", "gaussian", "gaussian_like", "uniform_int", "pad", "norm", "denorm", "denorm_uint8", "equations", "models",
           "diffusion_models", "optim", "training", "guidance"]
