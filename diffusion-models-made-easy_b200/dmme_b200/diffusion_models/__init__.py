from .ddpm import DDPM
from .ddim import DDIM
from .iddpm import IDDPM

__all__ = ["DDPM", "DDIM", "IDDPM"]
