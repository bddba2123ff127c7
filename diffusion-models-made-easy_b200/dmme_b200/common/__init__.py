from .noise import gaussian, gaussian_like, uniform_int, pad
from .norm import norm, denorm, denorm_uint8

__all__ = ["gaussian", "gaussian_like", "uniform_int", "pad", "norm", "denorm", "denorm_uint8"]
