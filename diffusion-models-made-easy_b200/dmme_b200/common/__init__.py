from .noise import gaussian, gaussian_like, uniform_int, pad

__all__ = ["gaussian", "gaussian_like", "uniform_int", "pad"]
