"""Schedule tables of the three diffusion processes.  They are built once, on the host, with the same
torch CPU ops as the reference so that they are bit-identical to its registered buffers
(SURVEY.md par. 8a: "bit-exact tables"); the per-step closed forms that consume them live in
csrc/sampler.cu."""
from . import ddpm, ddim, iddpm

__all__ = ["ddpm", "ddim", "iddpm"]
