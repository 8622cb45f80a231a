"""infodiffusion_b200 -- B200-native (sm_100a) implementation of InfoDiffusion's denoising hot path.

Public surface mirrors the reference's Python operator API:
    models.InfoDiff / AuxiliaryUNet / Encoder, sampling.DiffusionProcess, utils.compute_mmd
All arithmetic runs in hand-written CUDA kernels (libidf_b200.so, C ABI in include/idf_b200.h).
There is no CPU path and no fallback.
"""
from . import _lib  # noqa: F401
from .models import AuxiliaryUNet, Encoder, InfoDiff  # noqa: F401
from .sampling import DiffusionProcess  # noqa: F401
from .utils import compute_mmd  # noqa: F401

__all__ = ["AuxiliaryUNet", "Encoder", "InfoDiff", "DiffusionProcess", "compute_mmd"]
