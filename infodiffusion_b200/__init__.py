"""infodiffusion_b200 -- B200-native (sm_100a) implementation of InfoDiffusion's denoising hot path.

Public surface mirrors the reference's Python operator API:
    models.InfoDiff / Diff / AuxiliaryUNet / BottleneckAuxUNet / UNet / Encoder / LatentUNet,
    sampling.DiffusionProcess / TwoPhaseDiffusionProcess / LatentDiffusionProcess, utils.compute_mmd,
    optim.ClipAdamW (clip_grad_norm_ + AdamW), io (eval_fid / save_latent writers), distributed (batch sharding)
All arithmetic of the path runs in hand-written CUDA kernels (libidf_b200.so, C ABI in include/idf_b200.h).
There is no CPU path and no fallback.
"""
from . import _lib  # noqa: F401
from .models import AuxiliaryUNet, BottleneckAuxUNet, Diff, Encoder, InfoDiff, LatentUNet, UNet  # noqa: F401
from .optim import ClipAdamW  # noqa: F401
from .sampling import DiffusionProcess, LatentDiffusionProcess, TwoPhaseDiffusionProcess  # noqa: F401
from .utils import compute_mmd  # noqa: F401

__all__ = ["AuxiliaryUNet", "BottleneckAuxUNet", "UNet", "Encoder", "LatentUNet", "InfoDiff", "Diff", "DiffusionProcess",
           "TwoPhaseDiffusionProcess", "LatentDiffusionProcess", "compute_mmd", "ClipAdamW"]
