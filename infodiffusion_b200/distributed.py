"""One-process-per-GPU partitioning of the hot path.

Sampling, latent encoding and reverse DDIM are independent per sample (GroupNorm and attention never mix
samples), so a batch of N is cut into contiguous shards, every rank runs its shard with no communication
and a single all_gather assembles the result.  Training is data parallel: one all-reduce (average) of a
flat gradient buffer per step (train.allreduce_gradients).  Nothing here launches kernels; the functions
work with any torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import torch

from .layout import shard_range


def _world() -> Tuple[int, int]:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def local_slice(t: Optional[torch.Tensor], total: int) -> Optional[torch.Tensor]:
    """This rank's contiguous shard of a full-batch tensor (None passes through)."""
    if t is None:
        return None
    rank, world = _world()
    assert t.shape[0] == total, f"expected a full batch of {total}, got {t.shape[0]}"
    lo, hi = shard_range(total, rank, world)
    return t[lo:hi]


def gather_batch(local: torch.Tensor, total: int) -> torch.Tensor:
    """all_gather of per-rank shards (possibly of different sizes) back into the full batch, rank order =
    batch order.  Shards are padded to the largest one because all_gather needs equal shapes."""
    import torch.distributed as dist
    rank, world = _world()
    if world == 1:
        return local
    sizes = [hi - lo for lo, hi in (shard_range(total, r, world) for r in range(world))]
    assert local.shape[0] == sizes[rank]
    width = max(sizes)
    padded = local.new_zeros((width,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded.contiguous())
    return torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0)


def sharded_sampling(sample_fn: Callable[..., torch.Tensor], total: int, xT: Optional[torch.Tensor] = None,
                     a: Optional[torch.Tensor] = None, gather: bool = True) -> torch.Tensor:
    """Batch-sharded `DiffusionProcess.sampling` (reference sampling.py:89-101 run on every rank over its shard).

    `sample_fn(n_local, xT=..., a=...)` is the per-rank sampler (e.g. DiffusionProcess.sampling).  `xT` / `a`
    are FULL-batch tensors (identical on every rank, e.g. drawn from the reference seed) or None; with the same
    full-batch draws the gathered result equals the single-process result sample for sample."""
    rank, world = _world()
    lo, hi = shard_range(total, rank, world)
    out = sample_fn(hi - lo, xT=local_slice(xT, total), a=local_slice(a, total))
    return gather_batch(out, total) if gather else out


def sharded_map(fn: Callable[[torch.Tensor], Sequence[torch.Tensor]], x: torch.Tensor, gather: bool = True):
    """Apply a per-sample-independent function (Encoder.forward, reverse_sampling) to this rank's shard of the
    full batch `x` and gather every output (save_latent: z and x_T, run.py:416-443)."""
    total = x.shape[0]
    outs = fn(local_slice(x, total))
    single = torch.is_tensor(outs)
    outs = [outs] if single else list(outs)
    if gather:
        outs = [gather_batch(o, total) for o in outs]
    return outs[0] if single else tuple(outs)
