"""Fused train-step tail: gradient-norm clipping + AdamW over every parameter in three kernel launches
(reference run.py:177, 199-200: ``clip_grad_norm_(model.parameters(), 1.)`` then ``AdamW.step()``).

``ClipAdamW`` is a ``torch.optim.Optimizer`` (param_groups, ``lr`` per group, ``zero_grad``), so the reference's
LR schedulers (CosineAnnealingLR wrapped by GradualWarmupScheduler, run.py:182-185) drive it unchanged.  The
update rule and its defaults are torch.optim.AdamW's; the clip coefficient is clip_grad_norm_'s
``min(1, max_norm / (total_norm + 1e-6))`` with the L2 norm taken over all parameters of all groups.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _lib
from ._lib import ClipAdamWArgs

_CHUNK = 4096


class ClipAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_norm=1.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.max_norm = float(max_norm)
        self.total_norm = None         # device scalar after step(): what clip_grad_norm_ returns
        self._tables = {}

    def _group_tables(self, gi: int, ps: List[torch.Tensor]):
        """Static per-group tables (parameter / moment pointers, sizes, chunk map); gradients change per step."""
        key = (gi, tuple(id(p) for p in ps))
        tb = self._tables.get(gi)
        if tb is not None and tb["key"] == key:
            return tb
        if tb is not None and tb.get("uniform_step") is not None:       # the set changed: back to per-parameter counts
            for p in tb["ps"]:
                self.state[p]["step"] = tb["uniform_step"]
        dev = ps[0].device
        for p in ps:
            if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                raise RuntimeError("ClipAdamW needs contiguous fp32 CUDA parameters")
            st = self.state[p]
            if "exp_avg" not in st:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        ct, co = [], []
        for i, p in enumerate(ps):
            n = (p.numel() + _CHUNK - 1) // _CHUNK
            ct += [i] * n
            co += list(range(n))
        i64 = lambda v: torch.tensor(v, dtype=torch.int64, device=dev)
        i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
        tb = dict(key=key, n=len(ct), ps=list(ps), uniform_step=None,
                  params=i64([p.data_ptr() for p in ps]),
                  m=i64([self.state[p]["exp_avg"].data_ptr() for p in ps]),
                  v=i64([self.state[p]["exp_avg_sq"].data_ptr() for p in ps]),
                  numel=i64([p.numel() for p in ps]), ct=i32(ct), co=i32(co),
                  partial=torch.zeros(max(len(ct), 1), dtype=torch.float32, device=dev),
                  grads_host=torch.zeros(len(ps), dtype=torch.int64).pin_memory(),
                  grads=torch.zeros(len(ps), dtype=torch.int64, device=dev),
                  bc_host=torch.zeros(len(ps), 2, dtype=torch.float32).pin_memory(),
                  bc=torch.zeros(len(ps), 2, dtype=torch.float32, device=dev))
        self._tables[gi] = tb
        return tb

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = _lib.load()
        groups = []
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if ps:
                groups.append((gi, group, ps))
        if not groups:
            return loss
        dev = groups[0][2][0].device
        stream = torch.cuda.current_stream(dev).cuda_stream
        if len(groups) > 1 and self.max_norm > 0:
            raise NotImplementedError("global-norm clipping across several param groups")
        for gi, group, ps in groups:
            tb = self._group_tables(gi, ps)
            if tb.get("staged") is not None:
                tb["staged"].synchronize()                # the previous step's async copies read the pinned tables
            gh = tb["grads_host"]
            ptrs = []
            for p in ps:
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = p.grad = g.float().contiguous()
                ptrs.append(g.data_ptr())
            gh.numpy()[:] = ptrs                           # one bulk conversion instead of a tensor store per parameter
            tb["grads"].copy_(gh, non_blocking=True)
            b1, b2 = group["betas"]
            # one step count per parameter, like torch.optim.AdamW.  While the same set of parameters keeps receiving
            # gradients (same table) and they all share one count, a single counter stands in for all of them.
            bch = tb["bc_host"]
            if tb.get("uniform_step") is not None:
                tb["uniform_step"] += 1
                k = tb["uniform_step"]
                bch[:, 0] = 1.0 - b1 ** k
                bch[:, 1] = 1.0 - b2 ** k
            else:
                steps = []
                for p in ps:
                    st = self.state[p]
                    st["step"] = st.get("step", 0) + 1
                    steps.append(st["step"])
                if min(steps) == max(steps):
                    tb["uniform_step"] = steps[0]
                    bch[:, 0] = 1.0 - b1 ** steps[0]
                    bch[:, 1] = 1.0 - b2 ** steps[0]
                else:
                    bch.copy_(torch.tensor([[1.0 - b1 ** k, 1.0 - b2 ** k] for k in steps], dtype=torch.float32))
            tb["bc"].copy_(bch, non_blocking=True)
            tb["staged"] = torch.cuda.Event()
            tb["staged"].record(torch.cuda.current_stream(dev))
            norm_out = torch.empty(2, dtype=torch.float32, device=dev)
            a = ClipAdamWArgs()
            a.params, a.grads, a.exp_avg, a.exp_avg_sq = (tb[k].data_ptr() for k in ("params", "grads", "m", "v"))
            a.numel, a.chunk_tensor, a.chunk_offset = tb["numel"].data_ptr(), tb["ct"].data_ptr(), tb["co"].data_ptr()
            a.n_chunks = tb["n"]
            a.lr, a.beta1, a.beta2, a.eps, a.weight_decay = group["lr"], b1, b2, group["eps"], group["weight_decay"]
            a.bias_corrections = tb["bc"].data_ptr()
            a.max_norm = self.max_norm
            a.partial, a.norm_out = tb["partial"].data_ptr(), norm_out.data_ptr()
            _lib.check(lib.idf_clip_adamw(C.byref(a), stream))
            _lib.count_launch(3)
            _lib.bump_weight_epoch()              # parameters changed without torch noticing (raw-pointer kernel)
            self.total_norm = norm_out[0]
        return loss
