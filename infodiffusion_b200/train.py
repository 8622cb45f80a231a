"""Training-side host logic: closed-form parameter / modulation gradients of the fused AdaGN op and
(next) the backward plans of the networks.  Heavy arithmetic stays in libidf_b200.so."""
from __future__ import annotations

from typing import Dict, Optional

import torch


def adagn_param_grads(sums: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                      s_t: Optional[torch.Tensor] = None, b_t: Optional[torch.Tensor] = None,
                      s_z: Optional[torch.Tensor] = None, b_z: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Gradients of y = silu(((xhat*gamma + beta)(1+s_t) + b_t)(1+s_z) + b_z) w.r.t. everything but x, from
    the per-(sample, channel) sums S1 = sum_hw dv, S2 = sum_hw dv*xhat the backward kernel returns
    (dv = dL/d pre-activation).  sums: [B, C, 2]; gamma/beta: [C]; modulation tensors: [B, C] or None."""
    S1, S2 = sums[..., 0], sums[..., 1]
    q = gamma[None, :] * S2 + beta[None, :] * S1          # sum_hw dv * (xhat*gamma + beta)
    T = (1 + s_t) if s_t is not None else torch.ones_like(S1)
    Z = (1 + s_z) if s_z is not None else torch.ones_like(S1)
    out = {"gamma": (T * Z * S2).sum(0), "beta": (T * Z * S1).sum(0)}
    if s_t is not None:
        out["s_t"] = Z * q
        out["b_t"] = Z * S1
    if s_z is not None:
        out["s_z"] = T * q + (b_t if b_t is not None else 0) * S1
        out["b_z"] = S1
    return out
