"""CPU emulation of the pad-flat layout and of the implicit-GEMM K-block semantics (test helpers)."""
from __future__ import annotations

import torch


def to_padflat(x: torch.Tensor) -> torch.Tensor:
    """NCHW -> pad-flat [B*(H+1)*(W+1), C] with zero pad row / column."""
    B, C, H, W = x.shape
    buf = torch.zeros(B, H + 1, W + 1, C, dtype=x.dtype)
    buf[:, :H, :W, :] = x.permute(0, 2, 3, 1)
    return buf.reshape(B * (H + 1) * (W + 1), C)


def from_padflat(m: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
    C = m.shape[1]
    return m.reshape(B, H + 1, W + 1, C)[:, :H, :W, :].permute(0, 3, 1, 2).contiguous()


def interior_mask(B: int, H: int, W: int) -> torch.Tensor:
    m = torch.zeros(B, H + 1, W + 1, dtype=torch.bool)
    m[:, :H, :W] = True
    return m.reshape(-1)


def emulate_igemm(srcs, kblocks, wp, bias, B, H, W, residual=None):
    """out[r, n] = sum_kb A_kb[r + off, c0:c0+64] . wp[n, 64kb:64kb+64] + bias[n] (+ residual), interior rows only.
    Rows outside a source read as zero (what TMA's out-of-bounds fill does)."""
    rows = B * (H + 1) * (W + 1)
    out = torch.zeros(rows, wp.shape[0], dtype=torch.float64)
    r = torch.arange(rows)
    for kb, (si, c0, off) in enumerate(kblocks):
        src = srcs[si].double()
        idx = r + off
        ok = (idx >= 0) & (idx < src.shape[0])
        a = torch.zeros(rows, 64, dtype=torch.float64)
        a[ok] = src[idx[ok], c0:c0 + 64]
        out += a @ wp[:, 64 * kb:64 * kb + 64].double().t()
    out += bias.double()[None, :]
    if residual is not None:
        out += residual.double()
    out[~interior_mask(B, H, W)] = 0
    return out


def space_to_depth(x_pf: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
    """pad-flat [B,H,W,C] -> 4 stacked pad-flat phase maps [4 * B*(H/2+1)*(W/2+1), C]."""
    x = from_padflat(x_pf, B, H, W)
    phases = [to_padflat(x[:, :, py::2, px::2]) for py in (0, 1) for px in (0, 1)]
    return torch.cat(phases, 0)
