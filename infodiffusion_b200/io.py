"""Output stage either side of the sampling path: the files `run.py --mode eval_fid`
and `--mode save_latent` leave behind, in the formats their consumers read.

* eval_fid (run.py:282-295): every sample is clipped to [-1, 1], mapped to [0, 1] and written by
  torchvision.utils.save_image as `sample-%06d.png`; calc_fid.py:12 reads the folder.  Here the whole batch is
  quantised on the device by one kernel (idf_to_uint8_hwc, same fp32 arithmetic => same bytes) and the PNGs are
  encoded on the host from one D2H copy; an .npz of the uint8 batch is offered as the bulk alternative.
* save_latent (run.py:416-443): np.savez('<model>_<exp>_latent', all_a=..., all_attr=...), consumed by
  eval_disentanglement.py:368.
"""
from __future__ import annotations

import os
import struct
import zlib
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib


def images_to_uint8(x: torch.Tensor) -> torch.Tensor:
    """fp32 NCHW in [-1, 1] (CUDA) -> uint8 NHWC (CUDA), bit-identical to the reference's clip / normalise /
    save_image quantisation."""
    if not x.is_cuda:
        raise RuntimeError("images_to_uint8 needs a CUDA tensor: infodiffusion_b200 has no CPU path")
    x = x.contiguous().float()
    B, C_, H, W = x.shape
    out = torch.empty(B, H, W, C_, dtype=torch.uint8, device=x.device)
    lib = _lib.load()
    _lib.check(lib.idf_to_uint8_hwc(x.data_ptr(), out.data_ptr(), B, C_, H, W, torch.cuda.current_stream(x.device).cuda_stream))
    _lib.count_launch()
    return out


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def encode_png(hwc: np.ndarray, level: int = 1) -> bytes:
    """Minimal PNG encoder (8-bit gray / RGB, filter 0, one IDAT).  Decodes to exactly `hwc`."""
    assert hwc.dtype == np.uint8 and hwc.ndim == 3 and hwc.shape[2] in (1, 3)
    H, W, C_ = hwc.shape
    raw = np.zeros((H, 1 + W * C_), dtype=np.uint8)          # filter byte 0 in front of every scanline
    raw[:, 1:] = hwc.reshape(H, W * C_)
    ihdr = struct.pack(">IIBBBBB", W, H, 8, 0 if C_ == 1 else 2, 0, 0, 0)
    return b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", ihdr) + _chunk(b"IDAT", zlib.compress(raw.tobytes(), level)) + _chunk(b"IEND", b"")


def save_eval_images(batch: torch.Tensor, root: str, first_index: int = 0, limit: Optional[int] = None) -> int:
    """Write `sample-%06d.png` for every image of `batch` (fp32 NCHW in [-1, 1]) whose running index is below
    `limit` (run.py:288-295: args.sampling_number).  Returns the number of files written."""
    os.makedirs(root, exist_ok=True)
    u8 = images_to_uint8(batch).cpu().numpy()               # one D2H copy for the batch
    n = 0
    for k in range(u8.shape[0]):
        idx = first_index + k
        if limit is not None and idx >= limit:
            break
        with open(os.path.join(root, f"sample-{idx:06d}.png"), "wb") as f:
            f.write(encode_png(u8[k]))
        n += 1
    return n


def save_samples_npz(path: str, batches: Sequence[torch.Tensor]) -> None:
    """Bulk alternative to the PNG folder: one compressed .npz with `images` uint8 [N, H, W, C]."""
    np.savez_compressed(path, images=np.concatenate([images_to_uint8(b).cpu().numpy() for b in batches]))


def save_latents_npz(path: str, all_a: Sequence, all_attr: Sequence) -> None:
    """`{model}_{exp}_latent.npz` with keys all_a / all_attr (run.py:439-443)."""
    a = np.concatenate([t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t) for t in all_a])
    attr = np.concatenate([t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t) for t in all_attr])
    np.savez(path, all_a=a, all_attr=attr)
