"""Runs the UMMA shifted-descriptor probe for several row shifts and both base_offset conventions."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from infodiffusion_b200 import _lib  # noqa: E402

lib = _lib.load()
_lib.check(lib.idf_init())
dev = "cuda:0"
a = (torch.arange(256 * 64, device=dev, dtype=torch.float32).reshape(256, 64) % 509 - 254) / 4   # exact in bf16? values/4 up to 63.5
a = a.to(torch.bfloat16).contiguous()
b = torch.eye(64, device=dev, dtype=torch.bfloat16).contiguous()
for shift in (0, 1, 2, 7, 8, 9, 16, 65, 66, 73, 127, 128):
    for mode in (0, 1):
        out = torch.full((128, 64), float("nan"), device=dev)
        _lib.check(lib.idf_debug_shift_probe(a.data_ptr(), b.data_ptr(), out.data_ptr(), shift, mode,
                                             torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        ref = a[shift:shift + 128].float()
        bad = int((out != ref).sum())
        # if wrong, does it match some OTHER row shift / permutation?
        note = ""
        if bad:
            rows_ok = int((out == ref).all(dim=1).sum())
            note = f" rows fully right: {rows_ok}/128"
        print(f"shift {shift:3d} mode {mode}: mismatches {bad}{note}")
