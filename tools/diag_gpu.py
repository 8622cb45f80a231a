"""Layout diagnostics for the tcgen05 kernels (run on the GPU box when a parity test fails).
Uses identity-like weights and unique input values so that a wrong descriptor / swizzle shows up as a
readable permutation instead of noise."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from infodiffusion_b200 import _lib, layout  # noqa: E402
from tests.test_gpu_ops import BF, DEV, pf, run_conv, unpf  # noqa: E402


def main():
    lib = _lib.load()
    _lib.check(lib.idf_init())
    torch.manual_seed(0)
    # --- probe 1: 1x1 conv, identity weights, K = 64: out must equal x
    for (cin, cout, bn) in [(64, 64, 64), (128, 128, 128), (64, 128, 128)]:
        B, H = 1, 8
        x = torch.arange(B * cin * H * H, device=DEV, dtype=torch.float32).reshape(B, cin, H, H) % 251 - 125
        w = torch.zeros(cout, cin, device=DEV)
        for i in range(cout):
            w[i, i % cin] = 1.0
        b = torch.zeros(cout, device=DEV)
        try:
            out = run_conv(lib, [pf(x)], layout.taps1x1(cin), w.to(BF).contiguous(), b, B, H, cout, bn)
        except Exception as e:  # noqa: BLE001
            print(f"probe1 {cin}->{cout}: EXCEPTION {e}")
            return
        got = unpf(out, B, H, H)
        ref = x[:, [i % cin for i in range(cout)]]
        bad = (got != ref)
        print(f"probe1 1x1 identity {cin}->{cout} bn={bn}: mismatches {int(bad.sum())} / {bad.numel()}")
        if bad.any():
            idx = bad.nonzero()[:8]
            for n, c, y, xx in idx.tolist():
                print(f"   out[c={c},y={y},x={xx}] = {got[n, c, y, xx].item():.1f}, want {ref[n, c, y, xx].item():.1f}")
    # --- probe 2: 3x3, single non-zero tap per test
    B, H, cin = 1, 8, 64
    x = torch.randn(B, cin, H, H, device=DEV).to(BF).float()
    for tap in (0, 4, 8):
        w = torch.zeros(64, cin, 3, 3, device=DEV)
        ky, kx = divmod(tap, 3)
        for i in range(64):
            w[i, i, ky, kx] = 1.0
        out = run_conv(lib, [pf(x)], layout.taps3x3(cin, H, H), layout.pack_conv3x3(w).to(BF).contiguous(),
                       torch.zeros(64, device=DEV), B, H, 64, 64)
        ref = torch.nn.functional.conv2d(x, w, padding=1)
        err = (unpf(out, B, H, H) - ref).abs().max().item()
        print(f"probe2 3x3 tap {tap}: max err {err:.3e}")
    print("diag done")


if __name__ == "__main__":
    main()
