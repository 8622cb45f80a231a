"""Fused-AdaGN conv timings at the benchmark shapes (development helper): plain, fused (xf0), fused without MMAs (xf5),
fused with idle transform (xf1)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
import conv_microbench as mb
import os
lib, _lib = mb.lib, mb._lib
if os.environ.get("IDF_XF_MT"):            # forced tiles per work item (A/B of the tile-outer kernels)
    _lib.check(lib.idf_set_option(b"conv_force_mt", int(os.environ["IDF_XF_MT"])))
    print("forced MT", os.environ["IDF_XF_MT"])
SHAPES = [(64, 64, 64), (128, 64, 64), (192, 64, 64), (128, 128, 32), (128, 128, 16), (128, 128, 8)]
if os.environ.get("IDF_XF_SHORT"):
    SHAPES = [(64, 64, 64), (128, 64, 64), (64, 64, 32)]
for (cin, cout, H) in SHAPES:
    line = f"{cin:3d}->{cout:3d}@{H:2d}: "
    h, keep = mb.make(cin, cout, H, False, True, False)
    line += f"plain {mb.timeit(h):6.1f} |"
    lib.idf_conv_plan_destroy(h); del keep
    for dbg in (0, 5, 1):
        _lib.check(lib.idf_set_option(b"xf_debug", dbg))
        h, keep = mb.make(cin, cout, H, False, True, False, xf=True)
        line += f" xf{dbg} {mb.timeit(h):6.1f} |"
        lib.idf_conv_plan_destroy(h); del keep
    _lib.check(lib.idf_set_option(b"xf_debug", 0))
    print(line, flush=True)
