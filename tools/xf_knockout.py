"""Fused-AdaGN conv with parts of the pipeline knocked out (development helper): which stage bounds the item period?
xf_debug 0 = everything, 1 = transform skipped, 2 = affine only (no tanh), 5 = no MMAs; skip = epilogue skipped."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
import conv_microbench as mb
lib, _lib = mb.lib, mb._lib
for (cin, cout, H) in [(64, 64, 64), (128, 64, 64)]:
    for skip in (False, True):
        line = f"{cin:3d}->{cout:3d}@{H:2d} {'no-epilogue' if skip else 'full      '}: "
        h, keep = mb.make(cin, cout, H, False, True, skip)
        line += f"plain {mb.timeit(h):6.1f} |"
        lib.idf_conv_plan_destroy(h); del keep
        for dbg in (0, 2, 1, 5):
            _lib.check(lib.idf_set_option(b"xf_debug", dbg))
            h, keep = mb.make(cin, cout, H, False, True, skip, xf=True)
            line += f" xf{dbg} {mb.timeit(h):6.1f} |"
            lib.idf_conv_plan_destroy(h); del keep
        _lib.check(lib.idf_set_option(b"xf_debug", 0))
        print(line, flush=True)
