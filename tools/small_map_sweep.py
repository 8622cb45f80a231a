"""Column-tile width x tiles per work item x CTA pairing on the small maps (development helper): the 8x8 / 16x16 layers
are latency-bound, does a finer split over the 148 SMs shorten them?"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
import conv_microbench as mb
lib, _lib = mb.lib, mb._lib
for (cin, cout, H) in [(128, 128, 8), (256, 128, 8), (128, 128, 16), (256, 128, 16), (128, 128, 32)]:
    for res in (False, True):
        line = f"{cin:3d}->{cout:3d}@{H:2d}{'+res' if res else '    '}: "
        for bn in (128, 64):
            for pair in (1, 0):
                _lib.check(lib.idf_set_option(b"conv_pair", pair))
                for mt in (1, 2, 4):
                    if mt == 4 and bn == 128:
                        continue
                    _lib.check(lib.idf_set_option(b"conv_force_mt", mt))
                    try:
                        h, keep = mb.make(cin, cout, H, res, True, False, bn=bn)
                        us = mb.timeit(h, n=50)
                        lib.idf_conv_plan_destroy(h)
                        del keep
                        line += f"bn{bn} p{pair} mt{mt} {us:5.1f} | "
                    except Exception:
                        line += f"bn{bn} p{pair} mt{mt}  n/a | "
                    finally:
                        _lib.check(lib.idf_set_option(b"conv_force_mt", 0))
        _lib.check(lib.idf_set_option(b"conv_pair", 1))
        print(line, flush=True)
