"""Per-shape timing of the conv kernel at the benchmark batch (CUDA events), with the epilogue and the
GroupNorm partials switched off in turn: separates main-loop (TMA + UMMA) time from epilogue time."""
import ctypes as C
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from infodiffusion_b200 import _lib, layout  # noqa: E402
from infodiffusion_b200._lib import ConvDesc  # noqa: E402

lib = _lib.load()
_lib.check(lib.idf_init())
dev = "cuda:0"
BF = torch.bfloat16
B = int(os.environ.get("IDF_MB_BATCH", "256"))
PEAK = 1406.4
PAIR_DEFAULT = int(os.environ.get("IDF_CONV_PAIR", "1"))
_lib.check(lib.idf_set_option(b"conv_pair", PAIR_DEFAULT))


def make(cin, cout, H, residual, stats, skip, xf=False, bn=None):
    rows = B * (H + 1) * (H + 1)
    x = torch.randn(rows, cin, device=dev).to(BF)
    w = (torch.randn(cout, 9 * cin, device=dev) * 0.02).to(BF)
    b = torch.randn(cout, device=dev)
    out = torch.zeros(rows, cout, device=dev, dtype=BF)
    res = torch.randn(rows, cout, device=dev).to(BF) if residual else None
    st = torch.zeros(2 * ((rows + 127) // 128) * 4 * cout * 2, device=dev) if stats else None    # records A and B per 32-row window
    d = ConvDesc()
    d.n_src = 1
    d.src[0], d.src_rows[0], d.src_ld[0] = x.data_ptr(), rows, cin
    kb = layout.taps3x3(cin, H, H)
    d.num_kb = len(kb)
    for k, (si, c0, off) in enumerate(kb):
        d.kb_src[k], d.kb_c0[k], d.kb_rowoff[k] = si, c0, off
    bn = bn or (128 if cout % 128 == 0 else 64)
    d.weight, d.cout_pad, d.block_n, d.cout, d.bias = w.data_ptr(), cout, bn, cout, b.data_ptr()
    d.batch, d.H, d.W, d.epilogue = B, H, H, 0
    d.out, d.out_ld = out.data_ptr(), cout
    if res is not None:
        d.residual, d.res_ld = res.data_ptr(), cout
    if st is not None:
        d.stats_out = st.data_ptr()
    coef = None
    if xf:      # fused AdaGN on the A operand: per-image (A, B) coefficients, SiLU on
        coef = torch.stack([1.0 + 0.1 * torch.randn(B, cin, device=dev), 0.1 * torch.randn(B, cin, device=dev)], dim=-1).contiguous()
        d.xf_coef, d.xf_ctot, d.xf_silu = coef.data_ptr(), cin, 1
        for k, (si, c0, off) in enumerate(kb):
            d.kb_xf[k] = c0
    _lib.check(lib.idf_set_option(b"conv_debug_skip_epilogue", int(skip)))
    h = C.c_void_p()
    _lib.check(lib.idf_conv_plan_create(C.byref(d), C.byref(h)))
    _lib.check(lib.idf_set_option(b"conv_debug_skip_epilogue", 0))
    return h, (x, w, b, out, res, st, d, coef)


def timeit(h, n=20):
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        _lib.check(lib.idf_conv_run(h, s))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        _lib.check(lib.idf_conv_run(h, s))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3   # us


def main():
    print(f"batch {B}; us per launch (TFLOP/s, % of {PEAK} sustained)")
    for (cin, cout, H) in [(64, 64, 64), (128, 64, 64), (192, 64, 64), (128, 128, 64), (128, 128, 32), (256, 128, 32),
                           (128, 128, 16), (256, 128, 16), (128, 128, 8)]:
        fl = 2.0 * B * H * H * 9 * cin * cout
        line = f"{cin:3d}->{cout:3d} @{H:2d}: "
        for name, (res, stats, skip) in {"full": (False, True, False), "+res": (True, True, False),
                                         "nostats": (False, False, False), "mainloop": (False, False, True)}.items():
            h, keep = make(cin, cout, H, res, stats, skip)
            us = timeit(h)
            lib.idf_conv_plan_destroy(h)
            del keep
            line += f"{name} {us:7.1f} ({fl / us / 1e6:6.0f}, {fl / us / 1e6 / PEAK * 100:4.1f}%)  "
        if os.environ.get("IDF_MB_XF"):
            for dbg in (0, 2, 5, 1):
                _lib.check(lib.idf_set_option(b"xf_debug", dbg))
                h, keep = make(cin, cout, H, False, True, False, xf=True)
                us = timeit(h)
                lib.idf_conv_plan_destroy(h)
                del keep
                line += f"xf{dbg} {us:7.1f}  "
            _lib.check(lib.idf_set_option(b"xf_debug", 0))
        # planner's choice against every forced number of 128-row tiles per work unit, as single CTAs and as CTA pairs
        for pair in (() if os.environ.get("IDF_MB_QUICK") else (0, 1)):
            _lib.check(lib.idf_set_option(b"conv_pair", pair))
            line += "| pair " if pair else "| single "
            for mt in (1, 2, 4):
                if mt == 4 and cout % 128 == 0:
                    continue
                _lib.check(lib.idf_set_option(b"conv_force_mt", mt))
                try:
                    h, keep = make(cin, cout, H, False, True, False)
                    us = timeit(h)
                    lib.idf_conv_plan_destroy(h)
                    del keep
                    line += f"MT={mt} {us:6.1f}  "
                except Exception as e:       # configuration does not fit in shared memory
                    line += f"MT={mt}   n/a  "
                finally:
                    _lib.check(lib.idf_set_option(b"conv_force_mt", 0))
        _lib.check(lib.idf_set_option(b"conv_pair", PAIR_DEFAULT))
        print(line, flush=True)


if __name__ == "__main__":
    main()
