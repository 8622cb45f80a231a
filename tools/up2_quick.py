"""Folded nearest-x2 + 3x3 conv (idf_conv_desc.up2) against upsample2x + the 9-tap conv at the benchmark shapes."""
import ctypes as C
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent))
import conv_microbench as mb
from infodiffusion_b200 import layout
from infodiffusion_b200._lib import ConvDesc
lib, _lib, B, dev, BF = mb.lib, mb._lib, mb.B, mb.dev, mb.BF
for (cin, cout, H) in [(128, 128, 32), (128, 128, 16), (128, 128, 8)]:
    rows_in, rows_up = B * (H + 1) ** 2, B * (2 * H + 1) ** 2
    x = torch.randn(rows_in, cin, device=dev).to(BF)
    w = (torch.randn(4 * cout, 4 * cin, device=dev) * 0.02).to(BF)
    b = torch.randn(4 * cout, device=dev)
    out = torch.zeros(rows_up, cout, device=dev, dtype=BF)
    st = torch.zeros(2 * ((rows_in + 127) // 128) * 4 * 4 * cout * 2, device=dev)
    d = ConvDesc()
    d.n_src = 1
    d.src[0], d.src_rows[0], d.src_ld[0] = x.data_ptr(), rows_in, cin
    kb = layout.taps_up2(cin, H, H)
    d.num_kb = len(kb)
    for k, (si, c0, off) in enumerate(kb):
        d.kb_src[k], d.kb_c0[k], d.kb_rowoff[k] = si, c0, off
    d.weight, d.cout_pad, d.block_n, d.cout, d.bias = w.data_ptr(), 4 * cout, cout, cout, b.data_ptr()
    d.batch, d.H, d.W, d.epilogue, d.up2 = B, H, H, 0, 1
    d.out, d.out_ld, d.stats_out = out.data_ptr(), cout, st.data_ptr()
    h = C.c_void_p()
    _lib.check(lib.idf_conv_plan_create(C.byref(d), C.byref(h)))
    us_up2 = mb.timeit(h)
    lib.idf_conv_plan_destroy(h)
    h9, keep = mb.make(cin, cout, 2 * H, False, True, False)
    us9 = mb.timeit(h9)
    lib.idf_conv_plan_destroy(h9)
    del keep
    up = torch.zeros(rows_up, cin, device=dev, dtype=BF)
    s = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        _lib.check(lib.idf_upsample2x(x.data_ptr(), up.data_ptr(), B, H, H, cin, s))
    e0.record()
    for _ in range(20):
        _lib.check(lib.idf_upsample2x(x.data_ptr(), up.data_ptr(), B, H, H, cin, s))
    e1.record(); torch.cuda.synchronize()
    us_cp = e0.elapsed_time(e1) / 20 * 1e3
    print(f"{cin}->{cout} {H}->{2*H}: folded {us_up2:6.1f} us | upsample2x {us_cp:6.1f} + 9-tap conv {us9:6.1f} = {us_cp + us9:6.1f} us", flush=True)
