"""Phase timeline of the conv kernel's pipeline for the first CTA (development tool; needs the trace build:
nvcc ... -DIDF_CONV_TRACE linked as tools/ab/libidf_trace.so and IDF_LIB_AB pointing to it).
usage: IDF_LIB_AB=tools/ab/libidf_trace.so python tools/conv_trace.py cin cout H xf(0/1) [xf_debug]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
import conv_microbench as mb  # noqa: E402

cin, cout, H, xf = (int(v) for v in sys.argv[1:5])
dbg = int(sys.argv[5]) if len(sys.argv) > 5 else 0
mb._lib.check(mb.lib.idf_set_option(b"xf_debug", dbg))
trace = torch.zeros(64, 8, dtype=torch.int64, device="cuda:0")
h, keep = mb.make(cin, cout, H, False, True, False, xf=bool(xf))
d = keep[6]
# re-create the plan with the trace buffer attached
mb.lib.idf_conv_plan_destroy(h)
d.out_f32 = trace.data_ptr()
import ctypes as C
h = C.c_void_p()
mb._lib.check(mb.lib.idf_conv_plan_create(C.byref(d), C.byref(h)))
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    mb._lib.check(mb.lib.idf_conv_run(h, s))
torch.cuda.synchronize()
t = trace.cpu().numpy()
t0 = t[0, 0]
names = ["A issue", "A landed(xf)", "xf done", "mma start", "mma issued", "acc full", "-", "-"]
print("item  " + "  ".join(f"{n:>12s}" for n in names[:6]) + "   (us since the first TMA issue)")
for i in range(16):
    if t[i, 0] == 0:
        break
    print(f"{i:4d}  " + "  ".join(f"{(t[i, k] - t0) / 1e3:12.2f}" if t[i, k] else f"{'':>12s}" for k in range(6)))
# first drain warp's chunks (trace rows 32 + 4*item + tile): TMEM load issued / data in registers / previous TMA store
# has read the tile / statistics warp done with the tile / tile written + TMA store issued
print("\nfirst drain warp, per (item, tile):  tmem_ld  ld_done  store_read  stats_done  staged   (us since the first TMA issue)")
for i in range(6):
    for m in range(4):
        r = 32 + 4 * i + m
        if t[r, 0] == 0:
            continue
        print(f"  item {i} tile {m}: " + "  ".join(f"{(t[r, k] - t0) / 1e3:9.2f}" for k in range(5)))
