"""One AdaGN apply launch per shape, bracketed by cudaProfilerStart/Stop (ncu --profile-from-start off)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from infodiffusion_b200 import _lib  # noqa: E402
from infodiffusion_b200._lib import AdaGNArgs  # noqa: E402

lib = _lib.load()
_lib.check(lib.idf_init())
dev = "cuda:0"
B, H, Cc = 256, 64, 64
rows = B * (H + 1) * (H + 1)
x = torch.randn(rows, Cc, device=dev).to(torch.bfloat16)
out = torch.zeros_like(x)
tiles = (rows + 127) // 128
stats = torch.rand(2 * tiles * 4 * Cc * 2, device=dev) + 1.0
gamma, beta = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
a = AdaGNArgs()
a.src0, a.c0, a.out = x.data_ptr(), Cc, out.data_ptr()
a.batch, a.H, a.W = B, H, H
a.gamma, a.beta, a.eps, a.apply_silu = gamma.data_ptr(), beta.data_ptr(), 1e-5, 1
a.stats0 = stats.data_ptr()
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(lib.idf_adagn_silu_fwd(C.byref(a), st))
torch.cuda.synchronize()
torch.cuda.profiler.start()
_lib.check(lib.idf_adagn_silu_fwd(C.byref(a), st))
out2 = torch.empty_like(x)
out2.copy_(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
