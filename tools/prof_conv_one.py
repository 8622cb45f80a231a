"""One conv shape under ncu: `ncu --set full --import-source on -k regex:conv_halo -s 3 -c 1 python tools/prof_conv_one.py 64 64 64`
(arguments: cin cout H [residual 0/1] [fused AdaGN 0/1])."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
import conv_microbench as mb  # noqa: E402  (prints its table only when run as a script)

cin, cout, H = (int(v) for v in sys.argv[1:4])
res = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
xf = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
h, keep = mb.make(cin, cout, H, res, True, False, xf=xf)
s = torch.cuda.current_stream().cuda_stream
for _ in range(5):
    mb._lib.check(mb.lib.idf_conv_run(h, s))
torch.cuda.synchronize()
print("ran", cin, cout, H, res, "batch", mb.B)
