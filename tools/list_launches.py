"""Per-launch CUDA-event timing of one UNet evaluation at the benchmark shape: kernel class, milliseconds,
algorithmic FLOPs / bytes and the achieved rate of every launch (eager, events on the launching stream)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from infodiffusion_b200.models import InfoDiff  # noqa: E402
from infodiffusion_b200.sampling import DiffusionProcess  # noqa: E402

B = int(os.environ.get("IDF_PROF_BATCH", "256"))
dev = "cuda:0"
if os.environ.get("IDF_FUSE"):
    from infodiffusion_b200 import engine
    engine.FUSE_ADAGN = True
args = bench.make_args_ns(bench.T_STEPS)
torch.manual_seed(64)
model = InfoDiff(args, "cpu", (3, 64, 64)).to(dev).eval()
model.device = dev
proc = DiffusionProcess(args, model, dev, (3, 64, 64))
s = proc._sampler("ddim", B)
s.x.normal_(); s.set_latent(torch.randn(B, bench.A_DIM, device=dev)); s.noise.normal_(); s.step.fill_(50)
plan = s.plans[-1]
for _ in range(2):
    plan.run()
torch.cuda.synchronize()
reps = 5
acc = None
for _ in range(reps):
    r = plan.run_timed()
    acc = [list(x) for x in r] if acc is None else [[a[0], a[1] + b[1], a[2], a[3]] for a, b in zip(acc, r)]
tot = {}
for i, (tag, ms, fl, by) in enumerate(acc):
    ms /= reps
    rate = f"{fl / ms / 1e9:8.1f} TF/s" if fl else (f"{by / ms / 1e6:8.1f} GB/s" if by else "")
    print(f"{i:4d} {tag:16s} {ms * 1e3:9.1f} us  flops {fl / 1e9:9.2f} G  bytes {by / 1e6:8.1f} MB  {rate}")
    tot[tag] = tot.get(tag, 0.0) + ms
print({k: round(v, 3) for k, v in tot.items()}, "total ms", round(sum(tot.values()), 3))
