"""One training step (eager launches, no CUDA graphs) bracketed by cudaProfilerStart/Stop:
ncu --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file out.csv python tools/train_launches.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from infodiffusion_b200.models import InfoDiff  # noqa: E402
from infodiffusion_b200 import train as T  # noqa: E402
from infodiffusion_b200.optim import ClipAdamW  # noqa: E402

B = 32
dev = "cuda:0"
args = bench.make_args_ns(1000)
torch.manual_seed(64)
model = InfoDiff(args, "cpu", (3, 64, 64)).to(dev)
model.device = dev
for n in ("alpha_bars", "betas", "alphas", "alpha_prev_bars"):
    setattr(model, n, getattr(model, n).to(dev))
model.train()
opt = ClipAdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-5, max_norm=1.0)
x = (torch.rand(B, 3, 64, 64, device=dev) * 2 - 1)
T.USE_GRAPHS = False


def step():
    loss = model.loss_fn(args, x)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
