"""Host time needed to ENQUEUE one training step vs the GPU time of the step (is the step host- or device-bound?)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from infodiffusion_b200.models import InfoDiff  # noqa: E402
from infodiffusion_b200.optim import ClipAdamW  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda:0"
args = bench.make_args_ns(1000)
args.mode = "train"
torch.manual_seed(64)
model = InfoDiff(args, "cpu", (3, 64, 64)).to(dev)
model.device = dev
for n in ("alpha_bars", "betas", "alphas", "alpha_prev_bars"):
    setattr(model, n, getattr(model, n).to(dev))
model.train()
params = [p for p in model.parameters() if p.requires_grad]
opt = ClipAdamW(params, lr=1e-4, weight_decay=1e-5, max_norm=1.0)
x = (torch.rand(B, 3, 64, 64) * 2 - 1).to(dev)


def step(parts=None):
    t0 = time.perf_counter()
    loss = model.loss_fn(args, x)
    t1 = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    t2 = time.perf_counter()
    opt.step()
    t3 = time.perf_counter()
    if parts is not None:
        parts.append((t1 - t0, t2 - t1, t3 - t2))


for _ in range(4):
    step()
torch.cuda.synchronize()
# host enqueue time: let the GPU idle first, enqueue, do not wait
parts = []
t0 = time.perf_counter()
for _ in range(5):
    step(parts)
t_host = (time.perf_counter() - t0) / 5
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 5
f = sum(p[0] for p in parts) / 5 * 1e3
b = sum(p[1] for p in parts) / 5 * 1e3
o = sum(p[2] for p in parts) / 5 * 1e3
print(f"B={B}: host enqueue {t_host * 1e3:.2f} ms/step (loss_fn {f:.2f}, backward {b:.2f}, optimizer {o:.2f}); with final sync {t_all * 1e3:.2f} ms/step")
