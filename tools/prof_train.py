"""Kernel-time breakdown of one training step (torch profiler, CUDA activities)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from infodiffusion_b200.models import InfoDiff  # noqa: E402
from infodiffusion_b200 import train as T  # noqa: E402

B = 32
dev = "cuda:0"
args = bench.make_args_ns(1000)
torch.manual_seed(64)
model = InfoDiff(args, "cpu", (3, 64, 64)).to(dev)
model.device = dev
for n in ("alpha_bars", "betas", "alphas", "alpha_prev_bars"):
    setattr(model, n, getattr(model, n).to(dev))
model.train()
if os.environ.get("IDF_TRAIN_DROPOUT") is not None:      # what does the mask generator cost?
    model.backbone.dropout_p = model.encoder.dropout_p = float(os.environ["IDF_TRAIN_DROPOUT"])
params = [p for p in model.parameters() if p.requires_grad]
from infodiffusion_b200.optim import ClipAdamW  # noqa: E402
opt = ClipAdamW(params, lr=1e-4, weight_decay=1e-5, max_norm=1.0)      # fused clip_grad_norm_(1.0) + AdamW (run.py:199-200)
x = (torch.rand(B, 3, 64, 64, device=dev) * 2 - 1)
T.USE_GRAPHS = "--graphs" in sys.argv


def step():
    loss = model.loss_fn(args, x)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) / 5 * 1e3)
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
if not os.environ.get("IDF_NO_TABLE"):
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
