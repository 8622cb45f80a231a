"""One eager UNet evaluation + fused DDIM update at the benchmark shape, bracketed by
cudaProfilerStart/Stop (use with `ncu --profile-from-start off`)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from infodiffusion_b200.models import InfoDiff  # noqa: E402
from infodiffusion_b200.sampling import DiffusionProcess  # noqa: E402

B = int(os.environ.get("IDF_PROF_BATCH", "256"))
chunk = int(os.environ.get("IDF_SAMPLE_CHUNK", "0")) or None
dev = "cuda:0"
args = bench.make_args_ns(bench.T_STEPS)
args.sample_chunk = chunk
torch.manual_seed(64)
model = InfoDiff(args, "cpu", (3, 64, 64)).to(dev).eval()
model.device = dev
proc = DiffusionProcess(args, model, dev, (3, 64, 64))
s = proc._sampler("ddim", B)
s.x.normal_()
s.set_latent(torch.randn(B, bench.A_DIM, device=dev))
s.noise.normal_()
s.step.fill_(50)
for _ in range(2):
    s._enqueue()
torch.cuda.synchronize()
torch.cuda.profiler.start()
s._enqueue()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one UNet evaluation:", s.n_launch, "launches, batch", B, "chunk", chunk)
