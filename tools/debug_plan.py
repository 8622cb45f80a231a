"""Run the sampler plan op by op with a sync after each, report the first failing launch and its conv geometry."""
import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from infodiffusion_b200 import engine, _lib
from infodiffusion_b200.models import InfoDiff
from infodiffusion_b200.sampling import DiffusionProcess
B = int(os.environ.get("IDF_PROF_BATCH", "32"))
engine.FUSE_ADAGN = bool(os.environ.get("IDF_FUSE"))
dev = "cuda:0"
args = bench.make_args_ns(bench.T_STEPS)
torch.manual_seed(64)
model = InfoDiff(args, "cpu", (3, 64, 64)).to(dev).eval()
model.device = dev
proc = DiffusionProcess(args, model, dev, (3, 64, 64))
s = proc._sampler("ddim", B)
s.x.normal_(); s.set_latent(torch.randn(B, bench.A_DIM, device=dev)); s.noise.normal_(); s.step.fill_(50)
plan = s.plans[-1]
st = torch.cuda.current_stream().cuda_stream
convs = [d for d in plan.keep if isinstance(d, _lib.ConvDesc)]
ci = 0
for i, ((fn, a), m) in enumerate(zip(plan.ops, plan.meta)):
    desc = None
    if m["tag"].startswith("conv_igemm"):
        desc = convs[ci]; ci += 1
    try:
        _lib.check(fn(*a, st))
        torch.cuda.synchronize()
    except Exception as e:
        print("FAILED at op", i, m, repr(e)[:200])
        if desc is not None:
            print("conv: n_src", desc.n_src, "src_ld", list(desc.src_ld)[:desc.n_src], "num_kb", desc.num_kb, "H", desc.H, "cout", desc.cout,
                  "block_n", desc.block_n, "epi", desc.epilogue, "res", bool(desc.residual),
                  "kb", [(desc.kb_src[k], desc.kb_c0[k], desc.kb_rowoff[k], desc.kb_xf[k]) for k in range(desc.num_kb)])
        sys.exit(1)
print("all", len(plan.ops), "ops ok")
