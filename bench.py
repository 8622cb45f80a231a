"""Headline benchmark: 100-step DDIM sampling of 64x64 images (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one complete DDIM-100 sampling pass over one batch of 256 synthetic images per GPU
(InfoDiff a_dim=256, T=100, random-init weights, xT and a ~ N(0,1)).  Sampling is per-sample
independent, so N GPUs run N batch shards with no data-path collective until the final gather
("scaling": "weak": 256 images per GPU, 256*N per job).

 value  : images/s with xT and a already resident in HBM (device-timed with CUDA events, max over ranks)
 e2e    : same metric through the public API (DiffusionProcess.sampling) with pinned HOST inputs and a
          host read-back of the samples inside the timed region
 roofline     : dominant kernel class (tcgen05 implicit-GEMM conv), algorithmic FLOPs / CUDA-event time
 cpu_baseline : the CPU oracle (oracle/, a port of the reference's PyTorch path) on this box's cores,
                bounded sample, extrapolated to DDIM-100
--impl reference times that CPU path alone and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

A_DIM, T_STEPS, BATCH = 256, 100, 256
METRIC = "ddim100_64x64_images_per_sec"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("IDF_SAMPLE_CHUNK", "0")) or None)
    ap.add_argument("--lanes", type=int, default=int(os.environ.get("IDF_SAMPLE_LANES", "1")),
                    help="streams the micro-batches of a step are spread over (needs --chunk < batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pdl", type=int, default=int(os.environ.get("IDF_PDL", "0")),
                    help="1: launch conv / AdaGN kernels with programmatic dependent launch")
    ap.add_argument("--fuse-adagn", action="store_true", help="fold every AdaGN into its consumer conv (A/B comparison)")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-throughput measurement")
    ap.add_argument("--train-batch", type=int, default=32)
    return ap.parse_args()


def ncu_traffic(kind: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the captured kernel, from the newest committed
    `ncu --set full` raw page under profiles/ (None if there is none)."""
    import csv
    files = sorted((ROOT / "profiles").glob(f"*_ncu_{kind}_raw.csv"))
    if not files:
        return None, None
    rows = list(csv.reader(files[-1].open()))
    if len(rows) < 3:
        return None, None
    hdr, units = rows[0], rows[1]
    try:
        ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    except ValueError:
        return None, None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = [float(r[ir]) * scale.get(units[ir], 1.0) + float(r[iw]) * scale.get(units[iw], 1.0) for r in rows[2:]]
    return sum(vals) / len(vals), f"{files[-1].name}: mean over {len(vals)} captured launches of {rows[2][ik].split('(')[0]}"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


def make_args_ns(T):
    import types
    return types.SimpleNamespace(beta1=1e-5, betaT=1e-2, diffusion_steps=T, input_size=64, is_bottleneck=False,
                                 unets_channels=64, encoder_channels=64, a_dim=A_DIM, mmd_weight=0.1, kld_weight=0.0,
                                 is_latent=False, mode="eval_fid", prior="regular", batch_size=BATCH, use_C=False,
                                 C_max=25.0, epochs=1, deterministic=True, model="diff", split_step=0)


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU path (oracle port of the reference), bounded sample -> DDIM-100 images/s
# ------------------------------------------------------------------------------------------------
def cpu_ddim_rate(batch: int, min_seconds: float, max_evals: int):
    from oracle import infodiff_oracle as orc
    from infodiffusion_b200.models import InfoDiff
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(64)
    m = InfoDiff(make_args_ns(T_STEPS), "cpu", (3, 64, 64))
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    sch = orc.Schedule.make(1e-5, 1e-2, T_STEPS)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(batch, 3, 64, 64, generator=g)
    a = torch.randn(batch, A_DIM, generator=g)
    fn = orc.infodiff_eps_fn(sd, a)
    it = orc.ddim_steps(sch, fn, x)
    with torch.no_grad():
        next(it)                                     # warm-up evaluation (oneDNN primitive creation)
        t0 = time.perf_counter()
        n = 0
        for _ in it:
            n += 1
            if n >= max_evals or time.perf_counter() - t0 >= min_seconds:
                break
        dt = time.perf_counter() - t0
    per_eval = dt / n
    return batch / (per_eval * T_STEPS), cores, f"{n} DDIM steps of the T={T_STEPS} schedule at batch {batch} ({dt:.1f} s), x{T_STEPS}/{n} extrapolated"


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, sample, cores = [], "", 1
    for i in range(a.warmup + a.steps):
        v, cores, sample = cpu_ddim_rate(32, 6.0, 12)      # one step = a bounded sample: <= 6 s or 12 UNet evaluations
        if i >= a.warmup:
            vals.append(v)
    v = statistics.mean(vals)
    ms = 1000.0 * a.batch / v
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"DDIM-{T_STEPS} sampling, 64x64x3, a_dim {A_DIM}, CPU oracle port of the reference path"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------
# training throughput (secondary metric)
# ------------------------------------------------------------------------------------------------
def train_throughput(a, dev, world, rank, steps=6, warmup=3):
    import torch.distributed as dist
    from infodiffusion_b200 import _lib
    from infodiffusion_b200.models import InfoDiff
    from infodiffusion_b200.optim import ClipAdamW
    from infodiffusion_b200.train import GradSync, set_grad_sync
    B = a.train_batch
    args = make_args_ns(1000)
    args.mode = "train"
    torch.manual_seed(64)
    model = InfoDiff(args, "cpu", (3, 64, 64)).to(dev)
    model.device = dev
    for n in ("alpha_bars", "betas", "alphas", "alpha_prev_bars"):
        setattr(model, n, getattr(model, n).to(dev))
    model.train()
    params = [p for p in model.parameters() if p.requires_grad]
    opt = ClipAdamW(params, lr=1e-4, weight_decay=1e-5, max_norm=1.0)     # clip_grad_norm_(1.0) + AdamW (run.py:199-200)
    sync = GradSync(world) if world > 1 else None
    set_grad_sync(sync)
    g = torch.Generator().manual_seed(7 + rank)
    x_h = (torch.rand(B, 3, 64, 64, generator=g) * 2 - 1).pin_memory()

    def step():
        x = x_h.to(dev, non_blocking=True)
        loss = model.loss_fn(args, x)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if sync is not None:
            sync.finish(params)                  # all-reduces were started inside backward
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = _lib.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    set_grad_sync(None)
    flops = 3 * (15.318e9 + 11.774e9) * B * world * steps      # fwd + bwd ~ 3 x fwd
    return {"metric": "train_samples_per_sec", "value": world * B * steps / (ms / 1000.0), "unit": "samples/s",
            "ms_per_step": ms / steps, "batch_per_gpu": B, "n_gpus": world, "final_loss": float(loss.detach()),
            "tflops": flops / (ms / 1000.0) / 1e12, "idf_launches_per_step": (_lib.launches() - l0) // steps,
            "config": "InfoDiff a_dim 256, T=1000, bf16 kernels / fp32 params, dropout 0.1, fused clip(1.0)+AdamW(1e-4, "
                      "wd 1e-5), data-parallel all-reduce overlapped with backward" + (" (NCCL)" if world > 1 else " (single GPU)")}


# ------------------------------------------------------------------------------------------------
# GPU path
# ------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch.distributed as dist
    from infodiffusion_b200 import _lib
    from infodiffusion_b200.models import InfoDiff
    from infodiffusion_b200.sampling import DiffusionProcess

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product path has no CPU fallback)"
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner to fd 1 while the communicator is
        # created, so fd 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(dev))
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    B = a.batch
    args = make_args_ns(T_STEPS)
    args.sample_chunk = a.chunk
    args.sample_lanes = a.lanes
    torch.manual_seed(64)
    model = InfoDiff(args, "cpu", (3, 64, 64)).to(dev).eval()
    model.device = dev
    proc = DiffusionProcess(args, model, dev, (3, 64, 64))
    g = torch.Generator().manual_seed(1000 + rank)
    xT_h = torch.randn(B, 3, 64, 64, generator=g).pin_memory()
    a_h = torch.randn(B, A_DIM, generator=g).pin_memory()
    xT_d, a_d = xT_h.to(dev), a_h.to(dev)
    out_h = torch.empty(B, 3, 64, 64).pin_memory()
    gathered = torch.empty(world * B, 3, 64, 64, device=dev) if world > 1 else None

    def job_device():
        x = proc.sampling(B, xT=xT_d, a=a_d)
        if world > 1:
            dist.all_gather_into_tensor(gathered, x)
        return x

    def job_e2e():
        x = proc.sampling(B, xT=xT_h.to(dev, non_blocking=True), a=a_h.to(dev, non_blocking=True))
        out_h.copy_(x, non_blocking=True)
        return x

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(job, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launches()
        e0.record()
        for _ in range(k):
            job()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, _lib.launches() - l0

    for _ in range(max(a.warmup, 3)):
        job_device()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, launches = timed(job_device, a.steps)
    job_e2e()
    ms_e2e, _ = timed(job_e2e, a.steps)
    clk = clocks.stop() if rank == 0 else None
    value = world * B * a.steps / (ms / 1000.0)
    e2e = world * B * a.steps / (ms_e2e / 1000.0)

    # ---- per-kernel-class timing of one UNet evaluation (CUDA events around every launch)
    roof, breakdown = None, None
    if rank == 0:
        pk = peaks()
        s = proc._sampler("ddim", B)
        for _ in range(2):
            for p in s.plans:
                p.run()
        torch.cuda.synchronize()
        acc = {}
        reps = 3
        for _ in range(reps):
            for p in s.plans:
                for tag, t_ms, fl, by in p.run_timed():
                    d = acc.setdefault(tag, dict(ms=0.0, flops=0, bytes=0, n=0))
                    d["ms"] += t_ms; d["flops"] += fl; d["bytes"] += by; d["n"] += 1
        tot = sum(d["ms"] for d in acc.values())
        c = acc["conv_igemm"]
        tfs = c["flops"] / (c["ms"] * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "conv_igemm_kernel (tcgen05 implicit GEMM)", "achieved": tfs,
                "peak": pk["tf_sus"], "unit": "TFLOP/s", "frac": tfs / pk["tf_sus"], "traffic": ncu_traffic("conv")[0],
                "traffic_source": ncu_traffic("conv")[1],
                "peak_source": pk["src"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                "share_of_step": c["ms"] / tot, "launches_per_unet_eval": c["n"] // reps,
                "avg_launch_ms": c["ms"] / c["n"]}
        breakdown = {k: {"ms_per_unet_eval": v["ms"] / reps, "launches": v["n"] // reps} for k, v in acc.items()}
        if "adagn" in acc:       # the stand-alone AdaGN kernels, HBM-bound (default; --fuse-adagn folds them into the convs)
            g_ = acc["adagn"]
            gbs = g_["bytes"] / (g_["ms"] * 1e-3) / 1e9
            breakdown["adagn"].update({"bound": "hbm", "achieved_GBps": gbs, "peak_GBps": pk["hbm"], "frac": gbs / pk["hbm"],
                                       "traffic": ncu_traffic("adagn")[0], "traffic_source": ncu_traffic("adagn")[1]})
        else:
            roof["note"] = ("AdaGN+SiLU is applied to the conv's A operand in shared memory (transform warps): the 73 "
                            "normalised activations are never written to or read from HBM; only the per-image "
                            "coefficient kernels (adagn_coef) remain")
    # ---- secondary metric: training throughput (BASELINE configs[2]: a_dim 256, T=1000, batch 32/GPU,
    #      loss_fn + backward + grad all-reduce + clip_grad_norm + AdamW), through the public API
    train = None
    if not a.no_train:
        train = train_throughput(a, dev, world, rank)
    cpu = None
    if rank == 0 and not a.no_cpu_baseline:
        v, cores, sample = cpu_ddim_rate(32, 15.0, 40)     # ~15 s of host work on all cores
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        ws_gb = sum(w.bytes for w in proc._sampler("ddim", B).lane_ws) / 1e9
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"eval_fid-style DDIM-{T_STEPS} sampling, batch {B}/GPU, 64x64x3, InfoDiff a_dim {A_DIM} "
                                   f"(BASELINE configs[1]); random-init weights",
                       "global_batch": world * B, "parallelism": f"batch-sharded x{world}, final all_gather",
                       "sample_chunk": a.chunk or B, "sample_lanes": a.lanes,
                       "l2": f"no explicit flush: {ws_gb:.1f} GB of activations are rewritten per UNet evaluation "
                             f"(>> 126 MB L2) and every step consumes the previous step's output"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": xT_h.numel() * 4 + a_h.numel() * 4,
                    "d2h_bytes_per_step": out_h.numel() * 4},
            "gpu_launches": launches, "clocks": clk, "roofline": roof, "kernel_breakdown": breakdown,
            "cpu_baseline": cpu, "train": train,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    # stdout carries exactly one JSON line: NCCL's version / debug banner (NCCL_DEBUG=VERSION|INFO) goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        if a.fuse_adagn:
            from infodiffusion_b200 import engine
            engine.FUSE_ADAGN = True
        if os.environ.get("IDF_XF_DEBUG"):
            from infodiffusion_b200 import _lib
            _lib.check(_lib.load().idf_set_option(b"xf_debug", int(os.environ["IDF_XF_DEBUG"])))
        if a.pdl:
            from infodiffusion_b200 import _lib
            _lib.check(_lib.load().idf_set_option(b"pdl", 1))
        run_ours(a)
