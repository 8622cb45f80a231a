"""Pins the CPU oracle against the UNMODIFIED reference and writes tests/golden/*.npz.

Run in the build container only (the reference checkout is not available on the GPU box):

    python oracle/make_golden.py            # needs /root/reference

For every case it (1) runs the reference's own modules (models.py / sampling.py / utils.py imported
from /root/reference) on seeded inputs, (2) runs oracle/infodiff_oracle.py on the same inputs and
asserts agreement (bit-exact where the op order is identical, <= 2e-6 rel-L2 otherwise), and
(3) stores the REFERENCE outputs as the golden vectors.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import contextlib
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF))

import models as ref_models  # noqa: E402  (the reference)
import sampling as ref_sampling  # noqa: E402
import utils as ref_utils  # noqa: E402

from oracle import infodiff_oracle as orc  # noqa: E402
from oracle.golden_util import (SEED, make_args, perturb_state_dict, rand_inputs, rel_l2, state_digest,  # noqa: E402
                                step_noise)

OUT = ROOT / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)
torch.set_num_threads(8)


@contextlib.contextmanager
def patched_randn_like(queue):
    """Replace torch.randn_like by a FIFO of prepared tensors (reference draw order preserved)."""
    real = torch.randn_like

    def fake(t, **kw):
        v = queue.pop(0)
        assert v.shape == t.shape, (v.shape, t.shape)
        return v.to(t.dtype)
    torch.randn_like = fake
    try:
        yield
    finally:
        torch.randn_like = real


@contextlib.contextmanager
def silenced():
    import io
    old = sys.stdout
    sys.stdout = io.StringIO()
    try:
        yield
    finally:
        sys.stdout = old


def ref_model(a_dim, T, **kw):
    args = make_args(a_dim=a_dim, diffusion_steps=T, **kw)
    torch.manual_seed(SEED)
    m = ref_models.InfoDiff(args, "cpu", (3, 64, 64))
    raw_digest = state_digest(m.state_dict())
    sd = perturb_state_dict(m.state_dict())
    m.load_state_dict(sd, strict=True)
    m.eval()
    return args, m, sd, raw_digest


def check(name, got, want, tol):
    err = rel_l2(got, want)
    print(f"  oracle vs reference  {name:28s} rel-L2 = {err:.3e}")
    assert err <= tol, (name, err)


def main():
    meta = {}
    # ------------------------------------------------------------------ 1. constructor parity
    for a_dim in (32, 256):
        args = make_args(a_dim=a_dim, diffusion_steps=1000)
        torch.manual_seed(SEED)
        m = ref_models.InfoDiff(args, "cpu", (3, 64, 64))
        sd = m.state_dict()
        meta[f"state_a{a_dim}_T1000"] = dict(digest=state_digest(sd), nkeys=len(sd),
                                              nparams=int(sum(p.numel() for p in m.parameters())))
    print("constructor digests", meta)

    # ------------------------------------------------------------------ 2. backbone eps
    args, m, sd, dg = ref_model(32, 1000)
    x, t, a = rand_inputs(2, 32, 1000)
    with torch.no_grad():
        eps_ref = m.backbone(x, t, a)
        trace = {}
        eps_or = orc.aux_unet_forward(sd, x, t, a, trace=trace)
    check("backbone eps", eps_or, eps_ref, 1e-6)
    print("  eps std", float(eps_ref.std()))
    tr_names = sorted(trace)
    tr_rms = np.array([float(trace[k].pow(2).mean().sqrt()) for k in tr_names])
    np.savez_compressed(OUT / "backbone_a32_T1000.npz", eps=eps_ref.numpy(), trace_names=np.array(tr_names),
                        trace_rms=tr_rms)

    # ------------------------------------------------------------------ 3. encoder
    enc_noise = torch.randn(2, 32, generator=torch.Generator().manual_seed(5))
    with torch.no_grad(), patched_randn_like([enc_noise.clone()]):
        a_r, aq_r, mu_r, lv_r = m.encoder(x)
    with torch.no_grad():
        a_o, aq_o, mu_o, lv_o = orc.encoder_forward(sd, x, noise=enc_noise)
    for n, g, w in (("a", a_o, a_r), ("a_q", aq_o, aq_r), ("mu", mu_o, mu_r), ("log_var", lv_o, lv_r)):
        check("encoder " + n, g, w, 1e-6)
    np.savez_compressed(OUT / "encoder_a32.npz", a=a_r.numpy(), a_q=aq_r.numpy(), mu=mu_r.numpy(),
                        log_var=lv_r.numpy())

    # ------------------------------------------------------------------ 4. MMD value + grad
    mmd = {}
    for D in (32, 256):
        g = torch.Generator().manual_seed(11 + D)
        xs = torch.randn(32, D, generator=g)
        ys = (torch.randn(32, D, generator=g) * 0.7 + 0.2).requires_grad_(True)
        v_ref = ref_utils.compute_mmd(xs, ys)
        (g_ref,) = torch.autograd.grad(v_ref, ys)
        ys2 = ys.detach().clone().requires_grad_(True)
        v_or = orc.compute_mmd(xs, ys2)
        (g_or,) = torch.autograd.grad(v_or, ys2)
        check(f"mmd D={D} value", v_or, v_ref, 1e-6)
        check(f"mmd D={D} grad", g_or, g_ref, 1e-5)
        mmd[f"v{D}"] = v_ref.detach().numpy()
        mmd[f"g{D}"] = g_ref.numpy()
    np.savez_compressed(OUT / "mmd.npz", **mmd)

    # ------------------------------------------------------------------ 5. training loss (forward value)
    gl = torch.Generator().manual_seed(21)
    xb = torch.rand(4, 3, 64, 64, generator=gl) * 2 - 1
    idx = torch.randint(0, 1000, (4,), generator=gl)
    eps = torch.randn(4, 3, 64, 64, generator=gl)
    encn = torch.randn(4, 32, generator=gl)
    prior = torch.randn(4, 32, generator=gl)
    real_randint = torch.randint
    torch.randint = lambda *a_, **k_: idx.clone()
    try:
        with torch.no_grad(), silenced(), patched_randn_like([eps.clone(), encn.clone(), prior.clone()]):
            loss_ref = m.loss_fn(args, xb)
    finally:
        torch.randint = real_randint
    sch = orc.Schedule.make(args.beta1, args.betaT, args.diffusion_steps)
    with torch.no_grad():
        terms = orc.infodiff_loss(sd, sch, xb, idx, eps, encn, prior, args.mmd_weight, args.kld_weight,
                                  args.diffusion_steps)
    check("loss_fn value", terms["loss"], loss_ref, 1e-6)
    np.savez_compressed(OUT / "loss_a32.npz", loss=loss_ref.numpy(), eps_term=terms["eps"].numpy(),
                        rec_term=terms["rec"].numpy(), mmd_term=terms["mmd"].numpy())

    # ------------------------------------------------------------------ 6-8. sampler trajectories, T = 10
    T = 10
    args10, m10, sd10, _ = ref_model(32, T)
    sch10 = orc.Schedule.make(args10.beta1, args10.betaT, T)
    xT, _, a2 = rand_inputs(2, 32, T, seed=8)
    xT = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(3))
    shape = tuple(xT.shape)
    keep_steps = (T - 1, T // 2, 0)

    def run_ref(deterministic):
        args10.deterministic = deterministic
        proc = ref_sampling.DiffusionProcess(args10, m10, "cpu", (3, 64, 64))
        if deterministic:
            q = [step_noise(i, shape) for i in reversed(range(T)) if i > 0]
            gen = proc._ddim_one_diffusion_step(xT, a2)
        else:
            q = [step_noise(i, shape) for i in reversed(range(T)) if i > 0]
            gen = proc._ddpm_one_diffusion_step(xT, a2)
        xs = []
        with torch.no_grad(), patched_randn_like(q):
            for xx in gen:
                xs.append(xx)
        return xs

    for kind, det in (("ddim", True), ("ddpm", False)):
        xs_ref = run_ref(det)
        rec = []
        with torch.no_grad():
            orc.sample(sd10, sch10, xT, a2, det, noise_fn=lambda i, like: step_noise(i, shape), record=rec)
        out = {}
        for k, (idx_k, eps_k, x_k) in enumerate(rec):
            check(f"{kind} x after idx={idx_k}", x_k, xs_ref[k], 2e-6)
            if idx_k in keep_steps:
                out[f"x_{idx_k}"] = xs_ref[k].numpy()
                out[f"eps_{idx_k}"] = eps_k.numpy()     # oracle eps (the reference does not expose it)
        out["x_rms"] = np.array([float(v.pow(2).mean().sqrt()) for v in xs_ref])
        np.savez_compressed(OUT / f"{kind}10_a32.npz", **out)

    # reverse DDIM: the reference drops `a` and re-encodes x_t every step (sampling.py:84)
    x0 = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(4)) * 2 - 1
    args10.deterministic = True
    proc = ref_sampling.DiffusionProcess(args10, m10, "cpu", (3, 64, 64))
    encq = [torch.zeros(2, 32) for _ in range(T)]   # randn_like(mu) draws inside the encoder; a_q unused (kld=0)
    with torch.no_grad(), patched_randn_like(encq):
        xT_ref = proc.reverse_sampling(x0, a2)
    with torch.no_grad():
        xT_or = orc.reverse_sample(sd10, sch10, x0, a=None, enc_noise_fn=lambda xx: torch.zeros(2, 32))
        xT_or_a = orc.reverse_sample(sd10, sch10, x0, a=a2)
    check("reverse ddim (re-encode)", xT_or, xT_ref, 2e-6)
    np.savez_compressed(OUT / "reverse10_a32.npz", xT_reencode=xT_ref.numpy(), xT_given_a=xT_or_a.numpy())

    (OUT / "meta.json").write_text(json.dumps(meta, indent=1))
    print("golden files written to", OUT)


if __name__ == "__main__":
    main()
