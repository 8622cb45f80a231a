"""Golden vectors for the remaining network / sampler variants of the hot path (SURVEY section 8a rows a7,
a10, a12, a14, a16, a22, a23): BottleneckAuxUNet, the vanilla UNet, the Diff wrapper, the two-phase sampler
and the latent (MLP) diffusion path.  Same contract as oracle/make_golden.py: run the reference's own modules
from /root/reference on seeded inputs, assert that oracle/infodiff_oracle.py agrees, store the REFERENCE
outputs under tests/golden/.  Build container only.  TEST INFRASTRUCTURE ONLY.

One documented patch of the reference is needed: UNet.__init__ passes `crossattn=False` to ResBlock, whose
constructor does not take it (models.py:32-33 vs modules.py:207), so the vanilla UNet cannot be constructed at
HEAD.  The script substitutes a ResBlock subclass that accepts and ignores the keyword; nothing
else of the reference is touched.
"""
from __future__ import annotations

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
import modules as ref_modules  # noqa: E402
import sampling as ref_sampling  # noqa: E402

from oracle import infodiff_oracle as orc  # noqa: E402
from oracle.golden_util import SEED, make_args, perturb_state_dict, rand_inputs, rel_l2, state_digest, step_noise  # noqa: E402
from oracle.make_golden import check, patched_randn_like  # noqa: E402

OUT = ROOT / "tests" / "golden"
torch.set_num_threads(8)


class _ResBlockTakingCrossattn(ref_modules.ResBlock):
    def __init__(self, in_ch, out_ch, tdim, dropout, attn=False, crossattn=False):
        super().__init__(in_ch, out_ch, tdim, dropout, attn=attn)


def main():
    meta_path = OUT / "meta.json"
    meta = json.loads(meta_path.read_text())
    ref_models.ResBlock = _ResBlockTakingCrossattn        # used by UNet.__init__ and by isinstance() in forward
    x, t, a = rand_inputs(2, 32, 1000)

    # ------------------------------------------------------------------ BottleneckAuxUNet (a12) inside InfoDiff
    args = make_args(a_dim=32, diffusion_steps=1000, is_bottleneck=True)
    torch.manual_seed(SEED)
    m = ref_models.InfoDiff(args, "cpu", (3, 64, 64))
    meta["state_bottleneck_a32_T1000"] = dict(digest=state_digest(m.state_dict()), nkeys=len(m.state_dict()))
    sd = perturb_state_dict(m.state_dict())
    m.load_state_dict(sd, strict=True)
    m.eval()
    with torch.no_grad():
        eps_ref = m.backbone(x, t, a)
        eps_or = orc.bottleneck_unet_forward(sd, x, t, a)
    check("bottleneck eps", eps_or, eps_ref, 1e-6)
    np.savez_compressed(OUT / "bottleneck_a32_T1000.npz", eps=eps_ref.numpy())

    # ------------------------------------------------------------------ vanilla UNet (a7, a10), InfoDiff widths
    torch.manual_seed(SEED)
    u = ref_models.UNet(T=1000, ch=64, ch_mult=[1, 2, 2, 2], shape=(3, 64, 64))
    meta["state_unet_1222_T1000"] = dict(digest=state_digest(u.state_dict()), nkeys=len(u.state_dict()))
    sdu = perturb_state_dict({"backbone." + k: v for k, v in u.state_dict().items()})
    u.load_state_dict({k[len("backbone."):]: v for k, v in sdu.items()}, strict=True)
    u.eval()
    with torch.no_grad():
        eps_ref = u(x, t)
        eps_or = orc.unet_forward(sdu, x, t)
    check("unet eps", eps_or, eps_ref, 1e-6)
    np.savez_compressed(OUT / "unet_1222_T1000.npz", eps=eps_ref.numpy())

    # ------------------------------------------------------------------ vanilla UNet at the widths Diff hard-wires
    # (models.py:746, ch_mult [1,2,4,8]: 512 channels, GroupNorm over 1024, attention heads of 256 and 512)
    torch.manual_seed(SEED)
    uw = ref_models.UNet(T=1000, ch=64, ch_mult=[1, 2, 4, 8], shape=(3, 64, 64))
    meta["state_unet_1248_T1000"] = dict(digest=state_digest(uw.state_dict()), nkeys=len(uw.state_dict()))
    sdw = perturb_state_dict({"backbone." + k: v for k, v in uw.state_dict().items()})
    uw.load_state_dict({k[len("backbone."):]: v for k, v in sdw.items()}, strict=True)
    uw.eval()
    with torch.no_grad():
        eps_ref = uw(x, t)
        eps_or = orc.unet_forward(sdw, x, t)
    check("unet [1,2,4,8] eps", eps_or, eps_ref, 1e-6)
    np.savez_compressed(OUT / "unet_1248_T1000.npz", eps=eps_ref.numpy())
    del uw, sdw

    # ------------------------------------------------------------------ Diff wrapper (a16) + two-phase sampler (a22)
    T = 6
    args6 = make_args(a_dim=32, diffusion_steps=T, model="vanilla", split_step=2)
    torch.manual_seed(SEED)
    info = ref_models.InfoDiff(args6, "cpu", (3, 64, 64))
    sd1 = perturb_state_dict(info.state_dict())
    info.load_state_dict(sd1, strict=True)
    info.eval()
    torch.manual_seed(SEED + 1)
    van = ref_models.Diff(args6, "cpu", (3, 64, 64))       # builds the [1,2,4,8] UNet; swap in the InfoDiff widths
    torch.manual_seed(SEED + 1)
    van.backbone = ref_models.UNet(T=T, ch=64, ch_mult=[1, 2, 2, 2], shape=(3, 64, 64))
    meta["state_diff_unet_1222_T6_seed65"] = dict(digest=state_digest(van.state_dict()), nkeys=len(van.state_dict()))
    sd2 = perturb_state_dict(van.state_dict(), seed=4321)
    van.load_state_dict(sd2, strict=True)
    van.eval()
    xT = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(3))
    _, _, a2 = rand_inputs(2, 32, T, seed=8)
    shape = tuple(xT.shape)
    sch = orc.Schedule.make(args6.beta1, args6.betaT, T)
    with torch.no_grad():
        e_ref = van(xT, 3)
        e_or = orc.vanilla_eps_fn(sd2)(xT, 3)
    check("Diff.forward(x, idx)", e_or, e_ref, 1e-6)
    out = {"diff_eps_idx3": e_ref.numpy()}
    for kind, det in (("ddim", True), ("ddpm", False)):
        args6.deterministic = det
        proc = ref_sampling.TwoPhaseDiffusionProcess(args6, info, van, "cpu", (3, 64, 64))
        q = [step_noise(i, shape) for i in reversed(range(T)) if i > 0]
        with torch.no_grad(), patched_randn_like(q):
            x_ref = proc.sampling(2, xT=xT, a=a2)
        x_or = orc.two_phase_sample(orc.infodiff_eps_fn(sd1, a2), orc.vanilla_eps_fn(sd2), sch, xT, det, args6.split_step,
                                    noise_fn=lambda i, like: step_noise(i, shape))
        check(f"two-phase {kind} (bug-compatible)", x_or, x_ref, 2e-6)
        x_fix = orc.two_phase_sample(orc.infodiff_eps_fn(sd1, a2), orc.vanilla_eps_fn(sd2), sch, xT, det,
                                     args6.split_step, noise_fn=lambda i, like: step_noise(i, shape), fixed=True)
        out[f"{kind}_x0"] = x_ref.numpy()
        out[f"{kind}_x0_fixed"] = x_fix.numpy()             # oracle only: the reference cannot produce it
    np.savez_compressed(OUT / "twophase6_a32.npz", **out)

    # ------------------------------------------------------------------ latent path (a14, a16, a23)
    T = 10
    D = 32
    argsl = make_args(a_dim=D, diffusion_steps=T, model="vanilla", is_latent=True)
    torch.manual_seed(SEED)
    lat = ref_models.Diff(argsl, "cpu", (1, D, D))
    meta["state_latent_a32_T10"] = dict(digest=state_digest(lat.state_dict()), nkeys=len(lat.state_dict()))
    lat.load_state_dict(perturb_state_dict(lat.state_dict()), strict=True)
    # MLPLNAct registers linear_emb twice (`linear_emb` and `cond_layers.1`, models.py:113-115): the two keys alias
    # one tensor, so read the state back after loading to get one consistent set of values
    sdl = {k: v.clone() for k, v in lat.state_dict().items()}
    lat.eval()
    g = torch.Generator().manual_seed(17)
    z = torch.randn(5, D, generator=g)
    tz = torch.randint(0, T, (5,), generator=g)
    with torch.no_grad():
        e_ref = lat.backbone(z, tz)
        e_or = orc.latent_unet_forward(sdl, z, tz)
    check("latent unet eps", e_or, e_ref, 1e-6)
    out = {"eps": e_ref.numpy()}
    zshape = tuple(z.shape)
    for kind, det in (("ddim", True), ("ddpm", False)):
        argsl.deterministic = det
        proc = ref_sampling.LatentDiffusionProcess(argsl, lat, "cpu")
        q = [step_noise(i, zshape) for i in reversed(range(T)) if i > 0]
        with torch.no_grad(), patched_randn_like(q):
            z_ref = proc.sampling(5, xT=z)
        z_or = orc.latent_sample(sdl, orc.Schedule.make(argsl.beta1, argsl.betaT, T), z, det,
                                 noise_fn=lambda i, like: step_noise(i, zshape))
        check(f"latent {kind} z0", z_or, z_ref, 2e-6)
        out[f"{kind}_z0"] = z_ref.numpy()
    np.savez_compressed(OUT / "latent10_a32.npz", **out)

    # ------------------------------------------------------------------ loss_fn with the KLD term and the control constant
    argsk = make_args(a_dim=32, diffusion_steps=1000, kld_weight=0.5, mmd_weight=0.1, use_C=True, C_max=25.0, epochs=4)
    torch.manual_seed(SEED)
    mk = ref_models.InfoDiff(argsk, "cpu", (3, 64, 64))
    sdk = perturb_state_dict(mk.state_dict())
    mk.load_state_dict(sdk, strict=True)
    mk.eval()
    gl = torch.Generator().manual_seed(23)
    xb = torch.rand(4, 3, 64, 64, generator=gl) * 2 - 1
    idx = torch.randint(0, 1000, (4,), generator=gl)
    eps = torch.randn(4, 3, 64, 64, generator=gl)
    encn = torch.randn(4, 32, generator=gl)
    prior = torch.randn(4, 32, generator=gl)
    real_randint = torch.randint
    torch.randint = lambda *a_, **k_: idx.clone()
    try:
        from oracle.make_golden import silenced
        with torch.no_grad(), silenced(), patched_randn_like([eps.clone(), encn.clone(), prior.clone()]):
            loss_ref = mk.loss_fn(argsk, xb, curr_epoch=2)
    finally:
        torch.randint = real_randint
    sch = orc.Schedule.make(argsk.beta1, argsk.betaT, 1000)
    with torch.no_grad():
        terms = orc.infodiff_loss(sdk, sch, xb, idx, eps, encn, prior, 0.1, 0.5, 1000, use_C=True, C_max=25.0, epochs=4,
                                  curr_epoch=2)
    check("loss_fn kld + use_C", terms["loss"], loss_ref, 1e-6)
    np.savez_compressed(OUT / "loss_kld_a32.npz", loss=loss_ref.numpy(), kld=terms["kld"].numpy(), mmd=terms["mmd"].numpy())

    meta_path.write_text(json.dumps(meta, indent=1))
    print("variant golden files written to", OUT)


if __name__ == "__main__":
    main()
