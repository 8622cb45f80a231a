"""The CPU oracle reproduces the golden vectors minted from the unmodified reference
(oracle/make_golden.py).  This is what pins the oracle; no GPU needed."""
import numpy as np
import torch

from oracle import infodiff_oracle as orc
from oracle.golden_util import SEED, make_args, perturb_state_dict, rand_inputs, rel_l2, step_noise

_cache = {}


def model_sd(a_dim, T):
    """Weights are regenerated from the seed through our reference-identical constructors."""
    key = (a_dim, T)
    if key not in _cache:
        from infodiffusion_b200.models import InfoDiff
        torch.manual_seed(SEED)
        m = InfoDiff(make_args(a_dim=a_dim, diffusion_steps=T), "cpu", (3, 64, 64))
        _cache[key] = perturb_state_dict(m.state_dict())
    return _cache[key]


def test_backbone_eps(golden_dir):
    g = np.load(golden_dir / "backbone_a32_T1000.npz")
    sd = model_sd(32, 1000)
    x, t, a = rand_inputs(2, 32, 1000)
    with torch.no_grad():
        eps = orc.aux_unet_forward(sd, x, t, a)
    assert rel_l2(eps, torch.from_numpy(g["eps"])) < 1e-6
    assert float(eps.std()) > 0.1       # not the vacuous 1e-5-gain output


def test_encoder(golden_dir):
    g = np.load(golden_dir / "encoder_a32.npz")
    sd = model_sd(32, 1000)
    x, _, _ = rand_inputs(2, 32, 1000)
    noise = torch.randn(2, 32, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        a, a_q, mu, lv = orc.encoder_forward(sd, x, noise=noise)
    for name, v in (("a", a), ("a_q", a_q), ("mu", mu), ("log_var", lv)):
        assert rel_l2(v, torch.from_numpy(g[name])) < 1e-6, name


def test_mmd(golden_dir):
    g = np.load(golden_dir / "mmd.npz")
    for D in (32, 256):
        gen = torch.Generator().manual_seed(11 + D)
        xs = torch.randn(32, D, generator=gen)
        ys = (torch.randn(32, D, generator=gen) * 0.7 + 0.2).requires_grad_(True)
        v = orc.compute_mmd(xs, ys)
        (gr,) = torch.autograd.grad(v, ys)
        assert rel_l2(v.detach(), torch.from_numpy(g[f"v{D}"])) < 1e-6
        assert rel_l2(gr, torch.from_numpy(g[f"g{D}"])) < 1e-5


def test_loss_value(golden_dir):
    g = np.load(golden_dir / "loss_a32.npz")
    sd = model_sd(32, 1000)
    gl = torch.Generator().manual_seed(21)
    xb = torch.rand(4, 3, 64, 64, generator=gl) * 2 - 1
    idx = torch.randint(0, 1000, (4,), generator=gl)
    eps = torch.randn(4, 3, 64, 64, generator=gl)
    encn = torch.randn(4, 32, generator=gl)
    prior = torch.randn(4, 32, generator=gl)
    sch = orc.Schedule.make(1e-5, 1e-2, 1000)
    with torch.no_grad():
        terms = orc.infodiff_loss(sd, sch, xb, idx, eps, encn, prior, 0.1, 0.0, 1000)
    assert rel_l2(terms["loss"], torch.from_numpy(g["loss"])) < 1e-6


def test_sampler_trajectories(golden_dir):
    T = 10
    sd = model_sd(32, T)
    sch = orc.Schedule.make(1e-5, 1e-2, T)
    _, _, a = rand_inputs(2, 32, T, seed=8)
    xT = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(3))
    shape = tuple(xT.shape)
    for kind, det in (("ddim", True), ("ddpm", False)):
        g = np.load(golden_dir / f"{kind}10_a32.npz")
        rec = []
        orc.sample(sd, sch, xT, a, det, noise_fn=lambda i, like: step_noise(i, shape), record=rec)
        for idx, eps, x in rec:
            if f"x_{idx}" in g:
                assert rel_l2(x, torch.from_numpy(g[f"x_{idx}"])) < 2e-6, (kind, idx)
        rms = np.array([float(x.pow(2).mean().sqrt()) for _, _, x in rec])
        assert np.allclose(rms, g["x_rms"], rtol=1e-5)


def test_reverse_ddim(golden_dir):
    T = 10
    g = np.load(golden_dir / "reverse10_a32.npz")
    sd = model_sd(32, T)
    sch = orc.Schedule.make(1e-5, 1e-2, T)
    _, _, a = rand_inputs(2, 32, T, seed=8)
    x0 = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(4)) * 2 - 1
    xT = orc.reverse_sample(sd, sch, x0, a=None, enc_noise_fn=lambda xx: torch.zeros(2, 32))
    assert rel_l2(xT, torch.from_numpy(g["xT_reencode"])) < 2e-6
    xTa = orc.reverse_sample(sd, sch, x0, a=a)
    assert rel_l2(xTa, torch.from_numpy(g["xT_given_a"])) < 2e-6
    assert rel_l2(xTa, xT) > 1e-4      # the two variants really differ


def test_variant_goldens_and_constructor_digests(golden_dir):
    """Bottleneck / vanilla UNet / Diff / two-phase / latent: the product's constructors reproduce the reference's
    initial state bit for bit (digests minted by oracle/make_golden_variants.py) and the oracle reproduces the
    reference's outputs on the regenerated weights."""
    import json
    from infodiffusion_b200 import models
    from oracle.golden_util import state_digest
    meta = json.loads((golden_dir / "meta.json").read_text())
    x, t, a = rand_inputs(2, 32, 1000)

    args = make_args(a_dim=32, diffusion_steps=1000, is_bottleneck=True)
    torch.manual_seed(SEED)
    m = models.InfoDiff(args, "cpu", (3, 64, 64))
    assert state_digest(m.state_dict()) == meta["state_bottleneck_a32_T1000"]["digest"]
    sd = perturb_state_dict(m.state_dict())
    with torch.no_grad():
        eps = orc.bottleneck_unet_forward(sd, x, t, a)
    assert rel_l2(eps, torch.from_numpy(np.load(golden_dir / "bottleneck_a32_T1000.npz")["eps"])) < 1e-6

    torch.manual_seed(SEED)
    u = models.UNet(T=1000, ch=64, ch_mult=[1, 2, 2, 2], shape=(3, 64, 64))
    assert state_digest(u.state_dict()) == meta["state_unet_1222_T1000"]["digest"]
    sdu = perturb_state_dict({"backbone." + k: v for k, v in u.state_dict().items()})
    with torch.no_grad():
        eps = orc.unet_forward(sdu, x, t)
    assert rel_l2(eps, torch.from_numpy(np.load(golden_dir / "unet_1222_T1000.npz")["eps"])) < 1e-6

    args6 = make_args(a_dim=32, diffusion_steps=6, model="vanilla", split_step=2)
    torch.manual_seed(SEED + 1)
    van = models.Diff(args6, "cpu", (3, 64, 64))
    torch.manual_seed(SEED + 1)
    van.backbone = models.UNet(T=6, ch=64, ch_mult=[1, 2, 2, 2], shape=(3, 64, 64))
    assert state_digest(van.state_dict()) == meta["state_diff_unet_1222_T6_seed65"]["digest"]


def test_kld_control_constant_loss_golden(golden_dir):
    """InfoDiff.loss_fn with kld_weight != 0, use_C (reference models.py:648-668): oracle vs the reference's value."""
    from infodiffusion_b200.models import InfoDiff
    kw = dict(kld_weight=0.5, mmd_weight=0.1, use_C=True, C_max=25.0, epochs=4)
    args = make_args(a_dim=32, diffusion_steps=1000, **kw)
    torch.manual_seed(SEED)
    sd = perturb_state_dict(InfoDiff(args, "cpu", (3, 64, 64)).state_dict())
    gl = torch.Generator().manual_seed(23)
    xb = torch.rand(4, 3, 64, 64, generator=gl) * 2 - 1
    idx = torch.randint(0, 1000, (4,), generator=gl)
    eps = torch.randn(4, 3, 64, 64, generator=gl)
    encn = torch.randn(4, 32, generator=gl)
    prior = torch.randn(4, 32, generator=gl)
    sch = orc.Schedule.make(args.beta1, args.betaT, 1000)
    with torch.no_grad():
        terms = orc.infodiff_loss(sd, sch, xb, idx, eps, encn, prior, 0.1, 0.5, 1000, use_C=True, C_max=25.0, epochs=4, curr_epoch=2)
    g = np.load(golden_dir / "loss_kld_a32.npz")
    assert abs(float(terms["loss"]) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    assert abs(float(terms["kld"]) - float(g["kld"])) <= 1e-6 * abs(float(g["kld"]))
