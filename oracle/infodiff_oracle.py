"""CPU oracle for the InfoDiffusion denoising hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (plain torch-CPU tensor arithmetic driven by a
``state_dict``) of the reference's AdaNorm-conditioned UNet, its Encoder, the MMD prior
loss, the training loss and the DDPM / DDIM / reverse-DDIM step loops.  Each function
cites the reference file:line it follows (paths relative to the reference checkout).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module, and only as the checker / the timed CPU baseline.
The product path (``infodiffusion_b200``) never imports it and has no CPU fallback.

Parity status: the reference ships no tests, golden vectors or known-answer fixtures
 => "parity unpinned by the reference".  The oracle is
instead pinned against the reference ITSELF: ``oracle/make_golden.py`` imports the
unmodified reference modules from /root/reference in the build container, checks every
function below against them on seeded inputs (bit-exact or <=1e-6), and commits the
resulting vectors under ``tests/golden/``.

The arithmetic itself (conv2d, group_norm, softmax, linear) lives in PyTorch/ATen, a
third-party dependency the reference does not pin (no requirements.txt is shipped); the
container pin is torch 2.11.0.  The oracle calls the same ATen CPU ops the reference's
nn.Modules dispatch to, in the same order, in fp32 (or fp64 when asked).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------
@dataclass
class UNetCfg:
    """Static structure of AuxiliaryUNet / Encoder (models.py:238, 425; InfoDiff models.py:619-627)."""
    T: int = 1000
    ch: int = 64
    ch_mult: Tuple[int, ...] = (1, 2, 2, 2)
    attn: Tuple[int, ...] = (2,)
    num_res_blocks: int = 2
    a_dim: int = 32
    shape: Tuple[int, int, int] = (3, 64, 64)


# --------------------------------------------------------------------------------------
# leaf arithmetic
# --------------------------------------------------------------------------------------
def sinusoid_table(T: int, d_model: int) -> Tensor:
    """Frozen [T, d_model] table with interleaved (sin, cos) pairs -- modules.py:13-20."""
    freq = torch.arange(0, d_model, step=2) / torch.Tensor([d_model]) * math.log(10000)
    freq = torch.exp(-freq)
    ang = torch.arange(T).float()[:, None] * freq[None, :]
    return torch.stack([torch.sin(ang), torch.cos(ang)], dim=-1).view(T, d_model)


def _gn(x: Tensor, sd: SD, key: str) -> Tensor:
    # nn.GroupNorm(32, C), eps=1e-5 -- modules.py:132,214,219,225,265,278,284,335,340
    return F.group_norm(x, 32, sd[key + ".weight"], sd[key + ".bias"], 1e-5)


def _conv(x: Tensor, sd: SD, key: str, stride: int = 1, padding: int = 1) -> Tensor:
    return F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], stride=stride, padding=padding)


def _lin(x: Tensor, sd: SD, key: str) -> Tensor:
    return F.linear(x, sd[key + ".weight"], sd[key + ".bias"])


def time_embedding(sd: SD, pfx: str, t: Tensor) -> Tensor:
    """Embedding lookup -> Linear -> SiLU -> Linear -- modules.py:22-27, 36-38."""
    e = sd[pfx + "timembedding.0.weight"][t]
    e = _lin(e, sd, pfx + "timembedding.1")
    e = F.silu(e)
    return _lin(e, sd, pfx + "timembedding.3")


def attn_block(sd: SD, pfx: str, x: Tensor) -> Tensor:
    """Single-head self attention over H*W tokens with residual -- modules.py:145-164."""
    B, C, H, W = x.shape
    h = _gn(x, sd, pfx + "group_norm")
    q = _conv(h, sd, pfx + "proj_q", padding=0)
    k = _conv(h, sd, pfx + "proj_k", padding=0)
    v = _conv(h, sd, pfx + "proj_v", padding=0)
    q = q.permute(0, 2, 3, 1).reshape(B, H * W, C)
    k = k.reshape(B, C, H * W)
    w = torch.bmm(q, k) * (int(C) ** (-0.5))
    w = F.softmax(w, dim=-1)
    v = v.permute(0, 2, 3, 1).reshape(B, H * W, C)
    h = torch.bmm(w, v).reshape(B, H, W, C).permute(0, 3, 1, 2)
    h = _conv(h, sd, pfx + "proj", padding=0)
    return x + h


def _shortcut(sd: SD, pfx: str, x: Tensor) -> Tensor:
    if pfx + "shortcut.weight" in sd:  # 1x1 conv when in_ch != out_ch -- modules.py:290-293
        return _conv(x, sd, pfx + "shortcut", padding=0)
    return x


def aux_res_block(sd: SD, pfx: str, x: Tensor, temb: Tensor, aemb: Optional[Tensor]) -> Tensor:
    """AuxResBlock (aemb given) / ResBlock (aemb None) in eval mode -- modules.py:309-328, 247-258.

    Dropout (modules.py:280,286) is the identity in eval mode; parity runs use eval mode
    (SURVEY.md H4).
    """
    h = _conv(F.silu(_gn(x, sd, pfx + "block1.0")), sd, pfx + "block1.2")
    t_out = _lin(F.silu(temb), sd, pfx + "temb_proj.1")[:, :, None, None]
    scale, shift = torch.chunk(t_out, 2, dim=1)
    h = _gn(h, sd, pfx + "block2.0") * (1 + scale) + shift
    if aemb is not None and pfx + "aemb_proj.1.weight" in sd:     # plain ResBlock has no aemb_proj (modules.py:206-258)
        a_out = _lin(F.silu(aemb), sd, pfx + "aemb_proj.1")[:, :, None, None]
        scale, shift = torch.chunk(a_out, 2, dim=1)
        h = h * (1 + scale) + shift
    h = _conv(F.silu(h), sd, pfx + "block2.3")
    h = _conv(F.silu(_gn(h, sd, pfx + "block3.0")), sd, pfx + "block3.3")
    h = h + _shortcut(sd, pfx, x)
    if pfx + "attn.proj.weight" in sd:
        h = attn_block(sd, pfx + "attn.", h)
    return h


def res_block_encoder(sd: SD, pfx: str, x: Tensor) -> Tensor:
    """ResBlock_encoder in eval mode -- modules.py:361-366."""
    h = _conv(F.silu(_gn(x, sd, pfx + "block1.0")), sd, pfx + "block1.2")
    h = _conv(F.silu(_gn(h, sd, pfx + "block2.0")), sd, pfx + "block2.3")
    h = h + _shortcut(sd, pfx, x)
    if pfx + "attn.proj.weight" in sd:
        h = attn_block(sd, pfx + "attn.", h)
    return h


def _count(sd: SD, pfx: str) -> int:
    idx = {int(k[len(pfx):].split(".")[0]) for k in sd if k.startswith(pfx)}
    return (max(idx) + 1) if idx else 0


# --------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------
def _unet_body(sd: SD, pfx: str, x: Tensor, temb: Tensor, aemb: Optional[Tensor],
               trace: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """head -> down -> middle -> up (cat skip) -> tail, shared by UNet / AuxiliaryUNet / BottleneckAuxUNet
    (models.py:62-88, 296-326, 391-421).  Blocks without `aemb_proj` ignore `aemb`."""
    h = _conv(x, sd, pfx + "head")
    hs = [h]
    if trace is not None:
        trace["head"] = h
    for i in range(_count(sd, pfx + "downblocks.")):       # models.py:307-309
        p = f"{pfx}downblocks.{i}."
        if p + "main.weight" in sd:                        # DownSample, modules.py:73-75
            h = _conv(h, sd, p + "main", stride=2)
        else:
            h = aux_res_block(sd, p, h, temb, aemb)
        hs.append(h)
        if trace is not None:
            trace[f"down{i}"] = h
    for i in range(_count(sd, pfx + "middleblocks.")):     # models.py:312-316
        h = aux_res_block(sd, f"{pfx}middleblocks.{i}.", h, temb, aemb)
        if trace is not None:
            trace[f"mid{i}"] = h
    for i in range(_count(sd, pfx + "upblocks.")):         # models.py:319-322
        p = f"{pfx}upblocks.{i}."
        if p + "main.weight" in sd:                        # UpSample, modules.py:88-93
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(h, sd, p + "main")
        else:
            h = torch.cat([h, hs.pop()], dim=1)
            h = aux_res_block(sd, p, h, temb, aemb)
        if trace is not None:
            trace[f"up{i}"] = h
    assert len(hs) == 0
    return _conv(F.silu(_gn(h, sd, pfx + "tail.0")), sd, pfx + "tail.2")  # models.py:280-284, 323


def aux_unet_forward(sd: SD, x: Tensor, t: Tensor, a: Tensor, pfx: str = "backbone.",
                     trace: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """AuxiliaryUNet.forward(x, t, a) -> eps -- models.py:296-326."""
    aemb = _lin(a, sd, pfx + "fc_a")                       # models.py:298 (no activation before)
    temb = time_embedding(sd, pfx + "time_embedding.", t)  # models.py:301
    return _unet_body(sd, pfx, x, temb, aemb, trace)


def bottleneck_unet_forward(sd: SD, x: Tensor, t: Tensor, a: Tensor, pfx: str = "backbone.") -> Tensor:
    """BottleneckAuxUNet.forward(x, t, a) -> eps -- models.py:391-421: fc_a = SiLU -> Linear (336-339); only the
    two middle AuxResBlocks see aemb, down / up blocks are plain ResBlocks."""
    aemb = _lin(F.silu(a), sd, pfx + "fc_a.1")
    temb = time_embedding(sd, pfx + "time_embedding.", t)
    return _unet_body(sd, pfx, x, temb, aemb)


def unet_forward(sd: SD, x: Tensor, t: Tensor, pfx: str = "backbone.") -> Tensor:
    """UNet.forward(x, t) -> eps -- models.py:62-88 (unconditional; plain ResBlocks everywhere)."""
    temb = time_embedding(sd, pfx + "time_embedding.", t)
    return _unet_body(sd, pfx, x, temb, None)


def encoder_forward(sd: SD, x: Tensor, pfx: str = "encoder.",
                    noise: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Encoder.forward(x) -> (a, a_q, mu, log_var) -- models.py:488-518.

    ``noise`` replaces the ``torch.randn_like(mu)`` draw of models.py:515 (drawn from the
    global generator when None, like the reference).
    """
    h = _conv(x, sd, pfx + "head")
    hs = [h]
    for i in range(_count(sd, pfx + "downblocks.")):
        p = f"{pfx}downblocks.{i}."
        if p + "main.weight" in sd:
            h = _conv(h, sd, p + "main", stride=2)
        else:
            h = res_block_encoder(sd, p, h)
        hs.append(h)
    for i in range(_count(sd, pfx + "middleblocks.")):
        h = res_block_encoder(sd, f"{pfx}middleblocks.{i}.", h)
    for i in range(_count(sd, pfx + "upblocks.")):
        p = f"{pfx}upblocks.{i}."
        if p + "main.weight" in sd:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(h, sd, p + "main")
        else:
            h = torch.cat([h, hs.pop()], dim=1)
            h = res_block_encoder(sd, p, h)
    assert len(hs) == 0
    h = _conv(F.silu(_gn(h, sd, pfx + "tail.0")), sd, pfx + "tail.2")
    h = torch.flatten(h, start_dim=1)                      # models.py:510
    a = _lin(h, sd, pfx + "fc_a")
    mu = _lin(a, sd, pfx + "fc_mu")
    log_var = _lin(a, sd, pfx + "fc_var")
    if noise is None:
        noise = torch.randn_like(mu)
    a_q = mu + noise * torch.exp(0.5 * log_var)            # models.py:515
    return a, a_q, mu, log_var


def timestep_embedding(t: Tensor, dim: int, max_period: int = 10000) -> Tensor:
    """[cos | sin] sinusoid used by LatentUNet -- modules.py:41-60."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def latent_unet_forward(sd: SD, x: Tensor, t: Tensor, pfx: str = "backbone.",
                        num_time_emb_channels: int = 64) -> Tensor:
    """LatentUNet.forward(x, t): 10-layer skip-MLP eps-net over z, eval mode -- models.py:223-234, 147-163."""
    temb = timestep_embedding(t, num_time_emb_channels)
    temb = _lin(temb, sd, pfx + "time_embed.0")
    temb = _lin(F.silu(temb), sd, pfx + "time_embed.2")
    n_layers = _count(sd, pfx + "layers.")
    h = x
    for i in range(n_layers):
        p = f"{pfx}layers.{i}."
        if i >= 1:                                         # skip_layers = 1..n-1, models.py:186, 230-232
            h = torch.cat([h, x], dim=1)
        h = _lin(h, sd, p + "linear")
        if p + "linear_emb.weight" in sd:                  # use_cond, condition_bias = 1 (models.py:219)
            cond = _lin(F.silu(temb), sd, p + "linear_emb")
            h = h * (1 + cond)
        if p + "norm.weight" in sd:
            h = F.layer_norm(h, (h.shape[-1],), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
        if i != n_layers - 1:                              # last layer: activation None (models.py:195-200)
            h = F.silu(h)
    return h


# --------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------
def compute_kernel(x: Tensor, y: Tensor) -> Tensor:
    """exp(-mean_d((x-y)^2)/D) over all pairs -- utils.py:75-83."""
    dim = x.shape[1]
    d2 = (x[:, None, :] - y[None, :, :]).pow(2).mean(dim=2)
    return torch.exp(-d2 / dim * 1.0)


def compute_mmd(x: Tensor, y: Tensor) -> Tensor:
    """Biased V-statistic MMD, diagonals included -- utils.py:86-90."""
    return compute_kernel(x, x).mean() + compute_kernel(y, y).mean() - 2 * compute_kernel(x, y).mean()


@dataclass
class Schedule:
    """Noise schedule exactly as InfoDiff.__init__ / DiffusionProcess.__init__ build it
    (models.py:615-618; sampling.py:12-15)."""
    betas: Tensor
    alphas: Tensor
    alpha_bars: Tensor
    alpha_prev_bars: Tensor

    @staticmethod
    def make(beta1: float, betaT: float, T: int) -> "Schedule":
        betas = torch.linspace(start=beta1, end=betaT, steps=T)
        alphas = 1 - betas
        alpha_bars = torch.cumprod(1 - torch.linspace(start=beta1, end=betaT, steps=T), dim=0)
        alpha_prev_bars = torch.cat([torch.Tensor([1]), alpha_bars[:-1]])
        return Schedule(betas, alphas, alpha_bars, alpha_prev_bars)


def q_sample(sch: Schedule, x: Tensor, idx: Tensor, eps: Tensor) -> Tensor:
    """x_t = sqrt(abar_t) x + sqrt(1-abar_t) eps -- models.py:702-704."""
    ab = sch.alpha_bars[idx][:, None, None, None]
    return torch.sqrt(ab) * x + torch.sqrt(1 - ab) * eps


def infodiff_loss(sd: SD, sch: Schedule, x: Tensor, idx: Tensor, eps: Tensor, enc_noise: Tensor,
                  prior: Tensor, mmd_weight: float, kld_weight: float, T: int, use_C: bool = False, C_max: float = 25.0,
                  epochs: int = 1, curr_epoch: int = 0) -> Dict[str, Tensor]:
    """InfoDiff.loss_fn (prior='regular') with all random draws injected -- models.py:632-723."""
    x_t = q_sample(sch, x, idx, eps)
    a, a_q, mu, log_var = encoder_forward(sd, x, noise=enc_noise)          # models.py:710, on CLEAN x
    use_q = (kld_weight != 0)                                              # models.py:714-721
    out = aux_unet_forward(sd, x_t, idx, a_q if use_q else a)
    loss_eps = (out - eps).square().mean()                                 # models.py:640
    x_0 = torch.sqrt(1 / sch.alphas[0]) * (x - sch.betas[0] / torch.sqrt(1 - sch.alpha_bars[0]) * out)  # 644
    loss_rec = (x_0 - x).square().mean() / T                               # models.py:645-646
    loss = loss_eps + loss_rec
    terms = {"eps": loss_eps, "rec": loss_rec, "a": a, "out": out}
    if mmd_weight != 0:
        mmd = compute_mmd(prior, mu if kld_weight != 0 else a)             # models.py:659 / 682
        loss = loss + mmd_weight * mmd
        terms["mmd"] = mmd
    if kld_weight != 0:
        kld = torch.sum(-0.5 * torch.sum(1 + log_var - mu ** 2 - log_var.exp(), dim=1), dim=0)  # 663 / 687
        if use_C:                                                          # models.py:664-668 / 688-692
            c_max = torch.tensor([C_max], dtype=torch.float32)
            cc = torch.clamp(c_max / epochs * curr_epoch, torch.zeros(1), c_max)
            loss = loss + kld_weight * (kld - cc.squeeze(dim=0)).abs()
        else:
            loss = loss + kld_weight * kld
        terms["kld"] = kld
    terms["loss"] = loss
    return terms


# --------------------------------------------------------------------------------------
# samplers.  eps_fn(x, idx:int) -> eps.  Noise is injected through ``noise_fn(idx, like)``
# (defaults to torch.randn_like, drawn at the same point of the step as the reference).
# --------------------------------------------------------------------------------------
NoiseFn = Callable[[int, Tensor], Tensor]


def _randn(idx: int, like: Tensor) -> Tensor:
    return torch.randn_like(like)


def ddpm_steps(sch: Schedule, eps_fn, x: Tensor, noise_fn: NoiseFn = _randn) -> Iterator[Tuple[int, Tensor, Tensor]]:
    """sampling.py:23-39 -- noise drawn BEFORE the model call; zeros at idx == 0."""
    for idx in reversed(range(len(sch.alpha_bars))):
        noise = torch.zeros_like(x) if idx == 0 else noise_fn(idx, x)
        sqrt_tilde_beta = torch.sqrt((1 - sch.alpha_prev_bars[idx]) / (1 - sch.alpha_bars[idx]) * sch.betas[idx])
        eps = eps_fn(x, idx)
        mu = torch.sqrt(1 / sch.alphas[idx]) * (x - sch.betas[idx] / torch.sqrt(1 - sch.alpha_bars[idx]) * eps)
        x = mu + sqrt_tilde_beta * noise
        yield idx, eps, x


def ddim_steps(sch: Schedule, eps_fn, x: Tensor, noise_fn: NoiseFn = _randn) -> Iterator[Tuple[int, Tensor, Tensor]]:
    """sampling.py:41-60 -- eta = 0.01, 'current' abar is alpha_prev_bars[idx]; noise drawn AFTER the model call."""
    eta = 0.01
    apb, ab, betas = sch.alpha_prev_bars, sch.alpha_bars, sch.betas
    for idx in reversed(range(len(ab))):
        eps = eps_fn(x, idx)
        x_0 = (x - torch.sqrt(1 - apb[idx]) * eps) / torch.sqrt(apb[idx])
        if idx == 0:
            x = x_0
        else:
            noise = noise_fn(idx, x)
            sigma = eta * torch.sqrt((1 - apb[idx - 1]) / (1 - ab[idx - 1])) * torch.sqrt(betas[idx - 1])
            x = torch.sqrt(apb[idx - 1]) * x_0 + torch.sqrt(1 - apb[idx - 1] - sigma ** 2) * eps
            x = x + sigma * noise
        yield idx, eps, x


def ddim_reverse_steps(sch: Schedule, eps_fn, x: Tensor) -> Iterator[Tuple[int, Optional[Tensor], Tensor]]:
    """sampling.py:62-73 -- x0 -> xT, T-2 model calls, idx == 0 yields x unchanged."""
    apb = sch.alpha_prev_bars
    for idx in range(len(sch.alpha_bars) - 1):
        if idx == 0:
            yield idx, None, x
        else:
            eps = eps_fn(x, idx)
            x_0 = (x - torch.sqrt(1 - apb[idx]) * eps) / torch.sqrt(apb[idx])
            x = torch.sqrt(apb[idx + 1]) * x_0 + torch.sqrt(1 - apb[idx + 1]) * eps
            yield idx, eps, x


def infodiff_eps_fn(sd: SD, a: Optional[Tensor], enc_noise_fn: Optional[Callable[[Tensor], Tensor]] = None):
    """InfoDiff.forward(x, idx:int, a) as the samplers call it -- models.py:705-723.

    With ``a is None`` the Encoder is re-run on the (noisy) x every call, which is what
    DiffusionProcess.reverse_sampling triggers because it drops ``a`` (sampling.py:84).
    """
    def fn(x: Tensor, idx: int) -> Tensor:
        t = torch.full((x.shape[0],), idx, dtype=torch.long)
        if a is None:
            noise = enc_noise_fn(x) if enc_noise_fn is not None else None
            aa, _, _, _ = encoder_forward(sd, x, noise=noise)
        else:
            aa = a
        return aux_unet_forward(sd, x, t, aa)
    return fn


@torch.no_grad()
def sample(sd: SD, sch: Schedule, xT: Tensor, a: Tensor, deterministic: bool,
           noise_fn: NoiseFn = _randn, record: Optional[List] = None) -> Tensor:
    """DiffusionProcess.sampling with xT and a given -- sampling.py:89-101."""
    steps = ddim_steps if deterministic else ddpm_steps
    x = xT
    for idx, eps, x in steps(sch, infodiff_eps_fn(sd, a), xT, noise_fn):
        if record is not None:
            record.append((idx, eps, x))
    return x


@torch.no_grad()
def reverse_sample(sd: SD, sch: Schedule, x0: Tensor, a: Optional[Tensor] = None,
                   enc_noise_fn=None, record: Optional[List] = None) -> Tensor:
    """DiffusionProcess.reverse_sampling -- sampling.py:81-87.  ``a=None`` is the reference's
    (bug-compatible) behaviour; passing ``a`` gives the a-honouring variant."""
    x = x0
    for idx, eps, x in ddim_reverse_steps(sch, infodiff_eps_fn(sd, a, enc_noise_fn), x0):
        if record is not None:
            record.append((idx, eps, x))
    return x


# --------------------------------------------------------------------------------------
# other samplers
# --------------------------------------------------------------------------------------
@torch.no_grad()
def two_phase_sample(eps_fn_1, eps_fn_2, sch: Schedule, xT: Tensor, deterministic: bool, split_step: int,
                     noise_fn: NoiseFn = _randn, fixed: bool = False, record: Optional[List] = None) -> Tensor:
    """TwoPhaseDiffusionProcess.sampling -- sampling.py:127-204.

    eps_fn_1(x, idx) is the InfoDiff branch (a bound by the caller), eps_fn_2(x, idx) the vanilla branch.
    The reference passes the step counter ``t`` into the generator BY VALUE when the generator is created
    (``t = 0``; sampling.py:198-201), so ``t <= split_step`` is evaluated with t == 0 at every step and
    diffusion_fn_2 is used throughout whenever split_step >= 0.  ``fixed=True`` gives the
    evidently intended behaviour: the counter advances, steps with t <= split_step use fn_2, later ones fn_1.
    """
    steps = ddim_steps if deterministic else ddpm_steps
    counter = {"t": 0}

    def eps_fn(x: Tensor, idx: int) -> Tensor:
        t = counter["t"] if fixed else 0
        counter["t"] += 1
        return eps_fn_2(x, idx) if t <= split_step else eps_fn_1(x, idx)
    x = xT
    for idx, eps, x in steps(sch, eps_fn, xT, noise_fn):
        if record is not None:
            record.append((idx, eps, x))
    return x


def latent_eps_fn(sd: SD, pfx: str = "backbone."):
    """Diff.forward(x, idx:int) with a LatentUNet backbone -- models.py:764-779."""
    def fn(x: Tensor, idx: int) -> Tensor:
        t = torch.full((x.shape[0],), idx, dtype=torch.long)
        return latent_unet_forward(sd, x, t, pfx)
    return fn


def vanilla_eps_fn(sd: SD, pfx: str = "backbone."):
    """Diff.forward(x, idx:int) with a UNet backbone -- models.py:764-779."""
    def fn(x: Tensor, idx: int) -> Tensor:
        t = torch.full((x.shape[0],), idx, dtype=torch.long)
        return unet_forward(sd, x, t, pfx)
    return fn


@torch.no_grad()
def latent_sample(sd: SD, sch: Schedule, zT: Tensor, deterministic: bool, noise_fn: NoiseFn = _randn,
                  record: Optional[List] = None) -> Tensor:
    """LatentDiffusionProcess.sampling -- sampling.py:232-291 (same step formulas over [B, a_dim])."""
    steps = ddim_steps if deterministic else ddpm_steps
    x = zT
    for idx, eps, x in steps(sch, latent_eps_fn(sd), zT, noise_fn):
        if record is not None:
            record.append((idx, eps, x))
    return x
