"""CPU tests of the host-side logic: layout math, K-block tables, weight packing, coefficient tables,
constructor / state_dict parity and the C-ABI surface.  No GPU needed."""
import ctypes
import json
import re
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

from infodiffusion_b200 import layout
from oracle import infodiff_oracle as orc
from oracle.golden_util import SEED, make_args, state_digest
from tests.helpers import emulate_igemm, from_padflat, interior_mask, space_to_depth, to_padflat

ROOT = Path(__file__).resolve().parent.parent


def test_padflat_roundtrip():
    x = torch.randn(3, 5, 8, 8)
    pf = to_padflat(x)
    assert pf.shape == (layout.padflat_rows(3, 8, 8), 5)
    assert torch.equal(from_padflat(pf, 3, 8, 8), x)


@pytest.mark.parametrize("cin,cout,H", [(64, 64, 8), (128, 64, 16), (192, 128, 8)])
def test_conv3x3_kblocks_match_conv2d(cin, cout, H):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, cin, H, H, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) * 0.05
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    out = emulate_igemm([to_padflat(x)], layout.taps3x3(cin, H, H), layout.pack_conv3x3(w), b, 2, H, H)
    assert torch.allclose(from_padflat(out, 2, H, H), ref, atol=1e-9)


def test_stride2_kblocks_match_conv2d():
    g = torch.Generator().manual_seed(1)
    B, C, H = 3, 64, 16
    x = torch.randn(B, C, H, H, generator=g)
    w = torch.randn(C, C, 3, 3, generator=g) * 0.05
    b = torch.randn(C, generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=2, padding=1)
    Ho = H // 2
    ph = space_to_depth(to_padflat(x), B, H, H)
    kb = layout.taps_stride2(C, Ho, Ho, layout.padflat_rows(B, Ho, Ho))
    out = emulate_igemm([ph], kb, layout.pack_conv3x3(w), b, B, Ho, Ho)
    assert torch.allclose(from_padflat(out, B, Ho, Ho), ref, atol=1e-9)


def test_fused_shortcut_kblocks():
    """conv3(act) + shortcut1x1(cat(h, skip)) as one K loop over three sources."""
    g = torch.Generator().manual_seed(2)
    B, H, c_h, c_s, cout = 2, 8, 128, 64, 128
    act = torch.randn(B, cout, H, H, generator=g)
    h = torch.randn(B, c_h, H, H, generator=g)
    s = torch.randn(B, c_s, H, H, generator=g)
    w3 = torch.randn(cout, cout, 3, 3, generator=g) * 0.05
    wsc = torch.randn(cout, c_h + c_s, 1, 1, generator=g) * 0.05
    b3, bsc = torch.randn(cout, generator=g), torch.randn(cout, generator=g)
    ref = F.conv2d(act.double(), w3.double(), b3.double(), padding=1) + \
        F.conv2d(torch.cat([h, s], 1).double(), wsc.double(), bsc.double())
    kb = layout.taps3x3(cout, H, H, 0) + layout.taps1x1(c_h, 1) + layout.taps1x1(c_s, 2)
    wp = torch.cat([layout.pack_conv3x3(w3), layout.pack_conv1x1(wsc)], 1)
    out = emulate_igemm([to_padflat(act), to_padflat(h), to_padflat(s)], kb, wp, b3.double() + bsc.double(), B, H, H)
    assert torch.allclose(from_padflat(out, B, H, H), ref, atol=1e-9)


def test_head_im2col_packing():
    """head conv as a K=64 GEMM over 3x3 patches with k = tap*C + c."""
    g = torch.Generator().manual_seed(3)
    B, C, H = 2, 3, 8
    x = torch.randn(B, C, H, H, generator=g)
    w = torch.randn(64, C, 3, 3, generator=g)
    b = torch.randn(64, generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    xp = F.pad(x, (1, 1, 1, 1))
    patches = torch.zeros(B, 64, H, H)
    for tap in range(9):
        ky, kx = divmod(tap, 3)
        patches[:, tap * C:(tap + 1) * C] = xp[:, :, ky:ky + H, kx:kx + H]
    out = emulate_igemm([to_padflat(patches)], [(0, 0, 0)], layout.pad_cols(layout.pack_conv3x3(w), 64), b, B, H, H)
    assert torch.allclose(from_padflat(out, B, H, H), ref, atol=1e-9)


@pytest.mark.parametrize("kind", ["ddpm", "ddim", "reverse"])
def test_step_coefficients_match_oracle_formulas(kind):
    """The collapsed (cx, ce, cn) table reproduces the reference's step formulas on random tensors."""
    from infodiffusion_b200.sampling import make_schedule, step_coefficients
    T = 20
    sch = orc.Schedule.make(1e-5, 1e-2, T)
    coef = step_coefficients(kind, *make_schedule(1e-5, 1e-2, T))
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 3, 8, 8, generator=g)
    noises = {i: torch.randn(2, 3, 8, 8, generator=g) for i in range(T)}
    eps_of = {i: torch.randn(2, 3, 8, 8, generator=g) for i in range(T)}
    steps = {"ddpm": orc.ddpm_steps, "ddim": orc.ddim_steps}
    if kind == "reverse":
        it = orc.ddim_reverse_steps(sch, lambda xx, i: eps_of[i], x)
    else:
        it = steps[kind](sch, lambda xx, i: eps_of[i], x, lambda i, like: noises[i])
    mine = x.clone()
    for idx, eps, x_ref in it:
        if eps is None:
            continue
        cx, ce, cn = coef[idx].tolist()
        nz = noises[idx] if (kind != "reverse" and idx > 0) else torch.zeros_like(x)
        mine = cx * mine + ce * eps + cn * nz
        assert torch.allclose(mine, x_ref, rtol=2e-5, atol=2e-6), (kind, idx)


@pytest.mark.parametrize("a_dim", [32, 256])
def test_constructor_matches_reference_state_dict(a_dim, golden_dir):
    """Same seed -> same 969-key state_dict as the reference's InfoDiff (digest pinned by make_golden.py)."""
    from infodiffusion_b200.models import InfoDiff
    meta = json.loads((golden_dir / "meta.json").read_text())[f"state_a{a_dim}_T1000"]
    torch.manual_seed(SEED)
    m = InfoDiff(make_args(a_dim=a_dim, diffusion_steps=1000), "cpu", (3, 64, 64))
    sd = m.state_dict()
    assert len(sd) == meta["nkeys"]
    assert sum(p.numel() for p in m.parameters()) == meta["nparams"]
    assert state_digest(sd) == meta["digest"]


def test_blocks_have_no_cpu_forward():
    from infodiffusion_b200.models import InfoDiff
    m = InfoDiff(make_args(a_dim=32, diffusion_steps=10), "cpu", (3, 64, 64)).eval()
    with pytest.raises(RuntimeError):
        m.backbone(torch.zeros(1, 3, 64, 64), torch.zeros(1, dtype=torch.long), torch.zeros(1, 32))
    with pytest.raises(RuntimeError):
        m.backbone.downblocks[0](torch.zeros(1, 64, 64, 64), None, None)


def test_shard_range_partitions():
    for total, world in [(256, 8), (10, 4), (3, 8), (512, 2)]:
        spans = [layout.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_c_abi_exports_every_declared_symbol():
    """libidf_b200.so loads and exports each function include/idf_b200.h declares (no compute calls)."""
    from infodiffusion_b200 import _lib, build
    build.build()
    header = (ROOT / "include" / "idf_b200.h").read_text()
    declared = set(re.findall(r"\b(idf_[a-z0-9_]+)\s*\(", header))
    declared -= {"idf_status"}
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in idf_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().idf_version() >= 100


def test_product_never_imports_the_oracle_and_has_no_cpu_path():
    """The oracle is test infrastructure: nothing under infodiffusion_b200/ may import or execute it, and the
    network forwards must refuse CPU tensors instead of falling back."""
    import re
    pkg = ROOT / "infodiffusion_b200"
    for f in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        if f.suffix in (".py", ".cu", ".cuh"):
            txt = f.read_text()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f"{f.name} imports the oracle"
            assert "infodiff_oracle" not in txt, f"{f.name} references the oracle"
    from infodiffusion_b200.models import AuxiliaryUNet, Encoder, LatentUNet, UNet
    net = AuxiliaryUNet(T=10, ch=64, ch_mult=[1, 2, 2, 2], a_dim=8, shape=(3, 64, 64)).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 3, 64, 64), torch.zeros(1, dtype=torch.long), torch.zeros(1, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        Encoder(ch=64, ch_mult=[1, 2, 2, 2], a_dim=8, shape=(3, 64, 64)).eval()(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="CUDA"):
        UNet(T=10, ch=64, ch_mult=[1, 2, 2, 2], shape=(3, 64, 64)).eval()(torch.zeros(1, 3, 64, 64), torch.zeros(1, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA"):
        LatentUNet(T=10, shape=(1, 8, 8)).eval()(torch.zeros(2, 8), torch.zeros(2, dtype=torch.long))


def test_run_py_schedule_and_naming():
    """run.py mirrors the reference's LR schedule (GradualWarmupScheduler over CosineAnnealingLR, run.py:182-185: the
    sequence below was produced by the reference's own classes) and its experiment / folder naming (utils.py:49-61)."""
    import warnings
    import run as idf_run
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=1e-4)
    cos = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer=opt, T_max=6, eta_min=0, last_epoch=-1)
    warm = idf_run.GradualWarmupScheduler(optimizer=opt, multiplier=2., warm_epoch=1, after_scheduler=cos)
    got = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(6):
            got.append(opt.param_groups[0]["lr"])
            opt.step()
            warm.step()
    want = [0.0001, 0.0002, 0.00021435935394489817, 0.0002, 0.00016076951545867363, 0.00010717967697244908]
    assert all(abs(a - b) <= 1e-12 for a, b in zip(got, want)), got
    args = idf_run.parse_args(["--model", "diff", "--mode", "train", "--prior", "regular", "--dataset", "celeba", "--a_dim", "32",
                               "--kld_weight", "0.5", "--use_C", "--is_bottleneck"])
    assert idf_run.generate_exp_string(args) == "celeba_32d_0.5kld_25C_0.1mmd_bottleneck"
    assert idf_run.get_dataset_config(args) == (3, 64, 64) and args.unets_channels == 64
    args.dataset = "mnist"
    with pytest.raises(NotImplementedError):
        idf_run.get_dataset_config(args)


def test_fused_adagn_operand_kblocks_match_conv_over_normalised_concat():
    """engine.Plan._operand for a VAct (AdaGN of two channel-concatenated sources, applied inside the conv): the
    k-block order it produces must match the (tap, concatenated channel) packing of the conv weight, and kb_xf must
    point each k-block at its channels of the coefficient table -- checked by emulating the K loop on the CPU with
    the transform applied per k-block."""
    from infodiffusion_b200.engine import Act, Plan, VAct
    g = torch.Generator().manual_seed(5)
    B, H, c0, c1, cout = 2, 8, 128, 64, 64
    x0 = torch.randn(B, c0, H, H, generator=g)
    x1 = torch.randn(B, c1, H, H, generator=g)
    A = torch.randn(B, c0 + c1, generator=g)
    Bc = torch.randn(B, c0 + c1, generator=g)
    w = torch.randn(cout, c0 + c1, 3, 3, generator=g) * 0.05
    bias = torch.randn(cout, generator=g)
    normed = F.silu(torch.cat([x0, x1], 1).double() * A.double()[:, :, None, None] + Bc.double()[:, :, None, None])
    ref = F.conv2d(normed, w.double(), bias.double(), padding=1)
    s0, s1 = Act(to_padflat(x0), H, c0, 1, 0), Act(to_padflat(x1), H, c1, 1, 0)
    coef = torch.stack([A, Bc], dim=-1)                       # [B, C, 2] like idf_adagn_coef writes it
    srcs, kb, xf = Plan._operand(VAct([s0, s1], coef, True, H, c0 + c1), layout.tap_offsets3x3(H, H))
    assert [s.t.shape[1] for s in srcs] == [c0, c1] and len(kb) == 9 * (c0 + c1) // 64 == len(xf[2])
    rows = B * (H + 1) * (H + 1)
    img = torch.arange(rows) // ((H + 1) * (H + 1))
    inside = interior_mask(B, H, H)
    srcs_t = []
    for si, c0s, off in kb:                                    # apply the transform per k-block, as the kernel does
        cb = xf[2][len(srcs_t)]
        raw = srcs[si].t[:, c0s:c0s + 64].double()
        a_, b_ = coef[img, cb:cb + 64, 0].double(), coef[img, cb:cb + 64, 1].double()
        t = F.silu(raw * a_ + b_)
        t[~inside] = 0                                         # pad rows stay zero
        srcs_t.append(t)
    kb_t = [(k, 0, off) for k, (_, _, off) in enumerate(kb)]
    out = emulate_igemm(srcs_t, kb_t, layout.pack_conv3x3(w), bias, B, H, H)
    assert torch.allclose(from_padflat(out, B, H, H), ref, atol=1e-9)
