"""Which rounding dominates the bf16 path's eps error?    CPU-only emulation on the oracle.

The CUDA path stores activations and weights in bf16 and accumulates in fp32.  This script injects exactly those
roundings, one class at a time, into the fp32 oracle (test infrastructure; nothing here is product code) and reports
rel-L2(eps) against the un-rounded fp32 oracle on the parity inputs of tests/test_gpu_network.py::test_backbone_eps:

  W        conv / attention-projection weights rounded to bf16
  A        the operand entering every conv rounded to bf16 (= the AdaGN+SiLU output, or a raw activation)
  O_inner  outputs of the convs that feed ONLY a GroupNorm (block1 / block2 convs) rounded to bf16
  O_stream outputs on the residual / skip stream (block-closing convs, shortcuts, attention proj, head, down/up-sample)

    python tools/error_budget.py            # prints the table kept in DESIGN.md section 5
"""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import infodiff_oracle as orc  # noqa: E402
from oracle.golden_util import SEED, make_args, perturb_state_dict, rand_inputs, rel_l2  # noqa: E402
from infodiffusion_b200.models import InfoDiff  # noqa: E402  (constructor only: bit-identical init, no kernels)


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def run(sd, x, t, a, W=False, A=False, O_inner=False, O_stream=False):
    real = orc._conv

    def conv(xx, sd_, key, stride=1, padding=1):
        w, b = sd_[key + ".weight"], sd_[key + ".bias"]
        inner = key.endswith("block1.2") or (key.endswith("block2.3") and (key.rsplit(".", 2)[0] + ".block3.3.weight") in sd_)
        if A:
            xx = bf(xx)
        if W:
            w = bf(w)
        y = F.conv2d(xx.double(), w.double(), b.double(), stride=stride, padding=padding).float()
        if (O_inner and inner) or (O_stream and not inner and not key.endswith("tail.2")):
            y = bf(y)
        return y
    orc._conv = conv
    try:
        with torch.no_grad():
            return orc.aux_unet_forward(sd, x, t, a)
    finally:
        orc._conv = real


def main():
    args = make_args(a_dim=32, diffusion_steps=1000)
    torch.manual_seed(SEED)
    sd = perturb_state_dict(InfoDiff(args, "cpu", (3, 64, 64)).state_dict())
    x, t, a = rand_inputs(2, 32, 1000)
    with torch.no_grad():
        ref = orc.aux_unet_forward(sd, x, t, a)
    rows = [("weights only (W)", dict(W=True)),
            ("conv operands only (A)", dict(A=True)),
            ("GroupNorm-only conv outputs (O_inner)", dict(O_inner=True)),
            ("residual / skip stream (O_stream)", dict(O_stream=True)),
            ("all activations, fp32 weights (A + O_inner + O_stream)", dict(A=True, O_inner=True, O_stream=True)),
            ("everything = the CUDA path's storage (W + A + O_inner + O_stream)", dict(W=True, A=True, O_inner=True, O_stream=True)),
            ("fp32 residual / skip stream (W + A + O_inner)", dict(W=True, A=True, O_inner=True)),
            ("fp32 stream and fp32 GroupNorm inputs (W + A)", dict(W=True, A=True))]
    print(f"{'rounding injected into the fp32 oracle':70s} rel-L2(eps)")
    for name, kw in rows:
        print(f"{name:70s} {rel_l2(run(sd, x, t, a, **kw), ref):.3e}")


if __name__ == "__main__":
    main()
