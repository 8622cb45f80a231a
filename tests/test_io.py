"""Output stage: eval_fid PNG folder and save_latent .npz, against the reference's own
writers (torchvision.utils.save_image, np.savez as called in run.py:288-295, 439-443)."""
import io

import numpy as np
import pytest
import torch

from infodiffusion_b200 import io as idf_io


def test_png_encoder_round_trip_cpu():
    from PIL import Image
    g = np.random.default_rng(0)
    for shape in [(64, 64, 3), (5, 7, 3), (28, 28, 1), (1, 1, 3)]:
        img = g.integers(0, 256, size=shape, dtype=np.uint8)
        dec = np.asarray(Image.open(io.BytesIO(idf_io.encode_png(img))))
        assert np.array_equal(dec.reshape(shape), img)


def test_save_latents_npz_keys(tmp_path):
    a = [torch.randn(4, 8), torch.randn(3, 8)]
    attr = [np.arange(4), np.arange(3)]
    idf_io.save_latents_npz(str(tmp_path / "diff_exp_latent"), a, attr)
    z = np.load(tmp_path / "diff_exp_latent.npz")
    assert set(z.files) == {"all_a", "all_attr"} and z["all_a"].shape == (7, 8) and z["all_attr"].shape == (7,)
    assert np.array_equal(z["all_a"][:4], a[0].numpy())


def test_images_to_uint8_refuses_cpu():
    with pytest.raises(RuntimeError, match="CUDA"):
        idf_io.images_to_uint8(torch.zeros(1, 3, 4, 4))


@pytest.mark.gpu
def test_eval_images_equal_the_reference_writer(tmp_path):
    """Bytes of idf_to_uint8_hwc == the reference's clip / (x+1)/2 / save_image quantisation, and the PNG files decode
    to the same pixels as torchvision.utils.save_image writes (run.py:288-295)."""
    from PIL import Image
    from torchvision.utils import save_image
    g = torch.Generator().manual_seed(0)
    x = torch.randn(5, 3, 64, 64, generator=g) * 0.8            # some values beyond [-1, 1]
    x[0, 0, 0, :4] = torch.tensor([-1.0, 1.0, 0.0, 1e-8])
    u8 = idf_io.images_to_uint8(x.cuda()).cpu()
    ref = torch.clip(x, min=-1, max=1)
    ref = (ref + 1) / 2
    ref = ref.mul(255).add_(0.5).clamp_(0, 255).permute(0, 2, 3, 1).to(torch.uint8)      # torchvision save_image
    assert torch.equal(u8, ref)
    n = idf_io.save_eval_images(x.cuda(), str(tmp_path / "ours"), first_index=3, limit=7)
    assert n == 4 and not (tmp_path / "ours" / "sample-000007.png").exists()
    for k in range(4):
        img = torch.clip(x[k], min=-1, max=1)
        img = (img + 1) / 2
        save_image(img, str(tmp_path / f"ref{k}.png"))
        a = np.asarray(Image.open(tmp_path / "ours" / f"sample-{3 + k:06d}.png"))
        b = np.asarray(Image.open(tmp_path / f"ref{k}.png"))
        assert np.array_equal(a, b)
    idf_io.save_samples_npz(str(tmp_path / "bulk"), [x[:2].cuda(), x[2:].cuda()])
    assert np.array_equal(np.load(tmp_path / "bulk.npz")["images"], ref.numpy())
