"""GPU image / label edges: bit-exact against tensors produced by the reference's own host path (PIL, torchvision,
seg_model/utils/ext_transforms.py, acdc.encode_target, sample_integrated.postprocess; tests/golden/make_golden.py:g_io)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs CUDA")
    return torch.device("cuda")


def test_label_and_image_preprocessing_bit_exact(golden):
    from weatherconverter_b200 import image_io as io
    dev = _dev()
    d = golden("image_io.pt")
    enc = io.encode_label(d["label_ids"].to(dev))
    assert enc.dtype == torch.int64 and enc.shape == (1, 512, 512)
    assert torch.equal(enc.cpu(), d["encoded_label"].long())
    assert torch.equal(io.normalize_image(d["image_small"].to(dev)).cpu(), d["normalized"])
    for e in d["diffusion_inputs"]:
        got = io.diffusion_input(e["image"].to(dev)).cpu()
        assert torch.equal(got, e["tensor"]), float((got - e["tensor"]).abs().max())


def test_output_images_bit_exact(golden):
    from weatherconverter_b200 import image_io as io
    dev = _dev()
    d = golden("image_io.pt")
    assert torch.equal(io.ddpm_grid_uint8(d["ddpm_xt"].to(dev), 2).cpu(), d["ddpm_grid"])
    assert torch.equal(io.ddpm_grid_uint8(d["ddpm_xt"][:1].to(dev), 2).cpu(), d["ddpm_grid_single"])
    assert torch.equal(io.postprocess_uint8(d["legacy_xt"].to(dev)).cpu(), d["legacy_u8"])


def test_resize_idempotence_and_ragged_sizes():
    """Size-independent properties at sizes the golden does not hold: identity resize, monotone ramps stay monotone, constant
    images stay constant, and the GPU result equals the numpy oracle on odd sizes."""
    import numpy as np
    from oracle import image_io as oio
    from weatherconverter_b200 import image_io as io
    dev = _dev()
    rng = np.random.default_rng(3)
    for (h, w, oh, ow) in ((37, 53, 37, 53), (1080, 1920, 128, 227), (64, 48, 171, 128), (9, 7, 3, 2)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        got = io.resize_bilinear_u8(torch.from_numpy(img).to(dev), oh, ow).cpu().numpy()
        assert np.array_equal(got, oio.bilinear_resize_u8(img, oh, ow)), (h, w, oh, ow)
    const = torch.full((50, 70, 3), 137, dtype=torch.uint8, device=dev)
    assert bool((io.resize_bilinear_u8(const, 23, 31) == 137).all())
