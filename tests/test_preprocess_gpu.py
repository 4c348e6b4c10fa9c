"""GPU image preprocessing vs the reference recipe on Pillow/torchvision: bit-exact, through the C-ABI
(`sv_preproc_run_host`) with host uint8 images in and device pixels out."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess as P
from starvector_b200.preprocess import ImageTrainProcessor, SiglipImageProcessor, SimpleStarVectorProcessor

pytestmark = pytest.mark.gpu
CASES = [(224, 224, 3), (224, 224, 4), (64, 48, 4), (48, 64, 3), (1, 1, 3), (2, 5, 4), (300, 200, 4), (225, 223, 3),
         (640, 480, 3), (1000, 37, 3), (7, 900, 4), (512, 512, 4), (1536, 2048, 4)]


@pytest.fixture(scope="module")
def images():
    return [P.synthetic_image(h, w, c, seed=10 + i) for i, (h, w, c) in enumerate(CASES)]


def test_ragged_batch_is_bit_identical_to_the_reference_recipe(images):
    proc = ImageTrainProcessor(size=224)
    out = proc.batch(images)                                   # one call: 13 uploads, 2 launches
    assert out.shape == (len(images), 3, 224, 224) and out.dtype == torch.float32 and out.is_cuda
    assert proc.launch_count() == 2
    for i, a in enumerate(images):
        assert torch.equal(out[i].cpu(), P.reference_transform(a, 224, P.ALPHA_WHITE)), CASES[i]
    # per-image calls, sizes changing between calls (arena growth, coefficient cache), PIL input, strided view
    from PIL import Image

    for i in (6, 2, 12, 0):
        a = images[i]
        pil = Image.fromarray(a, "RGBA" if a.shape[2] == 4 else "RGB")
        assert torch.equal(proc(pil).cpu(), P.reference_transform(a, 224, P.ALPHA_WHITE))
    view = images[8][:, 100:420]
    assert torch.equal(proc(view).cpu(), P.reference_transform(np.ascontiguousarray(view), 224, P.ALPHA_WHITE))
    proc.close()


def test_bf16_output_is_the_rounded_fp32_tensor(images):
    proc = ImageTrainProcessor(size=224, dtype=torch.bfloat16)
    out = proc.batch(images[:8]).cpu()
    for i in range(8):
        assert torch.equal(out[i], P.reference_transform(images[i], 224, P.ALPHA_WHITE).to(torch.bfloat16))


def test_golden_fixture(golden_dir):
    g = torch.load(os.path.join(golden_dir, "preprocess_v1.pt"), weights_only=False)
    white, drop = ImageTrainProcessor(size=224), SimpleStarVectorProcessor(size=224)
    for case in g["cases"]:
        a = case["image"].numpy()
        got = (white(a) if case["alpha"] == P.ALPHA_WHITE else drop(images=a)["pixel_values"]).cpu()
        assert P.tensor_sha256(got) == case["sha256_f32"]
        assert P.tensor_sha256(got.to(torch.bfloat16)) == case["sha256_bf16"]


def test_simple_processor_drops_alpha(images):
    proc = SimpleStarVectorProcessor(size=224)
    out = proc(images=images[:8])["pixel_values"].cpu()
    for i in range(8):
        assert torch.equal(out[i], P.reference_transform(images[i], 224, P.ALPHA_DROP)), CASES[i]
    one = proc(images=images[2])["pixel_values"]
    assert one.shape == (3, 224, 224)


def test_siglip_processor_equals_the_pil_processor(images):
    import warnings

    from PIL import Image
    from transformers.models.siglip import SiglipImageProcessorPil

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = SiglipImageProcessorPil(size={"height": 384, "width": 384})
    proc = SiglipImageProcessor(size=384)
    pick = [images[i] for i in (2, 6, 8, 9, 11)]
    got = proc(images=pick, return_tensors="pt").pixel_values.cpu()
    for a, g in zip(pick, got):
        pil = Image.fromarray(a, "RGBA" if a.shape[2] == 4 else "RGB")
        assert torch.equal(g, ref(images=pil, return_tensors="pt").pixel_values[0]), a.shape


def test_size_independent_properties():
    """Full-size inputs: a constant image maps to a constant; an image already at the target size is only normalised."""
    proc = ImageTrainProcessor(size=224)
    lut = torch.from_numpy(P.normalize_lut())
    big = np.full((4096, 3000, 4), 0, np.uint8)
    big[..., :3] = (10, 200, 77)
    big[..., 3] = 255
    out = proc(big).cpu()                                      # opaque colour, white bars left/right of the centred image
    col = out[:, :, 112]
    for c, v in enumerate((10, 200, 77)):
        assert torch.all(col[c] == lut[c, v])
    assert torch.all(out[:, :, 0] == lut[:, 255][:, None])
    same = P.synthetic_image(224, 224, 3, seed=5)
    got = proc(same).cpu()
    want = torch.stack([lut[c][torch.from_numpy(same[..., c].astype(np.int64))] for c in range(3)])
    assert torch.equal(got, want)


def test_errors_are_python_exceptions():
    proc = ImageTrainProcessor(size=224)
    with pytest.raises(ValueError):
        proc(np.zeros((4, 4, 2), np.uint8))
    with pytest.raises(ValueError):
        proc(np.zeros((4, 4, 3), np.float32))
    with pytest.raises(ValueError):
        proc.batch([])
    with pytest.raises(ValueError):
        ImageTrainProcessor(size=0)
