"""The image pre/post-processing oracle (oracle/image_oracle.py) against the third-party pipeline the reference calls
(PIL Image.resize + torchvision ToTensor / Normalize / ToPILImage, IST/data/image_transform.py:5-31), live and through
the committed golden vectors (tests/golden/image_pipeline.npz, made by oracle/make_image_golden.py)."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import image_oracle as IO
from oracle import synth
from conftest import GOLDEN

MEAN = synth.IMAGENET_MEAN


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "image_pipeline.npz"))


@pytest.mark.parametrize("shape", [(37, 53, 64, 91), (64, 64, 128, 128), (64, 64, 32, 32), (50, 70, 23, 91), (100, 60, 100, 33),
                                   (33, 100, 77, 100), (97, 31, 5, 7), (16, 16, 1, 1), (3, 5, 40, 60), (128, 128, 256, 256)])
def test_resize_matches_pil_bit_for_bit(shape):
    h, w, oh, ow = shape
    img = np.random.Generator(np.random.PCG64(h * 1000 + w)).integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
    assert np.array_equal(IO.pil_resize_bilinear_u8(img, oh, ow), ref)


def test_resize_radar_like_frame():
    img = synth.radar_frame(128, 1)
    ref = np.asarray(Image.fromarray(img).resize((256, 256), Image.BILINEAR))
    assert np.array_equal(IO.pil_resize_bilinear_u8(img, 256, 256), ref)


def test_resize_target_rule():
    from torchvision import transforms
    for h, w, s in [(37, 53, 64), (53, 37, 64), (64, 64, 20), (100, 33, 50), (512, 512, 1024)]:
        pil = transforms.Resize(s)(Image.fromarray(np.zeros((h, w, 3), np.uint8)))
        assert IO.resize_target(h, w, s) == (pil.size[1], pil.size[0])


@pytest.mark.parametrize("size", [64, 20, 37])
def test_preparation_matches_product_cpu_transform_and_golden(gold, size):
    from ist_b200.data import ImageTransform
    img = gold["img_a"]
    got = IO.preparation(img, size, MEAN)
    assert np.array_equal(got, gold[f"a_prep_{size}"])
    assert np.array_equal(got, ImageTransform(size, MEAN).preparation(Image.fromarray(img)).numpy())


def test_golden_resizes(gold):
    for key in gold.files:
        if key.startswith("a_resize_"):
            oh, ow = (int(v) for v in key[len("a_resize_"):].split("x"))
            assert np.array_equal(IO.pil_resize_bilinear_u8(gold["img_a"], oh, ow), gold[key]), key


def test_post_preparation_and_handoff(gold):
    from ist_b200.data import ImageTransform
    x = gold["x_lo"]
    post = IO.post_preparation(x, MEAN)
    assert np.array_equal(post, gold["x_lo_post"])
    assert np.array_equal(post, np.asarray(ImageTransform(48, MEAN).post_preparation(torch.from_numpy(x.copy()))))
    assert np.array_equal(IO.hr_handoff(x, 96, MEAN), gold["x_hi"])
    # values outside [0,1] after de-normalisation clamp, and 8-bit conversion truncates (ToPILImage: mul(255).byte())
    assert post.min() == 0 and post.max() == 255


def test_coefficients_are_normalised():
    for n_in, n_out in [(512, 1024), (1024, 2048), (512, 300), (7, 5)]:
        bounds, kk = IO.resample_coeffs(n_in, n_out)
        assert (bounds[:, 0] >= 0).all() and (bounds[:, 0] + bounds[:, 1] <= n_in).all()
        assert np.abs(kk.sum(axis=1) - (1 << IO.PRECISION_BITS)).max() <= kk.shape[1]
