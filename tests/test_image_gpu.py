"""Device image pipeline (ist_image_* of include/ist_b200.h) against the oracle and against PIL / torchvision themselves:
bit-exact 8-bit results and bit-exact fp32 network images, including ragged sizes, batches and the coarse-to-fine hand-off."""
import ctypes
import os

import numpy as np
import pytest
import torch
from PIL import Image

from ist_b200 import _lib
from ist_b200.data import DeviceImageTransform, ImageTransform
from oracle import image_oracle as IO
from oracle import synth
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")
MEAN = synth.IMAGENET_MEAN


def _resize_dev(img_u8, oh, ow):
    lib = _lib.load()
    b, h, w, _ = img_u8.shape
    src = torch.from_numpy(img_u8).to(dev)
    out = torch.empty(b, oh, ow, 3, dtype=torch.uint8, device=dev)
    tmp = torch.empty(b, h, ow, 3, dtype=torch.uint8, device=dev)
    _lib.check(lib.ist_image_resize_u8(ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                                       ctypes.c_void_p(tmp.data_ptr()), b, h, w, oh, ow, _lib.stream_ptr()))
    return out.cpu().numpy()


@pytest.mark.parametrize("shape", [(37, 53, 64, 91), (64, 64, 128, 128), (64, 64, 32, 32), (50, 70, 23, 91), (100, 60, 100, 33),
                                   (33, 100, 77, 100), (97, 31, 5, 7), (16, 16, 1, 1), (3, 5, 40, 60), (64, 64, 64, 64)])
def test_resize_bit_exact_vs_pil_and_oracle(shape):
    h, w, oh, ow = shape
    rng = np.random.Generator(np.random.PCG64(h * 1000 + w))
    imgs = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
    got = _resize_dev(imgs, oh, ow)
    for i in range(2):
        assert np.array_equal(got[i], np.asarray(Image.fromarray(imgs[i]).resize((ow, oh), Image.BILINEAR)))
        assert np.array_equal(got[i], IO.pil_resize_bilinear_u8(imgs[i], oh, ow))


def test_golden_vectors():
    gold = np.load(os.path.join(GOLDEN, "image_pipeline.npz"))
    img = gold["img_a"]
    for size in (64, 20, 37):
        T = DeviceImageTransform(size, MEAN, dev)
        assert np.array_equal(T.preparation(Image.fromarray(img)).cpu().numpy(), gold[f"a_prep_{size}"])
    T48, T96 = DeviceImageTransform(48, MEAN, dev), DeviceImageTransform(96, MEAN, dev)
    x = torch.from_numpy(gold["x_lo"]).to(dev)
    assert np.array_equal(np.asarray(T48.post_preparation(x)), gold["x_lo_post"])
    assert np.array_equal(T96.handoff(x.unsqueeze(0))[0].cpu().numpy(), gold["x_hi"])


@pytest.mark.parametrize("hw", [(64, 64), (37, 53), (5, 3), (1, 1), (128, 96)])
def test_prep_and_post_bit_exact_with_ragged_sizes_and_batches(hw):
    h, w = hw
    rng = np.random.Generator(np.random.PCG64(h * 77 + w))
    T = DeviceImageTransform(max(h, w), MEAN, dev)
    imgs = rng.integers(0, 256, (3, h, w, 3), dtype=np.uint8)
    x = T.prep_u8(torch.from_numpy(imgs).to(dev)).cpu().numpy()
    for i in range(3):
        assert np.array_equal(x[i], IO.prep_u8(imgs[i], MEAN))
    xs = (x + rng.normal(0, 50, x.shape)).astype(np.float32)
    post = T.post_u8(torch.from_numpy(xs).to(dev)).cpu().numpy()
    cpuT = ImageTransform(max(h, w), MEAN)
    for i in range(3):
        assert np.array_equal(post[i], IO.post_preparation(xs[i], MEAN))
        assert np.array_equal(post[i], np.asarray(cpuT.post_preparation(torch.from_numpy(xs[i].copy()))))
    # exact 8-bit images survive the round trip prep -> post (the pipeline's fixed point)
    rt = T.post_u8(T.prep_u8(torch.from_numpy(imgs).to(dev))).cpu().numpy()
    assert np.abs(rt.astype(int) - imgs.astype(int)).max() <= 1


def test_preparation_matches_the_cpu_transform_on_radar_frames():
    for size, src in [(64, synth.radar_frame(48, 1)), (32, synth.lidar_frame(64, 2)), (96, synth.smooth_frame(96, 3))]:
        pil = Image.fromarray(src)
        a = DeviceImageTransform(size, MEAN, dev).preparation(pil).cpu()
        b = ImageTransform(size, MEAN).preparation(pil)
        assert torch.equal(a, b)


@pytest.mark.parametrize("sizes", [(256, 512), (512, 1024)])
def test_full_size_handoff_matches_pil_pipeline(sizes):
    lo, hi = sizes
    x = torch.from_numpy(synth.preprocess(synth.radar_frame(lo, 4)))
    x = x + 30.0 * torch.from_numpy(np.random.Generator(np.random.PCG64(9)).standard_normal(tuple(x.shape)).astype(np.float32))
    got = DeviceImageTransform(hi, MEAN, dev).handoff(x.to(dev))
    pil = ImageTransform(lo, MEAN).post_preparation(x[0].clone())
    ref = ImageTransform(hi, MEAN).preparation(pil)
    n_bad = int((got[0].cpu() != ref).sum())
    assert n_bad == 0, f"{n_bad} of {ref.numel()} values differ from the PIL pipeline"
    # the fused hand-off equals its three pieces run one by one
    T = DeviceImageTransform(hi, MEAN, dev)
    assert torch.equal(T.prep_u8(T.resize_u8(T.post_u8(x.to(dev)))), got)


def test_bad_arguments_raise():
    T = DeviceImageTransform(64, MEAN, dev)
    with pytest.raises(_lib.IstError):
        T.upload(Image.fromarray(np.zeros((8, 8), np.uint8)))           # not RGB
    with pytest.raises(_lib.IstError):
        DeviceImageTransform(64, MEAN, "cpu")
    lib = _lib.load()
    assert lib.ist_image_resize_u8(None, None, None, 1, 4, 4, 8, 8, None) != 0
