"""Generates tests/golden/image_pipeline.npz from the real third-party pipeline the reference calls (TEST INFRASTRUCTURE).

The reference's ImageTransform (IST/data/image_transform.py:5-31) is torchvision transforms over PIL images; its
`transforms.Scale` is the pre-0.12 name of `transforms.Resize` (absent from torchvision 0.26), every other line is used
verbatim below. Run in the build container:  python oracle/make_image_golden.py
Outputs pin oracle/image_oracle.py (CPU test) and the CUDA kernels (GPU test) to PIL / torchvision results.
"""
import os
import sys

import numpy as np
import torch
from PIL import Image
from torchvision import transforms

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402

MEAN = synth.IMAGENET_MEAN


def reference_transform(image_size):
    prep = transforms.Compose([
        transforms.Resize(image_size),                                  # transforms.Scale(image_size), image_transform.py:9
        transforms.ToTensor(),
        transforms.Lambda(lambda x: x[torch.LongTensor([2, 1, 0])]),
        transforms.Normalize(mean=MEAN, std=[1, 1, 1]),
        transforms.Lambda(lambda x: x.mul_(255)),
    ])
    post1 = transforms.Compose([
        transforms.Lambda(lambda x: x.mul_(1. / 255)),
        transforms.Normalize(mean=[(-1) * x for x in MEAN], std=[1, 1, 1]),
        transforms.Lambda(lambda x: x[torch.LongTensor([2, 1, 0])]),
    ])

    def post(t):
        t = post1(t.clone())
        t[t > 1] = 1
        t[t < 0] = 0
        return transforms.ToPILImage()(t)
    return prep, post


def main():
    rng = np.random.Generator(np.random.PCG64(2024))
    out = {}
    # 1. ragged random image, up- and down-scaling through Resize(int)
    img = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    out["img_a"] = img
    for size in (64, 20, 37):
        prep, post = reference_transform(size)
        out[f"a_prep_{size}"] = prep(Image.fromarray(img)).numpy()
    # 2. plain PIL resizes (both axes, one axis, extreme ratios)
    for k, (oh, ow) in enumerate([(64, 91), (23, 53), (37, 11), (111, 160), (5, 7)]):
        out[f"a_resize_{oh}x{ow}"] = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
    # 3. post_preparation of an out-of-range optimised image, then the coarse-to-fine hand-off 48 -> 96
    prep48, post48 = reference_transform(48)
    prep96, _ = reference_transform(96)
    radar = synth.radar_frame(48, 5)
    x = prep48(Image.fromarray(radar)).numpy()
    x = (x + rng.normal(0, 60, x.shape)).astype(np.float32)
    out["x_lo"] = x
    pil = post48(torch.from_numpy(x))
    out["x_lo_post"] = np.asarray(pil)
    out["x_hi"] = prep96(pil).numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "image_pipeline.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
