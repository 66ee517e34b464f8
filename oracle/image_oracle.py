"""CPU restatement of the image pre/post-processing around the IST optimisation loop (TEST INFRASTRUCTURE — not shipped,
not on the product path; only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it).

What it restates, with the reference call sites:
  * `ImageTransform.preparation` (IST/data/image_transform.py:8-14): `transforms.Scale(size)` (= Resize: smaller edge to
    `size`, PIL bilinear), `ToTensor` (uint8 HWC -> float CHW / 255), RGB->BGR, `Normalize(mean, [1,1,1])`, `mul_(255)`;
  * `ImageTransform.post_preparation` (image_transform.py:16-31): `mul_(1/255)`, `Normalize(-mean, 1)`, BGR->RGB, clamp to
    [0, 1], `ToPILImage` (`pic.mul(255).byte()`: truncation);
  * the coarse-to-fine hand-off (IST/model/engine/hr_transfer_style.py:21-27): the low-resolution result goes through
    post_preparation (8-bit RGB), is re-read by `preparation` at HRDATA.IMG_SIZE and becomes the initial image.

The resize itself lives in a third-party dependency that is not vendored in the reference: Pillow (the reference pins no
version; this image has 12.2.0), `Image.resize(size, BILINEAR)` -> `ImagingResample` (src/libImaging/Resample.c). Its
published algorithm for 8-bit images, restated here in numpy integer arithmetic:
  - per output coordinate xx: centre = (xx + 0.5) * scale, support = max(scale, 1) (bilinear filter support 1.0),
    xmin = int(centre - support + 0.5) clipped at 0, xmax = int(centre + support + 0.5) clipped at the input size, weights
    w = triangle((x + xmin - centre + 0.5) / max(scale, 1)), normalised by their sum (all in double);
  - fixed point: k = int(±0.5 + w * 2^22); out = clip8((2^21 + sum(pixel * k)) >> 22);
  - horizontal pass first (only if the width changes), then the vertical pass on its 8-bit result (only if the height changes).
Parity is pinned in tests/test_image_oracle.py against PIL / torchvision themselves, run in the test.
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2


def resize_target(h, w, size):
    """Output (h, w) of torchvision `Resize(int)`: the smaller edge becomes `size`, the other keeps the aspect ratio
    (truncated); an image already at that size is returned unchanged (torchvision/transforms/functional.py `resize`)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def resample_coeffs(in_size, out_size):
    """(bounds [out,2] int32 (xmin, count), coeffs [out,ksize] int32) of Pillow's precompute_coeffs + normalize_coeffs_8bpc
    for the bilinear filter over the full input range."""
    scale = float(in_size) / float(out_size)
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        x = np.arange(xmax, dtype=np.float64)
        arg = np.abs((x + xmin - center + 0.5) * ss)
        w = np.where(arg < 1.0, 1.0 - arg, 0.0)
        ww = 0.0
        for v in w:                       # same left-to-right double sum as the C loop
            ww += float(v)
        if ww != 0.0:
            w = w / ww
        k = np.where(w < 0, -0.5 + w * (1 << PRECISION_BITS), 0.5 + w * (1 << PRECISION_BITS))
        kk[xx, :xmax] = k.astype(np.int64).astype(np.int32)   # C cast: truncation toward zero
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def _resample_axis0(img, out_size):
    """Resample axis 0 of a uint8 array [n, ...] to out_size."""
    bounds, kk = resample_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for j in range(n):
            acc += src[x0 + j] * int(kk[xx, j])
        out[xx] = _clip8(acc)
    return out


def pil_resize_bilinear_u8(img, out_h, out_w):
    """uint8 [H,W,C] -> uint8 [out_h,out_w,C] exactly as PIL `Image.resize((out_w, out_h), Image.BILINEAR)`."""
    img = np.ascontiguousarray(img)
    h, w = img.shape[:2]
    if w != out_w:
        img = np.swapaxes(_resample_axis0(np.swapaxes(img, 0, 1), out_w), 0, 1)
    if h != out_h:
        img = _resample_axis0(np.ascontiguousarray(img), out_h)
    return np.ascontiguousarray(img)


def prep_u8(rgb, mean_bgr):
    """uint8 RGB [H,W,3] (already resized) -> float32 [3,H,W]: ToTensor, BGR, Normalize, x255 (image_transform.py:10-13)."""
    x = rgb.astype(np.float32).transpose(2, 0, 1) / np.float32(255)      # ToTensor: .div(255) in fp32
    x = x[[2, 1, 0]]
    m = np.asarray(mean_bgr, dtype=np.float32).reshape(3, 1, 1)
    x = (x - m) / np.float32(1)
    return (x * np.float32(255)).astype(np.float32)


def preparation(rgb, size, mean_bgr):
    """`ImageTransform(size, mean).preparation` on a uint8 RGB array."""
    h, w = rgb.shape[:2]
    oh, ow = resize_target(h, w, size)
    if (oh, ow) != (h, w):
        rgb = pil_resize_bilinear_u8(rgb, oh, ow)
    return prep_u8(rgb, mean_bgr)


def post_preparation(x, mean_bgr):
    """float32 [3,H,W] -> uint8 RGB [H,W,3] (image_transform.py:16-31)."""
    x = np.asarray(x, dtype=np.float32) * np.float32(1.0 / 255)
    m = np.asarray([(-1) * v for v in mean_bgr], dtype=np.float32).reshape(3, 1, 1)
    x = (x - m) / np.float32(1)
    x = x[[2, 1, 0]]
    x = np.where(x > 1, np.float32(1), x)
    x = np.where(x < 0, np.float32(0), x)
    return (x * np.float32(255)).astype(np.uint8).transpose(1, 2, 0).copy()    # ToPILImage: mul(255).byte()


def hr_handoff(x_lo, size, mean_bgr):
    """hr_transfer_style.py:21-27 for the optimised image: 8-bit clamp, resize to `size`, re-preprocess."""
    return preparation(post_preparation(x_lo, mean_bgr), size, mean_bgr)
