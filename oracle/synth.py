"""Synthetic, seed-reproducible inputs for the IST Gatys path (TEST INFRASTRUCTURE — not shipped, not on the product path).

The reference's real weights (`vgg_conv.pth`, IST/util/download_models.sh:3) and radar/lidar frames are not available
offline, so parity runs on:
  * Kaiming-normal (fan_in, ReLU gain) VGG19 weights with zero bias and the reference's state-dict keys
    (`conv{b}_{i}.weight` [Cout,Cin,3,3], `.bias` [Cout]; IST/model/meta_arch/vgg.py:24-39). PyTorch's default conv init
    would make the deep style losses vanish, so it is not used.
  * radar-like frames (sparse 255 points on black), lidar-like style frames (points + axis-aligned segments) and smooth
    frames (bicubic-upsampled noise), as uint8 RGB arrays, per SURVEY 8d.
Everything is generated with numpy's PCG64 (stable across numpy versions and machines), never with torch's RNG.
"""
import numpy as np

# (name, cin, cout) in the order of cfg.MODEL.VGG.CONV_LAYERS_DICT (IST/config/defaults.py:22-40)
VGG19_CONVS = [
    ("conv1_1", 3, 64), ("conv1_2", 64, 64),
    ("conv2_1", 64, 128), ("conv2_2", 128, 128),
    ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256), ("conv3_4", 256, 256),
    ("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512), ("conv4_4", 512, 512),
    ("conv5_1", 512, 512), ("conv5_2", 512, 512), ("conv5_3", 512, 512), ("conv5_4", 512, 512),
]
IMAGENET_MEAN = [0.40760392, 0.45795686, 0.48501961]  # cfg.DATA.IMAGENET_MEAN, IST/config/defaults.py:86 (BGR order)


def vgg_state_dict(seed=0, upto=None, bias_std=0.0):
    """dict name -> float32 ndarray; one independent PCG64 stream per layer so `upto` does not change earlier layers."""
    out = {}
    for i, (name, cin, cout) in enumerate(VGG19_CONVS):
        rng = np.random.Generator(np.random.PCG64([seed, i]))
        std = np.sqrt(2.0 / (cin * 9))
        out[name + ".weight"] = (rng.standard_normal((cout, cin, 3, 3)) * std).astype(np.float32)
        if bias_std > 0:
            out[name + ".bias"] = (rng.standard_normal(cout) * bias_std).astype(np.float32)
        else:
            out[name + ".bias"] = np.zeros(cout, dtype=np.float32)
        if upto is not None and name == upto:
            break
    return out


def radar_frame(size, seed, h=None, w=None):
    """uint8 [H,W,3]: black background, ~600*(S/256)^2 points at 255, replicated to RGB (IST/main.py:185,206 `.convert('RGB')`)."""
    h = h or size
    w = w or size
    rng = np.random.Generator(np.random.PCG64([7, seed]))
    n = int(round(600 * (h * w) / 256.0 ** 2))
    img = np.zeros((h, w), dtype=np.uint8)
    ys = rng.integers(0, h, n)
    xs = rng.integers(0, w, n)
    img[ys, xs] = 255
    return np.repeat(img[:, :, None], 3, axis=2)


def lidar_frame(size, seed, h=None, w=None):
    """uint8 [H,W,3]: ~3000*(S/256)^2 points plus 40 axis-aligned segments of length S/16..S/4."""
    h = h or size
    w = w or size
    rng = np.random.Generator(np.random.PCG64([11, seed]))
    n = int(round(3000 * (h * w) / 256.0 ** 2))
    img = np.zeros((h, w), dtype=np.uint8)
    img[rng.integers(0, h, n), rng.integers(0, w, n)] = 255
    s = min(h, w)
    for _ in range(40):
        ln = int(rng.integers(max(1, s // 16), max(2, s // 4)))
        y0 = int(rng.integers(0, h))
        x0 = int(rng.integers(0, w))
        if rng.integers(0, 2) == 0:
            img[y0, x0:min(w, x0 + ln)] = 255
        else:
            img[y0:min(h, y0 + ln), x0] = 255
    return np.repeat(img[:, :, None], 3, axis=2)


def smooth_frame(size, seed, h=None, w=None):
    """uint8 [H,W,3]: uniform noise [3,S/8,S/8] bicubic-upsampled to S and clamped (no exact ties, no flat background)."""
    import torch
    import torch.nn.functional as F
    h = h or size
    w = w or size
    rng = np.random.Generator(np.random.PCG64([13, seed]))
    low = rng.random((1, 3, max(2, h // 8), max(2, w // 8))).astype(np.float32)
    up = F.interpolate(torch.from_numpy(low), size=(h, w), mode="bicubic", align_corners=False).clamp(0, 1)
    return (up[0].permute(1, 2, 0).numpy() * 255.0 + 0.5).astype(np.uint8)


def preprocess(rgb_u8):
    """uint8 [H,W,3] RGB -> float32 [1,3,H,W]: ToTensor, RGB->BGR, -mean, x255 (IST/data/image_transform.py:8-14, no resize)."""
    x = rgb_u8.astype(np.float32) / np.float32(255.0)          # ToTensor
    x = np.transpose(x, (2, 0, 1))[[2, 1, 0]]                  # BGR
    x = x - np.asarray(IMAGENET_MEAN, dtype=np.float32)[:, None, None]
    x = x * np.float32(255.0)
    return np.ascontiguousarray(x[None]).astype(np.float32)


def postprocess_float(x):
    """float [3,H,W] (network space) -> float RGB in [0,255] before uint8 truncation (image_transform.py:16-31); for PSNR."""
    y = x.astype(np.float64) / 255.0 + np.asarray(IMAGENET_MEAN, dtype=np.float64)[:, None, None]
    y = y[[2, 1, 0]]
    return np.clip(y, 0.0, 1.0) * 255.0


def psnr(a, b):
    mse = float(np.mean((postprocess_float(a) - postprocess_float(b)) ** 2))
    return float("inf") if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)
