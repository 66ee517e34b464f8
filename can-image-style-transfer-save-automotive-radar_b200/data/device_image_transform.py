"""ImageTransform with the reference's method names (IST/data/image_transform.py:5-31) computed on the GPU.

`preparation` (Scale -> ToTensor -> BGR -> Normalize -> x255) and `post_preparation` (x1/255 -> +mean -> RGB -> clamp ->
ToPILImage) run as CUDA kernels of libist_b200.so on 8-bit device images, bit-identical to the torchvision / PIL pipeline
(the PIL bilinear resize is Pillow's two-pass fixed-point resampling, reproduced in integer arithmetic). Only the 8-bit
image crosses PCIe (3 bytes per pixel instead of 12), and the coarse-to-fine hand-off of
IST/model/engine/hr_transfer_style.py:21-27 (`handoff`) never leaves the device.
"""
import ctypes

import numpy as np
import torch
from PIL import Image

from .. import _lib


def _u8ptr(t):
    if not (t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous()):
        raise _lib.IstError("expected a contiguous uint8 CUDA tensor")
    return ctypes.c_void_p(t.data_ptr())


class DeviceImageTransform:

    def __init__(self, image_size, imagenet_mean, device):
        self.image_size = int(image_size)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.IstError("DeviceImageTransform runs on a CUDA device (the CPU pipeline is data.ImageTransform)")
        if len(imagenet_mean) != 3:
            raise ValueError("imagenet_mean must have three entries (BGR order, IST/config/defaults.py:86)")
        self.mean = (ctypes.c_double * 3)(*[float(m) for m in imagenet_mean])
        self.lib = _lib.load()

    # ---- pieces -------------------------------------------------------------------------------------------------------
    def resize_target(self, h, w):
        oh, ow = ctypes.c_int(), ctypes.c_int()
        _lib.check(self.lib.ist_image_resize_target(int(h), int(w), self.image_size, ctypes.byref(oh), ctypes.byref(ow)))
        return oh.value, ow.value

    def upload(self, image):
        """PIL image / uint8 HWC array -> uint8 [1,H,W,3] on the device (through pinned memory)."""
        if isinstance(image, Image.Image):
            if image.mode != "RGB":
                # ToTensor keeps the image's own channels; the IST loop only ever feeds RGB (IST/main.py:185,206 `.convert('RGB')`)
                raise _lib.IstError(f"expected an RGB image, got mode {image.mode}: call .convert('RGB') as IST/main.py does")
            arr = np.asarray(image)
        else:
            arr = np.asarray(image)
        if arr.dtype != np.uint8 or arr.ndim != 3 or arr.shape[2] != 3:
            raise _lib.IstError(f"expected uint8 [H,W,3], got {arr.dtype} {arr.shape}")
        host = torch.from_numpy(np.array(arr, copy=True, order="C")).pin_memory()
        return host.to(self.device, non_blocking=True).unsqueeze(0)

    def resize_u8(self, rgb):
        """uint8 [B,H,W,3] device -> Scale(image_size) of it (same tensor if it already has that size)."""
        with torch.cuda.device(self.device):
            b, h, w, _ = rgb.shape
            oh, ow = self.resize_target(h, w)
            if (oh, ow) == (h, w):
                return rgb
            out = torch.empty(b, oh, ow, 3, dtype=torch.uint8, device=self.device)
            tmp = torch.empty(b, h, ow, 3, dtype=torch.uint8, device=self.device) if (oh != h and ow != w) else None
            _lib.check(self.lib.ist_image_resize_u8(_u8ptr(rgb), _u8ptr(out), _u8ptr(tmp) if tmp is not None else None, b, h, w,
                                                    oh, ow, _lib.stream_ptr(self.device)))
            return out

    def prep_u8(self, rgb):
        """uint8 [B,H,W,3] device -> float32 [B,3,H,W] (ToTensor, BGR, Normalize, x255)."""
        with torch.cuda.device(self.device):
            b, h, w, _ = rgb.shape
            x = torch.empty(b, 3, h, w, dtype=torch.float32, device=self.device)
            _lib.check(self.lib.ist_image_prep_u8(_u8ptr(rgb), _lib.ptr(x), b, h, w, self.mean, _lib.stream_ptr(self.device)))
            return x

    def post_u8(self, x):
        """float32 [B,3,H,W] (or [3,H,W]) device -> uint8 RGB [B,H,W,3] device."""
        x = x.detach()
        if x.dim() == 3:
            x = x.unsqueeze(0)
        x = x.contiguous().float()
        with torch.cuda.device(self.device):
            b, _, h, w = x.shape
            rgb = torch.empty(b, h, w, 3, dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.ist_image_post_u8(_lib.ptr(x), _u8ptr(rgb), b, h, w, self.mean, _lib.stream_ptr(self.device)))
            return rgb

    # ---- the reference's two methods ------------------------------------------------------------------------------------
    def preparation(self, image):
        """PIL image (or uint8 HWC array, or uint8 [H,W,3] device tensor) -> float32 [3,h,w] device tensor."""
        rgb = image.unsqueeze(0) if (torch.is_tensor(image) and image.dim() == 3) else (image if torch.is_tensor(image) else self.upload(image))
        return self.prep_u8(self.resize_u8(rgb.contiguous()))[0]

    def post_preparation(self, tensor):
        """float32 [3,H,W] tensor -> PIL image; only the 8-bit image is copied to the host."""
        rgb = self.post_u8(tensor.to(self.device))[0]
        return Image.fromarray(rgb.cpu().numpy(), "RGB")

    # ---- coarse-to-fine hand-off ------------------------------------------------------------------------------------------
    def handoff(self, x_lo):
        """hr_transfer_style.py:21-27 for the optimised image, on the device: float32 [B,3,h,w] at the low resolution ->
        float32 [B,3,H,W] at `image_size` (8-bit clamp, bilinear resize, re-preprocess)."""
        x_lo = x_lo.detach().contiguous().float()
        with torch.cuda.device(self.device):
            b, _, h, w = x_lo.shape
            oh, ow = self.resize_target(h, w)
            nbytes = int(self.lib.ist_image_handoff_workspace(b, h, w, oh, ow))
            work = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            x_hi = torch.empty(b, 3, oh, ow, dtype=torch.float32, device=self.device)
            _lib.check(self.lib.ist_image_handoff(_lib.ptr(x_lo), _lib.ptr(x_hi), _u8ptr(work), nbytes, b, h, w, oh, ow, self.mean,
                                                  _lib.stream_ptr(self.device)))
            return x_hi
