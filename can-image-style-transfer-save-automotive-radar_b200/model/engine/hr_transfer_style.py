"""Coarse-to-fine stage with the reference's signature (IST/model/engine/hr_transfer_style.py:11-33): re-preprocess the
content, the style and the (8-bit, clamped) low-resolution result at HRDATA.IMG_SIZE, then optimise HRLOSS.MAX_ITER
evaluations from that up-scaled initial image with a fresh optimiser. `optimized_image` may be the PIL image the reference
passes or the float32 [1,3,h,w] device tensor of the previous stage; both give the same initial image bit for bit (the
tensor goes through the same 8-bit clamp and bilinear resize, on the device)."""
import os

from torch.autograd import Variable

import torch

from ...data import DeviceImageTransform
from ...util.logger import setup_logger
from .utils import transform_image, optimize

logger = setup_logger('style-transfer', False)


def do_hr_transfer_style(cfg, model, content_image, style_image, optimized_image, device, return_tensor=False, save=True):
    logger.info("Start transferring to high resolution.")
    image_transformer = DeviceImageTransform(cfg.HRDATA.IMG_SIZE, cfg.DATA.IMAGENET_MEAN, device)

    # transform images
    content_image = transform_image(image_transformer, content_image, device)
    style_image = transform_image(image_transformer, style_image, device)
    if torch.is_tensor(optimized_image) and optimized_image.is_floating_point():
        optimized_image = image_transformer.handoff(optimized_image if optimized_image.dim() == 4 else optimized_image.unsqueeze(0))
    else:
        optimized_image = transform_image(image_transformer, optimized_image, device)
    optimized_image = Variable(optimized_image.type_as(content_image.data), requires_grad=True)

    optimized_image = optimize(model, content_image, style_image, optimized_image, cfg, cfg.HRLOSS.MAX_ITER)

    out_image = image_transformer.post_preparation(optimized_image.data[0])
    if save:
        os.makedirs(cfg.OUTPUT.DIR, exist_ok=True)
        out_image.save(cfg.OUTPUT.DIR + cfg.OUTPUT.HR_FILE_NAME)
    if return_tensor:
        return out_image, optimized_image.data
    return out_image
