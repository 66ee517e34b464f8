"""Coarse-to-fine stage with the reference's signature (IST/model/engine/hr_transfer_style.py:11-33): re-preprocess the
content, the style and the (8-bit, clamped) low-resolution result at HRDATA.IMG_SIZE, then optimise HRLOSS.MAX_ITER
evaluations from that up-scaled initial image with a fresh optimiser."""
import os

from torch.autograd import Variable

from ...data import ImageTransform
from ...util.logger import setup_logger
from .utils import transform_image, optimize

logger = setup_logger('style-transfer', False)


def do_hr_transfer_style(cfg, model, content_image, style_image, optimized_image, device):
    logger.info("Start transferring to high resolution.")
    image_transformer = ImageTransform(cfg.HRDATA.IMG_SIZE, cfg.DATA.IMAGENET_MEAN)

    # transform images
    content_image = transform_image(image_transformer, content_image, device)
    style_image = transform_image(image_transformer, style_image, device)
    optimized_image = transform_image(image_transformer, optimized_image, device)
    optimized_image = Variable(optimized_image.type_as(content_image.data), requires_grad=True)

    optimized_image = optimize(model, content_image, style_image, optimized_image, cfg, cfg.HRLOSS.MAX_ITER)

    out_image = image_transformer.post_preparation(optimized_image.data[0].cpu().squeeze())
    os.makedirs(cfg.OUTPUT.DIR, exist_ok=True)
    out_image.save(cfg.OUTPUT.DIR + cfg.OUTPUT.HR_FILE_NAME)
    return out_image
