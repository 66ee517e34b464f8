"""Per-frame orchestration with the reference's signature (IST/model/engine/transfer_style.py:11-44)."""
import os

from torch.autograd import Variable

from ...data import ImageTransform
from ...util.logger import setup_logger
from .utils import transform_image, optimize_new

logger = setup_logger('style-transfer', False)


def do_transfer_style(cfg, model, content_image, style_image, device, content_only=False, style_only=False, opt='LBFGS',
                      saliency_map=False):
    logger.info("Start transferring.")
    if saliency_map:
        raise NotImplementedError("saliency maps are a debug utility of the reference (utils.py:104-161), outside the B200 hot path")
    image_transformer = ImageTransform(cfg.DATA.IMG_SIZE, cfg.DATA.IMAGENET_MEAN)

    # transform images
    content_image = transform_image(image_transformer, content_image, device)
    style_image = transform_image(image_transformer, style_image, device)
    optimized_image = Variable(content_image.data.clone(), requires_grad=True)

    optimized_image = optimize_new(model, content_image, style_image, optimized_image, cfg, cfg.LOSS.MAX_ITER,
                                   content_only, style_only, opt)

    out_image = image_transformer.post_preparation(optimized_image.data[0].cpu().squeeze())
    os.makedirs(cfg.OUTPUT.DIR, exist_ok=True)
    out_image.save(cfg.OUTPUT.DIR + cfg.OUTPUT.FILE_NAME)
    return out_image
