"""Per-frame orchestration with the reference's signature (IST/model/engine/transfer_style.py:11-44), plus a batched variant
for the frame loop of IST/main.py:186-238 (independent frames of equal size optimised side by side on one GPU).
Pre- and post-processing run on the device (data.DeviceImageTransform, bit-identical to the reference's torchvision / PIL
pipeline): 8-bit images go up, 8-bit images come back."""
import os

import torch
from torch.autograd import Variable

from ...data import DeviceImageTransform
from ...util.logger import setup_logger
from .utils import transform_image, optimize_new

logger = setup_logger('style-transfer', False)


def do_transfer_style(cfg, model, content_image, style_image, device, content_only=False, style_only=False, opt='LBFGS',
                      saliency_map=False, return_tensor=False, save=True):
    """`return_tensor=True` (not in the reference) also returns the optimised float32 [1,3,h,w] device tensor, which
    `do_hr_transfer_style` accepts in place of the PIL image to keep the coarse-to-fine hand-off on the device."""
    logger.info("Start transferring.")
    if saliency_map:
        raise NotImplementedError("saliency maps are a debug utility of the reference (utils.py:104-161), outside the B200 hot path")
    image_transformer = DeviceImageTransform(cfg.DATA.IMG_SIZE, cfg.DATA.IMAGENET_MEAN, device)

    # transform images
    content_image = transform_image(image_transformer, content_image, device)
    style_image = transform_image(image_transformer, style_image, device)
    optimized_image = Variable(content_image.data.clone(), requires_grad=True)

    optimized_image = optimize_new(model, content_image, style_image, optimized_image, cfg, cfg.LOSS.MAX_ITER,
                                   content_only, style_only, opt)

    out_image = image_transformer.post_preparation(optimized_image.data[0])
    if save:      # transfer_style.py:43 writes OUTPUT.DIR + FILE_NAME for every frame; batch drivers with several ranks pass save=False
        os.makedirs(cfg.OUTPUT.DIR, exist_ok=True)
        out_image.save(cfg.OUTPUT.DIR + cfg.OUTPUT.FILE_NAME)
    if return_tensor:
        return out_image, optimized_image.data
    return out_image


def do_transfer_style_batch(cfg, model, content_images, style_image, device, style_tensor=None):
    """Several content frames against one shared style image (IST/main.py:184-238 processes them one at a time). Every frame
    is its own optimisation problem with its own L-BFGS state — results equal the per-frame calls — but the frames share each
    kernel launch, which fills the GPU on the deep VGG layers. All frames must have the same size after the transform.
    `style_tensor` (the already transformed style image) may be passed to reuse the cached Gram targets across calls."""
    logger.info("Start transferring a batch of %d frames." % len(content_images))
    image_transformer = DeviceImageTransform(cfg.DATA.IMG_SIZE, cfg.DATA.IMAGENET_MEAN, device)
    contents = [transform_image(image_transformer, im, device) for im in content_images]
    shapes = {tuple(c.shape) for c in contents}
    if len(shapes) != 1:
        raise ValueError("do_transfer_style_batch: frames of different sizes cannot share a batch: %s" % sorted(shapes))
    content = torch.cat(contents, dim=0).contiguous()
    if style_tensor is None:
        style_tensor = transform_image(image_transformer, style_image, device)
    optimized = Variable(content.data.clone(), requires_grad=True)
    optimized = optimize_new(model, content, style_tensor, optimized, cfg, cfg.LOSS.MAX_ITER)
    from PIL import Image
    out = image_transformer.post_u8(optimized.data).cpu().numpy()
    return [Image.fromarray(out[i], "RGB") for i in range(out.shape[0])]
