"""The optimisation loop with the reference's signatures (IST/model/engine/utils.py:11-161).

``optimize`` keeps the reference's contract — targets from the style / content images, a fresh L-BFGS with torch's
defaults, `max_iterations` counted in closure evaluations and checked between optimizer steps, the optimised image
updated in place and returned — but runs the closure (VGG forward, Gram / content losses, image gradient) and the
optimiser on the GPU inside libist_b200.so: one CUDA graph and one host synchronisation per `optimizer.step`.
Style Gram targets are cached per (style image, layer set) instead of being recomputed for every frame
(the reference recomputes them at utils.py:19 although IST/main.py:184 shares one style image).
"""
import torch
import torch.nn as nn
from torch import optim
from torch.autograd import Variable

from ... import _lib
from ...lbfgs import DeviceLBFGS
from ...util.logger import setup_logger
from ..meta_arch import GramMatrix, GramMSELoss, StyleTransfer  # noqa: F401  (re-exported like the reference)

logger = setup_logger('style-transfer', False)


def transform_image(image_transformer, image, device):
    image_transformed = image_transformer.preparation(image)
    image_transformed = Variable(image_transformed.unsqueeze(0).to(device))
    return image_transformed


def _fused_spec(model, cfg):
    """(style_keys, style_w, content_keys, content_w) when the model is the Gatys configuration the fused closure covers:
    GramMSELoss on the style layers followed by nn.MSELoss on the content layers (IST/main.py:35-41)."""
    layers, fns, ws = list(model.loss_layers), list(model.loss_functions), list(model.loss_weights)
    if not (len(layers) == len(fns) == len(ws)):
        return None
    ns = len(cfg.LOSS.STYLE_LAYERS)
    if layers != list(cfg.LOSS.STYLE_LAYERS) + list(cfg.LOSS.CONTENT_LAYERS):
        return None
    if not all(isinstance(f, GramMSELoss) for f in fns[:ns]):
        return None
    if not all(type(f) is nn.MSELoss and f.reduction == 'mean' for f in fns[ns:]):
        return None
    if any(k.startswith('pool') for k in layers):
        return None
    return layers[:ns], ws[:ns], layers[ns:], ws[ns:]


def _deepest(vgg, keys):
    return max(keys, key=lambda k: vgg.out_seq.index(k))


def style_targets(vgg, style_image, style_keys):
    """Gram targets of the style image, cached on the VGG module for the style tensor last used (same tensor object, same
    version counter, same layer set and weights). The cache holds a reference to the tensor, so its storage cannot be
    recycled for a different image while the entry is alive."""
    cache = vgg.__dict__.setdefault('_style_target_cache', {})
    hit = cache.get('entry')
    if hit is not None:
        ref, version, keys, token, grams = hit
        if ref is style_image and version == style_image._version and keys == tuple(style_keys) and token == vgg._token():
            return grams
    b, _, hs, ws = style_image.shape
    with torch.no_grad():
        splan = vgg.plan(b, hs, ws, _deepest(vgg, style_keys), device=style_image.device)
        splan.forward(style_image.contiguous().float(), _deepest(vgg, style_keys))
        grams = [splan.gram(k) for k in style_keys]
    cache['entry'] = (style_image, style_image._version, tuple(style_keys), vgg._token(), grams)   # one live style (main.py:184)
    return grams


def optimize(model, content_image, style_image, optimized_image, cfg, max_iterations):
    spec = _fused_spec(model, cfg)
    if spec is None or not optimized_image.is_cuda:
        if not optimized_image.is_cuda:
            raise _lib.IstError("MODEL.DEVICE must be a CUDA device: this package has no CPU path (use the oracle for CPU runs)")
        return _optimize_autograd(model, content_image, style_image, optimized_image, cfg, max_iterations)
    style_keys, style_w, content_keys, content_w = spec
    vgg = model.vgg_model
    x = optimized_image.data
    if not (x.dtype == torch.float32 and x.is_contiguous()):
        raise _lib.IstError("optimized_image must be contiguous float32")
    nb, _, h, w = x.shape
    deepest = _deepest(vgg, style_keys + content_keys)

    # compute optimization targets (utils.py:19-21)
    grams = style_targets(vgg, style_image, style_keys)
    plan = vgg.plan(nb, h, w, deepest, device=x.device)
    plan.set_loss(style_keys, style_w, content_keys, content_w)
    for k, g in enumerate(grams):
        plan.set_style_target(k, g[0])
    content = content_image.data.contiguous().float()
    if content.shape[0] != nb:
        content = content.expand(nb, -1, -1, -1).contiguous()
    plan.forward(content, _deepest(vgg, content_keys))
    for k in range(len(content_keys)):
        plan.capture_content_target(k)

    # create optimizer (utils.py:24) and run (utils.py:28-43). The device optimiser (history buffers + the captured CUDA graph
    # of one step) is kept on the plan and reset to the state of a fresh optim.LBFGS([x]) for every call.
    optimizer = getattr(plan, "_lbfgs", None)
    loss_key = (tuple(style_keys), tuple(float(v) for v in style_w), tuple(content_keys), tuple(float(v) for v in content_w))
    if optimizer is None or optimizer.h is None or getattr(plan, "_lbfgs_key", None) != loss_key:
        if optimizer is not None:
            optimizer.close()                 # the captured graph bakes in the loss configuration
        optimizer = DeviceLBFGS(plan)
        plan._lbfgs = optimizer
        plan._lbfgs_key = loss_key
    optimizer.reset()
    iterations = [0]
    while iterations[0] < max_iterations:
        evals, _ = optimizer.step(x)
        iterations[0] += evals
        if evals == 0:
            raise _lib.IstError("L-BFGS made no closure evaluation")
    model.last_losses = optimizer.last_losses()
    model.last_evals = iterations[0]
    return optimized_image


def _optimize_autograd(model, content_image, style_image, optimized_image, cfg, max_iterations):
    """The reference loop verbatim in structure (utils.py:17-45) for loss configurations the fused closure does not cover;
    VGG, Gram and the losses still run in the CUDA library through their autograd wrappers."""
    style_targets_ = [GramMatrix()(A).detach() for A in model.vgg_model(style_image, cfg.LOSS.STYLE_LAYERS)]
    content_targets = [A.detach() for A in model.vgg_model(content_image, cfg.LOSS.CONTENT_LAYERS)]
    targets = style_targets_ + content_targets
    optimizer = optim.LBFGS([optimized_image])
    iterations = [0]
    while iterations[0] < max_iterations:
        def closure():
            optimizer.zero_grad()
            outputs = model.vgg_model(optimized_image, model.loss_layers)
            layer_losses = [model.loss_weights[a] * model.loss_functions[a](A, targets[a]) for a, A in enumerate(outputs)]
            loss = sum(layer_losses)
            loss.backward()
            iterations[0] += 1
            return loss
        optimizer.step(closure)
    return optimized_image


def optimize_new(model, content_image, style_image, optimized_image, cfg, max_iterations, content_only=False,
                 style_only=False, opt="LBFGS"):
    """utils.py:47-102. The default branch is `optimize`; the content-only / style-only branches are the reference's
    single-step, negated-loss debug variants and run through the autograd wrappers."""
    device = torch.device(cfg.MODEL.DEVICE)
    if (content_only is False) & (style_only is False):
        return optimize(model, content_image, style_image, optimized_image, cfg, max_iterations)
    elif content_only:
        targets = [A.detach() for A in model.vgg_model(content_image, cfg.LOSS.CONTENT_LAYERS)]
        loss_layers = cfg.LOSS.CONTENT_LAYERS
        loss_functions = [nn.MSELoss().to(device)] * len(cfg.LOSS.CONTENT_LAYERS)
        loss_weights = cfg.LOSS.CONTENT_WEIGHTS
    else:
        targets = [GramMatrix()(A).detach() for A in model.vgg_model(style_image, cfg.LOSS.STYLE_LAYERS)]
        loss_layers = cfg.LOSS.STYLE_LAYERS
        loss_functions = [GramMSELoss().to(device)] * len(cfg.LOSS.STYLE_LAYERS)
        loss_weights = cfg.LOSS.STYLE_WEIGHTS
    optimizer = optim.LBFGS([optimized_image]) if opt == "LBFGS" else optim.Adam([optimized_image])
    iterations = [0]
    while iterations[0] < max_iterations:
        def closure():
            optimizer.zero_grad()
            outputs = model.vgg_model(optimized_image, loss_layers)
            layer_losses = [loss_weights[a] * loss_functions[a](A, targets[a]) for a, A in enumerate(outputs)]
            loss = - sum(layer_losses)
            loss.backward()
            iterations[0] += 1
            return loss
        optimizer.step(closure)
        break
    return optimized_image
