"""Shared helpers of the GPU parity tests (imported by tests only)."""
import numpy as np
import torch

from oracle import ist_oracle as O
from oracle import synth

STYLE, CONTENT = O.STYLE_LAYERS, O.CONTENT_LAYERS


def strict_fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / (b.norm() + 1e-300)).item()


def build_model(dev, state_np=None):
    from ist_b200.config import get_cfg_defaults
    from ist_b200.main import get_model
    cfg = get_cfg_defaults()
    cfg.MODEL.DEVICE = str(dev)
    state_np = state_np if state_np is not None else synth.vgg_state_dict(0)
    model, _ = get_model(cfg, {k: torch.from_numpy(v) for k, v in state_np.items()})
    return cfg, model


def prepare_plan(model, cfg, content, style, batch=None):
    """Plan with the Gatys loss configuration and targets for (content, style), through the product API."""
    from ist_b200.model.engine.utils import style_targets
    vgg = model.vgg_model
    nb = batch or content.shape[0]
    plan = vgg.plan(nb, content.shape[2], content.shape[3], "relu5_1", device=content.device)
    plan.set_loss(cfg.LOSS.STYLE_LAYERS, cfg.LOSS.STYLE_WEIGHTS, cfg.LOSS.CONTENT_LAYERS, cfg.LOSS.CONTENT_WEIGHTS)
    for k, g in enumerate(style_targets(vgg, style, cfg.LOSS.STYLE_LAYERS)):
        plan.set_style_target(k, g[0])
    c = content if content.shape[0] == nb else content.expand(nb, -1, -1, -1).contiguous()
    plan.forward(c, "relu4_2")
    plan.capture_content_target(0)
    return plan


def frames(size, dev, kind="radar", h=None, w=None, cseed=1, sseed=2):
    mk = synth.radar_frame if kind == "radar" else synth.smooth_frame
    content = torch.from_numpy(synth.preprocess(mk(size, cseed, h=h, w=w))).to(dev)
    style = torch.from_numpy(synth.preprocess(synth.lidar_frame(size, sseed, h=h, w=w))).to(dev)
    return content, style


def noise_like(x, seed=3, scale=20.0):
    n = np.random.Generator(np.random.PCG64(seed)).standard_normal(tuple(x.shape)).astype(np.float32) * scale
    return torch.from_numpy(n).to(x.device)
