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


def frames(size, dev, kind="radar", h=None, w=None, cseed=1, sseed=2, style_kind="lidar"):
    mk = synth.radar_frame if kind == "radar" else synth.smooth_frame
    content = torch.from_numpy(synth.preprocess(mk(size, cseed, h=h, w=w))).to(dev)
    mks = synth.lidar_frame if style_kind == "lidar" else synth.smooth_frame
    style = torch.from_numpy(synth.preprocess(mks(size, sseed, h=h, w=w))).to(dev)
    return content, style


def noise_like(x, seed=3, scale=20.0):
    n = np.random.Generator(np.random.PCG64(seed)).standard_normal(tuple(x.shape)).astype(np.float32) * scale
    return torch.from_numpy(n).to(x.device)


def flip_aware_parity(plan, x, st64, st32, t64, t32, upto="relu5_1"):
    """SURVEY 8d parity protocol at one evaluation point x (fp32 [1,3,H,W] on the GPU). Compares the CUDA closure
    (plan.loss_and_grad, through the C ABI) with the fp64 oracle on the same device and returns a dict:
      loss_rel        max relative error of the six weighted layer losses + total vs fp64 (entries that are exactly 0 in the
                      oracle must be exactly 0 in ours)
      grad_rel, grad_cos1, grad_maxabs     plain image-gradient error vs fp64: rel-L2, 1 - cosine, max|diff| / max|g|
      flips, units    ReLU sign + pool argmax decisions of OUR forward that differ from the fp64 oracle's forward
      grad_rel_masked image-gradient rel-L2 vs the fp64 oracle evaluated WITH OUR MASKS (pure kernel arithmetic)
      loss_rel_masked same for the losses
      ref_*           the same quantities for the oracle's own fp32 run (cuDNN / cuBLAS fp32, TF32 off) vs its fp64 run
    """
    losses, grad = plan.loss_and_grad(x)
    losses, grad = losses.clone(), grad.clone()
    ours_masks = plan.masks(upto)
    x64 = x.double()
    l64, tot64, g64 = O.loss_and_grad(st64, x64, t64, full=False)
    m64 = O.forward_masks(st64, x64, upto)
    l32, tot32, g32 = O.loss_and_grad(st32, x, t32, full=False)
    m32 = O.forward_masks(st32, x, upto)

    def loss_rel(ours, ref, exact_zeros=True):
        # relative error per entry; an entry that is exactly 0 in the oracle (the content loss at x0) must be exactly 0 in ours.
        # With imposed masks that entry is ~1e-20 of the total instead of 0: entries below 1e-9 of the total are compared
        # against that floor.
        ours, ref = np.asarray(ours, dtype=np.float64), np.asarray(ref, dtype=np.float64)
        if exact_zeros:
            assert np.all(ours[ref == 0.0] == 0.0), (ours, ref)
        den = np.maximum(np.abs(ref), 1e-9 * np.abs(ref[-1]))
        return float(np.max(np.abs(ours - ref) / den))

    def grad_stats(g, ref):
        g, ref = g.double().flatten(), ref.double().flatten()
        return (((g - ref).norm() / ref.norm()).item(), 1.0 - float(torch.dot(g, ref) / (g.norm() * ref.norm())),
                ((g - ref).abs().max() / ref.abs().max()).item())

    def flips(a, b):
        mm = O.mask_mismatches(a, b)
        return sum(v[0] for v in mm.values()), sum(v[1] for v in mm.values()), {k: v[0] for k, v in mm.items() if v[0]}

    ours_l = losses[0].double().cpu().numpy()
    ref_l = np.array(l64 + [tot64])
    r = {}
    r["loss_rel"] = loss_rel(ours_l, ref_l)
    r["loss_rel_each"] = (np.abs(ours_l - ref_l) / np.maximum(np.abs(ref_l), 1e-300)).tolist()
    r["grad_rel"], r["grad_cos1"], r["grad_maxabs"] = grad_stats(grad, g64)
    r["flips"], r["units"], r["flips_by_layer"] = flips(ours_masks, m64)
    lm, totm, gm = O.loss_and_grad_masked(st64, x64, t64, ours_masks)
    r["loss_rel_masked"] = loss_rel(ours_l, np.array(lm + [totm]), exact_zeros=False)
    r["grad_rel_masked"] = grad_stats(grad, gm)[0]
    del gm, ours_masks
    r["ref_loss_rel"] = loss_rel(np.array(l32 + [tot32]), ref_l)
    r["ref_grad_rel"], r["ref_grad_cos1"], r["ref_grad_maxabs"] = grad_stats(g32, g64)
    r["ref_flips"], _, r["ref_flips_by_layer"] = flips(m32, m64)
    lm, totm, gm = O.loss_and_grad_masked(st64, x64, t64, m32)
    r["ref_grad_rel_masked"] = grad_stats(g32, gm)[0]
    r["total_loss"] = tot64
    return r


def parity_row(tag, r):
    return (f"{tag:<26s} losses {r['loss_rel']:.1e} (masked {r['loss_rel_masked']:.1e}) | grad rel-L2 {r['grad_rel']:.2e} 1-cos {r['grad_cos1']:.1e} "
            f"max {r['grad_maxabs']:.1e} | flips {r['flips']} of {r['units']} | grad rel-L2 with equal masks {r['grad_rel_masked']:.2e}"
            f"  ||  oracle fp32 vs fp64: losses {r['ref_loss_rel']:.1e} grad {r['ref_grad_rel']:.2e} flips {r['ref_flips']} "
            f"equal-mask grad {r['ref_grad_rel_masked']:.2e}")
