"""Can the data-gradient run with fewer than three MMAs per product?  (test infrastructure: executes oracle/ on the CPU)

VERDICT r01 item 4 asks for the backward as 2 passes (dY hi/lo x ONE 16-bit weight plane) or 1 pass, to be kept only if the
gradient error with equal masks stays <= 1e-4. With the ReLU / pool decisions fixed (the forward's, float64) the backward is a
linear map; this script applies it in float64 with the WEIGHTS (2-pass scheme) or weights AND incoming gradients (1-pass
scheme) rounded to fp16 / bf16 per layer — exactly the information a tensor-core pass with those operand planes sees, with
exact accumulation — and reports the rel-L2 distance of the image gradient from the unrounded float64 backward.

    python tools/cpu_backward_precision_study.py [SIZE=128]
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ist_oracle as O  # noqa: E402
from oracle import synth  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
torch.set_num_threads(8)
state = O.state_to_torch(synth.vgg_state_dict(0, upto="conv5_1"), torch.float64)
SEQ, OUT = O.FORWARD_SEQ[:O.OUT_SEQ.index("relu5_1") + 1], O.OUT_SEQ[:O.OUT_SEQ.index("relu5_1") + 1]


def rnd(t, fmt):
    if fmt is None:
        return t
    # per-tensor power-of-two scaling into the format's normal range, as the kernels scale their weight planes
    s = 2.0 ** (13 - int(torch.floor(torch.log2(t.abs().max()))))
    dt = torch.float16 if fmt == "fp16" else torch.bfloat16
    return (t * s).to(dt).double() / s


def backward(x, targets, w_fmt, g_fmt):
    """image gradient with the forward in float64 and the backward's operands rounded per layer"""
    xs = [x]
    masks = O.forward_masks(state, x)
    feats = {}
    prev = x
    for name, out in zip(SEQ, OUT):
        if "conv" in name:
            prev = F.conv2d(prev, state[name + ".weight"], state[name + ".bias"], padding=1) * masks[out].double()
        else:
            b, c, h, w = prev.shape
            win = prev[:, :, :2 * (h // 2), :2 * (w // 2)].reshape(b, c, h // 2, 2, w // 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(b, c, h // 2, w // 2, 4)
            prev = torch.gather(win, 4, masks[out].long().clamp(max=3).unsqueeze(-1)).squeeze(-1)
        feats[out] = prev
    # seeds dL/dF at the loss layers (exact, float64 autograd on the losses only)
    seeds = {}
    keys = O.STYLE_LAYERS + O.CONTENT_LAYERS
    w = O.STYLE_WEIGHTS + O.CONTENT_WEIGHTS
    for a, k in enumerate(keys):
        f = feats[k].detach().clone().requires_grad_(True)
        loss = w[a] * (O.gram_mse_loss(f, targets[a]) if a < 5 else F.mse_loss(f, targets[a]))
        loss.backward()
        seeds[k] = f.grad
    g = None
    for i in range(len(SEQ) - 1, -1, -1):
        name, out = SEQ[i], OUT[i]
        if out in seeds:
            g = seeds[out] if g is None else g + seeds[out]
        if g is None:
            continue
        if "conv" in name:
            g = g * masks[out].double()                                   # dY of this conv (what the kernel stores as planes)
            g = F.conv_transpose2d(rnd(g, g_fmt), rnd(state[name + ".weight"], w_fmt), padding=1)
        else:
            b, c, ho, wo = g.shape
            h, wd = feats[OUT[i - 1]].shape[2:]
            full = torch.zeros(b, c, ho, wo, 4, dtype=g.dtype)
            full.scatter_(4, masks[out].long().clamp(max=3).unsqueeze(-1), g.unsqueeze(-1))
            up = torch.zeros(b, c, h, wd, dtype=g.dtype)
            up[:, :, :2 * ho, :2 * wo] = full.reshape(b, c, ho, wo, 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(b, c, 2 * ho, 2 * wo)
            g = up
    return g


for kind in ("radar", "smooth"):
    mk = synth.radar_frame if kind == "radar" else synth.smooth_frame
    content = torch.from_numpy(synth.preprocess(mk(size, 1))).double()
    style = torch.from_numpy(synth.preprocess(synth.lidar_frame(size, 2))).double()
    targets = O.compute_targets(state, content, style, full=False)
    x = content + 20.0 * torch.randn(content.shape, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
    ref = backward(x, targets, None, None)
    _, _, g_auto = O.loss_and_grad(state, x, targets, full=False)
    print(f"{size}^2 {kind}: manual float64 backward vs autograd {float((ref - g_auto).norm() / g_auto.norm()):.1e}")
    for label, wf, gf in (("2 passes: dY hi/lo x fp16 weight plane", "fp16", None), ("2 passes: dY hi/lo x bf16 weight plane", "bf16", None),
                          ("1 pass: fp16 dY x fp16 weights", "fp16", "fp16"), ("1 pass: bf16 dY x bf16 weights (IST_B200_PASSES_BWD=1)", "bf16", "bf16")):
        g = backward(x, targets, wf, gf)
        print(f"    {label:58s} gradient rel-L2 with equal masks {float((g - ref).norm() / ref.norm()):.2e}")
