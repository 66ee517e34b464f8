"""Plan-level parity + timing on a B200 (run under gpurun): closure losses and image gradient of the CUDA path vs the
oracle (oracle/ist_oracle.py) in fp64 and fp32 on the same GPU, then CUDA-event timing of the closure.
Usage: python tools/gpu_plan_check.py [sizes...]   (default 64 256 512)
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ist_b200  # noqa: E402
from oracle import ist_oracle as O  # noqa: E402
from oracle import synth  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")

LAYERS = []
for f, o in zip(O.FORWARD_SEQ, O.OUT_SEQ):
    if f.startswith("conv"):
        cin, cout = [(a, b) for (n, a, b) in synth.VGG19_CONVS if n == f][0]
        LAYERS.append((0, cin, cout, f, o))
    else:
        LAYERS.append((1, 0, 0, f, o))
    if o == "relu5_1":
        break


def stats(name, g, ref):
    g = g.double().flatten()
    ref = ref.double().flatten()
    r2 = ((g - ref).norm() / ref.norm()).item()
    cos = 1.0 - (torch.dot(g, ref) / (g.norm() * ref.norm())).item()
    mx = ((g - ref).abs().max() / ref.abs().max()).item()
    print(f"    {name:34s} rel-L2 {r2:.3e}  1-cos {cos:.3e}  max/max {mx:.3e}")
    return r2


def run(size, kind):
    print(f"=== size {size} input {kind} ===")
    state_np = synth.vgg_state_dict(seed=0, upto="conv5_1")
    st32 = O.state_to_torch(state_np, torch.float32, dev)
    st64 = O.state_to_torch(state_np, torch.float64, dev)
    mk = synth.radar_frame if kind == "radar" else synth.smooth_frame
    content = torch.from_numpy(synth.preprocess(mk(size, 1))).to(dev)
    style = torch.from_numpy(synth.preprocess(synth.lidar_frame(size, 2))).to(dev)
    gen = torch.Generator().manual_seed(3)
    x1 = content + (torch.randn(content.shape, generator=gen) * 20.0).to(dev)

    plan = ist_b200.Plan(LAYERS, 1, size, size)
    plan.load_state_dict(st32)
    plan.set_loss(O.STYLE_LAYERS, O.STYLE_WEIGHTS, O.CONTENT_LAYERS, O.CONTENT_WEIGHTS)
    # targets through our own kernels (style Gram targets + content features)
    plan.forward(style, "relu5_1")
    our_grams = []
    for k, key in enumerate(O.STYLE_LAYERS):
        our_grams.append(plan.gram(key))
        plan.set_style_target(k, our_grams[-1])
    plan.forward(content, "relu4_2")
    plan.capture_content_target(0)
    print(f"  plan bytes {plan.nbytes / 1e6:.1f} MB")

    t64 = O.compute_targets(st64, content.double(), style.double(), full=False)
    t32 = O.compute_targets(st32, content, style, full=False)
    for k, key in enumerate(O.STYLE_LAYERS):
        stats(f"style target {key} ours", our_grams[k], t64[k])
        stats(f"style target {key} oracle-fp32", t32[k], t64[k])
    for name, x in (("P0=content", content), ("P1=content+N(0,20)", x1)):
        losses, grad = plan.loss_and_grad(x)
        torch.cuda.synchronize()
        l64, tot64, g64 = O.loss_and_grad(st64, x.double(), t64, full=False)
        l32, tot32, g32 = O.loss_and_grad(st32, x, t32, full=False)
        ours = losses[0].tolist()
        print(f"  {name}")
        print("    losses ours   ", ["%.6e" % v for v in ours])
        print("    losses fp64   ", ["%.6e" % v for v in l64 + [tot64]])
        print("    losses fp32ref", ["%.6e" % v for v in l32 + [tot32]])
        rel = [abs(a - b) / (abs(b) + 1e-30) for a, b in zip(ours, l64 + [tot64]) if abs(b) > 0]
        rel32 = [abs(a - b) / (abs(b) + 1e-30) for a, b in zip(l32 + [tot32], l64 + [tot64]) if abs(b) > 0]
        print(f"    max loss rel err ours {max(rel):.3e}   oracle-fp32 {max(rel32):.3e}")
        stats("grad ours vs fp64", grad, g64)
        stats("grad oracle-fp32 vs fp64", g32, g64)
        stats("grad ours vs oracle-fp32", grad, g32)
        if not torch.isfinite(grad).all():
            print("    NON-FINITE gradient!")
    # timing
    x = x1.clone()
    grad = torch.empty_like(x)
    losses = torch.empty(1, 7, device=dev)
    for _ in range(3):
        plan.loss_and_grad(x, grad, losses)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        plan.loss_and_grad(x, grad, losses)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gf = {64: 6.2, 128: 24.8, 256: 99.24, 512: 396.95, 1024: 1587.8}.get(size, 396.95 * (size / 512) ** 2)
    print(f"  closure (fwd+loss+bwd) {ms:.3f} ms/eval -> {1000 / ms:.1f} evals/s, {gf / ms:.1f} TFLOP/s algorithmic")
    # oracle fp32 on GPU timing (cuDNN/cuBLAS, TF32 off and on)
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        for _ in range(2):
            O.loss_and_grad(st32, x, t32, full=True)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(5):
            O.loss_and_grad(st32, x, t32, full=True)
        torch.cuda.synchronize()
        print(f"  torch closure on this GPU (cudnn tf32={tf32}): {(time.time() - t0) / 5 * 1e3:.3f} ms/eval")
    torch.backends.cudnn.allow_tf32 = False
    plan.close()


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [64, 256, 512]
    for s in sizes:
        run(s, "radar")
    run(sizes[-2] if len(sizes) > 1 else sizes[0], "smooth")
