"""Why does the reference disagree with itself after 20 evaluations?  (test infrastructure: executes oracle/ on the CPU)

Runs the oracle's optimize() (torch.optim.LBFGS defaults, IST/model/engine/utils.py:17-45) in fp32 and fp64 on the same
frame and prints, per closure evaluation, the PSNR between the two iterates, the first curvature pair's y.s in both
precisions and the style image class. SURVEY 7.3 H2 measured 39 dB (fp32 vs fp64, smooth class, 20 evals) with a *smooth
style image*; the GPU tests use the lidar-like style (sparse points), which makes the gradient at x0 much rougher.

    python tools/cpu_trajectory_study.py [SIZE] [EVALS]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ist_oracle as O  # noqa: E402
from oracle import synth  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
evals = int(sys.argv[2]) if len(sys.argv) > 2 else 20
torch.set_num_threads(8)
state_np = synth.vgg_state_dict(0, upto="conv5_1")


def run(dtype, content, style):
    st = O.state_to_torch(state_np, dtype)
    c, s = content.to(dtype), style.to(dtype)
    x = c.clone().requires_grad_(True)
    targets = O.compute_targets(st, c, s, full=False)
    opt = torch.optim.LBFGS([x])
    xs, n = [], [0]
    while n[0] < evals:
        def closure():
            opt.zero_grad()
            loss = sum(O.layer_losses(st, x, targets, full=False))
            loss.backward()
            n[0] += 1
            xs.append(x.detach().clone().double().numpy()[0])
            return loss
        opt.step(closure)
    xs.append(x.detach().clone().double().numpy()[0])
    stt = opt.state[opt._params[0]]
    return xs, stt


for ckind, skind in (("smooth", "lidar"), ("smooth", "smooth"), ("radar", "lidar")):
    mk = {"smooth": synth.smooth_frame, "radar": synth.radar_frame, "lidar": synth.lidar_frame}
    content = torch.from_numpy(synth.preprocess(mk[ckind](size, 1)))
    style = torch.from_numpy(synth.preprocess(mk[skind](size, 2)))
    x32, s32 = run(torch.float32, content, style)
    x64, s64 = run(torch.float64, content, style)
    ps = [synth.psnr(a, b) for a, b in zip(x32, x64)]
    print(f"== {size}^2 content {ckind}, style {skind}: PSNR(fp32 iterate, fp64 iterate) at the k-th evaluation point")
    print("   " + " ".join(f"{k}:{p:.1f}" for k, p in enumerate(ps) if np.isfinite(p)))
    print(f"   history pairs kept after {evals} evals: fp32 {len(s32['old_dirs'])}, fp64 {len(s64['old_dirs'])}; "
          f"first ro fp32 {float(s32['ro'][0]):.3e} fp64 {float(s64['ro'][0]):.3e}")
