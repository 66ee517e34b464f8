"""Reference PyTorch GPU path (cuDNN / cuBLAS / torch.optim.LBFGS through the oracle restatement) timed on the same B200, for
DESIGN.md: python tools/gpu_reference_gpu_path.py [SIZE] [EVALS]   (test/bench infrastructure: executes oracle/)"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ist_oracle as O  # noqa: E402
from oracle import synth  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
evals = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = torch.device("cuda:0")
state = O.state_to_torch(synth.vgg_state_dict(0), torch.float32, dev)
style = torch.from_numpy(synth.preprocess(synth.lidar_frame(size, 2))).to(dev)
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    for rep in range(2):
        content = torch.from_numpy(synth.preprocess(synth.radar_frame(size, 1000 + rep))).to(dev)
        x = content.clone().requires_grad_(True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, n = O.optimize(state, content, style, x, evals, full=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"reference PyTorch GPU path, {size}x{size}, cudnn.allow_tf32={tf32}: {n} closure evals in {dt:.3f} s -> {n / dt:.1f} evals/s "
          f"({1e3 * dt / n:.2f} ms/eval), torch {torch.__version__}")
    # closure only (no optimiser): forward + losses + backward
    targets = O.compute_targets(state, content, style, full=True)
    xx = (content + 20 * torch.randn_like(content))
    for _ in range(3):
        O.loss_and_grad(state, xx, targets, full=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        O.loss_and_grad(state, xx, targets, full=True)
    torch.cuda.synchronize()
    print(f"    closure only: {(time.perf_counter() - t0) / 20 * 1e3:.2f} ms/eval")
