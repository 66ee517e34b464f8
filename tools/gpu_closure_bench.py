"""Closure-only timing loop for profiling (run under gpurun / ncu): python tools/gpu_closure_bench.py SIZE REPS [BATCH]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ist_b200  # noqa: E402
from oracle import synth  # noqa: E402  (synthetic weights/frames only; no oracle compute)

from tools.gpu_plan_check import LAYERS  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
state = {k: torch.from_numpy(v).to(dev) for k, v in synth.vgg_state_dict(0, upto="conv5_1").items()}
content = torch.from_numpy(synth.preprocess(synth.radar_frame(size, 1))).to(dev).repeat(nb, 1, 1, 1).contiguous()
style = torch.from_numpy(synth.preprocess(synth.lidar_frame(size, 2))).to(dev).repeat(nb, 1, 1, 1).contiguous()
plan = ist_b200.Plan(LAYERS, nb, size, size)
plan.load_state_dict(state)
SL = ['relu1_1', 'relu2_1', 'relu3_1', 'relu4_1', 'relu5_1']
plan.set_loss(SL, [1e3 / n ** 2 for n in [64, 128, 256, 512, 512]], ['relu4_2'], [0.5])
plan.forward(style, "relu5_1")
for k, key in enumerate(SL):
    plan.set_style_target(k, plan.gram(key)[0])
plan.forward(content, "relu4_2")
plan.capture_content_target(0)
x = content + 20 * torch.randn_like(content)
grad = torch.empty_like(x)
losses = torch.empty(nb, 7, device=dev)
for _ in range(3):
    plan.loss_and_grad(x, grad, losses)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    plan.loss_and_grad(x, grad, losses)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"size {size} batch {nb}: {ms:.3f} ms/closure, {nb * 1000 / ms:.1f} evals/s, losses {losses[0].tolist()}")
