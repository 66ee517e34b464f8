"""Per-launch CUDA-event timing of one closure (run under gpurun): python tools/gpu_layer_times.py SIZE [BATCH]
Prints every kernel launch of ist_plan_loss_and_grad in order with its algorithmic FLOPs / bytes, time and rate."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ist_b200  # noqa: E402
from oracle import synth  # noqa: E402  (synthetic weights/frames only)
from tools.gpu_plan_check import LAYERS  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
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
lib = ist_b200.load()
for _ in range(3):
    plan.loss_and_grad(x)
torch.cuda.synchronize()
REPS = 5
maxr = 4096
acc = None
for rep in range(REPS):
    lib.ist_profile_begin()
    plan.loss_and_grad(x)
    names = ctypes.create_string_buffer(maxr * 40)
    flops = (ctypes.c_double * maxr)()
    nbytes = (ctypes.c_double * maxr)()
    ms = (ctypes.c_float * maxr)()
    n = ctypes.c_int(0)
    ist_b200._lib.check(lib.ist_profile_end(maxr, names, flops, nbytes, ms, ctypes.byref(n)))
    rows = [(names.raw[i * 40:(i + 1) * 40].split(b"\0")[0].decode(), flops[i], nbytes[i], ms[i]) for i in range(n.value)]
    if acc is None:
        acc = [[r[0], r[1], r[2], [r[3]]] for r in rows]
    else:
        for a, r in zip(acc, rows):
            a[3].append(r[3])
tot = 0.0
print(f"{'#':>3} {'kernel':24s} {'GFLOP':>8s} {'MB':>8s} {'us(min)':>8s} {'us(med)':>8s} {'TF/s':>7s} {'GB/s':>7s}")
for i, (nm, fl, by, t) in enumerate(acc):
    t = sorted(t)
    tmin, tmed = t[0] * 1e3, t[len(t) // 2] * 1e3
    tot += tmed
    print(f"{i:3d} {nm:24s} {fl / 1e9:8.3f} {by / 1e6:8.2f} {tmin:8.1f} {tmed:8.1f} {fl / (tmed * 1e-6) / 1e12 if fl else 0:7.1f} {by / (tmed * 1e-6) / 1e9:7.0f}")
print(f"sum of medians: {tot / 1e3:.3f} ms (eager, includes launch gaps only via event placement)")
