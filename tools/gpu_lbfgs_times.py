"""Per-kernel timing of optimizer.step (eager, IST_B200_NO_GRAPH=1 must be set): python tools/gpu_lbfgs_times.py [SIZE] [STEPS_BEFORE]
Profiles one step() after STEPS_BEFORE warm-up steps (so the history holds min(100, 20*STEPS_BEFORE) pairs)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
assert os.environ.get("IST_B200_NO_GRAPH") == "1", "run with IST_B200_NO_GRAPH=1"
import ist_b200  # noqa: E402
from ist_b200.lbfgs import DeviceLBFGS  # noqa: E402
from oracle import synth  # noqa: E402
from tools.gpu_plan_check import LAYERS  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
before = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda:0")
state = {k: torch.from_numpy(v).to(dev) for k, v in synth.vgg_state_dict(0, upto="conv5_1").items()}
content = torch.from_numpy(synth.preprocess(synth.radar_frame(size, 1))).to(dev)
style = torch.from_numpy(synth.preprocess(synth.lidar_frame(size, 2))).to(dev)
plan = ist_b200.Plan(LAYERS, 1, size, size)
plan.load_state_dict(state)
SL = ['relu1_1', 'relu2_1', 'relu3_1', 'relu4_1', 'relu5_1']
plan.set_loss(SL, [1e3 / n ** 2 for n in [64, 128, 256, 512, 512]], ['relu4_2'], [0.5])
plan.forward(style, "relu5_1")
for k, key in enumerate(SL):
    plan.set_style_target(k, plan.gram(key)[0])
plan.forward(content, "relu4_2")
plan.capture_content_target(0)
x = content.clone()
opt = DeviceLBFGS(plan)
for _ in range(before):
    opt.step(x)
lib = ist_b200.load()
lib.ist_profile_begin()
opt.step(x)
maxr = 8192
names = ctypes.create_string_buffer(maxr * 40)
flops = (ctypes.c_double * maxr)(); nbytes = (ctypes.c_double * maxr)(); ms = (ctypes.c_float * maxr)(); n = ctypes.c_int(0)
ist_b200._lib.check(lib.ist_profile_end(maxr, names, flops, nbytes, ms, ctypes.byref(n)))
agg = {}
for i in range(n.value):
    nm = names.raw[i * 40:(i + 1) * 40].split(b"\0")[0].decode()
    a = agg.setdefault(nm, [0.0, 0.0, 0])
    a[0] += ms[i]; a[1] += nbytes[i]; a[2] += 1
tot = sum(a[0] for a in agg.values())
print(f"one optimizer.step (20 evals) after {before} steps, {size}x{size}: {tot:.2f} ms total = {tot / 20:.3f} ms/eval")
for nm, (t, b, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"  {nm:24s} {c:4d} launches  {t / 20 * 1e3:8.1f} us/eval  {b / max(t, 1e-9) / 1e6:8.0f} GB/s")
