"""BASELINE configs[2]: the coarse-to-fine schedule 512 -> 1024 -> 2048 on one B200 through the frame-level API
(do_transfer_style, then do_hr_transfer_style per extra size with the device-resident hand-off), timed per stage.

    python tools/gpu_schedule.py [sizes...]        default 512 1024 2048; LOSS.MAX_ITER 300, HRLOSS.MAX_ITER 500
"""
import os
import sys
import tempfile
import time

import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ist_b200  # noqa: E402,F401
from ist_b200.config import get_cfg_defaults  # noqa: E402
from ist_b200.main import get_model  # noqa: E402
from ist_b200.model.engine import do_hr_transfer_style, do_transfer_style  # noqa: E402
from oracle import synth  # noqa: E402  (synthetic weights/frames only)

sizes = [int(a) for a in sys.argv[1:]] or [512, 1024, 2048]
dev = torch.device("cuda:0")
cfg = get_cfg_defaults()
cfg.MODEL.DEVICE = "cuda:0"
cfg.OUTPUT.DIR = tempfile.mkdtemp() + "/"
model, _ = get_model(cfg, {k: torch.from_numpy(v) for k, v in synth.vgg_state_dict(0).items()})
content = Image.fromarray(synth.radar_frame(512, 1))          # real frames are 512x512 8-bit BEV images
style = Image.fromarray(synth.lidar_frame(512, 2))


def stage(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


for rep in range(2):                                           # first pass builds plans / graphs, second is the measurement
    total_t, total_e = 0.0, 0
    cfg.DATA.IMG_SIZE = sizes[0]
    (img, x), dt = stage(lambda: do_transfer_style(cfg, model, content, style, dev, return_tensor=True))
    rows = [(sizes[0], model.last_evals, dt, float(model.last_losses[0, -1]))]
    for s in sizes[1:]:
        cfg.HRDATA.IMG_SIZE = s
        (img, x), dt = stage(lambda: do_hr_transfer_style(cfg, model, content, style, x, dev, return_tensor=True))
        rows.append((s, model.last_evals, dt, float(model.last_losses[0, -1])))
    print(f"pass {rep} ({'cold: plan + graph construction included' if rep == 0 else 'warm'}):")
    for s, e, dt, loss in rows:
        print(f"  stage {s:5d}^2: {e:4d} closure evals in {dt * 1e3:9.1f} ms -> {e / dt:7.1f} evals/s ({dt / e * 1e3:6.2f} ms/eval incl. "
              f"pre/post-processing, PNG encode), final loss {loss:.4e}")
        total_t += dt
        total_e += e
    print(f"  schedule total: {total_t:.2f} s for {total_e} evals; output {img.size}; "
          f"peak torch memory {torch.cuda.max_memory_allocated() / 2 ** 30:.2f} GiB, "
          f"device memory in use {(torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 2 ** 30:.1f} GiB")
