"""Frame pipeline (IST/main.py:184-238 with decode / upload / optimise / download / encode overlapped) and the
multi-device guard of the C ABI."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import synth
from gpu_common import build_model, frames, noise_like, prepare_plan, strict_fp32

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


@pytest.fixture(scope="module")
def model_cfg():
    strict_fp32()
    return build_model(dev)


def test_pipeline_equals_per_frame_calls_and_skips_existing(model_cfg, tmp_path):
    from ist_b200.model.engine import do_transfer_style
    from ist_b200.pipeline import FramePipeline
    cfg, model = model_cfg
    cfg = cfg.clone()
    cfg.DATA.IMG_SIZE = 64
    cfg.LOSS.MAX_ITER = 40
    cfg.OUTPUT.DIR = str(tmp_path / "single") + "/"
    src = tmp_path / "radar"
    os.makedirs(src)
    sizes = [(64, 64), (64, 64), (64, 64), (64, 80), (64, 64)]          # the fourth frame has another aspect ratio
    paths = []
    for i, (h, w) in enumerate(sizes):
        p = str(src / ("%03d.png" % i))
        Image.fromarray(synth.radar_frame(64, 1000 + i, h=h, w=w), "RGB").save(p)
        paths.append(p)
    style = Image.fromarray(synth.lidar_frame(64, 2), "RGB")
    singles = [np.asarray(do_transfer_style(cfg, model, Image.open(p).convert("RGB"), style, dev)) for p in paths]
    out_dir = str(tmp_path / "out")
    pipe = FramePipeline(cfg, model, dev, style, out_dir, frames_per_batch=2, prefetch=2, keep_results=True)
    done = pipe.run(paths)
    assert done == [0, 1, 2, 3, 4] and pipe.stats["written"] == 5 and pipe.stats["evals"] == 5 * 40
    for i, p in enumerate(paths):
        got = np.asarray(Image.open(pipe.out_path(p)))
        assert got.shape == singles[i].shape and np.array_equal(got, singles[i]), i      # batched + pipelined == one by one
        assert np.array_equal(pipe.results[i].cpu().numpy(), singles[i])
    # a restarted job recomputes only what is missing
    os.remove(pipe.out_path(paths[2]))
    done2 = pipe.run(paths, skip_existing=True)
    assert done2 == [2] and pipe.stats["skipped"] == 4
    assert np.array_equal(np.asarray(Image.open(pipe.out_path(paths[2]))), singles[2])
    pipe.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_plan_on_a_device_that_is_not_current(model_cfg):
    """MODEL.DEVICE = 'cuda:1' while cuda:0 is the current device (the reference supports this through torch): plans,
    per-op entry points and the optimiser bind to their tensor's device; results equal the cuda:0 run bit for bit."""
    from ist_b200.model.engine.utils import optimize
    cfg, model = model_cfg
    d1 = torch.device("cuda:1")
    assert torch.cuda.current_device() == 0
    cfg1, model1 = build_model(d1)
    content, style = frames(64, dev, "radar")
    x = content + noise_like(content)
    l0, g0 = prepare_plan(model, cfg, content, style).loss_and_grad(x)
    l1, g1 = prepare_plan(model1, cfg1, content.to(d1), style.to(d1)).loss_and_grad(x.to(d1))
    assert torch.cuda.current_device() == 0
    assert torch.equal(l0.cpu(), l1.cpu()) and torch.equal(g0.cpu(), g1.cpu())
    xa = content.clone().requires_grad_(True)
    optimize(model, content, style, xa, cfg, 40)
    xb = content.to(d1).clone().requires_grad_(True)
    optimize(model1, content.to(d1), style.to(d1), xb, cfg1, 40)
    assert torch.equal(xa.detach().cpu(), xb.detach().cpu())
    feats = model1.vgg_model(content.to(d1), ["relu2_1"])
    assert feats[0].device == d1
