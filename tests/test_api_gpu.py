"""The drop-in plugin API on the GPU: VGG.forward(input, out_keys) under autograd, GramMatrix / GramMSELoss modules, the
non-fused optimisation loop, and the frame-level functions (do_transfer_style / do_hr_transfer_style) with PIL images."""
import numpy as np
import pytest
import torch
import torch.nn as nn
from PIL import Image

from ist_b200 import _lib
from ist_b200.model.engine import do_hr_transfer_style, do_transfer_style
from ist_b200.model.engine.utils import _optimize_autograd
from ist_b200.model.meta_arch import GramMatrix, GramMSELoss
from oracle import ist_oracle as O
from oracle import synth
from gpu_common import build_model, frames, noise_like, rel_l2, strict_fp32

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


@pytest.fixture(scope="module")
def model_cfg():
    strict_fp32()
    return build_model(dev)


def test_vgg_forward_keys_order_and_errors(model_cfg):
    cfg, model = model_cfg
    vgg = model.vgg_model
    content, _ = frames(64, dev, "smooth")
    with torch.no_grad():
        outs = vgg(content, ["relu4_2", "relu1_1", "pool_5"])
    assert [tuple(o.shape) for o in outs] == [(1, 512, 8, 8), (1, 64, 64, 64), (1, 512, 2, 2)]
    with pytest.raises(KeyError):
        vgg(content, ["relu9_9"])
    with pytest.raises(_lib.IstError):
        vgg(content.cpu(), ["relu1_1"])
    state = O.state_to_torch(synth.vgg_state_dict(0), torch.float64, dev)
    ref = O.vgg_forward(state, content.double(), ["relu4_2", "relu1_1", "pool_5"])
    for a, b in zip(outs, ref):
        assert rel_l2(a, b) < 5e-6


def test_autograd_path_matches_oracle(model_cfg):
    """The reference's closure written with the public modules (utils.py:31-36) and differentiated by autograd."""
    cfg, model = model_cfg
    content, style = frames(64, dev, "smooth")
    x = (content + noise_like(content)).requires_grad_(True)
    vgg = model.vgg_model
    with torch.no_grad():
        st = [GramMatrix()(A) for A in vgg(style, cfg.LOSS.STYLE_LAYERS)]
        ct = [A for A in vgg(content, cfg.LOSS.CONTENT_LAYERS)]
    targets = st + ct
    outs = vgg(x, model.loss_layers)
    ll = [model.loss_weights[a] * model.loss_functions[a](A, targets[a]) for a, A in enumerate(outs)]
    loss = sum(ll)
    loss.backward()
    state = O.state_to_torch(synth.vgg_state_dict(0, upto="conv5_1"), torch.float64, dev)
    t64 = O.compute_targets(state, content.double(), style.double(), full=False)
    l64, tot64, g64 = O.loss_and_grad(state, x.detach().double(), t64, full=False)
    ours = np.array([float(v) for v in ll] + [float(loss)])
    ref = np.array(l64 + [tot64])
    assert np.all(np.abs(ours - ref) / np.abs(ref) < 1e-4)
    assert rel_l2(x.grad, g64) < 2e-3
    # pool outputs and non-Gatys losses route through the generic backward as well
    x2 = (content + noise_like(content, seed=5)).requires_grad_(True)
    f = vgg(x2, ["pool_2", "relu3_3"])
    (f[0].pow(2).mean() + f[1].abs().mean()).backward()
    xr = x2.detach().double().requires_grad_(True)
    fr = O.vgg_forward(state, xr, ["pool_2", "relu3_3"], full=False)
    (fr[0].pow(2).mean() + fr[1].abs().mean()).backward()
    assert rel_l2(x2.grad, xr.grad) < 2e-3


def test_gram_modules(golden):
    m = golden["modules"]
    feat = torch.from_numpy(np.ascontiguousarray(np.tile(m["feat"], (1, 8, 1, 1)))).to(dev)      # 8 -> 64 channels
    G = GramMatrix()(feat)
    ref = O.gram_matrix(feat.double())
    assert rel_l2(G, ref) < 6e-7
    np.testing.assert_allclose(G[:, :8, :8].cpu().numpy(), m["gram"], rtol=1e-5, atol=1e-5)     # the reference's own output
    x = feat[:1].clone().requires_grad_(True)
    tgt = (ref[:1] * 0.5).float()
    l = GramMSELoss()(x, tgt)
    l.backward()
    xr = feat[:1].double().clone().requires_grad_(True)
    lr = O.gram_mse_loss(xr, tgt.double())
    lr.backward()
    assert abs(float(l) - float(lr)) < 1e-5 * abs(float(lr)) and rel_l2(x.grad, xr.grad) < 2e-5


def test_unfused_loop_runs(model_cfg):
    cfg, model = model_cfg
    content, style = frames(64, dev, "smooth")
    x = content.clone().requires_grad_(True)
    _optimize_autograd(model, content, style, x, cfg, 20)
    assert torch.isfinite(x).all() and not torch.equal(x.detach(), content)


def test_frame_functions_with_pil(model_cfg, tmp_path):
    cfg, model = model_cfg
    cfg = cfg.clone()
    cfg.DATA.IMG_SIZE = 64
    cfg.HRDATA.IMG_SIZE = 96
    cfg.LOSS.MAX_ITER = 20
    cfg.HRLOSS.MAX_ITER = 20
    cfg.OUTPUT.DIR = str(tmp_path) + "/"
    content = Image.fromarray(synth.radar_frame(80, 1))
    style = Image.fromarray(synth.lidar_frame(80, 2))
    out = do_transfer_style(cfg, model, content, style, dev)
    assert out.size == (64, 64) and (tmp_path / cfg.OUTPUT.FILE_NAME).exists()
    hr = do_hr_transfer_style(cfg, model, content, style, out, dev)
    assert hr.size == (96, 96) and (tmp_path / cfg.OUTPUT.HR_FILE_NAME).exists()
    assert np.asarray(hr).std() > 0


def test_device_handoff_equals_pil_handoff(model_cfg, tmp_path):
    """Coarse-to-fine with the previous stage's tensor kept on the device (8-bit clamp + bilinear resize + re-preprocessing
    as kernels) gives the same images as the reference's route through a PIL image (hr_transfer_style.py:21-27)."""
    cfg, model = model_cfg
    cfg = cfg.clone()
    cfg.DATA.IMG_SIZE = 48
    cfg.HRDATA.IMG_SIZE = 96
    cfg.LOSS.MAX_ITER = 20
    cfg.HRLOSS.MAX_ITER = 20
    cfg.OUTPUT.DIR = str(tmp_path) + "/"
    content = Image.fromarray(synth.radar_frame(64, 3))
    style = Image.fromarray(synth.lidar_frame(64, 2))
    out, x = do_transfer_style(cfg, model, content, style, dev, return_tensor=True)
    assert tuple(x.shape) == (1, 3, 48, 48) and x.is_cuda
    hr_pil = do_hr_transfer_style(cfg, model, content, style, out, dev)
    hr_dev, x_hr = do_hr_transfer_style(cfg, model, content, style, x, dev, return_tensor=True)
    assert tuple(x_hr.shape) == (1, 3, 96, 96)
    assert np.array_equal(np.asarray(hr_pil), np.asarray(hr_dev))


def test_batched_frames_equal_per_frame_calls(model_cfg, tmp_path):
    """do_transfer_style_batch == one do_transfer_style per frame (independent problems; at this size every layer's work split
    is the same for 1 and 3 frames, so the kernels round identically — see DESIGN.md 3 for large images)."""
    from ist_b200.model.engine import do_transfer_style_batch
    cfg, model = model_cfg
    cfg = cfg.clone()
    cfg.DATA.IMG_SIZE = 64
    cfg.LOSS.MAX_ITER = 20
    cfg.OUTPUT.DIR = str(tmp_path) + "/"
    style = Image.fromarray(synth.lidar_frame(80, 2))
    contents = [Image.fromarray(synth.radar_frame(80, s)) for s in (1, 5, 9)]
    single = [np.asarray(do_transfer_style(cfg, model, c, style, dev)) for c in contents]
    batch = [np.asarray(o) for o in do_transfer_style_batch(cfg, model, contents, style, dev)]
    assert len(batch) == 3
    for a, b in zip(single, batch):
        assert np.array_equal(a, b)
