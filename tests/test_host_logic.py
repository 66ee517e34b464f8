"""Host-side logic that needs no GPU: layer table from the config, fused-closure detection, module contracts, sharding."""
import numpy as np
import pytest
import torch
import torch.nn as nn

from ist_b200 import _lib
from ist_b200.config import get_cfg_defaults
from ist_b200.model import build_model
from ist_b200.model.engine.utils import _fused_spec
from ist_b200.model.meta_arch import GramMSELoss, StyleTransfer
from ist_b200.parallel import frames_per_rank, shard_indices
from ist_b200.plan import layer_table
from oracle import synth


def test_layer_table():
    cfg = get_cfg_defaults()
    full = layer_table(cfg.MODEL.VGG)
    assert len(full) == 21 and full[0][:3] == (0, 3, 64) and full[2][0] == 1
    t = layer_table(cfg.MODEL.VGG, "relu5_1")
    assert len(t) == 17 and t[-1][3:] == ("conv5_1", "relu5_1") and t[-1][1:3] == (512, 512)
    assert [l[4] for l in t if l[0] == 1] == ["pool_1", "pool_2", "pool_3", "pool_4"]


def test_vgg_module_contract():
    cfg = get_cfg_defaults()
    vgg = build_model(cfg)
    sd = vgg.state_dict()
    assert len(sd) == 32 and sum(v.numel() for v in sd.values()) == 20024384
    assert tuple(sd["conv1_1.weight"].shape) == (64, 3, 3, 3) and tuple(sd["conv5_4.bias"].shape) == (512,)
    vgg.load_state_dict({k: torch.from_numpy(v) for k, v in synth.vgg_state_dict(0).items()})
    for p in vgg.parameters():
        p.requires_grad = False
    with pytest.raises(KeyError):
        build_model(cfg, pool="avg")                 # the reference only defines max pooling (vgg.py:20-22)


def test_fused_spec_detection():
    cfg = get_cfg_defaults()
    vgg = build_model(cfg)
    layers = cfg.LOSS.STYLE_LAYERS + cfg.LOSS.CONTENT_LAYERS
    ws = cfg.LOSS.STYLE_WEIGHTS + cfg.LOSS.CONTENT_WEIGHTS
    m = StyleTransfer(vgg, layers, [GramMSELoss()] * 5 + [nn.MSELoss()], ws)
    spec = _fused_spec(m, cfg)
    assert spec is not None and spec[0] == cfg.LOSS.STYLE_LAYERS and spec[2] == ["relu4_2"] and spec[3] == [0.5]
    m2 = StyleTransfer(vgg, layers, [GramMSELoss()] * 5 + [nn.L1Loss()], ws)
    assert _fused_spec(m2, cfg) is None
    m3 = StyleTransfer(vgg, layers[:-1], [GramMSELoss()] * 5, ws[:-1])
    assert _fused_spec(m3, cfg) is None


def test_shard_indices_is_a_permutation():
    for n in (0, 1, 7, 256):
        for world in (1, 2, 3, 8):
            parts = [shard_indices(n, r, world) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert all(len(p) <= frames_per_rank(n, world) for p in parts) if n else True
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_synthetic_frames():
    r = synth.radar_frame(64, 1)
    assert r.shape == (64, 64, 3) and r.dtype == np.uint8 and set(np.unique(r)) <= {0, 255}
    assert (r[..., 0] == r[..., 1]).all()
    x = synth.preprocess(r)
    assert x.shape == (1, 3, 64, 64) and x.dtype == np.float32
    assert -130 < x.min() < -100 and 130 < x.max() < 160
    s = synth.smooth_frame(0, 1, h=48, w=80)
    assert s.shape == (48, 80, 3)
    assert synth.psnr(x[0], x[0]) == float("inf")


def test_ptr_rejects_cpu_tensors():
    with pytest.raises(_lib.IstError):
        _lib.ptr(torch.zeros(4))
