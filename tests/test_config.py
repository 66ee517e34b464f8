"""The config tree is the API of the path: same keys and values as the reference's IST/config/defaults.py (captured from the
reference itself by oracle/make_golden.py into tests/golden/cfg_defaults.json)."""
import json
import os

import pytest

from ist_b200.config import CfgNode, get_cfg_defaults

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def plain(n):
    if isinstance(n, dict):
        return {k: plain(v) for k, v in n.items()}
    if isinstance(n, (list, tuple)):
        return [plain(v) for v in n]
    return n


def test_defaults_equal_reference():
    ref = json.load(open(os.path.join(GOLDEN, "cfg_defaults.json")))
    assert plain(get_cfg_defaults()) == ref


def test_loss_weights_values():
    cfg = get_cfg_defaults()
    assert cfg.LOSS.STYLE_WEIGHTS == [0.244140625, 0.06103515625, 0.0152587890625, 0.003814697265625, 0.003814697265625]
    assert cfg.LOSS.CONTENT_WEIGHTS == [0.5] and cfg.LOSS.MAX_ITER == 300 and cfg.HRLOSS.MAX_ITER == 500
    assert len(cfg.MODEL.VGG.FORWARD_SEQ) == len(cfg.MODEL.VGG.OUT_SEQ) == 21


def test_clone_freeze_merge(tmp_path):
    cfg = get_cfg_defaults()
    c2 = cfg.clone()
    c2.DATA.IMG_SIZE = 256
    assert cfg.DATA.IMG_SIZE == 512
    c2.merge_from_list(["LOSS.MAX_ITER", "60", "MODEL.DEVICE", "cuda:1"])
    assert c2.LOSS.MAX_ITER == 60 and c2.MODEL.DEVICE == "cuda:1"
    y = tmp_path / "c.yaml"
    y.write_text("HRDATA:\n  IMG_SIZE: 1024\n")
    c2.merge_from_file(str(y))
    assert c2.HRDATA.IMG_SIZE == 1024
    with pytest.raises(KeyError):
        c2.merge_from_list(["LOSS.NOPE", "1"])
    c2.freeze()
    with pytest.raises(AttributeError):
        c2.DATA.IMG_SIZE = 1
    assert isinstance(c2.MODEL, CfgNode)
