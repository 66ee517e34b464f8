"""Pins oracle/perceptual_oracle.py (a restatement of CycleGAN/models.py:397-476) against tests/golden/perceptual.npz, the
outputs of the UNMODIFIED reference class on seeded inputs (oracle/make_perceptual_golden.py), and checks the host-side
helpers of the product module that need no GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import perceptual_oracle as PO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "perceptual.npz"))


@pytest.mark.parametrize("tag", sorted(PO.CASES))
@pytest.mark.parametrize("sfx,dtype,tol", [("f32", torch.float32, 2e-5), ("f64", torch.float64, 1e-10)])
def test_oracle_against_reference(gold, tag, sfx, dtype, tol):
    b, h, w, sl, cl, ws, wc = PO.CASES[tag]
    pred, content, style = PO.images(b, h, w)
    for name, arr in (("pred", pred), ("content", content), ("style", style)):
        assert np.array_equal(arr, gold[f"{tag}_{name}"])              # the seeded inputs regenerate bit for bit
    state = {k: torch.from_numpy(v).to(dtype) for k, v in PO.vgg16_state(0).items()}
    loss, grad = PO.loss_and_grad(state, torch.from_numpy(pred).to(dtype), torch.from_numpy(content).to(dtype),
                                  torch.from_numpy(style).to(dtype), cl, sl, ws, wc)
    ref_l, ref_g = float(gold[f"{tag}_loss_{sfx}"]), gold[f"{tag}_grad_{sfx}"]
    assert abs(float(loss) - ref_l) <= tol * abs(ref_l)
    assert np.linalg.norm(grad.numpy() - ref_g) <= tol * np.linalg.norm(ref_g)


def test_vgg16_table_and_state_dict_keys():
    from ist_b200.model.perceptual import convert_state_dict, vgg16_cfg
    v = vgg16_cfg().MODEL.VGG
    assert len(v.FORWARD_SEQ) == len(v.OUT_SEQ) == 18 and list(v.CONV_LAYERS_DICT[0]) == PO.vgg16_conv_names()
    assert v.OUT_SEQ[:3] == ['relu1_1', 'relu1_2', 'pool_1'] and v.OUT_SEQ[-1] == 'pool_5'
    # torchvision vgg16.features indices of the 13 convs
    tv = {f"features.{i}.weight": i for i in (0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28)}
    conv = convert_state_dict(tv)
    assert list(conv) == [n + ".weight" for n in PO.vgg16_conv_names()]
    assert convert_state_dict({"0.bias": 1, "classifier.0.weight": 2}) == {"conv1_1.bias": 1}
    with pytest.raises(KeyError):
        convert_state_dict({"features.1.weight": 0})                       # a ReLU has no parameters


def test_no_cpu_path():
    from ist_b200 import IstError
    from ist_b200.model.perceptual import PerceptualLoss
    with pytest.raises(IstError):
        PerceptualLoss(['3,3'], ['1,2'], 'cpu', [1.0], [1.0], state_dict={})
