"""`PerceptualLoss` (CycleGAN/models.py:397-476 re-built on the IST kernels) against the golden vectors of the
unmodified reference and against the oracle evaluated live: loss 1e-4 relative, d loss / d pred 1e-3 rel-L2 (north_star's
tolerance; on these smooth inputs the reference's own fp32-vs-fp64 gradient error is 2e-7, so no mask-flip allowance)."""
import os

import numpy as np
import pytest
import torch

from oracle import perceptual_oracle as PO
from gpu_common import rel_l2, strict_fp32

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make(tag):
    from ist_b200.model.perceptual import PerceptualLoss
    b, h, w, sl, cl, ws, wc = PO.CASES[tag]
    state = {k: torch.from_numpy(v) for k, v in PO.vgg16_state(0).items()}
    return PerceptualLoss(cl, sl, dev, ws, wc, state_dict=state)


@pytest.mark.parametrize("tag", sorted(PO.CASES))
def test_against_reference_golden(tag):
    strict_fp32()
    gold = np.load(os.path.join(GOLDEN, "perceptual.npz"))
    pl = make(tag)
    pred = torch.from_numpy(gold[f"{tag}_pred"]).to(dev).requires_grad_(True)
    content, style = torch.from_numpy(gold[f"{tag}_content"]).to(dev), torch.from_numpy(gold[f"{tag}_style"]).to(dev)
    loss = pl.calculate_loss(pred, content, style)
    loss.backward()
    ref_l, ref_g = float(gold[f"{tag}_loss_f64"]), torch.from_numpy(gold[f"{tag}_grad_f64"])
    err = rel_l2(pred.grad.cpu(), ref_g)
    print(f"{tag}: loss {float(loss):.6f} vs {ref_l:.6f}; grad rel-L2 {err:.2e}")
    assert abs(float(loss) - ref_l) <= 1e-4 * abs(ref_l)
    assert err <= 1e-3


def test_through_a_generator_like_graph():
    """pred = tanh(conv(z)) as in CycleGAN's generator head: the loss gradient reaches upstream parameters through autograd."""
    strict_fp32()
    tag = "a"
    b, h, w, sl, cl, ws, wc = PO.CASES[tag]
    _, content, style = [torch.from_numpy(a).to(dev) for a in PO.images(b, h, w)]
    torch.manual_seed(0)
    head = torch.nn.Conv2d(4, 3, 3, padding=1).to(dev)
    z = torch.randn(b, 4, h, w, device=dev)
    pl = make(tag)
    loss = pl.calculate_loss(torch.tanh(head(z)), content, style)
    loss.backward()
    ours = head.weight.grad.detach().clone()
    # oracle: same graph in fp64
    head64 = torch.nn.Conv2d(4, 3, 3, padding=1).to(dev).double()
    head64.load_state_dict({k: v.double() for k, v in head.state_dict().items()})
    st64 = {k: torch.from_numpy(v).to(dev).double() for k, v in PO.vgg16_state(0).items()}
    l64 = PO.calculate_loss(st64, torch.tanh(head64(z.double())), content.double(), style.double(), cl, sl, ws, wc)
    l64.backward()
    assert abs(float(loss) - float(l64)) <= 1e-4 * abs(float(l64))
    assert rel_l2(ours, head64.weight.grad) <= 1e-3
