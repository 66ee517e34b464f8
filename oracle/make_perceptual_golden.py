"""Generate tests/golden/perceptual.npz by running the UNMODIFIED reference class `PerceptualLoss`
(/root/reference/CycleGAN/models.py:397-476, imported read-only).

TEST INFRASTRUCTURE. Run in the build container only (the reference tree does not exist on the GPU box):
    python oracle/make_perceptual_golden.py
The reference constructs its network with `torchvision.models.vgg16(pretrained=True)` (models.py:399), a download that is
impossible offline: `torchvision.models.vgg16` is replaced for the duration of the constructor by a function returning
the same architecture with the seeded synthetic weights of oracle/synth.py. Nothing else of the reference is touched and
nothing of it is copied: only its outputs on seeded inputs are stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402
from oracle import perceptual_oracle as PO  # noqa: E402

REF = "/root/reference/CycleGAN"
GOLD = os.path.join(ROOT, "tests", "golden")

def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    import torchvision.models as tvm
    import models as ref_models            # the reference module

    state = PO.vgg16_state(0)
    names = PO.vgg16_conv_names()

    def fake_vgg16(pretrained=True, **kw):
        net = real_vgg16(weights=None)
        convs = [m for m in net.features if isinstance(m, torch.nn.Conv2d)]
        assert len(convs) == len(names)
        with torch.no_grad():
            for m, n in zip(convs, names):
                m.weight.copy_(torch.from_numpy(state[n + '.weight']))
                m.bias.copy_(torch.from_numpy(state[n + '.bias']))
        return net

    real_vgg16 = tvm.vgg16
    out = {}
    for tag, (b, h, w, sl, cl, ws, wc) in PO.CASES.items():
        pred, content, style = PO.images(b, h, w)
        tvm.vgg16 = fake_vgg16
        try:
            pl = ref_models.PerceptualLoss(cl, sl, 'cpu', ws, wc)
        finally:
            tvm.vgg16 = real_vgg16
        for dt, sfx in ((torch.float32, 'f32'), (torch.float64, 'f64')):
            pl.net = [l.to(dt) for l in pl.net]
            p = torch.from_numpy(pred).to(dt).requires_grad_(True)
            loss = pl.calculate_loss(p, torch.from_numpy(content).to(dt), torch.from_numpy(style).to(dt))
            (g,) = torch.autograd.grad(loss, p)
            out[f"{tag}_loss_{sfx}"] = np.array(float(loss))
            out[f"{tag}_grad_{sfx}"] = g.numpy()
        out[f"{tag}_pred"], out[f"{tag}_content"], out[f"{tag}_style"] = pred, content, style
        print(tag, "loss f32", out[f"{tag}_loss_f32"], "f64", out[f"{tag}_loss_f64"],
              "grad rel fp32-vs-fp64", np.linalg.norm(out[f"{tag}_grad_f32"] - out[f"{tag}_grad_f64"]) / np.linalg.norm(out[f"{tag}_grad_f64"]))
    np.savez_compressed(os.path.join(GOLD, "perceptual.npz"), **out)
    print("wrote", os.path.join(GOLD, "perceptual.npz"))


if __name__ == "__main__":
    main()
