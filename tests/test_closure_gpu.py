"""Closure parity (IST/model/engine/utils.py:29-41): six weighted layer losses, their sum and the image gradient of the CUDA
path vs (a) golden vectors produced by the unmodified reference (tests/golden, fp64 = ground truth, fp32 = what the
reference itself reproduces) and (b) the oracle evaluated live on the same GPU.

Tolerances: losses 1e-4 relative (north_star asks 1e-3). Gradient: rel-L2 <= max(2e-3, 3 x the reference's own
fp32-vs-fp64 error on the same point) — the gradient error is quantised by ReLU / max-pool mask flips (SURVEY 7.3 H1:
one flipped unit in millions moves it by ~1e-3), so it is always reported and judged beside the reference's own floor."""
import numpy as np
import pytest
import torch

from oracle import ist_oracle as O
from oracle import synth
from gpu_common import build_model, frames, noise_like, prepare_plan, rel_l2, strict_fp32

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


@pytest.fixture(scope="module")
def model_cfg():
    strict_fp32()
    cfg, model = build_model(dev)
    return cfg, model


@pytest.mark.parametrize("tag", ["64", "48x80"])
def test_closure_against_reference_golden(model_cfg, golden, tag):
    cfg, model = model_cfg
    g = golden[tag]
    content, style = torch.from_numpy(g["content"]).to(dev), torch.from_numpy(g["style"]).to(dev)
    plan = prepare_plan(model, cfg, content, style)
    for pname, xk in (("p0", "content"), ("p1", "x1")):
        x = torch.from_numpy(g[xk]).to(dev)
        losses, grad = plan.loss_and_grad(x)
        ours = losses[0].double().cpu().numpy()
        ref64, ref32 = g[f"losses_{pname}_f64"], g[f"losses_{pname}_f32"]
        nz = np.abs(ref64) > 0
        assert np.all(np.abs(ours[nz] - ref64[nz]) / np.abs(ref64[nz]) < 1e-4), (ours, ref64)
        g64, g32 = torch.from_numpy(g[f"grad_{pname}_f64"]), torch.from_numpy(g[f"grad_{pname}_f32"])
        err = rel_l2(grad.cpu(), g64)
        floor = rel_l2(g32, g64)
        print(f"{tag} {pname}: grad rel-L2 ours {err:.2e}, reference fp32-vs-fp64 {floor:.2e}")
        assert err <= max(2e-3, 3 * floor), (err, floor)
    # at x0 = content the content loss and its seed are exactly zero (Appendix A of SURVEY)
    losses, _ = plan.loss_and_grad(content)
    assert float(losses[0, 5]) == 0.0


@pytest.mark.parametrize("size,kind", [(64, "radar"), (128, "smooth"), (256, "radar")])
def test_closure_against_live_oracle(model_cfg, size, kind):
    cfg, model = model_cfg
    content, style = frames(size, dev, kind)
    plan = prepare_plan(model, cfg, content, style)
    state_np = synth.vgg_state_dict(0, upto="conv5_1")
    st64, st32 = O.state_to_torch(state_np, torch.float64, dev), O.state_to_torch(state_np, torch.float32, dev)
    t64 = O.compute_targets(st64, content.double(), style.double(), full=False)
    t32 = O.compute_targets(st32, content, style, full=False)
    x = content + noise_like(content)
    losses, grad = plan.loss_and_grad(x)
    l64, tot64, g64 = O.loss_and_grad(st64, x.double(), t64, full=False)
    _, _, g32 = O.loss_and_grad(st32, x, t32, full=False)
    ours = losses[0].double().cpu().numpy()
    ref = np.array(l64 + [tot64])
    assert np.all(np.abs(ours - ref) / np.abs(ref) < 1e-4)
    err, floor = rel_l2(grad, g64), rel_l2(g32, g64)
    cos = 1.0 - float(torch.dot(grad.double().flatten(), g64.flatten()) / (grad.double().norm() * g64.norm()))
    print(f"{size} {kind}: grad rel-L2 ours {err:.2e} (1-cos {cos:.1e}), oracle fp32-vs-fp64 {floor:.2e}")
    assert err <= max(2e-3, 3 * floor) and cos < 1e-4


def test_feature_export_and_gram_targets(model_cfg, golden):
    """VGG.forward features and Gram targets vs the reference's golden corners / checksums (fp64)."""
    cfg, model = model_cfg
    g = golden["64"]
    content, style = torch.from_numpy(g["content"]).to(dev), torch.from_numpy(g["style"]).to(dev)
    keys = ["relu1_1", "pool_1", "relu3_1", "relu4_2", "pool_4", "relu5_1"]
    with torch.no_grad():
        feats = model.vgg_model(content, keys)
    for k, f in zip(keys, feats):
        ref = g[f"feat_{k}_corner_f64"]
        np.testing.assert_allclose(f[0, :4, :3, :3].cpu().numpy(), ref, rtol=2e-5, atol=2e-5 * np.abs(ref).max() + 1e-6)
        assert abs(float(f.double().sum()) - g[f"feat_{k}_sums_f64"][0]) <= 2e-5 * abs(g[f"feat_{k}_sums_f64"][0])
    from ist_b200.model.engine.utils import style_targets
    grams = style_targets(model.vgg_model, style, cfg.LOSS.STYLE_LAYERS)
    for k, G in enumerate(grams):
        ref = g[f"gram{k}_corner_f64"]
        np.testing.assert_allclose(G[0, :8, :8].cpu().numpy(), ref, rtol=2e-5, atol=2e-5 * np.abs(ref).max())
        assert abs(float(G.double().sum()) - g[f"gram{k}_sums_f64"][0]) <= 2e-5 * abs(g[f"gram{k}_sums_f64"][0])


@pytest.mark.parametrize("size", [64, 256])
def test_closure_is_bitwise_deterministic_and_batch_consistent(model_cfg, size):
    """256^2 is large enough for the conv kernels' stream-K split (several chunks per CTA, partial tiles exchanged between
    CTAs, three frames per launch): the per-frame partition must make a batch round exactly like single frames."""
    cfg, model = model_cfg
    content, style = frames(size, dev, "radar")
    c2, _ = frames(size, dev, "radar", cseed=5)
    plan1 = prepare_plan(model, cfg, content, style)
    x1 = content + noise_like(content)
    l_a, g_a = plan1.loss_and_grad(x1)
    l_a, g_a = l_a.clone(), g_a.clone()
    l_b, g_b = plan1.loss_and_grad(x1)
    assert torch.equal(l_a, l_b) and torch.equal(g_a, g_b)              # no float atomics anywhere
    plan_c2 = prepare_plan(model, cfg, c2, style)
    x2 = c2 + noise_like(c2, seed=4)
    l_c, g_c = plan_c2.loss_and_grad(x2)
    l_c, g_c = l_c.clone(), g_c.clone()
    # a batch of three independent frames gives each frame exactly its single-frame result
    both = torch.cat([content, c2, content])
    planb = prepare_plan(model, cfg, both, style)
    lb, gb = planb.loss_and_grad(torch.cat([x1, x2, x1]))
    assert torch.equal(lb[0], l_a[0]) and torch.equal(gb[0], g_a[0])
    assert torch.equal(lb[1], l_c[0]) and torch.equal(gb[1], g_c[0])
    assert torch.equal(lb[2], l_a[0]) and torch.equal(gb[2], g_a[0])


@pytest.mark.parametrize("h,w", [(70, 52), (37, 90)])
def test_closure_odd_sizes(model_cfg, h, w):
    """Image sizes that are no multiple of the 16 x 8 pixel tile (Scale keeps the aspect ratio, so e.g. 512 x 683 frames occur):
    partial and phantom tiles of the CTA-pair conv kernel and of the tensor-core first-conv kernels, odd pool edges."""
    cfg, model = model_cfg
    content, style = frames(max(h, w), dev, "smooth", h=h, w=w)
    plan = prepare_plan(model, cfg, content, style)
    state_np = synth.vgg_state_dict(0, upto="conv5_1")
    st64 = O.state_to_torch(state_np, torch.float64, dev)
    st32 = O.state_to_torch(state_np, torch.float32, dev)
    t64 = O.compute_targets(st64, content.double(), style.double(), full=False)
    t32 = O.compute_targets(st32, content, style, full=False)
    x = content + noise_like(content)
    losses, grad = plan.loss_and_grad(x)
    l64, tot64, g64 = O.loss_and_grad(st64, x.double(), t64, full=False)
    _, _, g32 = O.loss_and_grad(st32, x, t32, full=False)
    ours = losses[0].double().cpu().numpy()
    ref = np.array(l64 + [tot64])
    assert np.all(np.abs(ours - ref) / np.abs(ref) < 1e-4), (ours, ref)
    err, floor = rel_l2(grad, g64), rel_l2(g32, g64)
    print(f"{h}x{w}: grad rel-L2 ours {err:.2e}, oracle fp32-vs-fp64 {floor:.2e}")
    assert err <= max(2e-3, 3 * floor)
