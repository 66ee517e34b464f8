"""optimize() / L-BFGS on the GPU.

The no-line-search L-BFGS trajectory is chaotic (two fp32 runs of the REFERENCE that differ only in CPU
thread count are 37 dB apart after 20 evaluations, fp32 vs fp64 is 22 dB on radar-like frames), so the pixel-level PSNR gate
is stated relative to the reference's own fp32-vs-fp64 PSNR measured in the same test, and the robust checks are
(a) the optimiser arithmetic against torch.optim.LBFGS driven by the same CUDA closure over the first evaluations,
(b) the loss level reached, (c) evaluation counting, (d) bitwise run-to-run determinism."""
import pytest
import torch

from ist_b200.lbfgs import DeviceLBFGS
from ist_b200.model.engine.utils import optimize
from oracle import ist_oracle as O
from oracle import synth
from gpu_common import build_model, frames, prepare_plan, strict_fp32

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


@pytest.fixture(scope="module")
def model_cfg():
    strict_fp32()
    return build_model(dev)


def psnr(a, b):
    return synth.psnr(a.detach().cpu().numpy()[0], b.detach().cpu().numpy()[0])


def test_lbfgs_matches_torch_on_same_closure(model_cfg):
    """Device L-BFGS vs torch.optim.LBFGS, both driven by the same CUDA closure (max_iter=1 so that every optimizer.step
    exposes the loss of one evaluation): identical for the first evaluations, then separating only at rounding level."""
    cfg, model = model_cfg
    content, style = frames(128, dev, "smooth")
    plan = prepare_plan(model, cfg, content, style)
    xa = content.clone()
    opt = DeviceLBFGS(plan, max_iter=1)
    tr_a = []
    for _ in range(8):
        ev, l0 = opt.step(xa)
        assert ev == 1
        tr_a.append(l0)
    opt.close()
    xb = content.clone().requires_grad_(True)
    topt = torch.optim.LBFGS([xb], max_iter=1)
    losses = torch.empty(1, 7, device=dev)

    def closure():
        g = torch.empty_like(xb)
        plan.loss_and_grad(xb.data, g, losses)
        xb.grad = g
        return losses[0, 6].clone()
    tr_b = [float(topt.step(closure)) for _ in range(8)]
    rel = [abs(a - b) / abs(b) for a, b in zip(tr_a, tr_b)]
    print("loss trace rel diff:", ["%.1e" % r for r in rel])
    assert max(rel[:4]) < 1e-6 and max(rel[:6]) < 1e-4
    assert tr_a[-1] < 0.5 * tr_a[0]


def test_eval_counting_matches_reference_loop(model_cfg):
    """max_iterations counts closure evaluations and is checked between optimizer steps (utils.py:28,37,43):
    20 -> 20 evaluations, 50 -> 60, 45 -> 60."""
    cfg, model = model_cfg
    content, style = frames(64, dev, "smooth")
    for max_it, expect in ((20, 20), (50, 60), (45, 60)):
        x = content.clone().requires_grad_(True)
        out = optimize(model, content, style, x, cfg, max_it)
        assert out is x and model.last_evals == expect
        assert torch.isfinite(x).all()


@pytest.mark.parametrize("kind,style_kind", [("smooth", "lidar"), ("radar", "lidar"), ("smooth", "smooth")])
def test_optimize_against_oracle(model_cfg, kind, style_kind):
    """Free-running 20 evaluations. The pixel-level agreement any two runs of this loop can reach is set by the input class
    (tools/cpu_trajectory_study.py, profiles/r02_cpu_trajectory_study_128.log: the reference's own fp32-vs-fp64 PSNR after 20
    evaluations is 36.5 dB for a smooth content + smooth style pair — SURVEY 7.3 H2's 39 dB case —, 17.7 dB for smooth content
    with the sparse lidar-like style and 22.5 dB for radar + lidar, on the CPU; cuDNN's fp32 path on the GPU sits a few dB lower):
    the sparse style makes the gradient at x0 rough and the first curvature pairs noise. The gate is therefore relative to the
    reference's own self-PSNR measured in the same test; the teacher-forced tests pin the arithmetic itself."""
    cfg, model = model_cfg
    size = 128
    content, style = frames(size, dev, kind, style_kind=style_kind)
    state_np = synth.vgg_state_dict(0, upto="conv5_1")
    st32, st64 = O.state_to_torch(state_np, torch.float32, dev), O.state_to_torch(state_np, torch.float64, dev)
    x = content.clone().requires_grad_(True)
    optimize(model, content, style, x, cfg, 20)
    x32 = content.clone().requires_grad_(True)
    O.optimize(st32, content, style, x32, 20, full=False)
    x64 = content.double().clone().requires_grad_(True)
    O.optimize(st64, content.double(), style.double(), x64, 20, full=False)
    t64 = O.compute_targets(st64, content.double(), style.double(), full=False)
    l_ours = O.loss_and_grad(st64, x.detach().double(), t64, full=False)[1]
    l_32 = O.loss_and_grad(st64, x32.detach().double(), t64, full=False)[1]
    l_64 = O.loss_and_grad(st64, x64.detach(), t64, full=False)[1]
    l_0 = O.loss_and_grad(st64, content.double(), t64, full=False)[1]
    p_ours, p_floor = psnr(x, x64), psnr(x32, x64)
    print(f"{kind} content / {style_kind} style: PSNR ours-vs-fp64 {p_ours:.1f} dB, reference fp32-vs-fp64 {p_floor:.1f} dB; "
          f"loss after 20 evals ours {l_ours:.4e} fp32 {l_32:.4e} fp64 {l_64:.4e} (start {l_0:.4e})")
    # same loss level as the reference's own runs, far below the starting loss
    assert l_ours < 0.2 * l_0
    assert l_ours < 2.5 * max(l_32, l_64)
    # pixel agreement at least as good as what the reference reproduces of itself (minus 3 dB slack)
    assert p_ours > p_floor - 3.0


def test_optimize_is_bitwise_deterministic(model_cfg):
    cfg, model = model_cfg
    content, style = frames(64, dev, "radar")
    outs = []
    for _ in range(2):
        x = content.clone().requires_grad_(True)
        optimize(model, content, style, x, cfg, 40)
        outs.append(x.detach().clone())
    assert torch.equal(outs[0], outs[1])


def test_batched_frames_equal_single_frames(model_cfg):
    """Frames of a batch keep independent optimiser state: each equals its single-frame run bit for bit."""
    cfg, model = model_cfg
    c1, style = frames(64, dev, "radar", cseed=1)
    c2, _ = frames(64, dev, "smooth", cseed=7)
    singles = []
    for c in (c1, c2):
        x = c.clone().requires_grad_(True)
        optimize(model, c, style, x, cfg, 20)
        singles.append(x.detach().clone())
    both = torch.cat([c1, c2])
    xb = both.clone().requires_grad_(True)
    optimize(model, both, style, xb, cfg, 20)
    assert torch.equal(xb.detach()[0:1], singles[0]) and torch.equal(xb.detach()[1:2], singles[1])


def test_style_targets_are_cached(model_cfg):
    from ist_b200.model.engine.utils import style_targets
    cfg, model = model_cfg
    _, style = frames(64, dev, "radar")
    a = style_targets(model.vgg_model, style, cfg.LOSS.STYLE_LAYERS)
    b = style_targets(model.vgg_model, style, cfg.LOSS.STYLE_LAYERS)
    assert a is b
    style.add_(1.0)           # in-place change bumps the version counter -> recomputed
    c = style_targets(model.vgg_model, style, cfg.LOSS.STYLE_LAYERS)
    assert c is not a
