"""Flip-aware closure parity at the sizes BASELINE.json names: 512^2 (configs[1]), 1024^2 and 2048^2 (the coarse-to-fine
schedule, configs[2]) — IST/model/engine/utils.py:29-41 through ist_plan_loss_and_grad vs the fp64 oracle on the same GPU.

Protocol = SURVEY 8d / 7.3 H1. Evaluation points: P0 = x0 (the content image; content loss exactly 0, exact pool ties on
the constant radar background), P1 = content + N(0, 20^2) (seed 3), P2 = the oracle's own iterate after one optimizer.step
(20 evaluations). Per point: six weighted layer losses (<= 1e-4 relative; north_star asks 1e-3), image gradient rel-L2 /
cosine / max-abs, the number of ReLU-sign + pool-argmax decisions of our forward that differ from the fp64 forward, and
the gradient rel-L2 against the fp64 oracle evaluated WITH OUR MASKS, which isolates kernel arithmetic from mask flips and
must be <= 1e-4. The oracle's own fp32-vs-fp64 row (cuDNN / cuBLAS fp32, TF32 off) is printed beside ours: the plain
gradient error of both is a count of flipped units, each worth ~1e-3 / sqrt(pixels / 128^2).

Run with -s to see the table (profiles/r02_parity_fullsize.log is that output for the shipped build)."""
import pytest
import torch

from oracle import ist_oracle as O
from oracle import synth
from gpu_common import build_model, flip_aware_parity, frames, noise_like, parity_row, prepare_plan, strict_fp32

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")

LOSS_TOL = 1e-4
MASKED_GRAD_TOL = 1e-4


@pytest.fixture(scope="module")
def ctx():
    strict_fp32()
    cfg, model = build_model(dev)
    state_np = synth.vgg_state_dict(0, upto="conv5_1")
    st64, st32 = O.state_to_torch(state_np, torch.float64, dev), O.state_to_torch(state_np, torch.float32, dev)
    yield cfg, model, st64, st32
    model.vgg_model.release_plans()


def _check(tag, r):
    print(parity_row(tag, r))
    print(f"    flips by layer: ours {r['flips_by_layer']} | oracle fp32 {r['ref_flips_by_layer']} | loss rel err per entry "
          + " ".join(f"{v:.1e}" for v in r["loss_rel_each"]))
    assert r["loss_rel"] <= LOSS_TOL, r
    assert r["loss_rel_masked"] <= LOSS_TOL, r
    assert r["grad_rel_masked"] <= MASKED_GRAD_TOL, r
    # plain gradient: a count of flipped units; judged beside the reference's own fp32-vs-fp64 figure on the same point
    assert r["grad_rel"] <= max(2e-3, 3 * r["ref_grad_rel"]), r
    assert r["grad_cos1"] < 1e-4, r
    # our forward takes no more wrong decisions than a handful more than the reference's own fp32 forward
    assert r["flips"] <= 3 * r["ref_flips"] + 30, r


def _setup(ctx, size, kind):
    cfg, model, st64, st32 = ctx
    model.vgg_model.release_plans()
    torch.cuda.empty_cache()
    content, style = frames(size, dev, kind)
    plan = prepare_plan(model, cfg, content, style)
    t64 = O.compute_targets(st64, content.double(), style.double(), full=False)
    t32 = O.compute_targets(st32, content, style, full=False)
    return plan, content, style, t64, t32


@pytest.mark.parametrize("kind", ["radar", "smooth"])
def test_parity_512_p0_p1_p2(ctx, kind):
    cfg, model, st64, st32 = ctx
    plan, content, style, t64, t32 = _setup(ctx, 512, kind)
    _check(f"512 {kind} P0", flip_aware_parity(plan, content, st64, st32, t64, t32))
    _check(f"512 {kind} P1", flip_aware_parity(plan, content + noise_like(content), st64, st32, t64, t32))
    x2 = content.clone().requires_grad_(True)
    O.optimize(st32, content, style, x2, 20, full=False)
    _check(f"512 {kind} P2", flip_aware_parity(plan, x2.detach().contiguous(), st64, st32, t64, t32))


@pytest.mark.parametrize("kind", ["radar", "smooth"])
def test_parity_1024(ctx, kind):
    cfg, model, st64, st32 = ctx
    plan, content, style, t64, t32 = _setup(ctx, 1024, kind)
    _check(f"1024 {kind} P0", flip_aware_parity(plan, content, st64, st32, t64, t32))
    _check(f"1024 {kind} P1", flip_aware_parity(plan, content + noise_like(content), st64, st32, t64, t32))


@pytest.mark.parametrize("kind", ["radar", "smooth"])
def test_parity_2048(ctx, kind):
    cfg, model, st64, st32 = ctx
    plan, content, style, t64, t32 = _setup(ctx, 2048, kind)
    if kind == "radar":
        _check(f"2048 {kind} P0", flip_aware_parity(plan, content, st64, st32, t64, t32))
    _check(f"2048 {kind} P1", flip_aware_parity(plan, content + noise_like(content), st64, st32, t64, t32))
