"""L-BFGS / optimize() checks on a B200 (run under gpurun):
 1. device L-BFGS vs torch.optim.LBFGS driven by the SAME CUDA closure (isolates the optimiser arithmetic);
 2. optimize() (fused path) vs the oracle's optimize() in fp32 and fp64: PSNR + final loss, with the oracle's own
    fp32-vs-fp64 numbers beside ours;
 3. iterations/s of optimize() at 512^2.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ist_b200  # noqa: E402
from ist_b200.config import get_cfg_defaults  # noqa: E402
from ist_b200.lbfgs import DeviceLBFGS  # noqa: E402
from ist_b200.model import build_model  # noqa: E402
from ist_b200.model.engine.utils import optimize  # noqa: E402
from ist_b200.model.meta_arch import GramMSELoss, StyleTransfer  # noqa: E402
from oracle import ist_oracle as O  # noqa: E402
from oracle import synth  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")


def get_model(cfg, state_np):
    vgg = build_model(cfg).to(dev)
    vgg.load_state_dict({k: torch.from_numpy(v) for k, v in state_np.items()})
    for p in vgg.parameters():
        p.requires_grad = False
    fns = [GramMSELoss()] * len(cfg.LOSS.STYLE_LAYERS) + [torch.nn.MSELoss()] * len(cfg.LOSS.CONTENT_LAYERS)
    return StyleTransfer(vgg, cfg.LOSS.STYLE_LAYERS + cfg.LOSS.CONTENT_LAYERS, fns, cfg.LOSS.STYLE_WEIGHTS + cfg.LOSS.CONTENT_WEIGHTS)


def psnr(a, b):
    return synth.psnr(a.detach().cpu().numpy()[0], b.detach().cpu().numpy()[0])


def main():
    cfg = get_cfg_defaults()
    state_np = synth.vgg_state_dict(0)
    model = get_model(cfg, state_np)
    vgg = model.vgg_model
    st32 = O.state_to_torch(state_np, torch.float32, dev)
    st64 = O.state_to_torch(state_np, torch.float64, dev)

    for size, kind in ((128, "smooth"), (128, "radar"), (256, "smooth")):
        print(f"=== {size} {kind} ===")
        mk = synth.radar_frame if kind == "radar" else synth.smooth_frame
        content = torch.from_numpy(synth.preprocess(mk(size, 1))).to(dev)
        style = torch.from_numpy(synth.preprocess(synth.lidar_frame(size, 2))).to(dev)

        # ---- 1. optimiser arithmetic: same closure, our L-BFGS vs torch's ------------------------------------------------
        from ist_b200.model.engine.utils import style_targets
        plan = vgg.plan(1, size, size, "relu5_1")
        plan.set_loss(cfg.LOSS.STYLE_LAYERS, cfg.LOSS.STYLE_WEIGHTS, cfg.LOSS.CONTENT_LAYERS, cfg.LOSS.CONTENT_WEIGHTS)
        for k, g in enumerate(style_targets(vgg, style, cfg.LOSS.STYLE_LAYERS)):
            plan.set_style_target(k, g[0])
        plan.forward(content, "relu4_2")
        plan.capture_content_target(0)
        for max_iter, nsteps in ((1, 30), (20, 2)):
            xa = content.clone()
            opt = DeviceLBFGS(plan, max_iter=max_iter)
            tr_a = []
            for _ in range(nsteps):
                ev, l0 = opt.step(xa)
                tr_a.append(l0)
            fa = opt.last_losses()[0, -1].item()
            opt.close()
            xb = content.clone().requires_grad_(True)
            topt = torch.optim.LBFGS([xb], max_iter=max_iter)
            tr_b = []
            losses = torch.empty(1, 7, device=dev)

            def closure():
                g = torch.empty_like(xb)
                plan.loss_and_grad(xb.data, g, losses)
                xb.grad = g
                return losses[0, 6].clone()
            for _ in range(nsteps):
                tr_b.append(float(topt.step(closure)))
            rel = [abs(a - b) / abs(b) for a, b in zip(tr_a, tr_b)]
            print(f"  same-closure max_iter={max_iter}: first-loss trace rel diff: " + " ".join(f"{r:.1e}" for r in rel[:12]),
                  f"... last {rel[-1]:.1e}; PSNR(x ours, x torch-lbfgs) {psnr(xa, xb):.1f} dB")
            print(f"     losses ours {tr_a[-1]:.6e} torch {tr_b[-1]:.6e}")

        # ---- 2. end-to-end optimize() vs oracle -----------------------------------------------------------------------------
        for n_eval in (20, 60):
            x_ours = content.clone().requires_grad_(True)
            optimize(model, content, style, x_ours, cfg, n_eval)
            x32 = content.clone().requires_grad_(True)
            O.optimize(st32, content, style, x32, n_eval, full=False)
            x64 = content.double().clone().requires_grad_(True)
            O.optimize(st64, content.double(), style.double(), x64, n_eval, full=False)
            t64 = O.compute_targets(st64, content.double(), style.double(), full=False)
            lo = O.loss_and_grad(st64, x_ours.detach().double(), t64, full=False)[1]
            l32 = O.loss_and_grad(st64, x32.detach().double(), t64, full=False)[1]
            l64 = O.loss_and_grad(st64, x64.detach(), t64, full=False)[1]
            print(f"  optimize {n_eval} evals: PSNR ours-vs-fp64 {psnr(x_ours, x64):.1f} dB | oracle fp32-vs-fp64 {psnr(x32, x64):.1f} dB | "
                  f"ours-vs-oracle-fp32 {psnr(x_ours, x32):.1f} dB; final loss (fp64-evaluated) ours {lo:.5e} fp32 {l32:.5e} fp64 {l64:.5e}")

    # ---- 3. throughput at 512^2 ---------------------------------------------------------------------------------------------
    size = 512
    content = torch.from_numpy(synth.preprocess(synth.radar_frame(size, 1))).to(dev)
    style = torch.from_numpy(synth.preprocess(synth.lidar_frame(size, 2))).to(dev)
    for n_eval in (60, 300):
        x = content.clone().requires_grad_(True)
        torch.cuda.synchronize()
        t0 = time.time()
        optimize(model, content, style, x, cfg, n_eval)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"  optimize 512^2 {n_eval} evals: {dt:.3f} s -> {model.last_evals / dt:.1f} iters/s; final losses {model.last_losses[0].tolist()}")
    # torch reference loop on the same GPU (oracle, cuDNN) for orientation
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        x = content.clone().requires_grad_(True)
        torch.cuda.synchronize()
        t0 = time.time()
        _, n = O.optimize(st32, content, style, x, 60, full=True)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"  oracle optimize 512^2 60 evals on this GPU (cudnn tf32={tf32}): {dt:.3f} s -> {n / dt:.1f} iters/s")


if __name__ == "__main__":
    main()
