"""Generate tests/golden/*.npz|json by running the UNMODIFIED reference (imported read-only from /root/reference/IST).

TEST INFRASTRUCTURE. Run in the build container only (the reference tree does not exist on the GPU box):
    python oracle/make_golden.py
The reference needs `yacs` (not installed, no network): a 10-line stand-in for yacs.config.CfgNode is registered before
`import config`, which then returns the reference's real default tree. Nothing from the reference is copied:
only its outputs on seeded synthetic inputs are stored.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402

REF = "/root/reference/IST"
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)

    class CN(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

        def clone(self):
            def cp(n):
                if isinstance(n, CN):
                    return CN({k: cp(v) for k, v in n.items()})
                if isinstance(n, dict):
                    return {k: cp(v) for k, v in n.items()}
                if isinstance(n, list):
                    return [cp(v) for v in n]
                return n
            return cp(self)

        def freeze(self):
            pass

    y, yc = types.ModuleType("yacs"), types.ModuleType("yacs.config")
    yc.CfgNode = CN
    sys.modules["yacs"], sys.modules["yacs.config"] = y, yc
    from config import get_cfg_defaults
    from model import build_model
    from model.meta_arch import GramMatrix, GramMSELoss, StyleTransfer
    from model.engine.utils import optimize
    return get_cfg_defaults, build_model, GramMatrix, GramMSELoss, StyleTransfer, optimize


def to_plain(node):
    if isinstance(node, dict):
        return {k: to_plain(v) for k, v in node.items()}
    if isinstance(node, (list, tuple)):
        return [to_plain(v) for v in node]
    return node


def build_ref_model(get_cfg_defaults, build_model, GramMSELoss, StyleTransfer, dtype):
    cfg = get_cfg_defaults()
    cfg.MODEL.DEVICE = "cpu"
    vgg = build_model(cfg)
    state = synth.vgg_state_dict(seed=0)
    vgg.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()})
    for p in vgg.parameters():
        p.requires_grad = False
    if dtype == torch.float64:
        vgg.double()
    loss_layers = cfg.LOSS.STYLE_LAYERS + cfg.LOSS.CONTENT_LAYERS                       # main.py:35
    loss_functions = [GramMSELoss()] * len(cfg.LOSS.STYLE_LAYERS) + [torch.nn.MSELoss()] * len(cfg.LOSS.CONTENT_LAYERS)
    loss_weights = cfg.LOSS.STYLE_WEIGHTS + cfg.LOSS.CONTENT_WEIGHTS
    return cfg, StyleTransfer(vgg, loss_layers, loss_functions, loss_weights), state


def closure_eval(model, GramMatrix, cfg, content, style, x):
    """the closure body of the reference (utils.py:19-21, 29-36) at point x, using the reference's modules"""
    style_targets = [GramMatrix()(A).detach() for A in model.vgg_model(style, cfg.LOSS.STYLE_LAYERS)]
    content_targets = [A.detach() for A in model.vgg_model(content, cfg.LOSS.CONTENT_LAYERS)]
    targets = style_targets + content_targets
    xg = x.clone().requires_grad_(True)
    outs = model.vgg_model(xg, model.loss_layers)
    ll = [model.loss_weights[a] * model.loss_functions[a](A, targets[a]) for a, A in enumerate(outs)]
    loss = sum(ll)
    loss.backward()
    return targets, np.array([float(v) for v in ll] + [float(loss)]), xg.grad.detach().numpy()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    os.makedirs(GOLD, exist_ok=True)
    get_cfg_defaults, build_model, GramMatrix, GramMSELoss, StyleTransfer, optimize = import_reference()

    with open(os.path.join(GOLD, "cfg_defaults.json"), "w") as f:
        json.dump(to_plain(get_cfg_defaults()), f, indent=1, sort_keys=True)

    # ---- per-module vectors ------------------------------------------------------------------------------------------------
    rng = np.random.Generator(np.random.PCG64(123))
    feat = rng.standard_normal((2, 8, 5, 7)).astype(np.float32)
    tgt = rng.standard_normal((2, 8, 8)).astype(np.float32)
    G = GramMatrix()(torch.from_numpy(feat)).numpy()
    gl = float(GramMSELoss()(torch.from_numpy(feat), torch.from_numpy(tgt)))
    np.savez(os.path.join(GOLD, "modules.npz"), feat=feat, tgt=tgt, gram=G, gram_mse=np.float64(gl))

    # ---- closure vectors at two sizes, fp32 and fp64 -------------------------------------------------------------------------
    weights_checksum = None
    for tag, (h, w), kind in (("64", (64, 64), "radar"), ("48x80", (48, 80), "smooth")):
        out = {}
        mk = synth.radar_frame if kind == "radar" else synth.smooth_frame
        content_np = synth.preprocess(mk(0, 1, h=h, w=w))
        style_np = synth.preprocess(synth.lidar_frame(0, 2, h=h, w=w))
        noise = np.random.Generator(np.random.PCG64(3)).standard_normal(content_np.shape).astype(np.float32) * 20.0
        x1_np = content_np + noise
        out["content"], out["style"], out["x1"] = content_np, style_np, x1_np
        for dname, dtype in (("f32", torch.float32), ("f64", torch.float64)):
            cfg, model, state = build_ref_model(get_cfg_defaults, build_model, GramMSELoss, StyleTransfer, dtype)
            if weights_checksum is None:
                weights_checksum = {k: [float(np.sum(v, dtype=np.float64)), float(np.sum(v.astype(np.float64) ** 2))] for k, v in state.items()}
            content = torch.from_numpy(content_np).to(dtype)
            style = torch.from_numpy(style_np).to(dtype)
            for pname, xnp in (("p0", content_np), ("p1", x1_np)):
                targets, losses, grad = closure_eval(model, GramMatrix, cfg, content, style, torch.from_numpy(xnp).to(dtype))
                out[f"losses_{pname}_{dname}"] = losses
                out[f"grad_{pname}_{dname}"] = grad
            for k, t in enumerate(targets[:5]):
                t = t.numpy()
                out[f"gram{k}_corner_{dname}"] = t[0, :8, :8].copy()
                out[f"gram{k}_sums_{dname}"] = np.array([t.sum(dtype=np.float64), (t.astype(np.float64) ** 2).sum()])
            # features of the content image at a few keys (checksums + a corner) through VGG.forward
            feats = model.vgg_model(content, ["relu1_1", "pool_1", "relu3_1", "relu4_2", "pool_4", "relu5_1"])
            for key, ft in zip(["relu1_1", "pool_1", "relu3_1", "relu4_2", "pool_4", "relu5_1"], feats):
                ft = ft.detach().numpy()
                out[f"feat_{key}_sums_{dname}"] = np.array([ft.sum(dtype=np.float64), (ft.astype(np.float64) ** 2).sum()])
                out[f"feat_{key}_corner_{dname}"] = ft[0, :4, :3, :3].copy()
            # one optimizer.step(): optimize(..., max_iterations=20) == 20 closure evaluations (utils.py:28,43)
            if dname == "f32":
                x = torch.from_numpy(content_np).clone().requires_grad_(True)
                res = optimize(model, content, style, x, cfg, 20)
                out["opt20_f32"] = res.detach().numpy()
                tgs, l_end, _ = closure_eval(model, GramMatrix, cfg, content, style, res.detach())
                out["opt20_losses_f32"] = l_end
        np.savez_compressed(os.path.join(GOLD, f"closure_{tag}.npz"), **out)
        print("wrote closure_%s.npz" % tag, {k: v.shape for k, v in out.items() if k.startswith("losses")})
    with open(os.path.join(GOLD, "weights_checksum.json"), "w") as f:
        json.dump(weights_checksum, f, indent=1, sort_keys=True)
    print("golden vectors written to", GOLD)


if __name__ == "__main__":
    main()
