"""Kernel microbench, BASELINE configs[4]: every launch of one closure at SIZE x SIZE (default 1024) next to the cuDNN /
cuBLAS PyTorch op the reference runs for the same layer on the same B200, with TF32 on and off.

    python tools/gpu_microbench.py [SIZE] [REPS]

Ours: per-launch CUDA-event times of ist_plan_loss_and_grad (ist_profile_*), median of REPS closures.
Torch: F.conv2d(+bias)+relu / torch.nn.grad.conv2d_input + threshold_backward / max_pool2d fwd+bwd / bmm Gram + MSE fwd and
autograd backward / mse_loss, CUDA events, median of REPS after 3 warm-ups, fp32 NCHW as IST/model/meta_arch/vgg.py runs them.
"""
import ctypes
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ist_b200  # noqa: E402
from oracle import synth  # noqa: E402  (synthetic weights/frames only)
from tools.gpu_plan_check import LAYERS  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
state = {k: torch.from_numpy(v).to(dev) for k, v in synth.vgg_state_dict(0, upto="conv5_1").items()}
content = torch.from_numpy(synth.preprocess(synth.radar_frame(size, 1))).to(dev)
style = torch.from_numpy(synth.preprocess(synth.lidar_frame(size, 2))).to(dev)
SL = ['relu1_1', 'relu2_1', 'relu3_1', 'relu4_1', 'relu5_1']
SW = [1e3 / n ** 2 for n in [64, 128, 256, 512, 512]]


# ---- ours ---------------------------------------------------------------------------------------------------------------
def ours():
    plan = ist_b200.Plan(LAYERS, 1, size, size)
    plan.load_state_dict(state)
    plan.set_loss(SL, SW, ['relu4_2'], [0.5])
    plan.forward(style, "relu5_1")
    for k, key in enumerate(SL):
        plan.set_style_target(k, plan.gram(key)[0])
    plan.forward(content, "relu4_2")
    plan.capture_content_target(0)
    x = content + 20 * torch.randn_like(content)
    lib = ist_b200.load()
    for _ in range(3):
        plan.loss_and_grad(x)
    torch.cuda.synchronize()
    maxr, acc = 4096, None
    for _ in range(REPS):
        lib.ist_profile_begin()
        plan.loss_and_grad(x)
        names = ctypes.create_string_buffer(maxr * 40)
        flops, nbytes = (ctypes.c_double * maxr)(), (ctypes.c_double * maxr)()
        ms, n = (ctypes.c_float * maxr)(), ctypes.c_int(0)
        ist_b200._lib.check(lib.ist_profile_end(maxr, names, flops, nbytes, ms, ctypes.byref(n)))
        rows = [(names.raw[i * 40:(i + 1) * 40].split(b"\0")[0].decode(), flops[i], nbytes[i], ms[i]) for i in range(n.value)]
        if acc is None:
            acc = [[r[0], r[1], r[2], [r[3]]] for r in rows]
        else:
            for a, r in zip(acc, rows):
                a[3].append(r[3])
    plan.close()
    return [(a[0], a[1], a[2], sorted(a[3])[len(a[3]) // 2] * 1e3) for a in acc]     # us


# ---- torch ----------------------------------------------------------------------------------------------------------------
def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(REPS):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


def torch_ops(tf32):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = False          # the reference never enables it (torch default)
    torch.backends.cudnn.benchmark = True
    out = {}
    g = torch.Generator(device=dev).manual_seed(0)
    h = w = size
    with torch.no_grad():
        for kind, cin, cout, fname, oname in LAYERS:
            if kind == 0:
                x = torch.rand(1, cin, h, w, device=dev, generator=g) * 30
                wt, b = state[fname + ".weight"], state[fname + ".bias"]
                y = F.relu(F.conv2d(x, wt, b, padding=1))
                dy = torch.randn(1, cout, h, w, device=dev, generator=g)
                out[fname + ".fwd"] = timed(lambda: F.relu(F.conv2d(x, wt, b, padding=1)))
                out[fname + ".dgrad"] = timed(lambda: torch.nn.grad.conv2d_input(x.shape, wt, torch.ops.aten.threshold_backward(dy, y, 0), padding=1))
                if oname in SL:
                    A = torch.rand(1, cout, cout, device=dev, generator=g)

                    def gram_fwd():
                        Fm = y.view(1, cout, h * w)
                        G = torch.bmm(Fm, Fm.transpose(1, 2))
                        G.div_(h * w)
                        return F.mse_loss(G, A)
                    out[oname + ".gram_mse.fwd"] = timed(gram_fwd)
                    yr = y.clone().requires_grad_(True)

                    def gram_fwd_bwd():
                        with torch.enable_grad():
                            Fm = yr.view(1, cout, h * w)
                            G = torch.bmm(Fm, Fm.transpose(1, 2)) / (h * w)
                            loss = F.mse_loss(G, A)
                        yr.grad = None
                        loss.backward()
                    out[oname + ".gram_mse.fwd+bwd"] = timed(gram_fwd_bwd)
                if oname == "relu4_2":
                    t = torch.rand_like(y)
                    out["content_mse.fwd"] = timed(lambda: F.mse_loss(y, t))
                del x, y, dy
            else:
                c = [l for l in LAYERS[:LAYERS.index((kind, cin, cout, fname, oname))] if l[0] == 0][-1][2]
                x = torch.rand(1, c, h, w, device=dev, generator=g)
                out[fname + ".fwd"] = timed(lambda: F.max_pool2d(x, 2, 2))
                xr = x.clone().requires_grad_(True)
                with torch.enable_grad():
                    yp = F.max_pool2d(xr, 2, 2)
                dyp = torch.randn_like(yp)
                out[fname + ".bwd"] = timed(lambda: torch.autograd.grad(yp, xr, dyp, retain_graph=True))
                h, w = h // 2, w // 2
                del x, xr, yp, dyp
    return out


rows = ours()
t_tf32 = torch_ops(True)
t_fp32 = torch_ops(False)

conv_names = [l[3] for l in LAYERS if l[0] == 0]
pool_names = [l[3] for l in LAYERS if l[0] == 1]
fwd_iter, pool_iter, gram_iter = iter(conv_names), iter(pool_names), iter(SL)
dgrad_iter, route_iter = iter(reversed(conv_names)), iter(reversed(pool_names))
print(f"microbench {size}x{size}, batch 1, {REPS} reps, torch {torch.__version__}, {torch.cuda.get_device_name(0)}")
print(f"{'ours: kernel':22s} {'layer':18s} {'GFLOP':>8s} {'ours us':>9s} {'TF/s':>7s} | {'torch op(s)':34s} {'tf32 us':>9s} {'fp32 us':>9s} {'x tf32':>7s} {'x fp32':>7s}")
tot = {"ours": 0.0, "tf32": 0.0, "fp32": 0.0}
for name, fl, by, us in rows:
    layer, key, what = "", None, ""
    if name in ("conv_first_fwd", "conv_halo_fwd"):
        layer = next(fwd_iter); key = layer + ".fwd"; what = "conv2d+bias, relu"
    elif name in ("conv_halo_dgrad", "conv_first_dgrad"):
        layer = next(dgrad_iter); key = layer + ".dgrad"; what = "threshold_backward, conv2d_input"
    elif name == "maxpool_fwd":
        layer = next(pool_iter); key = layer + ".fwd"; what = "max_pool2d"
    elif name == "grad_route":
        layer = next(route_iter); key = layer + ".bwd"; what = "max_pool2d backward (+relu bwd ours)"
    elif name == "gram_syrk":
        layer = next(gram_iter); key = layer + ".gram_mse.fwd"; what = "bmm, div_, mse_loss"
    elif name == "content_partial":
        layer = "relu4_2"; key = "content_mse.fwd"; what = "mse_loss"
    a, b = (t_tf32.get(key), t_fp32.get(key)) if key else (None, None)
    tot["ours"] += us
    if a is not None:
        tot["tf32"] += a; tot["fp32"] += b
    print(f"{name:22s} {layer:18s} {fl / 1e9:8.3f} {us:9.1f} {fl / (us * 1e-6) / 1e12 if fl else 0:7.1f} | {what:34s} "
          + (f"{a:9.1f} {b:9.1f} {a / us:7.2f} {b / us:7.2f}" if a is not None else f"{'-':>9s} {'-':>9s}"))
gb_t = sum(t_tf32[k + ".gram_mse.fwd+bwd"] - t_tf32[k + ".gram_mse.fwd"] for k in SL)
gb_f = sum(t_fp32[k + ".gram_mse.fwd+bwd"] - t_fp32[k + ".gram_mse.fwd"] for k in SL)
print(f"{'(fused into dgrad)':22s} {'5 Gram backwards':18s} {'':8s} {'':9s} {'':7s} | {'autograd of bmm/div/mse (2 bmm each)':34s} {gb_t:9.1f} {gb_f:9.1f}")
tot["tf32"] += gb_t; tot["fp32"] += gb_f
print(f"sum over the closure: ours {tot['ours'] / 1e3:.3f} ms | torch ops tf32 {tot['tf32'] / 1e3:.3f} ms ({tot['tf32'] / tot['ours']:.2f}x) | "
      f"torch ops fp32 {tot['fp32'] / 1e3:.3f} ms ({tot['fp32'] / tot['ours']:.2f}x)")
