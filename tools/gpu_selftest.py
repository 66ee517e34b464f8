"""Per-op self-test of libist_b200.so on a B200, with diagnostics (run under gpurun; not a pytest file).

Compares every C-ABI op against a float64 torch evaluation of the same math on the GPU. Prints one line per case and,
for a failing conv case, structured probes (identity / single-tap weights) that localise TMA-vs-descriptor errors.
"""
import ctypes
import importlib.util
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location(
    "ist_lib", os.path.join(ROOT, "can-image-style-transfer-save-automotive-radar_b200", "_lib.py"))
L = importlib.util.module_from_spec(spec)
spec.loader.exec_module(L)
lib = L.load()
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
FAILS = []


def rel(a, b):
    a = a.double()
    b = b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item(), ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def report(name, got, ref, tol):
    torch.cuda.synchronize()
    r2, rmax = rel(got, ref)
    ok = (r2 < tol) and bool(torch.isfinite(got).all())
    print(f"{'PASS' if ok else 'FAIL'} {name:58s} rel-L2 {r2:.3e}  max/max {rmax:.3e}  (tol {tol:.0e})", flush=True)
    if not ok:
        FAILS.append(name)
    return ok


def conv_fwd(x, w, b):
    NB, cin, H, W = x.shape
    cout = w.shape[0]
    y = torch.empty(NB, cout, H, W, device=dev)
    L.check(lib.ist_op_conv3x3_relu_fwd(L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(y), NB, cin, cout, H, W, 1, L.stream_ptr()))
    return y


def conv_dgrad(dy, w, cin, passes=3):
    NB, cout, H, W = dy.shape
    dx = torch.empty(NB, cin, H, W, device=dev)
    L.check(lib.ist_op_conv3x3_dgrad(L.ptr(dy), L.ptr(w), L.ptr(dx), NB, cin, cout, H, W, passes, L.stream_ptr()))
    return dx


def probe_conv(cin, cout, H, W):
    """Localise a conv failure: which of {A tile rows, channel order, tap shift, B rows} is wrong."""
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.rand(1, cin, H, W, generator=g).to(dev) + 0.5
    b = torch.zeros(cout, device=dev)
    for tap in (4, 0, 2, 6, 8):
        w = torch.zeros(cout, cin, 3, 3, device=dev)
        n = min(cin, cout)
        w[torch.arange(n), torch.arange(n), tap // 3, tap % 3] = 1.0
        y = conv_fwd(x, w, b)
        ref = F.relu(F.conv2d(x.double(), w.double(), padding=1)).float()
        r2, _ = rel(y, ref)
        print(f"    probe identity tap {tap}: rel-L2 {r2:.3e}")
        if r2 > 1e-4 and tap == 4:
            d = (y - ref).abs()
            bad_c = (d.amax(dim=(0, 2, 3)) > 1e-3).nonzero().flatten().tolist()
            bad_y = (d.amax(dim=(0, 1, 3)) > 1e-3).nonzero().flatten().tolist()
            bad_x = (d.amax(dim=(0, 1, 2)) > 1e-3).nonzero().flatten().tolist()
            print(f"      bad channels {bad_c[:16]}.. ({len(bad_c)}), bad rows {bad_y[:16]} ({len(bad_y)}), bad cols {bad_x[:16]} ({len(bad_x)})")
            print("      y[0,:8,0,0] ", y[0, :8, 0, 0].tolist())
            print("      ref[0,:8,0,0]", ref[0, :8, 0, 0].tolist())
            print("      y[0,0,0,:8] ", y[0, 0, 0, :8].tolist())
            print("      ref[0,0,0,:8]", ref[0, 0, 0, :8].tolist())


def main():
    print("device:", torch.cuda.get_device_name(0), "cc", torch.cuda.get_device_capability(0))
    L.check(lib.ist_device_check())
    g = torch.Generator(device="cpu").manual_seed(0)

    def randn(*s, scale=1.0):
        return (torch.randn(*s, generator=g) * scale).to(dev)

    # ---- first conv -------------------------------------------------------------------------------------------
    for (NB, H, W) in [(1, 32, 32), (2, 24, 40), (1, 7, 9)]:
        x = randn(NB, 3, H, W, scale=60.0)
        w = randn(64, 3, 3, 3, scale=0.27)
        b = randn(64, scale=0.5)
        y = conv_fwd(x, w, b)
        ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), padding=1))
        report(f"conv_first fwd NB{NB} {H}x{W}", y, ref, 2e-6)
        dy = randn(NB, 64, H, W)
        dx = conv_dgrad(dy, w, 3)
        refd = torch.nn.grad.conv2d_input(x.double().shape, w.double(), dy.double(), padding=1)
        report(f"conv_first dgrad NB{NB} {H}x{W}", dx, refd, 3e-5)

    # ---- tcgen05 implicit GEMM conv ------------------------------------------------------------------------------
    cases = [(1, 64, 64, 16, 16), (1, 64, 64, 8, 16), (1, 64, 128, 32, 32), (2, 128, 128, 24, 40), (1, 128, 256, 16, 16),
             (1, 256, 256, 19, 21), (1, 256, 512, 8, 8), (1, 512, 512, 4, 4), (2, 512, 512, 2, 2), (1, 512, 512, 1, 1),
             (1, 64, 64, 96, 96)]
    for (NB, cin, cout, H, W) in cases:
        x = F.relu(randn(NB, cin, H, W, scale=40.0))
        w = randn(cout, cin, 3, 3, scale=(2.0 / (9 * cin)) ** 0.5)
        b = randn(cout, scale=0.5)
        try:
            y = conv_fwd(x, w, b)
            ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), padding=1))
            ok = report(f"conv_igemm fwd NB{NB} {cin}->{cout} {H}x{W}", y, ref, 6e-7)
            if not ok and (cin, cout) == (64, 64):
                probe_conv(cin, cout, H, W)
            dy = randn(NB, cout, H, W)
            dx = conv_dgrad(dy, w, cin, 3)
            refd = torch.nn.grad.conv2d_input(x.double().shape, w.double(), dy.double(), padding=1)
            report(f"conv_igemm dgrad(3-pass bf16) NB{NB} {cin}<-{cout} {H}x{W}", dx, refd, 5e-5)
            dx1 = conv_dgrad(dy, w, cin, 1)
            report(f"conv_igemm dgrad(1-pass bf16) NB{NB} {cin}<-{cout} {H}x{W}", dx1, refd, 1e-2)
        except L.IstError as e:
            print("FAIL", (NB, cin, cout, H, W), e)
            FAILS.append(str((NB, cin, cout, H, W)))
            raise

    # ---- pool / relu ------------------------------------------------------------------------------------------------
    for (NB, C, H, W) in [(1, 64, 16, 16), (2, 128, 9, 11), (1, 64, 32, 32)]:
        x = F.relu(randn(NB, C, H, W, scale=30.0))
        x[:, :, : H // 2, :] = x[:, :, : H // 2, :].round()      # exact ties
        x[:, : C // 2, :, : W // 2] = 0.0                        # constant background (radar-like)
        y = torch.empty(NB, C, H // 2, W // 2, device=dev)
        L.check(lib.ist_op_maxpool2x2_fwd(L.ptr(x), L.ptr(y), NB, C, H, W, L.stream_ptr()))
        xr = x.clone().requires_grad_(True)
        yr = F.max_pool2d(xr, 2, 2)
        report(f"maxpool fwd NB{NB} C{C} {H}x{W}", y, yr.detach(), 1e-6)
        dy = randn(NB, C, H // 2, W // 2)
        yr.backward(dy)
        dx = torch.empty_like(x)
        L.check(lib.ist_op_maxpool2x2_bwd(L.ptr(x), L.ptr(dy), L.ptr(dx), NB, C, H, W, L.stream_ptr()))
        report(f"maxpool bwd (ties, first-max) NB{NB} C{C} {H}x{W}", dx, xr.grad, 1e-7)
        dyf = randn(NB, C, H, W)
        dxr = torch.empty_like(x)
        L.check(lib.ist_op_relu_bwd(L.ptr(x), L.ptr(dyf), L.ptr(dxr), NB, C, H, W, L.stream_ptr()))
        report(f"relu bwd NB{NB} C{C} {H}x{W}", dxr, dyf * (x > 0), 1e-7)

    # ---- Gram / GramMSE / MSE ---------------------------------------------------------------------------------------
    for (NB, C, H, W) in [(1, 64, 32, 32), (1, 64, 10, 10), (2, 128, 16, 16), (1, 256, 16, 24), (1, 512, 8, 8), (1, 512, 2, 2),
                          (1, 64, 128, 128)]:
        x = F.relu(randn(NB, C, H, W, scale=30.0))
        G = torch.empty(NB, C, C, device=dev)
        L.check(lib.ist_op_gram(L.ptr(x), L.ptr(G), NB, C, H, W, L.stream_ptr()))
        Fm = x.double().view(NB, C, H * W)
        Gr = torch.bmm(Fm, Fm.transpose(1, 2)) / (H * W)
        report(f"gram NB{NB} C{C} {H}x{W}", G, Gr, 2e-6)
        # GramMSE forward/backward for each frame against its own b=1 reference
        tgt = (Gr[0] * 0.7 + 3.0).float().contiguous()
        wgt = 1e3 / C ** 2
        loss = torch.empty(NB, device=dev)
        dx = torch.empty_like(x)
        L.check(lib.ist_op_gram_mse(L.ptr(x), L.ptr(tgt), wgt, L.ptr(loss), L.ptr(dx), NB, C, H, W, L.stream_ptr()))
        xr = x.double().clone().requires_grad_(True)
        tot = 0
        lr = []
        for n in range(NB):
            Fn = xr[n:n + 1].view(1, C, H * W)
            Gn = torch.bmm(Fn, Fn.transpose(1, 2)) / (H * W)
            ln = wgt * F.mse_loss(Gn, tgt.double()[None])
            lr.append(ln.detach())
            tot = tot + ln
        tot.backward()
        report(f"gram_mse loss NB{NB} C{C} {H}x{W}", loss, torch.stack(lr), 5e-6)
        report(f"gram_mse grad NB{NB} C{C} {H}x{W}", dx, xr.grad, 2e-5)
    for (NB, C, H, W) in [(1, 512, 8, 8), (2, 64, 9, 7)]:
        x = F.relu(randn(NB, C, H, W, scale=30.0))
        t = F.relu(randn(NB, C, H, W, scale=30.0))
        loss = torch.empty(NB, 2, device=dev)
        dx = torch.empty_like(x)
        L.check(lib.ist_op_mse(L.ptr(x), L.ptr(t), 0.5, L.ptr(loss), L.ptr(dx), NB, C, H, W, L.stream_ptr()))
        lr = torch.stack([0.5 * F.mse_loss(x[n].double(), t[n].double()) for n in range(NB)])
        gr = 2 * 0.5 * (x.double() - t.double()) / (C * H * W)
        report(f"mse loss NB{NB} C{C} {H}x{W}", loss[:, 0], lr, 2e-6)
        report(f"mse grad NB{NB} C{C} {H}x{W}", dx, gr, 2e-6)

    print("FAILS:", FAILS)
    return 1 if FAILS else 0


if __name__ == "__main__":
    t0 = time.time()
    rc = main()
    print(f"selftest done in {time.time() - t0:.1f}s rc={rc}")
    sys.exit(rc)
