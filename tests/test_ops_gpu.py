"""Per-op parity of the CUDA kernels, called through the C ABI, against float64 evaluations of the reference's ops
(F.conv2d+relu, max_pool2d, bmm Gram, mse_loss and their autograd) on the same seeded inputs.
Tolerances (relative L2, stated per test): forward conv / Gram 6e-7 (fp32-class: fp16 hi/lo operands, promoted fp32
accumulation), data-gradient 5e-5 (bf16 hi/lo operands), elementwise / routing exact."""
import pytest
import torch
import torch.nn.functional as F

from ist_b200 import _lib
from gpu_common import rel_l2, strict_fp32

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


def randn(g, *s, scale=1.0):
    return (torch.randn(*s, generator=g) * scale).to(dev)


def conv_fwd(lib, x, w, b):
    nb, cin, h, wd = x.shape
    y = torch.empty(nb, w.shape[0], h, wd, device=dev)
    _lib.check(lib.ist_op_conv3x3_relu_fwd(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), nb, cin, w.shape[0], h, wd, 1, _lib.stream_ptr()))
    return y


def conv_dgrad(lib, dy, w, cin, passes=3):
    nb, cout, h, wd = dy.shape
    dx = torch.empty(nb, cin, h, wd, device=dev)
    _lib.check(lib.ist_op_conv3x3_dgrad(_lib.ptr(dy), _lib.ptr(w), _lib.ptr(dx), nb, cin, cout, h, wd, passes, _lib.stream_ptr()))
    return dx


@pytest.mark.parametrize("tc", [1, 0])
@pytest.mark.parametrize("nb,h,w", [(1, 32, 32), (2, 24, 40), (1, 7, 9), (1, 1, 1), (1, 37, 90), (2, 70, 52), (1, 129, 17)])
def test_first_conv(lib, nb, h, w, tc):
    """conv1_1 forward / data-gradient through the dispatch the plan uses: tc = 1 the shipped tensor-core kernels
    (conv_first_fwd_tc_kernel, conv_first_dgrad_tc_kernel), tc = 0 the CUDA-core kernels (IST_B200_CFF / CFD = cuda);
    sizes include non-multiples of the 16 x 8 pixel tile."""
    strict_fp32()
    _lib.check(lib.ist_set_option(b"first_conv_fwd_tc", tc))
    _lib.check(lib.ist_set_option(b"first_conv_dgrad_tc", tc))
    try:
        g = torch.Generator().manual_seed(1)
        x, wt, b = randn(g, nb, 3, h, w, scale=60.0), randn(g, 64, 3, 3, 3, scale=0.27), randn(g, 64, scale=0.5)
        ref = F.relu(F.conv2d(x.double(), wt.double(), b.double(), padding=1))
        n0 = lib.ist_launch_count()
        y = conv_fwd(lib, x, wt, b)
        assert rel_l2(y, ref) < 6e-7
        pre = F.conv2d(x.double(), wt.double(), b.double(), padding=1)
        assert int((((y > 0) != (pre > 0)) & (pre.abs() > 1e-4 * pre.abs().max())).sum()) == 0
        dy = randn(g, nb, 64, h, w)
        refd = torch.nn.grad.conv2d_input(x.double().shape, wt.double(), dy.double(), padding=1)
        assert rel_l2(conv_dgrad(lib, dy, wt, 3), refd) < 5e-5
        assert lib.ist_launch_count() > n0
    finally:
        _lib.check(lib.ist_set_option(b"first_conv_fwd_tc", 1))
        _lib.check(lib.ist_set_option(b"first_conv_dgrad_tc", 1))


@pytest.mark.parametrize("nb,cin,cout,h,w", [
    (1, 64, 64, 16, 16), (1, 64, 64, 8, 16), (1, 64, 128, 32, 32), (2, 128, 128, 24, 40), (1, 128, 256, 16, 16),
    (1, 256, 256, 19, 21), (1, 256, 512, 8, 8), (1, 512, 512, 4, 4), (2, 512, 512, 2, 2), (1, 512, 512, 1, 1), (1, 64, 64, 96, 96)])
def test_conv_igemm(lib, nb, cin, cout, h, w):
    strict_fp32()
    g = torch.Generator().manual_seed(cin + cout + h)
    x = F.relu(randn(g, nb, cin, h, w, scale=40.0))
    wt = randn(g, cout, cin, 3, 3, scale=(2.0 / (9 * cin)) ** 0.5)
    b = randn(g, cout, scale=0.5)
    ref = F.relu(F.conv2d(x.double(), wt.double(), b.double(), padding=1))
    y = conv_fwd(lib, x, wt, b)
    assert torch.isfinite(y).all() and rel_l2(y, ref) < 6e-7
    # ReLU sign pattern == reference's, up to |pre-activation| below the fp32 noise level
    pre = F.conv2d(x.double(), wt.double(), b.double(), padding=1)
    flips = ((y > 0) != (pre > 0)) & (pre.abs() > 1e-4 * pre.abs().max())
    assert int(flips.sum()) == 0
    dy = randn(g, nb, cout, h, w)
    refd = torch.nn.grad.conv2d_input(x.double().shape, wt.double(), dy.double(), padding=1)
    assert rel_l2(conv_dgrad(lib, dy, wt, cin, 3), refd) < 5e-5
    assert rel_l2(conv_dgrad(lib, dy, wt, cin, 1), refd) < 1e-2      # hi*hi only: bf16-level, for the record


def test_conv_translation_invariance(lib):
    """Equal patches give bit-equal outputs (ties must stay ties): a periodic image through the conv."""
    g = torch.Generator().manual_seed(9)
    tile = F.relu(randn(g, 1, 64, 8, 8, scale=30.0))
    x = tile.repeat(1, 1, 8, 8).contiguous()          # 64x64, period 8
    wt, b = randn(g, 128, 64, 3, 3, scale=0.06), randn(g, 128, scale=0.3)
    y = conv_fwd(lib, x, wt, b)
    inner = y[:, :, 8:56, 8:56]
    assert torch.equal(inner[:, :, 0:8, 0:8], inner[:, :, 8:16, 16:24])
    assert torch.equal(inner[:, :, 0:8, 0:8], inner[:, :, 32:40, 24:32])


@pytest.mark.parametrize("nb,c,h,w", [(1, 64, 16, 16), (2, 128, 9, 11), (1, 64, 32, 32), (1, 8, 5, 5)])
def test_pool_and_relu_routing(lib, nb, c, h, w):
    g = torch.Generator().manual_seed(h)
    x = F.relu(randn(g, nb, c, h, w, scale=30.0))
    x[:, :, : h // 2, :] = x[:, :, : h // 2, :].round()      # exact ties
    x[:, : c // 2, :, : w // 2] = 0.0                        # constant background, as in radar frames
    y = torch.empty(nb, c, h // 2, w // 2, device=dev)
    _lib.check(lib.ist_op_maxpool2x2_fwd(_lib.ptr(x), _lib.ptr(y), nb, c, h, w, _lib.stream_ptr()))
    xr = x.clone().requires_grad_(True)
    yr = F.max_pool2d(xr, 2, 2)
    assert rel_l2(y, yr.detach()) < 1e-6
    dy = randn(g, nb, c, h // 2, w // 2)
    yr.backward(dy)
    dx = torch.empty_like(x)
    _lib.check(lib.ist_op_maxpool2x2_bwd(_lib.ptr(x), _lib.ptr(dy), _lib.ptr(dx), nb, c, h, w, _lib.stream_ptr()))
    assert torch.equal(dx, xr.grad)                           # first-max tie rule, floor on odd sizes
    dyf = randn(g, nb, c, h, w)
    dxr = torch.empty_like(x)
    _lib.check(lib.ist_op_relu_bwd(_lib.ptr(x), _lib.ptr(dyf), _lib.ptr(dxr), nb, c, h, w, _lib.stream_ptr()))
    assert torch.equal(dxr, dyf * (x > 0))                    # ReLU'(0) = 0


@pytest.mark.parametrize("nb,c,h,w", [(1, 64, 32, 32), (1, 64, 10, 10), (2, 128, 16, 16), (1, 256, 16, 24), (1, 512, 8, 8),
                                      (1, 512, 2, 2), (1, 64, 128, 128)])
def test_gram_and_gram_mse(lib, nb, c, h, w):
    strict_fp32()
    g = torch.Generator().manual_seed(c + h)
    x = F.relu(randn(g, nb, c, h, w, scale=30.0))
    G = torch.empty(nb, c, c, device=dev)
    _lib.check(lib.ist_op_gram(_lib.ptr(x), _lib.ptr(G), nb, c, h, w, _lib.stream_ptr()))
    Fm = x.double().view(nb, c, h * w)
    Gr = torch.bmm(Fm, Fm.transpose(1, 2)) / (h * w)
    assert rel_l2(G, Gr) < 6e-7
    tgt = (Gr[0] * 0.7 + 3.0).float().contiguous()
    wgt = 1e3 / c ** 2
    loss = torch.empty(nb, device=dev)
    dx = torch.empty_like(x)
    _lib.check(lib.ist_op_gram_mse(_lib.ptr(x), _lib.ptr(tgt), wgt, _lib.ptr(loss), _lib.ptr(dx), nb, c, h, w, _lib.stream_ptr()))
    xr = x.double().clone().requires_grad_(True)
    lr = []
    for n in range(nb):
        Fn = xr[n:n + 1].view(1, c, h * w)
        lr.append(wgt * F.mse_loss(torch.bmm(Fn, Fn.transpose(1, 2)) / (h * w), tgt.double()[None]))
    sum(lr).backward()
    assert rel_l2(loss, torch.stack(lr).detach()) < 5e-6
    assert rel_l2(dx, xr.grad) < 2e-5
    # arbitrary upstream gradient through the Gram backward entry point
    dg = randn(g, nb, c, c)
    dx2 = torch.empty_like(x)
    _lib.check(lib.ist_op_gram_bwd(_lib.ptr(x), _lib.ptr(dg), _lib.ptr(dx2), nb, c, h, w, _lib.stream_ptr()))
    ref2 = torch.bmm((dg + dg.transpose(1, 2)).double(), Fm).view(nb, c, h, w) / (h * w)
    assert rel_l2(dx2, ref2) < 2e-5


@pytest.mark.parametrize("nb,c,h,w", [(1, 512, 8, 8), (2, 64, 9, 7)])
def test_mse(lib, nb, c, h, w):
    g = torch.Generator().manual_seed(4)
    x, t = F.relu(randn(g, nb, c, h, w, scale=30.0)), F.relu(randn(g, nb, c, h, w, scale=30.0))
    loss = torch.empty(nb, 2, device=dev)
    dx = torch.empty_like(x)
    _lib.check(lib.ist_op_mse(_lib.ptr(x), _lib.ptr(t), 0.5, _lib.ptr(loss), _lib.ptr(dx), nb, c, h, w, _lib.stream_ptr()))
    lr = torch.stack([0.5 * F.mse_loss(x[n].double(), t[n].double()) for n in range(nb)])
    assert rel_l2(loss[:, 0], lr) < 2e-6
    assert rel_l2(dx, (x.double() - t.double()) / (c * h * w)) < 2e-6
    _lib.check(lib.ist_op_mse(_lib.ptr(x), _lib.ptr(x.clone()), 0.5, _lib.ptr(loss), _lib.ptr(dx), nb, c, h, w, _lib.stream_ptr()))
    assert float(loss[:, 0].abs().max()) == 0.0 and float(dx.abs().max()) == 0.0
