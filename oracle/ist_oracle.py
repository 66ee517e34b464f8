"""CPU/PyTorch restatement of the reference's Gatys style-transfer path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs (cpu_baseline, `--impl reference`, and `gpu_reference`: the
reference's own PyTorch GPU path timed beside the product on the same B200) may import this module, and only as the checker
or the timed baseline — never as the product path (the product path is the CUDA library behind
include/ist_b200.h and raises when that library is missing).

What is restated (paths relative to the reference root, DJNing/Can-Image-Style-Transfer-Save-Automotive-Radar):
  * VGG.forward                IST/model/meta_arch/vgg.py:44-58 (layer table: IST/config/defaults.py:22-61)
  * GramMatrix / GramMSELoss   IST/model/meta_arch/gram_matrix.py:6-11, gram_mse_loss.py:6-8
  * content loss nn.MSELoss    IST/main.py:36-37
  * loss weights               IST/config/defaults.py:67-70
  * closure + optimize()       IST/model/engine/utils.py:17-45 (torch.optim.LBFGS with all defaults, :24)
  * ImageTransform             IST/data/image_transform.py:5-31 (transforms.Scale -> Resize, its modern name)
  * coarse-to-fine stage       IST/model/engine/hr_transfer_style.py:11-33
The arithmetic of the path lives in PyTorch (reference pins torch==1.7.1+cu110, docker/dockerfile:25; this image has
torch 2.11): F.conv2d, F.relu, F.max_pool2d, torch.bmm, F.mse_loss, autograd and torch.optim.LBFGS are called here exactly
at the reference's call sites.

Pinning: the reference has no tests and no golden vectors, so this restatement is pinned against the
reference ITSELF: oracle/make_golden.py imports the unmodified modules from /root/reference/IST, runs them on seeded
synthetic inputs and stores their outputs in tests/golden/*.npz; tests/test_oracle_golden.py checks this file against
those vectors on every CPU test run.
"""
import numpy as np
import torch
import torch.nn.functional as F

FORWARD_SEQ = [
    'conv1_1', 'conv1_2', 'pool_1',
    'conv2_1', 'conv2_2', 'pool_2',
    'conv3_1', 'conv3_2', 'conv3_3', 'conv3_4', 'pool_3',
    'conv4_1', 'conv4_2', 'conv4_3', 'conv4_4', 'pool_4',
    'conv5_1', 'conv5_2', 'conv5_3', 'conv5_4', 'pool_5',
]
OUT_SEQ = [
    'relu1_1', 'relu1_2', 'pool_1',
    'relu2_1', 'relu2_2', 'pool_2',
    'relu3_1', 'relu3_2', 'relu3_3', 'relu3_4', 'pool_3',
    'relu4_1', 'relu4_2', 'relu4_3', 'relu4_4', 'pool_4',
    'relu5_1', 'relu5_2', 'relu5_3', 'relu5_4', 'pool_5',
]
STYLE_LAYERS = ['relu1_1', 'relu2_1', 'relu3_1', 'relu4_1', 'relu5_1']
CONTENT_LAYERS = ['relu4_2']
STYLE_WEIGHTS = [1e3 / n ** 2 for n in [64, 128, 256, 512, 512]]
CONTENT_WEIGHTS = [5e-1]


def state_to_torch(state_np, dtype=torch.float32, device='cpu'):
    return {k: torch.from_numpy(np.asarray(v)).to(device=device, dtype=dtype) for k, v in state_np.items()}


def vgg_forward(state, x, out_keys, full=True):
    """vgg.py:44-58. `full=True` walks all 21 layers like the reference; `full=False` stops at the deepest requested key."""
    outputs = {}
    prev = x
    last = len(FORWARD_SEQ) - 1 if full else max(OUT_SEQ.index(k) for k in out_keys)
    for i in range(last + 1):
        name = FORWARD_SEQ[i]
        if name.find('conv') != -1:
            if name + '.weight' not in state:
                break
            outputs[OUT_SEQ[i]] = F.relu(F.conv2d(prev, state[name + '.weight'], state[name + '.bias'], padding=1))
        elif name.find('pool') != -1:
            outputs[OUT_SEQ[i]] = F.max_pool2d(prev, kernel_size=2, stride=2)
        prev = outputs[OUT_SEQ[i]]
    return [outputs[k] for k in out_keys]


def gram_matrix(x):
    """gram_matrix.py:6-11 (divides by h*w only)."""
    b, c, h, w = x.size()
    Fm = x.view(b, c, h * w)
    G = torch.bmm(Fm, Fm.transpose(1, 2))
    G.div_(h * w)
    return G


def gram_mse_loss(x, target):
    """gram_mse_loss.py:6-8."""
    return F.mse_loss(gram_matrix(x), target)


def compute_targets(state, content_image, style_image, style_layers=STYLE_LAYERS, content_layers=CONTENT_LAYERS, full=True):
    """utils.py:19-21."""
    style_targets = [gram_matrix(A).detach() for A in vgg_forward(state, style_image, style_layers, full)]
    content_targets = [A.detach() for A in vgg_forward(state, content_image, content_layers, full)]
    return style_targets + content_targets


def layer_losses(state, x, targets, style_layers=STYLE_LAYERS, content_layers=CONTENT_LAYERS,
                 weights=None, full=True):
    """utils.py:31-33: weighted per-layer losses, style layers first then content layers."""
    weights = weights if weights is not None else STYLE_WEIGHTS + CONTENT_WEIGHTS
    outs = vgg_forward(state, x, list(style_layers) + list(content_layers), full)
    fns = [gram_mse_loss] * len(style_layers) + [F.mse_loss] * len(content_layers)
    return [weights[a] * fns[a](A, targets[a]) for a, A in enumerate(outs)]


def loss_and_grad(state, x, targets, **kw):
    """closure body of utils.py:29-41 for one evaluation point: returns ([weighted layer losses], total, d total / d x)."""
    xg = x.detach().clone().requires_grad_(True)
    ll = layer_losses(state, xg, targets, **kw)
    loss = sum(ll)
    loss.backward()
    return [float(l.detach()) for l in ll], float(loss.detach()), xg.grad.detach()


# ---------------------------------------------------------------------------------------------------------------------
# Flip-aware parity (SURVEY 7.3 H1 / 8d "Parity protocol"): the image gradient is discontinuous in the forward activations
# through the ReLU sign maps and the max-pool argmax maps, so parity is reported as (i) the number of mask mismatches between
# two forwards and (ii) the gradient of the SAME closure evaluated with the masks imposed from outside — with equal masks
# the closure is a smooth function of its inputs and the remaining difference is kernel arithmetic alone.
# ---------------------------------------------------------------------------------------------------------------------
def _last_needed(keys):
    return max(OUT_SEQ.index(k) for k in keys)


def forward_masks(state, x, upto='relu5_1'):
    """The decisions vgg.py:52,54 take on x: {relu name: bool [b,C,H,W] (conv output > 0), pool name: uint8 [b,C,H/2,W/2]
    window position 0..3 (row-major, first maximum wins as in ATen's max_pool2d) or 4 where the pooled value is not
    positive (all four inputs are masked by the preceding ReLU; no gradient passes)}."""
    masks = {}
    prev = x
    for i in range(OUT_SEQ.index(upto) + 1):
        name = FORWARD_SEQ[i]
        if name.find('conv') != -1:
            prev = F.relu(F.conv2d(prev, state[name + '.weight'], state[name + '.bias'], padding=1))
            masks[OUT_SEQ[i]] = prev > 0
        else:
            h, w = prev.shape[2], prev.shape[3]
            prev, idx = F.max_pool2d(prev, kernel_size=2, stride=2, return_indices=True)
            ho, wo = prev.shape[2], prev.shape[3]
            iy, ix = idx // w, idx % w
            yo = torch.arange(ho, device=x.device).view(1, 1, ho, 1)
            xo = torch.arange(wo, device=x.device).view(1, 1, 1, wo)
            pos = ((iy - 2 * yo) * 2 + (ix - 2 * xo)).to(torch.uint8)
            masks[OUT_SEQ[i]] = torch.where(prev > 0, pos, torch.full_like(pos, 4))
    return masks


def mask_mismatches(a, b):
    """{layer: (mismatching units, units)} between two mask sets of forward_masks layout. A pool unit whose pooled value is
    not positive on both sides (code 4) matches whatever the positions."""
    out = {}
    for k in a:
        if k not in b:
            continue
        out[k] = (int((a[k] != b[k]).sum()), a[k].numel())
    return out


def vgg_forward_masked(state, x, masks, out_keys):
    """vgg.py:44-58 with the decisions imposed: relu(y) -> y * mask, max-pool -> the window element at the given position
    (code 4 reads position 0: its gradient is stopped by the ReLU mask of the layer below, which is all-zero there)."""
    outputs = {}
    prev = x
    for i in range(_last_needed(out_keys) + 1):
        name = FORWARD_SEQ[i]
        if name.find('conv') != -1:
            y = F.conv2d(prev, state[name + '.weight'], state[name + '.bias'], padding=1)
            outputs[OUT_SEQ[i]] = y * masks[OUT_SEQ[i]].to(y.dtype)
        else:
            b, c, h, w = prev.shape
            ho, wo = h // 2, w // 2
            win = prev[:, :, :2 * ho, :2 * wo].reshape(b, c, ho, 2, wo, 2).permute(0, 1, 2, 4, 3, 5).reshape(b, c, ho, wo, 4)
            pos = masks[OUT_SEQ[i]].to(torch.int64).clamp(max=3).unsqueeze(-1)
            outputs[OUT_SEQ[i]] = torch.gather(win, 4, pos).squeeze(-1)
        prev = outputs[OUT_SEQ[i]]
    return [outputs[k] for k in out_keys]


def loss_and_grad_masked(state, x, targets, masks, style_layers=STYLE_LAYERS, content_layers=CONTENT_LAYERS, weights=None):
    """loss_and_grad (utils.py:29-41) with the ReLU / pool decisions of `masks` instead of the ones x itself would take."""
    weights = weights if weights is not None else STYLE_WEIGHTS + CONTENT_WEIGHTS
    xg = x.detach().clone().requires_grad_(True)
    outs = vgg_forward_masked(state, xg, masks, list(style_layers) + list(content_layers))
    fns = [gram_mse_loss] * len(style_layers) + [F.mse_loss] * len(content_layers)
    ll = [weights[a] * fns[a](A, targets[a]) for a, A in enumerate(outs)]
    loss = sum(ll)
    loss.backward()
    return [float(l.detach()) for l in ll], float(loss.detach()), xg.grad.detach()


def optimize(state, content_image, style_image, optimized_image, max_iterations, full=True, trace=None, **kw):
    """utils.py:17-45 with torch.optim.LBFGS defaults. `optimized_image` is a leaf tensor updated in place.
    `trace`, if a list, receives (eval index, total loss) per closure evaluation."""
    targets = compute_targets(state, content_image, style_image, full=full,
                              **{k: v for k, v in kw.items() if k in ('style_layers', 'content_layers')})
    optimizer = torch.optim.LBFGS([optimized_image])
    iterations = [0]
    while iterations[0] < max_iterations:
        def closure():
            optimizer.zero_grad()
            ll = layer_losses(state, optimized_image, targets, full=full, **kw)
            loss = sum(ll)
            loss.backward()
            iterations[0] += 1
            if trace is not None:
                trace.append((iterations[0], float(loss)))
            return loss
        optimizer.step(closure)
    return optimized_image, iterations[0]


# ---------------------------------------------------------------------------------------------------------------------
# L-BFGS restated (torch/optim/lbfgs.py:333-537 of torch 2.11, line_search_fn=None), float64 bookkeeping optional.
# Used to check the device-side L-BFGS, whose two-loop recursion is algebraically rearranged.
# ---------------------------------------------------------------------------------------------------------------------
class LbfgsRestated:
    def __init__(self, lr=1.0, max_iter=20, max_eval=None, tolerance_grad=1e-7, tolerance_change=1e-9, history_size=100):
        self.lr, self.max_iter = lr, max_iter
        self.max_eval = max_eval if max_eval is not None else max_iter * 5 // 4
        self.tolerance_grad, self.tolerance_change, self.history_size = tolerance_grad, tolerance_change, history_size
        self.state = {'func_evals': 0, 'n_iter': 0}
        self.log = None      # set to a list to receive one dict per iteration: n_iter, d, t, H_diag, ys, accepted, hist, gtd, applied

    def step(self, x, closure):
        """x: flat tensor updated in place; closure() -> (loss float, flat grad tensor) evaluated at the current x."""
        st = self.state
        orig_loss, flat_grad = closure()
        loss = float(orig_loss)
        current_evals = 1
        st['func_evals'] += 1
        if float(flat_grad.abs().max()) <= self.tolerance_grad:
            return orig_loss
        d, t = st.get('d'), st.get('t')
        old_dirs, old_stps, ro = st.get('old_dirs'), st.get('old_stps'), st.get('ro')
        H_diag, prev_flat_grad, prev_loss = st.get('H_diag'), st.get('prev_flat_grad'), st.get('prev_loss')
        n_iter = 0
        while n_iter < self.max_iter:
            n_iter += 1
            st['n_iter'] += 1
            ys, accepted = None, False
            if st['n_iter'] == 1:
                d = flat_grad.neg()
                old_dirs, old_stps, ro = [], [], []
                H_diag = 1.0
            else:
                y = flat_grad.sub(prev_flat_grad)
                s = d.mul(t)
                ys = float(y.dot(s))
                accepted = ys > 1e-10
                if ys > 1e-10:
                    if len(old_dirs) == self.history_size:
                        old_dirs.pop(0); old_stps.pop(0); ro.pop(0)
                    old_dirs.append(y); old_stps.append(s); ro.append(1.0 / ys)
                    H_diag = ys / float(y.dot(y))
                num_old = len(old_dirs)
                al = [None] * num_old
                q = flat_grad.neg()
                for i in range(num_old - 1, -1, -1):
                    al[i] = float(old_stps[i].dot(q)) * ro[i]
                    q.add_(old_dirs[i], alpha=-al[i])
                d = r = torch.mul(q, H_diag)
                for i in range(num_old):
                    be_i = float(old_dirs[i].dot(r)) * ro[i]
                    r.add_(old_stps[i], alpha=al[i] - be_i)
            prev_flat_grad = flat_grad.clone()
            prev_loss = loss
            if st['n_iter'] == 1:
                t = min(1.0, 1.0 / float(flat_grad.abs().sum())) * self.lr
            else:
                t = self.lr
            gtd = float(flat_grad.dot(d))
            if self.log is not None:
                self.log.append(dict(n_iter=st['n_iter'], d=d.clone(), t=t, H_diag=H_diag, ys=ys, accepted=accepted,
                                     hist=len(old_dirs), gtd=gtd, applied=not (gtd > -self.tolerance_change)))
            if gtd > -self.tolerance_change:
                break
            x.add_(d, alpha=t)
            ls_func_evals = 0
            if n_iter != self.max_iter:
                l, flat_grad = closure()
                loss = float(l)
                ls_func_evals = 1
            current_evals += ls_func_evals
            st['func_evals'] += ls_func_evals
            if n_iter == self.max_iter:
                break
            if current_evals >= self.max_eval:
                break
            if float(flat_grad.abs().max()) <= self.tolerance_grad:
                break
            if float(d.mul(t).abs().max()) <= self.tolerance_change:
                break
            if abs(loss - prev_loss) < self.tolerance_change:
                break
        st.update(d=d, t=t, old_dirs=old_dirs, old_stps=old_stps, ro=ro, H_diag=H_diag,
                  prev_flat_grad=prev_flat_grad, prev_loss=prev_loss)
        return orig_loss


# ---------------------------------------------------------------------------------------------------------------------
# ImageTransform restated (image_transform.py:5-31) on PIL images
# ---------------------------------------------------------------------------------------------------------------------
class ImageTransform:
    def __init__(self, image_size, imagenet_mean):
        from torchvision import transforms
        self.preparation = transforms.Compose([
            transforms.Resize(image_size),                                 # transforms.Scale in the reference
            transforms.ToTensor(),
            transforms.Lambda(lambda x: x[torch.LongTensor([2, 1, 0])]),   # RGB -> BGR
            transforms.Normalize(mean=imagenet_mean, std=[1, 1, 1]),
            transforms.Lambda(lambda x: x.mul_(255)),
        ])
        self.post1 = transforms.Compose([
            transforms.Lambda(lambda x: x.mul_(1. / 255)),
            transforms.Normalize(mean=[(-1) * x for x in imagenet_mean], std=[1, 1, 1]),
            transforms.Lambda(lambda x: x[torch.LongTensor([2, 1, 0])]),
        ])
        self.post2 = transforms.Compose([transforms.ToPILImage()])

    def post_preparation(self, tensor):
        t = self.post1(tensor)
        t[t > 1] = 1
        t[t < 0] = 0
        return self.post2(t)
