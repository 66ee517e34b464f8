"""Python handle on an ``ist_plan`` (include/ist_b200.h): the VGG feature stack of the reference
(IST/model/meta_arch/vgg.py:5-58) for one (batch, H, W) with all buffers resident in HBM, plus the loss configuration
of ``StyleTransfer`` (IST/main.py:35-43) and the cached targets of ``optimize`` (IST/model/engine/utils.py:19-20).
"""
import ctypes

import torch

from . import _lib


def layer_table(cfg_vgg, upto_key=None):
    """[(kind, cin, cout, forward_name, out_name)] from cfg.MODEL.VGG (IST/config/defaults.py:22-61), optionally truncated
    after the layer whose OUT_SEQ name is `upto_key`."""
    convs = cfg_vgg.CONV_LAYERS_DICT[0]
    fseq, oseq = list(cfg_vgg.FORWARD_SEQ), list(cfg_vgg.OUT_SEQ)
    if len(fseq) != len(oseq):
        raise Exception("Forward and Output of layers of VGG must have the same length.")   # vgg.py:45-46
    out = []
    for f, o in zip(fseq, oseq):
        if f.find("conv") != -1:
            d = convs[f]
            if d["kernel"] != 3 or d["padding"] != 1:
                raise _lib.IstError(f"{f}: the B200 path implements kernel=3, padding=1 convolutions only")
            out.append((_lib.LAYER_CONV3X3_RELU, d["in_channels"], d["out_channels"], f, o))
        elif f.find("pool") != -1:
            out.append((_lib.LAYER_MAXPOOL2X2, 0, 0, f, o))
        if upto_key is not None and o == upto_key:
            break
    return out


class Plan:
    def __init__(self, layers, batch, H, W):
        self.lib = _lib.load()
        self.layers = list(layers)
        self.batch, self.H, self.W = int(batch), int(H), int(W)
        arr = (_lib.LayerDesc * len(layers))()
        for i, l in enumerate(layers):
            arr[i].kind, arr[i].cin, arr[i].cout = l[0], l[1], l[2]
        h = ctypes.c_void_p()
        _lib.check(self.lib.ist_plan_create(ctypes.byref(h), len(layers), arr, self.batch, self.H, self.W))
        self.h = h
        self.out_index = {l[4]: i for i, l in enumerate(layers)}
        self.conv_names = [l[3] for l in layers if l[0] == _lib.LAYER_CONV3X3_RELU]
        self.n_style = self.n_content = 0
        self.device = torch.device("cuda", torch.cuda.current_device())

    def close(self):
        opt = getattr(self, "_lbfgs", None)
        if opt is not None:
            opt.close()
            self._lbfgs = None
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.ist_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def nbytes(self):
        return int(self.lib.ist_plan_bytes(self.h))

    # ---- weights (vgg.load_state_dict, IST/main.py:30) ------------------------------------------------------------------
    def load_state_dict(self, state):
        for ci, name in enumerate(self.conv_names):
            w = state[name + ".weight"].detach().to(self.device, torch.float32).contiguous()
            b = state[name + ".bias"].detach().to(self.device, torch.float32).contiguous()
            _lib.check(self.lib.ist_plan_set_weights(self.h, ci, _lib.ptr(w), _lib.ptr(b), _lib.stream_ptr(self.device)))

    # ---- forward / features -------------------------------------------------------------------------------------------------
    def _check_x(self, x):
        if self.h is None:
            raise _lib.IstError("this plan was closed (evicted from the VGG module's plan cache or released); run VGG.forward again")
        if tuple(x.shape) != (self.batch, 3, self.H, self.W):
            raise _lib.IstError(f"plan is for input {(self.batch, 3, self.H, self.W)}, got {tuple(x.shape)}")

    def forward(self, x, upto_key):
        self._check_x(x)
        _lib.check(self.lib.ist_plan_forward(self.h, _lib.ptr(x), self.out_index[upto_key], _lib.stream_ptr(self.device)))

    def feature_shape(self, key):
        c, h, w = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(self.lib.ist_plan_feature_shape(self.h, self.out_index[key], ctypes.byref(c), ctypes.byref(h), ctypes.byref(w)))
        return c.value, h.value, w.value

    def feature(self, key):
        c, h, w = self.feature_shape(key)
        out = torch.empty(self.batch, c, h, w, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.ist_plan_get_feature(self.h, self.out_index[key], _lib.ptr(out), _lib.stream_ptr(self.device)))
        return out

    def pool_index(self, key):
        """uint8 [batch,C,h,w]: window position 0..3 each pooled element of pool layer `key` took its maximum from in the last
        forward (first maximum in row-major order), 4 where the pooled value is not positive."""
        c, h, w = self.feature_shape(key)
        out = torch.empty(self.batch, c, h, w, device=self.device, dtype=torch.uint8)
        _lib.check(self.lib.ist_plan_get_pool_index(self.h, self.out_index[key], _lib.ptr(out, torch.uint8), _lib.stream_ptr(self.device)))
        return out

    def masks(self, upto_key):
        """Every discontinuous decision of the last forward up to `upto_key`, in the oracle's `forward_masks` layout:
        {relu name: bool sign map [batch,C,h,w], pool name: uint8 argmax position}."""
        out = {}
        for l in self.layers:
            key = l[4]
            out[key] = self.feature(key) > 0 if l[0] == _lib.LAYER_CONV3X3_RELU else self.pool_index(key)
            if key == upto_key:
                break
        return out

    def gram(self, key):
        c, _, _ = self.feature_shape(key)
        out = torch.empty(self.batch, c, c, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.ist_plan_gram(self.h, self.out_index[key], _lib.ptr(out), _lib.stream_ptr(self.device)))
        return out

    # ---- losses -----------------------------------------------------------------------------------------------------------------
    def set_loss(self, style_keys, style_weights, content_keys, content_weights):
        ns, nc = len(style_keys), len(content_keys)
        sl = (ctypes.c_int * max(ns, 1))(*[self.out_index[k] for k in style_keys])
        sw = (ctypes.c_float * max(ns, 1))(*[float(w) for w in style_weights])
        cl = (ctypes.c_int * max(nc, 1))(*[self.out_index[k] for k in content_keys])
        cw = (ctypes.c_float * max(nc, 1))(*[float(w) for w in content_weights])
        _lib.check(self.lib.ist_plan_set_loss(self.h, ns, sl, sw, nc, cl, cw))
        self.n_style, self.n_content = ns, nc
        self.style_keys, self.content_keys = list(style_keys), list(content_keys)

    def set_style_target(self, slot, gram):
        g = gram.detach().to(self.device, torch.float32).contiguous()
        g = g[0] if g.dim() == 3 else g
        _lib.check(self.lib.ist_plan_set_style_target(self.h, slot, _lib.ptr(g.contiguous()), _lib.stream_ptr(self.device)))

    def capture_content_target(self, slot):
        _lib.check(self.lib.ist_plan_capture_content_target(self.h, slot, _lib.stream_ptr(self.device)))

    def loss_and_grad(self, x, grad=None, losses=None):
        """closure body (utils.py:29-41): returns (losses [batch, n+1], grad [batch,3,H,W]); last loss column = total."""
        self._check_x(x)
        if grad is None:
            grad = torch.empty_like(x)
        if losses is None:
            losses = torch.empty(self.batch, self.n_style + self.n_content + 1, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.ist_plan_loss_and_grad(self.h, _lib.ptr(x), _lib.ptr(grad), _lib.ptr(losses), _lib.stream_ptr(self.device)))
        return losses, grad

    def backward(self, seeds):
        """seeds: {out_key: dL/d(feature) fp32 NCHW}; returns dL/dx. Needs a preceding forward() on the same x."""
        keys = list(seeds.keys())
        ts = [seeds[k].detach().to(self.device, torch.float32).contiguous() for k in keys]
        idx = (ctypes.c_int * len(keys))(*[self.out_index[k] for k in keys])
        ptrs = (ctypes.c_void_p * len(keys))(*[t.data_ptr() for t in ts])
        grad = torch.empty(self.batch, 3, self.H, self.W, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.ist_plan_backward(self.h, len(keys), idx, ptrs, _lib.ptr(grad), _lib.stream_ptr(self.device)))
        return grad
