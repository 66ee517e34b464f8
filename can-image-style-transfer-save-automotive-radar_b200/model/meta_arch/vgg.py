"""VGG feature extractor with the reference's contract (IST/model/meta_arch/vgg.py:5-58), computed by the B200 plan.

Same constructor `(cfg, pool)`, same sub-module names (`conv{b}_{i}` -> state-dict keys `conv1_1.weight` ... as in
`vgg_conv.pth`), same `forward(input, out_keys) -> [Tensor]` ordered like `out_keys`, differentiable w.r.t. `input`.
The nn.Conv2d sub-modules only hold the parameters (so `.to(device)`, `load_state_dict`, `parameters()` behave as in
IST/main.py:25-32); the arithmetic runs in libist_b200.so. Weights are treated as frozen (main.py:31-32): no
weight-gradient kernel exists, and asking for one raises.
"""
import collections
import os

import torch
import torch.nn as nn

from ... import _lib
from ...plan import Plan, layer_table


class _VGGFeatures(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vgg, keys, x):
        plan = vgg._plan_for(x, keys)
        deepest = max(keys, key=lambda k: plan.out_index[k])
        plan.forward(x, deepest)
        ctx.plan, ctx.keys = plan, list(keys)
        ctx.x_ref = (x.data_ptr(), x._version)
        outs = tuple(plan.feature(k) for k in keys)
        return outs

    @staticmethod
    def backward(ctx, *grads):
        seeds = {}
        for k, g in zip(ctx.keys, grads):
            if g is None:
                continue
            seeds[k] = g if k not in seeds else seeds[k] + g
        if not seeds:
            return None, None, None
        if ctx.plan.last_forward_token is not ctx:
            raise _lib.IstError("VGG.forward was called again on this plan before backward; the plan holds one set of activations")
        return None, None, ctx.plan.backward(seeds)


class VGG(nn.Module):
    def __init__(self, cfg, pool='max'):
        super(VGG, self).__init__()
        self.cfg = cfg
        self.pool = pool
        vcfg = self.cfg.MODEL.VGG
        if len(vcfg.FORWARD_SEQ) != len(vcfg.OUT_SEQ):
            raise Exception("Forward and Output of layers of VGG must have the same length.")   # vgg.py:45-46
        if pool != 'max':
            # the reference defines pool layers only for pool == 'max' (vgg.py:20-22); anything else fails there with KeyError
            raise KeyError("pool layers are only defined for pool='max'")
        self.layers = {}
        for name, d in vcfg.CONV_LAYERS_DICT[0].items():
            conv = nn.Conv2d(in_channels=d['in_channels'], out_channels=d['out_channels'],
                             kernel_size=d['kernel'], padding=d['padding'])
            self.layers[name] = conv
            setattr(self, name, conv)           # registers the parameters under the reference's state-dict keys
        self.forward_seq = list(vcfg.FORWARD_SEQ)
        self.out_seq = list(vcfg.OUT_SEQ)
        # Plan cache, least recently used first. A plan pins every activation / gradient buffer of its size (0.35 GB at 512^2,
        # 5 GB at 2048^2) and, after optimize(), an L-BFGS history (0.63 GB / 10 GB) plus a captured graph; a directory of frames
        # with many aspect ratios or the 512 -> 1024 -> 2048 schedule would otherwise accumulate them until the device is full
        # (the reference frees everything per frame). IST_B200_MAX_PLANS overrides the bound.
        self._plans = collections.OrderedDict()
        self._max_plans = max(2, int(os.environ.get("IST_B200_MAX_PLANS", "6")))
        self._weights_token = None

    # -------------------------------------------------------------------------------------------------------------------
    def _token(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _plan_for(self, x, keys):
        """Plan cache keyed by (batch, H, W): built to the deepest layer ever requested, rebuilt deeper on demand."""
        if not x.is_cuda:
            raise _lib.IstError("the B200 VGG runs on CUDA tensors only (MODEL.DEVICE='cpu' is served by the oracle, not by this package)")
        for k in keys:
            if k not in self.out_seq:
                raise KeyError(k)                # same failure as vgg.py:58 for an unknown key
        b, c, h, w = x.shape
        need = max(self.out_seq.index(k) for k in keys)
        key = (int(b), int(h), int(w), x.device.index)
        plan = self._plans.get(key)
        if plan is None or plan.depth < need:
            if plan is not None:
                plan.close()
            table = layer_table(self.cfg.MODEL.VGG, self.out_seq[need])
            with torch.cuda.device(x.device):
                plan = Plan(table, b, h, w)
            plan.depth = need
            plan.weights_token = None
            plan.last_forward_token = None
            self._plans[key] = plan
            while len(self._plans) > self._max_plans:
                _, old = self._plans.popitem(last=False)      # least recently used; a later use of it raises (closed handle)
                old.close()
        self._plans.move_to_end(key)
        tok = self._token()
        if plan.weights_token != tok:
            plan.load_state_dict(self.state_dict())
            plan.weights_token = tok
        return plan

    def plan(self, batch, H, W, upto_key, device=None):
        """The cached plan for a given input size (used by engine.optimize for the fused closure)."""
        device = device if device is not None else next(self.parameters()).device
        dummy = torch.empty(batch, 3, H, W, device=device)
        return self._plan_for(dummy, [upto_key])

    def release_plans(self):
        for p in self._plans.values():
            p.close()
        self._plans = collections.OrderedDict()

    def forward(self, input, out_keys):
        if len(self.forward_seq) != len(self.out_seq):
            raise Exception("Forward and Output of layers of VGG must have the same length.")
        x = input.contiguous().float()
        keys = list(out_keys)
        if not keys:
            return []
        if torch.is_grad_enabled() and x.requires_grad:
            outs = _VGGFeatures.apply(self, tuple(keys), x)
            # the autograd node that owns the plan's activations (a later forward on the same plan invalidates it)
            self._plan_for(x, keys).last_forward_token = outs[0].grad_fn
            return list(outs)
        plan = self._plan_for(x, keys)
        deepest = max(keys, key=lambda k: plan.out_index[k])
        plan.forward(x, deepest)
        plan.last_forward_token = None
        return [plan.feature(k) for k in keys]
