"""GramMatrix with the reference's semantics (IST/model/meta_arch/gram_matrix.py:5-11): G = F F^T / (h*w) per batch
element, computed by the tcgen05 SYRK kernel; backward dF = (dG + dG^T) F / (h*w) by the tensor-core GEMM kernel."""
import torch
import torch.nn as nn

from ... import _lib


class _GramFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous().float()
        b, c, h, w = x.shape
        lib = _lib.load()
        g = torch.empty(b, c, c, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):          # the per-op entry points work on the current device
            _lib.check(lib.ist_op_gram(_lib.ptr(x), _lib.ptr(g), b, c, h, w, _lib.stream_ptr(x.device)))
        ctx.save_for_backward(x)
        return g

    @staticmethod
    def backward(ctx, dg):
        (x,) = ctx.saved_tensors
        b, c, h, w = x.shape
        dx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().ist_op_gram_bwd(_lib.ptr(x), _lib.ptr(dg.contiguous().float()), _lib.ptr(dx), b, c, h, w,
                                                   _lib.stream_ptr(x.device)))
        return dx


class GramMatrix(nn.Module):
    def forward(self, input):
        if input.shape[1] % 64 != 0 or (input.shape[1] > 64 and input.shape[1] % 128 != 0):
            raise _lib.IstError("the B200 Gram kernel needs C == 64 or C % 128 == 0 (VGG feature widths)")
        return _GramFn.apply(input)
