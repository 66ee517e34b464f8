"""Python handle on ``ist_lbfgs`` (include/ist_b200.h): device-side torch.optim.LBFGS semantics for the fused closure.
Stands in for ``optim.LBFGS([optimized_image])`` + ``optimizer.step(closure)`` of IST/model/engine/utils.py:24,43."""
import ctypes

import torch

from . import _lib


class DeviceLBFGS:
    def __init__(self, plan, lr=1.0, max_iter=20, max_eval=None, tolerance_grad=1e-7, tolerance_change=1e-9, history_size=100):
        self.lib = _lib.load()
        self.plan = plan                      # keeps the plan alive
        if max_eval is None:
            max_eval = max_iter * 5 // 4      # torch/optim/lbfgs.py default
        h = ctypes.c_void_p()
        _lib.check(self.lib.ist_lbfgs_create(ctypes.byref(h), plan.h, int(history_size), int(max_iter), int(max_eval),
                                             float(lr), float(tolerance_grad), float(tolerance_change)))
        self.h = h
        self.n_losses = plan.n_style + plan.n_content

    def reset(self):
        """Forget all optimiser state (equivalent to constructing a new optim.LBFGS([x])); keeps buffers and the captured graph."""
        _lib.check(self.lib.ist_lbfgs_reset(self.h, _lib.stream_ptr()))

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.ist_lbfgs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, x):
        """One optimizer.step(closure): returns (closure evaluations performed, loss of the first evaluation)."""
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
            raise _lib.IstError("optimized image must be a contiguous float32 CUDA tensor")
        self.plan._check_x(x)
        evals = ctypes.c_int(0)
        loss = ctypes.c_float(0.0)
        _lib.check(self.lib.ist_lbfgs_step(self.h, ctypes.c_void_p(x.data_ptr()), ctypes.byref(evals), ctypes.byref(loss),
                                           _lib.stream_ptr()))
        return evals.value, loss.value

    def last_losses(self):
        """[batch, n_losses + 1] weighted layer losses (+ total) of the most recent closure evaluation."""
        out = (ctypes.c_float * (self.plan.batch * (self.n_losses + 1)))()
        _lib.check(self.lib.ist_lbfgs_last_losses(self.h, out))
        return torch.tensor(list(out), dtype=torch.float32).view(self.plan.batch, self.n_losses + 1)
