"""Python handle on ``ist_lbfgs`` (include/ist_b200.h): device-side torch.optim.LBFGS semantics for the fused closure.
Stands in for ``optim.LBFGS([optimized_image])`` + ``optimizer.step(closure)`` of IST/model/engine/utils.py:24,43."""
import ctypes

import torch

from . import _lib


class DeviceLBFGS:
    def __init__(self, plan, lr=1.0, max_iter=20, max_eval=None, tolerance_grad=1e-7, tolerance_change=1e-9, history_size=100,
                 test_objective=None):
        """`plan`: the Plan whose closure (loss + image gradient) is minimised. `test_objective=(a, b, c)` (float32 CUDA tensors
        [batch, n], `plan` None) runs the same optimiser on f(x) = sum 0.5 a (x - b)^2 + c cos(x) per frame instead — the
        parity tests use it to reach rejected curvature pairs and every tolerance exit of torch/optim/lbfgs.py."""
        self.lib = _lib.load()
        self.plan = plan                      # keeps the plan alive
        if max_eval is None:
            max_eval = max_iter * 5 // 4      # torch/optim/lbfgs.py default
        h = ctypes.c_void_p()
        if test_objective is not None:
            a, b, c = [t.contiguous() for t in test_objective]
            if plan is not None or not all(t.is_cuda and t.dtype == torch.float32 and t.shape == a.shape and t.dim() == 2 for t in (a, b, c)):
                raise _lib.IstError("test_objective needs plan=None and three float32 CUDA tensors [batch, n]")
            self._keep = (a, b, c)
            self.batch, self.n = int(a.shape[0]), int(a.shape[1])
            self.device = a.device
            with torch.cuda.device(self.device):
                _lib.check(self.lib.ist_lbfgs_create_test(ctypes.byref(h), self.batch, self.n, _lib.ptr(a), _lib.ptr(b), _lib.ptr(c),
                                                          int(history_size), int(max_iter), int(max_eval), float(lr),
                                                          float(tolerance_grad), float(tolerance_change)))
            self.n_losses = 0
        else:
            self.batch, self.n = plan.batch, 3 * plan.H * plan.W
            self.device = plan.device
            with torch.cuda.device(self.device):
                _lib.check(self.lib.ist_lbfgs_create(ctypes.byref(h), plan.h, int(history_size), int(max_iter), int(max_eval),
                                                     float(lr), float(tolerance_grad), float(tolerance_change)))
            self.n_losses = plan.n_style + plan.n_content
        self.h = h
        self._trace = None

    # ---- diagnostics (parity tests) ------------------------------------------------------------------------------------------
    def enable_trace(self, capacity):
        """Record every closure evaluation (see ist_lbfgs_set_trace). Must precede the first step()."""
        mk = lambda *shape, dt=torch.float32: torch.zeros(*shape, device=self.device, dtype=dt)
        tr = dict(x=mk(capacity, self.batch, self.n), g=mk(capacity, self.batch, self.n), d=mk(capacity, self.batch, self.n),
                  sc=mk(capacity, self.batch, 16, dt=torch.float64))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ist_lbfgs_set_trace(self.h, _lib.ptr(tr["x"]), _lib.ptr(tr["g"]), _lib.ptr(tr["d"]), _lib.ptr(tr["sc"], torch.float64),
                                                    int(capacity)))
        self._trace = tr

    TRACE_FIELDS = ("loss", "computed", "active", "n_iter", "hist_len", "head", "accepted", "H_diag", "t", "gtd", "ys", "yy",
                    "applied", "func_evals", "step_evals", "new_slot")

    def trace(self):
        """(count, x, g, d, scalars) of the records so far; scalars[e, frame] is a dict-like row ordered as TRACE_FIELDS."""
        n = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ist_lbfgs_trace_count(self.h, ctypes.byref(n)))
        tr = self._trace
        return n.value, tr["x"][: n.value], tr["g"][: n.value], tr["d"][: n.value], tr["sc"][: n.value]

    def frame_state(self, frame):
        v = [ctypes.c_int(0) for _ in range(5)]
        _lib.check(self.lib.ist_lbfgs_frame_state(self.h, int(frame), *[ctypes.byref(c) for c in v]))
        return dict(zip(("func_evals", "n_iter", "hist_len", "active", "step_evals"), [c.value for c in v]))

    def reset(self):
        """Forget all optimiser state (equivalent to constructing a new optim.LBFGS([x])); keeps buffers and the captured graph."""
        _lib.check(self.lib.ist_lbfgs_reset(self.h, _lib.stream_ptr(self.device)))

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
        if self.plan is not None:
            self.plan._check_x(x)
        elif tuple(x.shape) != (self.batch, self.n):
            raise _lib.IstError(f"optimiser is for x of shape {(self.batch, self.n)}, got {tuple(x.shape)}")
        evals = ctypes.c_int(0)
        loss = ctypes.c_float(0.0)
        _lib.check(self.lib.ist_lbfgs_step(self.h, ctypes.c_void_p(x.data_ptr()), ctypes.byref(evals), ctypes.byref(loss),
                                           _lib.stream_ptr(self.device)))
        return evals.value, loss.value

    def last_losses(self):
        """[batch, n_losses + 1] weighted layer losses (+ total) of the most recent closure evaluation."""
        out = (ctypes.c_float * (self.batch * (self.n_losses + 1)))()
        _lib.check(self.lib.ist_lbfgs_last_losses(self.h, out))
        return torch.tensor(list(out), dtype=torch.float32).view(self.batch, self.n_losses + 1)
