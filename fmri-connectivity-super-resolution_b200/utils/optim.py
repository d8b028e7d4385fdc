"""Adam on the library's own kernel (SURVEY 8 f, row N3): ``torch.optim.Adam(model.parameters(), lr)`` of
training_gsr.ipynb:72 with every parameter tensor updated by one launch (``gsr_adam_step``) and the step counter
on the device -- no host synchronisation, capturable in ``GraphedTrainStep``'s CUDA graph."""
from __future__ import annotations

import ctypes as C

import torch

from .._lib import check, lib


class FlatAdam(torch.optim.Optimizer):
    """Same update rule and defaults as ``torch.optim.Adam`` (no amsgrad / maximize); state per parameter:
    ``exp_avg``, ``exp_avg_sq`` and the shared device scalar ``step``."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("FlatAdam: invalid hyper-parameter")
        # capturable: the step counter lives on the device (GraphedTrainStep checks the flag)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, capturable=True))
        self._tables = {}

    def _state(self, p, step):
        st = self.state[p]
        if not st:
            st["step"] = step
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()):
                    raise RuntimeError("FlatAdam: parameters and gradients must be contiguous CUDA float32 (no CPU path)")
            if "step_tensor" not in group:
                group["step_tensor"] = torch.zeros((), device=ps[0].device, dtype=torch.float32)
            if group["step_tensor"].device != ps[0].device:          # e.g. after load_state_dict(map_location="cpu")
                group["step_tensor"] = group["step_tensor"].to(ps[0].device, torch.float32)
            states = [self._state(p, group["step_tensor"]) for p in ps]
            for p, st in zip(ps, states):                             # loaded moments may sit on another device
                for k in ("exp_avg", "exp_avg_sq"):
                    if st[k].device != p.device or st[k].dtype != torch.float32 or not st[k].is_contiguous():
                        st[k] = st[k].to(p.device, torch.float32).contiguous()
            # The pointer tables are cached per set of (parameter, gradient, moment) buffers: every raw pointer a table
            # holds is part of its key, and the table keeps the tensors alive, so that a replaced state tensor
            # (load_state_dict) or a fresh .grad (zero_grad(set_to_none=True)) can never leave a stale pointer behind.
            key = (gi,) + tuple(x for p, st in zip(ps, states)
                                for x in (p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()))
            tab = self._tables.get(key)
            if tab is None:
                if len(self._tables) >= 8:
                    self._tables.pop(next(iter(self._tables)))
                n = len(ps)
                tab = self._tables[key] = (
                    (C.c_void_p * n)(*[p.data_ptr() for p in ps]), (C.c_void_p * n)(*[p.grad.data_ptr() for p in ps]),
                    (C.c_void_p * n)(*[s["exp_avg"].data_ptr() for s in states]),
                    (C.c_void_p * n)(*[s["exp_avg_sq"].data_ptr() for s in states]),
                    (C.c_longlong * n)(*[p.numel() for p in ps]), n,
                    [(p, p.grad, s["exp_avg"], s["exp_avg_sq"]) for p, s in zip(ps, states)])
            b1, b2 = group["betas"]
            check(lib().gsr_adam_step(tab[0], tab[1], tab[2], tab[3], tab[4], tab[5], group["step_tensor"].data_ptr(),
                                      float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                      float(group["weight_decay"]), torch.cuda.current_stream().cuda_stream),
                  "gsr_adam_step")
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables.clear()                      # the loaded state tensors are new buffers

    def __setstate__(self, state):
        super().__setstate__(state)
        self._tables = {}
