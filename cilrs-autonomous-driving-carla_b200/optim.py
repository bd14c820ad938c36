"""Fused Adam over the model's flat parameter arena.

`FusedAdam(model.parameters(), lr, weight_decay=...)` is a `torch.optim.Optimizer` with torch.optim.Adam's semantics
(L2-coupled weight decay, bias correction, eps outside the sqrt — notebook/notebook.ipynb:533-534,555) and state layout
(`state[p] = {step, exp_avg, exp_avg_sq}`, so `optimizer.state_dict()` stays interchangeable), but one kernel launch
updates every parameter: p, g, m, v are single contiguous fp32 buffers.
"""
import ctypes

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, model=None):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam: one parameter group (the whole model) is supported")
        if model is None:
            raise ValueError("FusedAdam needs model=<cilrs_b200.CILRS> (it updates the model's flat arena)")
        self.model = model
        plist = list(model.parameters())
        if len(plist) != len(params) or any(a is not b for a, b in zip(plist, params)):
            raise ValueError("FusedAdam must be given exactly model.parameters()")
        flat = model.flat_parameters()
        self._m = torch.zeros_like(flat)
        self._v = torch.zeros_like(flat)
        self._step = 0
        self._step_dev = torch.zeros(1, dtype=torch.long, device=flat.device) if flat.is_cuda else None
        for p, mv, vv in zip(plist, model._views(self._m), model._views(self._v)):
            self.state[p] = {"step": torch.tensor(0.0), "exp_avg": mv, "exp_avg_sq": vv}

    def _gather_grads(self):
        """Gradients normally already live in the model's flat arena (autograd hands out views of it); anything else
        (e.g. a clone made by grad accumulation) is copied in."""
        model = self.model
        g = model.flat_gradients()
        for p, view in zip(model.parameters(), model._views(g)):
            if p.grad is None:
                view.zero_()
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
        return g

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0, grad_scale_dev=None, grads_in_arena=False):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grp = self.param_groups[0]
        model = self.model
        flat = model.flat_parameters()
        if self._m.device != flat.device or self._m.data_ptr() == 0:
            raise RuntimeError("FusedAdam: the model moved to another device after the optimizer was created")
        g = model.flat_gradients() if grads_in_arena else self._gather_grads()
        self._step += 1
        if self._step_dev is None:
            raise RuntimeError("FusedAdam: create the optimizer after model.to('cuda')")
        _lib.call("cilrs_adam_step", flat, g, self._m, self._v, ctypes.c_longlong(flat.numel()), ctypes.c_float(grp["lr"]),
                  ctypes.c_float(grp["betas"][0]), ctypes.c_float(grp["betas"][1]), ctypes.c_float(grp["eps"]),
                  ctypes.c_float(grp["weight_decay"]), ctypes.c_longlong(0), self._step_dev, ctypes.c_float(grad_scale),
                  grad_scale_dev, _lib.stream_ptr())
        model.mark_parameters_changed()
        for st in self.state.values():
            st["step"] = torch.tensor(float(self._step))
        return loss
