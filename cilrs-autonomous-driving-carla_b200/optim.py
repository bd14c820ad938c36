"""Fused Adam over the model's flat parameter arena.

`FusedAdam(model.parameters(), lr, weight_decay=...)` is a `torch.optim.Optimizer` with torch.optim.Adam's semantics
(L2-coupled weight decay, bias correction, eps outside the sqrt — notebook/notebook.ipynb:533-534,555) and state layout
(`state[p] = {step, exp_avg, exp_avg_sq}`, so `optimizer.state_dict()` / `load_state_dict()` stay interchangeable with
torch.optim.Adam's), but one kernel launch updates every parameter: p, g, m, v are single contiguous fp32 buffers.

Every hyper-parameter the kernel uses (lr, betas, eps, weight_decay, the 1/world gradient scale) and the step number live in
DEVICE memory, so a CUDA graph captured around `step()` keeps following `param_groups[0]["lr"]` — the reference halves the rate
every 8 epochs with `StepLR` (notebook/notebook.ipynb:535-536,604) — and resumes correctly from a loaded optimizer state.
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
        if not flat.is_cuda:
            raise RuntimeError("FusedAdam: create the optimizer after model.to('cuda') (there is no CPU path)")
        self._m = torch.zeros_like(flat)
        self._v = torch.zeros_like(flat)
        self._step = 0                      # host mirror of the device step counter
        self._step_dev = torch.zeros(1, dtype=torch.long, device=flat.device)
        self._hyper = torch.zeros(8, dtype=torch.float32, device=flat.device)
        self._hyper_host = None             # the values last uploaded
        self._point_state_at_views()

    # ------------------------------------------------------------------------------------------
    def _point_state_at_views(self):
        model = self.model
        for p, mv, vv in zip(model.parameters(), model._views(self._m), model._views(self._v)):
            self.state[p] = {"step": torch.tensor(float(self._step)), "exp_avg": mv, "exp_avg_sq": vv}

    def _sync_hyper(self, grad_scale=1.0):
        """Upload (lr, betas, eps, weight_decay, grad_scale) when they differ from what the device holds. Stream-ordered, so a
        change made between two graph replays applies to the later one."""
        grp = self.param_groups[0]
        vals = (float(grp["lr"]), float(grp["betas"][0]), float(grp["betas"][1]), float(grp["eps"]), float(grp["weight_decay"]),
                float(grad_scale), 0.0, 0.0)
        if vals != self._hyper_host:
            staging = torch.tensor(vals, dtype=torch.float32).pin_memory()
            self._hyper.copy_(staging, non_blocking=True)
            self._staging = staging  # keep the pinned buffer alive until the copy has run
            self._hyper_host = vals

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
    def step(self, closure=None, grad_scale=1.0, grad_scale_dev=None, grads_in_arena=False, grads_bf16=None, zero_grad=False,
             arena_range=None, advance=True):
        """grad_scale / grad_scale_dev: host / device factors applied to the gradient first (1/world, clip coefficient).
        grads_bf16: bf16 gradient buffer (arena layout) used instead of the fp32 arena. zero_grad: also clear the fp32 arena
        (the next step's optimizer.zero_grad()). arena_range=(lo, hi): update only that element range of the arena (multiples of
        16); advance=False: a further range of the SAME optimizer step (the step counter is not advanced again)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        model = self.model
        flat = model.flat_parameters()
        if self._m.device != flat.device:
            raise RuntimeError("FusedAdam: the model moved to another device after the optimizer was created")
        g = model.flat_gradients() if (grads_in_arena or grads_bf16 is not None) else self._gather_grads()
        self._sync_hyper(grad_scale)
        if advance:
            self._step += 1
        lo, hi = arena_range if arena_range is not None else (0, flat.numel())
        if lo % 16 or hi % 16 or not (0 <= lo <= hi <= flat.numel()):
            raise ValueError("arena_range must be 16-element aligned and inside the arena")
        _lib.call("cilrs_adam_step_ex", flat[lo:hi], g[lo:hi], None if grads_bf16 is None else grads_bf16[lo:hi], self._m[lo:hi],
                  self._v[lo:hi], ctypes.c_longlong(hi - lo), self._hyper, self._step_dev, grad_scale_dev,
                  int(bool(zero_grad)) | (0 if advance else 2), _lib.stream_ptr())
        model.mark_parameters_changed()
        return loss

    # ------------------------------------------------------------------------------------------
    # checkpoint interchange with torch.optim.Adam (notebook/notebook.ipynb:642-646 saves optimizer.state_dict())
    # ------------------------------------------------------------------------------------------
    def state_dict(self):
        step = torch.tensor(float(self._step))
        for st in self.state.values():
            st["step"] = step.clone()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        """Accepts FusedAdam's own and torch.optim.Adam's state dicts: the loaded moments are copied into the flat m / v arenas
        the kernel reads, and the device step counter is set from the loaded step."""
        super().load_state_dict(state_dict)
        model = self.model
        plist = list(model.parameters())
        steps = set()
        with torch.no_grad():
            for p, mv, vv in zip(plist, model._views(self._m), model._views(self._v)):
                st = self.state.get(p)
                if not st:
                    mv.zero_()
                    vv.zero_()
                    steps.add(0)
                    continue
                mv.copy_(st["exp_avg"])
                vv.copy_(st["exp_avg_sq"])
                steps.add(int(float(st["step"])))
        if len(steps) != 1:
            raise ValueError("FusedAdam.load_state_dict: parameters have different step counts (%s); one shared step is "
                             "supported (the whole model is one parameter group)" % sorted(steps))
        self._step = steps.pop()
        self._step_dev.fill_(self._step)
        self._hyper_host = None  # param_groups may have been replaced: upload again
        self._point_state_at_views()
