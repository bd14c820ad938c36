"""Fused loss kernels behind the reference's `CILRSLoss` interface (notebook/notebook.ipynb:504-527) and the
README/config MSE recipe (configs/train_config.json:30-32).

`CILRSLoss()(pred_controls, target_controls, pred_speed, target_speed) -> (total_loss, dict)` as in the notebook; the six
scalars come from ONE kernel and one device buffer (the reference does 6 `.item()` syncs per step; here the dict
values are 0-dim device tensors unless `as_float=True`, which does a single D2H copy).
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib

_NAMES = ("total", "control", "steer", "throttle", "brake", "speed")


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred_controls, target_controls, pred_speed, target_speed, mode, ws):
        b = pred_controls.shape[0]
        dev = pred_controls.device
        out = torch.empty(6, dtype=torch.float32, device=dev)
        dctrl = torch.empty(b, 3, dtype=torch.float32, device=dev)
        dspd = torch.empty(b, dtype=torch.float32, device=dev)
        _lib.call("cilrs_loss", pred_controls.contiguous(), pred_speed.contiguous(), target_controls.contiguous().float(),
                  target_speed.contiguous().float(), b, mode, ctypes.c_float(ws[0]), ctypes.c_float(ws[1]), ctypes.c_float(ws[2]),
                  ctypes.c_float(ws[3]), ctypes.c_float(1.0), out, dctrl, dspd, _lib.stream_ptr())
        ctx.save_for_backward(dctrl, dspd)
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, gtotal, _gout):
        dctrl, dspd = ctx.saved_tensors
        return dctrl * gtotal, None, dspd * gtotal, None, None, None


class CILRSLoss(nn.Module):
    """mode='l1': steer_w*L1(steer) + throttle_w*L1(throttle) + brake_w*L1(brake) + speed_w*MSE(speed)  (notebook recipe)
    mode='mse': MSE(controls) + speed_w*MSE(speed)                                                   (README recipe)"""

    def __init__(self, steer_w=5.0, throttle_w=1.0, brake_w=1.0, speed_w=0.5, mode="l1", as_float=True):
        super().__init__()
        if mode not in ("l1", "mse"):
            raise ValueError("mode must be 'l1' or 'mse'")
        self.steer_w, self.throttle_w, self.brake_w, self.speed_w = steer_w, throttle_w, brake_w, speed_w
        self.mode = mode
        self.as_float = as_float

    def forward(self, pred_controls, target_controls, pred_speed, target_speed):
        if not pred_controls.is_cuda:
            raise RuntimeError("cilrs_b200.CILRSLoss runs on CUDA tensors only")
        total, out = _LossFn.apply(pred_controls, target_controls, pred_speed, target_speed, 1 if self.mode == "l1" else 0,
                                   (self.steer_w, self.throttle_w, self.brake_w, self.speed_w))
        if self.as_float:
            vals = out.tolist()  # one D2H copy / sync instead of six
            return total, dict(zip(_NAMES, vals))
        return total, {n: out[i] for i, n in enumerate(_NAMES)}


CMD_NAMES = {0: "FOLLOW", 1: "LEFT", 2: "RIGHT", 3: "STRAIGHT"}   # notebook/notebook.ipynb:583


class Validator:
    """Device-side accumulator behind `validate()` (notebook/notebook.ipynb:563-585): per batch ONE kernel adds the six loss
    scalars and the per-command steer absolute errors / counts to a 16-double device buffer; `result()` does the only D2H copy
    of the pass. The reference reads 6 `.item()`s and up to 4 masked `.cpu()` tensors per batch."""

    def __init__(self, criterion, device="cuda"):
        self.criterion = criterion
        self.acc = torch.zeros(16, dtype=torch.float64, device=device)

    def reset(self):
        self.acc.zero_()

    def update(self, pred_controls, target_controls, pred_speed, target_speed, command):
        c = self.criterion
        _lib.call("cilrs_validate_accumulate", pred_controls.contiguous().float(), pred_speed.contiguous().float(),
                  target_controls.contiguous().float(), target_speed.contiguous().float(), command.contiguous(),
                  int(pred_controls.shape[0]), 1 if c.mode == "l1" else 0, ctypes.c_float(c.steer_w), ctypes.c_float(c.throttle_w),
                  ctypes.c_float(c.brake_w), ctypes.c_float(c.speed_w), self.acc, _lib.stream_ptr())

    def result(self):
        """({total, control, steer, throttle, brake, speed} averaged over batches, {FOLLOW, LEFT, RIGHT, STRAIGHT} mean steer
        error, nan where a command never occurred) — the pair the reference's validate() returns."""
        a = self.acc.tolist()
        n = max(a[14], 1.0)
        losses = {k: a[i] / n for i, k in enumerate(_NAMES)}
        cmd_avg = {CMD_NAMES[k]: (a[6 + k] / a[10 + k] if a[10 + k] > 0 else float("nan")) for k in range(4)}
        return losses, cmd_avg


def validate(model, loader, criterion, device):
    """Drop-in for the notebook's validate(model, loader, criterion, device) (notebook/notebook.ipynb:563-585): eval-mode
    forward over the loader, mean losses and per-command steer MAE — with no host synchronisation inside the loop."""
    model.eval()
    v = Validator(criterion, device)
    with torch.no_grad():
        for imgs, speeds, cmds, tgts in loader:
            imgs, speeds, cmds, tgts = (imgs.to(device, non_blocking=True), speeds.to(device, non_blocking=True),
                                        cmds.to(device, non_blocking=True), tgts.to(device, non_blocking=True))
            pred_ctrl, pred_spd = model(imgs, speeds, cmds)
            v.update(pred_ctrl, tgts, pred_spd, speeds, cmds)
    return v.result()
