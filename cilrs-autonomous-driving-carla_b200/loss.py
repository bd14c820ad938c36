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
