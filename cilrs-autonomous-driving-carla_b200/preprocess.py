"""Host-side mirror of the reference's inference call sites (model/autonomous_drive.py:897-920):

  preprocess_image(image)                      -> f32 [1,3,88,200] on the device           (:897-902)
  predict_controls(model, image, kmh, cmd)     -> (steer, throttle, brake, pred_speed_kmh) (:908-920)
  InferenceSession                             -> the same, with the whole frame->controls path (H2D of the raw uint8
                                                  frame, K0 resize+normalise, CILRS forward, D2H of 4 floats) captured in
                                                  one CUDA graph for batch-1 latency.
"""
import numpy as np
import torch

from . import ops

IMG_WIDTH, IMG_HEIGHT = 200, 88
SPEED_NORM_FACTOR = 90.0


def _as_u8_tensor(image):
    if isinstance(image, np.ndarray):
        if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] not in (3, 4):
            raise ValueError("image must be uint8 [H,W,3|4]")
        return torch.from_numpy(np.ascontiguousarray(image))
    if isinstance(image, torch.Tensor) and image.dtype == torch.uint8 and image.dim() == 3:
        return image
    raise TypeError("image must be a uint8 ndarray / tensor [H,W,C]")


def preprocess_image(image, device="cuda", reverse=False):
    """cv2.resize(image,(200,88)) -> /255 -> CHW -> Normalize -> unsqueeze(0).to(device), computed on the device.
    `image`: uint8 [H,W,3|4] (RGB, or BGR(A) with reverse=True)."""
    t = _as_u8_tensor(image)
    t = t.to(device, non_blocking=True) if not t.is_cuda else t
    return ops.preprocess(t.unsqueeze(0), reverse=reverse, dst_hw=(IMG_HEIGHT, IMG_WIDTH))["f32"]


def predict_controls(model, image, speed_kmh, command_idx, device="cuda"):
    img = preprocess_image(image, device)
    speed = torch.tensor([min(speed_kmh / SPEED_NORM_FACTOR, 1.0)], dtype=torch.float32).to(device)
    command = torch.tensor([command_idx], dtype=torch.long).to(device)
    with torch.no_grad():
        controls, pred_speed = model(img, speed, command)
    out = torch.cat([controls.reshape(-1), pred_speed.reshape(-1)]).tolist()  # one D2H instead of four
    return out[0], out[1], out[2], out[3] * SPEED_NORM_FACTOR


class InferenceSession:
    """Batched frame -> controls path with static buffers and (optionally) one CUDA graph.

    session = InferenceSession(model, batch=1); steer, thr, brk, kmh = session.predict(frame_u8, speed_kmh, command)
    """

    def __init__(self, model, batch=1, src_hw=(600, 800), src_c=3, reverse=False, use_graph=True):
        if model.training:
            raise RuntimeError("InferenceSession needs model.eval()")
        self.model = model
        self.batch = batch
        self.reverse = reverse
        dev = model.flat_parameters().device
        self.h_frames = torch.empty(batch, src_hw[0], src_hw[1], src_c, dtype=torch.uint8).pin_memory()
        self.h_speed = torch.empty(batch, dtype=torch.float32).pin_memory()
        self.h_command = torch.empty(batch, dtype=torch.long).pin_memory()
        self.h_out = torch.empty(batch, 4, dtype=torch.float32).pin_memory()
        self.d_frames = torch.empty_like(self.h_frames, device=dev)
        self.d_speed = torch.empty(batch, dtype=torch.float32, device=dev)
        self.d_command = torch.zeros(batch, dtype=torch.long, device=dev)
        self.d_out = torch.empty(batch, 4, dtype=torch.float32, device=dev)
        self.s2d = model.input_s2d_buffer(batch)
        self.graph = None
        self.stream = torch.cuda.Stream(device=dev)
        with torch.no_grad():
            self.d_speed.zero_()
            self.d_frames.zero_()
            self._body()  # warm-up: builds the plan, packs weights, folds BN
            torch.cuda.synchronize(dev)
            if use_graph:
                with torch.cuda.stream(self.stream):
                    self._body()
                self.stream.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.stream):
                    self._body()
                self.graph = g

    def _body(self):
        from . import _lib
        b = self.batch
        _lib.call("cilrs_preprocess_u8", self.d_frames, b, self.d_frames.shape[1], self.d_frames.shape[2], self.d_frames.shape[3],
                  int(self.reverse), IMG_HEIGHT, IMG_WIDTH, None, None, self.s2d, _lib.stream_ptr())
        controls, pred_speed = self.model._launch_forward(None, self.d_speed, self.d_command, keep=False, s2d=self.s2d)
        self.d_out[:, :3].copy_(controls)
        self.d_out[:, 3].copy_(pred_speed)

    @torch.no_grad()
    def run(self):
        """h_frames / h_speed / h_command (pinned) -> h_out (pinned): H2D, kernels, D2H on the session stream, then sync."""
        with torch.cuda.stream(self.stream):
            self.d_frames.copy_(self.h_frames, non_blocking=True)
            self.d_speed.copy_(self.h_speed, non_blocking=True)
            self.d_command.copy_(self.h_command, non_blocking=True)
            if self.graph is not None:
                self.graph.replay()
            else:
                self._body()
            self.h_out.copy_(self.d_out, non_blocking=True)
        self.stream.synchronize()
        return self.h_out

    def predict(self, image, speed_kmh, command_idx):
        self.h_frames[0].copy_(_as_u8_tensor(image))
        self.h_speed[0] = min(speed_kmh / SPEED_NORM_FACTOR, 1.0)
        self.h_command[0] = int(command_idx)
        o = self.run()[0]
        return float(o[0]), float(o[1]), float(o[2]), float(o[3]) * SPEED_NORM_FACTOR
