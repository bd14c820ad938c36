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
        model._ensure(batch)
        self.h_frames = torch.empty(batch, src_hw[0], src_hw[1], src_c, dtype=torch.uint8).pin_memory()
        self.h_speed = torch.empty(batch, dtype=torch.float32).pin_memory()
        self.h_command = torch.empty(batch, dtype=torch.long).pin_memory()
        self.h_out = torch.empty(batch, 4, dtype=torch.float32).pin_memory()
        self.d_frames = torch.empty_like(self.h_frames, device=dev)
        self.d_speed = torch.empty(batch, dtype=torch.float32, device=dev)
        self.d_command = torch.zeros(batch, dtype=torch.long, device=dev)
        self.d_out = torch.empty(batch, 4, dtype=torch.float32, device=dev)
        self.s2d = model.input_s2d_buffer(batch)
        self.err = model.error_flag()
        self.h_err = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._plan_gen = model._plan_gen
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
        if self.model._plan_gen != self._plan_gen:
            raise RuntimeError("cilrs_b200.InferenceSession: the model's plan was rebuilt (a larger batch went through the module, "
                               "or it moved device) after this session was created; create a new InferenceSession")
        with torch.cuda.stream(self.stream):
            self.d_frames.copy_(self.h_frames, non_blocking=True)
            self.d_speed.copy_(self.h_speed, non_blocking=True)
            self.d_command.copy_(self.h_command, non_blocking=True)
            if self.graph is not None:
                self.graph.replay()
            else:
                self._body()
            self.h_out.copy_(self.d_out, non_blocking=True)
            self.h_err.copy_(self.err, non_blocking=True)
        self.stream.synchronize()
        if int(self.h_err[0]) != 0:   # the reference's gather(0, command) raises on an out-of-range command
            self.err.zero_()
            raise IndexError("cilrs_b200: a command index outside [0, 4) reached the model")
        return self.h_out

    def predict(self, image, speed_kmh, command_idx):
        self.h_frames[0].copy_(_as_u8_tensor(image))
        self.h_speed[0] = min(speed_kmh / SPEED_NORM_FACTOR, 1.0)
        self.h_command[0] = int(command_idx)
        o = self.run()[0]
        return float(o[0]), float(o[1]), float(o[2]), float(o[3]) * SPEED_NORM_FACTOR


class ShardedInference:
    """BASELINE.json configs[4]: batched CILRS inference (multi-agent rollout) sharded across the GPUs of one box. The batch
    dimension shards with no data-path collective: rank r owns frames [r*per_rank, (r+1)*per_rank); every rank runs its own
    `InferenceSession(batch=per_rank)` (one CUDA graph: H2D, K0, forward, D2H). `gather()` collects the [B,4] results on every
    rank (16 bytes per frame) when one host wants them; it is the only communication. World size 1 works without
    torch.distributed."""

    def __init__(self, model, global_batch, src_hw=(600, 800), src_c=3, reverse=False, process_group=None, use_graph=True):
        import torch.distributed as dist
        self.pg = process_group
        self.world, self.rank = 1, 0
        if dist.is_available() and dist.is_initialized():
            self.world, self.rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        if global_batch % self.world:
            raise ValueError("global_batch must be a multiple of the world size")
        self.global_batch = global_batch
        self.per_rank = global_batch // self.world
        self.session = InferenceSession(model, batch=self.per_rank, src_hw=src_hw, src_c=src_c, reverse=reverse, use_graph=use_graph)

    def shard(self):
        """(lo, hi) global frame indices this rank owns."""
        return self.rank * self.per_rank, (self.rank + 1) * self.per_rank

    def predict_batch(self, frames_u8, speed_kmh, command_idx):
        """frames_u8 uint8 [per_rank,H,W,C], speed_kmh float [per_rank], command_idx int [per_rank] (this rank's shard, host
        tensors / arrays) -> pinned float [per_rank,4] = (steer, throttle, brake, speed_kmh)."""
        s = self.session
        s.h_frames.copy_(torch.as_tensor(frames_u8))
        s.h_speed.copy_(torch.clamp(torch.as_tensor(speed_kmh, dtype=torch.float32) / SPEED_NORM_FACTOR, max=1.0))
        s.h_command.copy_(torch.as_tensor(command_idx, dtype=torch.long))
        out = s.run()
        out[:, 3] *= SPEED_NORM_FACTOR
        return out

    def gather(self, local_out):
        """[per_rank,4] of every rank -> [global_batch,4] on every rank (a 16 B/frame all_gather; identity at world size 1)."""
        if self.world == 1:
            return local_out.clone()
        import torch.distributed as dist
        dev = self.session.d_out.device
        full = torch.empty(self.global_batch, 4, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(full, local_out.to(dev, non_blocking=True), group=self.pg)
        return full.cpu()
