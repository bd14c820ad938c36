"""The fused training step (reference: train_one_epoch body, notebook/notebook.ipynb:545-555) and its data-parallel form.

    H2D -> [K0 normalise] -> forward -> loss -> zero_grad -> backward -> [allreduce] -> [clip] -> Adam -> repack

Everything after the H2D copies is device work launched through the C-ABI with no host synchronisation; with
`use_graph=True` the whole step (with N > 1 including the NCCL allreduces) is one CUDA-graph replay. Data parallelism shards the batch across ranks
(one process per GPU); the flat gradient arena is all-reduced over NCCL in five ranges that complete back-to-front
(heads+layer4, layer3, layer2, layer1, stem), each range's allreduce overlapping the backward of the next one.
BatchNorm statistics stay per rank (plain-DDP semantics; the reference has no SyncBN).
"""
import ctypes

import torch

from . import _lib
from .ddp import allreduce_ranges, backward_part_ranges, broadcast_parameters
from .model import MODE_TRAIN
from .optim import FusedAdam


class FusedTrainer:
    def __init__(self, model, batch, lr=2e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, loss="mse", steer_w=5.0,
                 throttle_w=1.0, brake_w=1.0, speed_w=0.05, grad_clip=0.0, process_group=None, use_graph=False, frames="f32",
                 async_parts=False, overlap_allreduce="first"):
        # overlap_allreduce: True = every backward part's range as soon as it is complete; False = one allreduce after the
        # backward; "first" = only the first part (heads + layer4, 64 % of the bytes) overlapped - it runs under layer3's
        # backward, whose 98-CTA conv grids leave SMs free for NCCL's CTAs - and one allreduce for the rest at the end
        if frames not in ("f32", "u8"):
            raise ValueError("frames: 'f32' ([B,3,88,200] normalised, as the reference's loader yields) or 'u8' ([B,88,200,3])")
        self.model = model
        self.batch = batch
        self.opt = FusedAdam(model.parameters(), lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, model=model)
        self.loss_mode = 1 if loss == "l1" else 0
        self.ws = (steer_w, throttle_w, brake_w, speed_w)
        self.grad_clip = float(grad_clip)
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.frames = frames
        dev = model.flat_parameters().device
        self.dev = dev
        if use_graph and model.dropout > 0:
            raise ValueError("use_graph replays one dropout mask; train with dropout through the eager step")
        model.train()
        model._ensure(batch)
        if self.world > 1:
            broadcast_parameters(model, 0, process_group)
        self.d_image = torch.zeros(batch, 3, 88, 200, dtype=torch.float32, device=dev) if frames == "f32" else None
        self.d_frames = torch.zeros(batch, 88, 200, 3, dtype=torch.uint8, device=dev) if frames == "u8" else None
        self.d_speed = torch.zeros(batch, dtype=torch.float32, device=dev)
        self.d_command = torch.zeros(batch, dtype=torch.long, device=dev)
        self.d_targets = torch.zeros(batch, 3, dtype=torch.float32, device=dev)
        self.controls = torch.zeros(batch, 3, dtype=torch.float32, device=dev)
        self.pred_speed = torch.zeros(batch, dtype=torch.float32, device=dev)
        self.dcontrols = torch.zeros(batch, 3, dtype=torch.float32, device=dev)
        self.dspeed = torch.zeros(batch, dtype=torch.float32, device=dev)
        self.loss6 = torch.zeros(6, dtype=torch.float32, device=dev)
        self.norm_ws = torch.zeros(1024, dtype=torch.float64, device=dev)
        self.norm_cnt = torch.zeros(4, dtype=torch.int32, device=dev)
        self.norm_out = torch.zeros(2, dtype=torch.float32, device=dev)
        self.s2d = model.input_s2d_buffer(batch)
        self.part_ranges = backward_part_ranges(model)
        # async_parts: enqueue each part's allreduce behind the model's gradient stream instead of joining that stream into the
        # caller's after every part (cilrs_model_backward_part_async). Measured on 2 GPUs: 3.38 ms/step when the host reads the
        # loss every step (3.52 joined) but 3.63 ms for back-to-back graph replays (3.41 joined) - hence off by default.
        self.async_parts = bool(async_parts)
        # overlap_allreduce=False: one allreduce of the whole gradient arena after the backward (no NCCL CTAs competing with the
        # convolutions for SMs, but the collective is fully exposed)
        self.overlap_allreduce = overlap_allreduce if overlap_allreduce == "first" else bool(overlap_allreduce)
        self.graph = None
        self.kernel_launches = None
        self.graph_error = None
        if use_graph:
            if self.world > 1:
                # the NCCL allreduces are captured into the graph with the kernels (PyTorch records them on the process
                # group's stream, forked from / joined into the capturing stream); any failure falls back to eager launches
                try:
                    self._capture()
                except Exception as ex:  # pragma: no cover - depends on the NCCL / driver combination
                    self.graph = None
                    self.graph_error = repr(ex)
                    torch.cuda.synchronize(self.dev)
            else:
                self._capture()

    # ------------------------------------------------------------------------------------------
    def _device_step(self):
        m = self.model
        b = self.batch
        sp = _lib.stream_ptr()
        if self.frames == "u8":
            _lib.call("cilrs_preprocess_u8", self.d_frames, b, 88, 200, 3, 0, 88, 200, None, None, self.s2d, sp)
            img, s2d = None, self.s2d
        else:
            img, s2d = self.d_image, None
        m._seed = (m._seed * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        _lib.call("cilrs_model_forward", m._handle, b, MODE_TRAIN, img, s2d, self.d_speed, self.d_command, self.controls,
                  self.pred_speed, 1, 1, ctypes.c_float(m.dropout), ctypes.c_ulonglong(m._seed), sp)
        _lib.call("cilrs_loss", self.controls, self.pred_speed, self.d_targets, self.d_speed, b, self.loss_mode,
                  ctypes.c_float(self.ws[0]), ctypes.c_float(self.ws[1]), ctypes.c_float(self.ws[2]), ctypes.c_float(self.ws[3]),
                  ctypes.c_float(1.0), self.loss6, self.dcontrols, self.dspeed, sp)
        g = m.flat_gradients()
        g.zero_()
        works = []
        if self.world > 1 and self.overlap_allreduce == "first":
            args = (self.dcontrols, self.dspeed, self.d_speed, self.d_command, ctypes.c_float(m.dropout), sp)
            _lib.call("cilrs_model_backward", m._handle, b, MODE_TRAIN, 0, *args)
            works.extend(allreduce_ranges(g, [self.part_ranges[0]], self.pg))
            for part in range(1, 5):
                _lib.call("cilrs_model_backward_part_async", m._handle, b, MODE_TRAIN, part, *args)
            _lib.call("cilrs_model_backward_join", m._handle, sp)
            works.extend(allreduce_ranges(g, [(0, self.part_ranges[0][0])], self.pg))
            for w in works:
                w.wait()
        elif self.world > 1 and not self.overlap_allreduce:
            _lib.call("cilrs_model_backward", m._handle, b, MODE_TRAIN, -1, self.dcontrols, self.dspeed, self.d_speed,
                      self.d_command, ctypes.c_float(m.dropout), sp)
            for w in allreduce_ranges(g, [(0, g.numel())], self.pg):
                w.wait()
        elif self.world > 1:
            # The backward runs in five parts; the allreduce of a part's gradient range is enqueued behind the model's
            # gradient stream (where the part's weight gradients finish) while the caller's stream already runs the next
            # part: the dgrad / BatchNorm chain never waits for a weight gradient or a collective. One join at the end.
            lib = _lib.lib()
            lib.cilrs_model_gradient_stream.restype = ctypes.c_void_p
            lib.cilrs_model_gradient_stream.argtypes = [ctypes.c_void_p]
            gs_ptr = lib.cilrs_model_gradient_stream(m._handle)
            gstream = torch.cuda.ExternalStream(gs_ptr, device=self.dev) if (gs_ptr and self.async_parts) else None
            for part in range(5):
                if gstream is not None:
                    _lib.call("cilrs_model_backward_part_async", m._handle, b, MODE_TRAIN, part, self.dcontrols, self.dspeed,
                              self.d_speed, self.d_command, ctypes.c_float(m.dropout), sp)
                    with torch.cuda.stream(gstream):
                        works.extend(allreduce_ranges(g, [self.part_ranges[part]], self.pg))
                else:
                    _lib.call("cilrs_model_backward", m._handle, b, MODE_TRAIN, part, self.dcontrols, self.dspeed, self.d_speed,
                              self.d_command, ctypes.c_float(m.dropout), sp)
                    works.extend(allreduce_ranges(g, [self.part_ranges[part]], self.pg))
            if gstream is not None:
                _lib.call("cilrs_model_backward_join", m._handle, sp)
            for w in works:
                w.wait()
        else:
            _lib.call("cilrs_model_backward", m._handle, b, MODE_TRAIN, -1, self.dcontrols, self.dspeed, self.d_speed,
                      self.d_command, ctypes.c_float(m.dropout), sp)
        scale_dev = None
        if self.grad_clip > 0:
            # norm of the averaged gradient = norm of the summed one / world
            _lib.call("cilrs_grad_sumsq", g, ctypes.c_longlong(g.numel()), self.norm_ws, self.norm_cnt,
                      ctypes.c_float(self.grad_clip * self.world), self.norm_out, sp)
            scale_dev = self.norm_out[1:]
        self.opt.step(grad_scale=1.0 / self.world, grad_scale_dev=scale_dev, grads_in_arena=True)
        _lib.call("cilrs_model_refresh", m._handle, 1, sp)
        self._after_step()

    def _after_step(self):
        # parameters and BatchNorm buffers changed through raw pointers (also on every graph replay): the bf16 operands were
        # repacked on the stream, but the next eval forward of the module must re-fold its BatchNorm
        m = self.model
        m.mark_parameters_changed(repacked=True)
        m._extra_b += 1
        m._fwd_gen += 1

    def _capture(self):
        """Warm up on a side stream, restore the state the warm-up mutated, then capture one step into a CUDA graph.
        (The Adam step number lives in device memory, so the same graph is valid for every step.)"""
        s = torch.cuda.Stream(device=self.dev)
        m = self.model
        m._refresh_if_needed(infer=False)
        state = (m.flat_parameters(), self.opt._m, self.opt._v, m._flat_buf, m._flat_nbt, self.opt._step_dev)
        saved = [t.clone() for t in state]
        step0 = self.opt._step
        torch.cuda.synchronize(self.dev)
        with torch.cuda.stream(s):
            for _ in range(2):
                self._device_step()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(g, stream=s):
                self._device_step()
            self.graph = g
        finally:
            # the warm-up steps (and a failed capture) must not leave a trace in the training state
            for t, sv in zip(state, saved):
                t.copy_(sv)
            self.opt._step = step0
            _lib.call("cilrs_model_refresh", m._handle, 1, _lib.stream_ptr())
            torch.cuda.synchronize(self.dev)

    # ------------------------------------------------------------------------------------------
    def load_batch(self, frames_or_image, speed, command, targets):
        """Asynchronous H2D (from pinned host tensors) or D2D copies into the static step inputs."""
        dst = self.d_frames if self.frames == "u8" else self.d_image
        dst.copy_(frames_or_image, non_blocking=True)
        self.d_speed.copy_(speed, non_blocking=True)
        self.d_command.copy_(command, non_blocking=True)
        self.d_targets.copy_(targets, non_blocking=True)

    def step(self):
        """One optimisation step on the loaded batch. Returns the device tensor of the 6 loss scalars
        (total, control, steer, throttle, brake, speed) — no host sync."""
        if self.graph is not None:
            self.graph.replay()
            self.opt._step += 1
            self._after_step()
        else:
            self._device_step()
        return self.loss6
