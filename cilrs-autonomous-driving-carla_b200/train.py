"""The fused training step (reference: train_one_epoch body, notebook/notebook.ipynb:545-555) and its data-parallel form.

    H2D -> [K0 normalise] -> forward -> loss -> backward -> [allreduce] -> [clip] -> Adam (+ zero_grad) -> repack

Everything after the H2D copies is device work launched through the C-ABI with no host synchronisation; with
`use_graph=True` the whole step (with N > 1 including the NCCL allreduces) is one CUDA-graph replay. Everything that varies
from step to step lives in device memory — the Adam step number and hyper-parameters (so `StepLR`, notebook.ipynb:535-536,604,
keeps working on the captured graph), the clip coefficient (`clip_grad_norm_(…, 1.0)`, :553-554) and the dropout mask counter
(`dropout=0.5`, :480) — so the reference's executed recipe runs on the replayed graph unchanged.

Data parallelism shards the batch across ranks (one process per GPU); the gradient arena is all-reduced over NCCL in groups
of the five ranges the backward completes back-to-front (heads+layer4, layer3, layer2, layer1, stem), each group's allreduce
overlapping the backward of the next one; `grad_comm="bf16"` exchanges bf16 gradients (half the bytes; Adam's moments and the
master weights stay fp32). BatchNorm statistics stay per rank (plain-DDP semantics; the reference has no SyncBN).
"""
import ctypes
import os

import torch

from . import _lib
from .ddp import SCHEDULES, allreduce_ranges, backward_part_ranges, broadcast_parameters
from .model import MODE_TRAIN
from .optim import FusedAdam


class FusedTrainer:
    def __init__(self, model, batch, lr=2e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, loss="mse", steer_w=5.0,
                 throttle_w=1.0, brake_w=1.0, speed_w=0.05, grad_clip=0.0, process_group=None, use_graph=False, frames="f32",
                 async_parts=False, overlap_allreduce="two", grad_comm="bf16", optimizer_in_backward=False, adam_beside_stem=False):
        # overlap_allreduce: a key of ddp.SCHEDULES (or True = "all", False = "tail"): which backward parts share an allreduce
        #   "all"   every part's range as soon as it is complete
        #   "two"   heads+layer4 | layer3 | layer2+layer1+stem  (the default: the exposed tail is 1.35 M of 22.4 M gradients)
        #   "first" heads+layer4 early, the rest after the backward    "tail" one allreduce after the backward
        if frames not in ("f32", "u8"):
            raise ValueError("frames: 'f32' ([B,3,88,200] normalised, as the reference's loader yields) or 'u8' ([B,88,200,3])")
        if grad_comm not in ("bf16", "fp32"):
            raise ValueError("grad_comm: 'bf16' or 'fp32'")
        if overlap_allreduce is True:
            overlap_allreduce = "all"
        elif overlap_allreduce is False:
            overlap_allreduce = "tail"
        if overlap_allreduce not in SCHEDULES:
            raise ValueError("overlap_allreduce: one of %s" % sorted(SCHEDULES))
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
        model.train()
        model._ensure(batch)
        model._refresh_if_needed(infer=False)   # the bf16 operands must exist before the first (eager) step
        self._plan_gen = model._plan_gen
        # dropout masks follow the optimizer's DEVICE step counter, so a replayed graph draws a new mask every step
        _lib.call("cilrs_model_set_dropout_counter", model._handle, self.opt._step_dev)
        if self.world > 1:
            broadcast_parameters(model, 0, process_group)
        # The step's static inputs are views of ONE buffer, and a second buffer of the same layout is the staging area of
        # prefetch_batch(): the next batch's H2D copies run on a copy stream under the current step, load_prefetched() is a single
        # D2D copy (the reference's DataLoader prefetches the same way: num_workers=2, pin_memory=True, notebook.ipynb:417-420).
        nb_img = batch * 88 * 200 * 3 * (4 if frames == "f32" else 1)
        offs, o = [], 0
        for nbytes in (nb_img, 4 * batch, 8 * batch, 12 * batch):
            offs.append(o)
            o += (nbytes + 255) // 256 * 256
        self._inputs = torch.zeros(o, dtype=torch.uint8, device=dev)
        self._stage = torch.zeros(o, dtype=torch.uint8, device=dev)

        def views(buf):
            img = buf[offs[0]:offs[0] + nb_img]
            img = img.view(torch.float32).view(batch, 3, 88, 200) if frames == "f32" else img.view(batch, 88, 200, 3)
            return (img, buf[offs[1]:offs[1] + 4 * batch].view(torch.float32), buf[offs[2]:offs[2] + 8 * batch].view(torch.long),
                    buf[offs[3]:offs[3] + 12 * batch].view(torch.float32).view(batch, 3))
        img, self.d_speed, self.d_command, self.d_targets = views(self._inputs)
        self.d_image = img if frames == "f32" else None
        self.d_frames = img if frames == "u8" else None
        self._stage_views = views(self._stage)
        self._copy_stream = None
        self._staged = None       # event: the staged batch has landed
        self._stage_free = None   # event: the staging buffer may be overwritten
        self.controls = torch.zeros(batch, 3, dtype=torch.float32, device=dev)
        self.pred_speed = torch.zeros(batch, dtype=torch.float32, device=dev)
        self.dcontrols = torch.zeros(batch, 3, dtype=torch.float32, device=dev)
        self.dspeed = torch.zeros(batch, dtype=torch.float32, device=dev)
        self.loss6 = torch.zeros(6, dtype=torch.float32, device=dev)
        self.norm_ws = torch.zeros(1024, dtype=torch.float64, device=dev)
        self.norm_cnt = torch.zeros(4, dtype=torch.int32, device=dev)
        self.norm_out = torch.zeros(2, dtype=torch.float32, device=dev)   # [sum of squares, clip coefficient]
        self.s2d = model.input_s2d_buffer(batch)
        self.err = model.error_flag()
        self.part_ranges = backward_part_ranges(model)
        self.schedule = SCHEDULES[overlap_allreduce]
        self.overlap_allreduce = overlap_allreduce
        # async_parts: enqueue each group's (conversion and) allreduce behind the model's gradient stream instead of joining
        # that stream into the caller's after every group: the dgrad / BatchNorm chain never waits for a weight gradient
        self.async_parts = bool(async_parts)
        self.grad_comm = grad_comm if self.world > 1 else "fp32"
        self.g16 = torch.zeros(model.flat_parameters().numel(), dtype=torch.bfloat16, device=dev) if self.grad_comm == "bf16" else None
        model.flat_gradients().zero_()   # from here on Adam / the bf16 conversion leave the arena zeroed for the next step
        self.graph = None
        self.graph_error = None
        self.skip_optimizer = False   # test hook: leave the exchanged gradient in place (no clip / Adam / repack / zeroing)
        self.optimizer_in_backward = optimizer_in_backward
        self._opt_in_bwd = bool(optimizer_in_backward) and self.world == 1 and grad_clip == 0   # the variant actually taken
        # opt-in (single GPU, no clipping): Adam of everything above the stem runs on the gradient stream while the stem's backward
        # (max-pool + BatchNorm backward + conv1 weight gradient: three HBM-bound launches, ~0.19 ms) is still on the main one.
        # Measured on B200: 2.81 ms against 2.79 ms - the stem's elementwise kernels hold every SM's registers (2 CTAs x 256 threads
        # x 128 registers), so Adam only starts when they drain (CUPTI timeline: +58 us late) and then shares HBM with conv1's wgrad.
        self._adam_beside_stem = bool(adam_beside_stem) and self.world == 1 and grad_clip == 0 and not self._opt_in_bwd
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
    def _backward_args(self):
        return (self.dcontrols, self.dspeed, self.d_speed, self.d_command, ctypes.c_float(self.model.dropout), _lib.stream_ptr())

    def _exchange(self, g, lo, hi):
        """Start the allreduce of gradient range [lo, hi) on the current stream's order; returns work handles."""
        if self.grad_comm == "bf16":
            _lib.call("cilrs_grad_to_bf16", g[lo:hi], self.g16[lo:hi], ctypes.c_longlong(hi - lo), 1, _lib.stream_ptr())
            return allreduce_ranges(self.g16, [(lo, hi)], self.pg)
        return allreduce_ranges(g, [(lo, hi)], self.pg)

    def _device_step(self):
        m = self.model
        b = self.batch
        sp = _lib.stream_ptr()
        if not self._opt_in_bwd:
            # bf16 operands of the weights the previous step's Adam wrote: the stem's on this stream, the trunk's on the model's
            # gradient stream beside K0 + the stem convolution (the forward waits for them after its stem) - 59 us off the chain
            _lib.call("cilrs_model_refresh_async", m._handle, sp)
        if self.frames == "u8":
            _lib.call("cilrs_preprocess_u8", self.d_frames, b, 88, 200, 3, 0, 88, 200, None, None, self.s2d, sp)
            img, s2d = None, self.s2d
        else:
            img, s2d = self.d_image, None
        m._seed = (m._seed * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        # forward with the loss (and its gradients w.r.t. controls / pred_speed) fused into the heads kernel
        _lib.call("cilrs_model_forward_loss", m._handle, b, MODE_TRAIN, img, s2d, self.d_speed, self.d_command, self.controls,
                  self.pred_speed, 1, ctypes.c_float(m.dropout), ctypes.c_ulonglong(m._seed), self.d_targets, self.loss_mode,
                  ctypes.c_float(self.ws[0]), ctypes.c_float(self.ws[1]), ctypes.c_float(self.ws[2]), ctypes.c_float(self.ws[3]),
                  ctypes.c_float(1.0), self.loss6, self.dcontrols, self.dspeed, sp)
        g = m.flat_gradients()   # zero on entry: the previous step's Adam / bf16 conversion cleared it (optimizer.zero_grad())
        if self.world > 1:
            works = []
            lib = _lib.lib()
            lib.cilrs_model_gradient_stream.restype = ctypes.c_void_p
            lib.cilrs_model_gradient_stream.argtypes = [ctypes.c_void_p]
            gs_ptr = lib.cilrs_model_gradient_stream(m._handle) if self.async_parts else None
            gstream = torch.cuda.ExternalStream(gs_ptr, device=self.dev) if gs_ptr else None
            for group in self.schedule:
                lo, hi = self.part_ranges[group[-1]][0], self.part_ranges[group[0]][1]
                if gstream is not None:
                    # everything the group wrote is complete in the ORDER OF the gradient stream: exchange it there
                    for part in group:
                        _lib.call("cilrs_model_backward_part_async", m._handle, b, MODE_TRAIN, part, *self._backward_args())
                    with torch.cuda.stream(gstream):
                        works.extend(self._exchange(g, lo, hi))
                else:
                    for part in group[:-1]:
                        _lib.call("cilrs_model_backward_part_async", m._handle, b, MODE_TRAIN, part, *self._backward_args())
                    _lib.call("cilrs_model_backward", m._handle, b, MODE_TRAIN, group[-1], *self._backward_args())  # joins the gradient stream
                    works.extend(self._exchange(g, lo, hi))
            if gstream is not None:
                _lib.call("cilrs_model_backward_join", m._handle, sp)
            if self.skip_optimizer or self.grad_clip > 0 or len(works) < 2:
                for w in works:
                    w.wait()
            else:
                # The last group's allreduce (layer2 + layer1 + stem with the default schedule: 6 % of the gradients) is the
                # only collective nothing hides. Adam already updates everything above it while that allreduce is in flight.
                for w in works[:-1]:
                    w.wait()
                cut = self.part_ranges[self.schedule[-1][0]][1]
                self.opt.step(grad_scale=1.0 / self.world, grads_in_arena=True, grads_bf16=self.g16, zero_grad=self.g16 is None,
                              arena_range=(cut, g.numel()))
                works[-1].wait()
                self.opt.step(grad_scale=1.0 / self.world, grads_in_arena=True, grads_bf16=self.g16, zero_grad=self.g16 is None,
                              arena_range=(0, cut), advance=False)
                self._after_step()
                return
        else:
            lib = _lib.lib()
            lib.cilrs_model_gradient_stream.restype = ctypes.c_void_p
            lib.cilrs_model_gradient_stream.argtypes = [ctypes.c_void_p]
            gs_ptr = lib.cilrs_model_gradient_stream(m._handle) if self._opt_in_bwd else None
            if gs_ptr and self.grad_clip == 0 and not self.skip_optimizer:
                # Optimizer in the backward (opt-in): once a part's gradients are complete (in the order of the model's gradient
                # stream) Adam and the bf16 repack of that part run THERE, under the dgrad chain of the lower layers. Layer4 +
                # heads (64 % of the parameters) and layer3 (30 %) go this way; only the last 6 % wait for the end of the step.
                # Measured on B200: 3.16 ms per step against 3.05 ms with Adam at the end - the gradient stream is not idle
                # capacity, its HBM traffic and SMs come out of the dgrad chain's share - hence off by default.
                gstream = torch.cuda.ExternalStream(gs_ptr, device=self.dev)
                self.opt._sync_hyper(1.0)
                for part in range(5):
                    _lib.call("cilrs_model_backward_part_async", m._handle, b, MODE_TRAIN, part, *self._backward_args())
                    if part < 2:
                        with torch.cuda.stream(gstream):
                            self.opt.step(grads_in_arena=True, zero_grad=True, arena_range=self.part_ranges[part], advance=(part == 0))
                            _lib.call("cilrs_model_refresh_part", m._handle, part, _lib.stream_ptr())
                _lib.call("cilrs_model_backward_join", m._handle, sp)
                self.opt.step(grads_in_arena=True, zero_grad=True, arena_range=(0, self.part_ranges[1][0]), advance=False)
                for part in (2, 3, 4):
                    _lib.call("cilrs_model_refresh_part", m._handle, part, sp)
                self._after_step(repacked=True)
                return
            if gs_ptr is None and self._adam_beside_stem and not self.skip_optimizer:
                gs = lib.cilrs_model_gradient_stream(m._handle)
                if gs:
                    gstream = torch.cuda.ExternalStream(gs, device=self.dev)
                    self.opt._sync_hyper(1.0)
                    for part in range(4):
                        _lib.call("cilrs_model_backward_part_async", m._handle, b, MODE_TRAIN, part, *self._backward_args())
                    # part 5 = the stem's max-pool / BatchNorm backward (HBM-bound, main stream); the gradient stream is made to wait
                    # for it, so Adam starts exactly when conv1's weight gradient (part 6: L2-bound, 5 % DRAM) does
                    _lib.call("cilrs_model_backward_part_async", m._handle, b, MODE_TRAIN, 5, *self._backward_args())
                    cut = self.part_ranges[3][0]      # the stem's parameters come first in the arena: [0, cut)
                    with torch.cuda.stream(gstream):  # ordered behind every gradient of layers 1-4 and the heads
                        self.opt.step(grads_in_arena=True, zero_grad=True, arena_range=(cut, g.numel()))
                    _lib.call("cilrs_model_backward_part_async", m._handle, b, MODE_TRAIN, 6, *self._backward_args())
                    _lib.call("cilrs_model_backward_join", m._handle, sp)
                    self.opt.step(grads_in_arena=True, zero_grad=True, arena_range=(0, cut), advance=False)
                    self._after_step()
                    return
            _lib.call("cilrs_model_backward", m._handle, b, MODE_TRAIN, -1, *self._backward_args())
        if self.skip_optimizer:
            return
        scale_dev = None
        if self.grad_clip > 0:
            # clip_grad_norm_ on the AVERAGED gradient: norm(avg) = norm(sum) / world, so the bound is scaled instead
            if self.g16 is not None:
                _lib.call("cilrs_grad_sumsq_bf16", self.g16, ctypes.c_longlong(g.numel()), self.norm_ws, self.norm_cnt,
                          ctypes.c_float(self.grad_clip * self.world), self.norm_out, sp)
            else:
                _lib.call("cilrs_grad_sumsq", g, ctypes.c_longlong(g.numel()), self.norm_ws, self.norm_cnt,
                          ctypes.c_float(self.grad_clip * self.world), self.norm_out, sp)
            scale_dev = self.norm_out[1:]
        self.opt.step(grad_scale=1.0 / self.world, grad_scale_dev=scale_dev, grads_in_arena=True, grads_bf16=self.g16,
                      zero_grad=self.g16 is None)
        self._after_step()

    def _after_step(self, repacked=False):
        # parameters and BatchNorm buffers changed through raw pointers (also on every graph replay): the next step repacks the
        # bf16 operands itself (first thing, beside the stem); a forward of the MODULE in between must repack (and, in eval
        # mode, re-fold its BatchNorm) first - unless the optimizer-in-backward variant already repacked on the stream
        m = self.model
        m.mark_parameters_changed(repacked=repacked)
        m._extra_b += 1
        m._fwd_gen += 1

    def _capture(self):
        """Warm up on a side stream, restore the state the warm-up mutated, then capture one step into a CUDA graph.
        (Step number, hyper-parameters, clip coefficient and dropout counter live in device memory, so the same graph is valid
        for every step.)"""
        # Measurement aid: CILRS_CAPTURE_PRIORITY=1 captures on a HIGH-priority stream. Kernel nodes inherit the priority of the
        # stream they were captured on; the model's gradient stream (weight gradients, repack) has the lowest priority, which is
        # also an ordinary stream's, so only a raised main chain makes the difference real. Measured on B200: 2.92-2.96 ms per
        # step against 2.73 ms - with the dgrad / BatchNorm chain always first in line the weight gradients fall behind and are
        # exposed at the end of every layer group. Equal priorities (the default) interleave them better.
        prio = -1 if os.environ.get("CILRS_CAPTURE_PRIORITY", "0") == "1" else 0
        s = torch.cuda.Stream(device=self.dev, priority=prio)
        m = self.model
        m._refresh_if_needed(infer=False)
        state = (m.flat_parameters(), self.opt._m, self.opt._v, m._flat_buf, m._flat_nbt, self.opt._step_dev)
        saved = [t.clone() for t in state]
        step0, seed0 = self.opt._step, m._seed
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
            m._seed = seed0
            m.flat_gradients().zero_()
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

    def prefetch_batch(self, frames_or_image, speed, command, targets):
        """Start copying the NEXT batch (pinned host tensors, or device tensors) into the staging buffer on the trainer's copy
        stream; returns at once. The copies overlap whatever the compute stream is doing - typically the current step.
        `load_prefetched()` then makes it the step's input. The host tensors must stay untouched until then."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.dev)
        cs = self._copy_stream
        if self._stage_free is not None:
            cs.wait_event(self._stage_free)     # the previous staged batch has been moved out
        else:
            cs.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(cs):
            for dst, src in zip(self._stage_views, (frames_or_image, speed, command, targets)):
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        self._staged = ev

    def load_prefetched(self):
        """The batch staged by prefetch_batch() becomes the step's input: one D2D copy on the current stream."""
        if self._staged is None:
            raise RuntimeError("cilrs_b200.FusedTrainer.load_prefetched: no batch was prefetched")
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self._staged)
        self._inputs.copy_(self._stage, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cur)
        self._stage_free = ev
        self._staged = None

    def step(self):
        """One optimisation step on the loaded batch. Returns the device tensor of the 6 loss scalars
        (total, control, steer, throttle, brake, speed) — no host sync."""
        if self.model._plan_gen != self._plan_gen:
            raise RuntimeError("cilrs_b200.FusedTrainer: the model's plan was rebuilt (a larger batch went through the module, or it "
                               "moved device) after this trainer was created; create a new FusedTrainer")
        if self.graph is not None:
            self.opt._sync_hyper(1.0 / self.world)   # lr / weight-decay changes (StepLR) reach the captured kernels here
            self.graph.replay()
            self.opt._step += 1
            self._after_step(repacked=self._opt_in_bwd)
        else:
            self._device_step()
        return self.loss6

    def read_loss(self):
        """Host copy of the 6 loss scalars (one D2H sync) — also the point where an out-of-range command raises, like the
        reference's gather would."""
        vals = torch.cat([self.loss6, self.err.float()]).tolist()
        if vals[6] != 0:
            self.err.zero_()
            raise IndexError("cilrs_b200: a command index outside [0, 4) reached the model")
        return dict(zip(("total", "control", "steer", "throttle", "brake", "speed"), vals[:6]))

    def close(self):
        """Drop the captured graph (it references the process group's communicator) before the process group is destroyed."""
        if self.graph is not None:
            torch.cuda.synchronize(self.dev)
            self.graph.reset()
            self.graph = None
