"""Drop-in `CILRS(nn.Module)` for the reference model (model/autonomous_drive.py:361-399 == notebook/notebook.ipynb:440-477).

Same constructor (`num_commands=4, dropout=0.0`), same `forward(image, speed, command) -> (controls, pred_speed)`,
same 250-key `state_dict()` layout (strict load of a reference checkpoint works), parameters are
ordinary fp32 leaf `nn.Parameter`s whose `.grad` is filled by `loss.backward()`, so `optim.Adam(model.parameters())`,
`clip_grad_norm_`, `model.train()/eval()`, `.to(device)` keep working (notebook/notebook.ipynb:480-555).

What is different is everything underneath: the sub-modules below only *hold* parameters (their own forward is never
used); forward/backward run the sm_100a kernels through the C-ABI (`cilrs_model_forward/backward`). All parameters are
views into one flat fp32 arena (and gradients into a second one) so the fused Adam and the gradient allreduce see one
contiguous buffer. There is no CPU path: a non-CUDA module raises.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib

IMG_H, IMG_W = 88, 200
MODE_TRAIN, MODE_FROZEN, MODE_INFER = 0, 1, 2


class _Holder(nn.Module):
    """Parameter container; compute happens in the fused CUDA plan, never here."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("cilrs_b200: sub-modules only hold parameters; call CILRS.forward")

    def __getitem__(self, idx):  # the reference containers are nn.Sequential: keep visual_encoder[1] etc. working
        return self._modules[str(idx)]

    def __len__(self):
        return len(self._modules)


class _Conv(_Holder):
    def __init__(self, cin, cout, k):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        nn.init.kaiming_normal_(self.weight, mode="fan_out", nonlinearity="relu")  # torchvision resnet init


class _BN(_Holder):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class _Downsample(_Holder):
    def __init__(self, cin, cout):
        super().__init__()
        self.add_module("0", _Conv(cin, cout, 1))
        self.add_module("1", _BN(cout))


class _BasicBlock(_Holder):
    def __init__(self, cin, cout, downsample):
        super().__init__()
        self.conv1 = _Conv(cin, cout, 3)
        self.bn1 = _BN(cout)
        self.conv2 = _Conv(cout, cout, 3)
        self.bn2 = _BN(cout)
        if downsample:
            self.downsample = _Downsample(cin, cout)


def _layer(cin, cout, blocks):
    mods = [_BasicBlock(cin, cout, cin != cout)] + [_BasicBlock(cout, cout, False) for _ in range(blocks - 1)]
    seq = _Holder()
    for i, mod in enumerate(mods):
        seq.add_module(str(i), mod)
    return seq


class _Slot(_Holder):
    """Index placeholder for parameter-free entries of the reference Sequentials (ReLU, MaxPool, ...)."""


def _mlp(spec):
    """spec: list of ('lin', in, out) | None, placed at the reference's Sequential indices."""
    seq = _Holder()
    for i, s in enumerate(spec):
        seq.add_module(str(i), nn.Linear(s[1], s[2]) if s else _Slot())
    return seq


class _Function(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, image, speed, command, *params):
        ctx.image = image if model.compute_dtype == "fp32" else None
        controls, pred_speed = model._launch_forward(image, speed, command, keep=True)
        ctx.model = model
        ctx.gen = model._fwd_gen
        ctx.speed = speed
        ctx.command = command
        ctx.batch = image.shape[0]
        ctx.mode = model._last_mode
        ctx.dropout = model._last_dropout
        return controls, pred_speed

    @staticmethod
    def backward(ctx, dcontrols, dspeed):
        model = ctx.model
        if ctx.gen != model._fwd_gen:
            raise RuntimeError("cilrs_b200: backward() must follow the forward() it belongs to (activations of one forward "
                               "are kept at a time)")
        b = ctx.batch
        dev = ctx.speed.device
        dcontrols = torch.zeros(b, 3, device=dev) if dcontrols is None else dcontrols.contiguous().float()
        dspeed = torch.zeros(b, device=dev) if dspeed is None else dspeed.contiguous().float()
        grads = model._fresh_grad_arena()
        if model.compute_dtype == "fp32":
            _lib.call("cilrs_model32_backward", model._handle, b, ctx.mode, dcontrols, dspeed, ctx.speed, ctx.command,
                      float(ctx.dropout), _lib.stream_ptr())
        else:
            _lib.call("cilrs_model_backward", model._handle, b, ctx.mode, -1, dcontrols, dspeed, ctx.speed, ctx.command,
                      float(ctx.dropout), _lib.stream_ptr())
        return (None, None, None, None) + tuple(model._views(grads))


class CILRS(nn.Module):
    """compute_dtype (opt-in keyword, the reference's two ctor arguments keep their meaning and defaults):
      "bf16" (default) - bf16 operands / fp32 accumulate on the tcgen05 tensor cores: the throughput mode (2e-2 parity)
      "fp32"           - fp32 storage and arithmetic end to end (csrc/fp32_path.cu): reproduces the reference's fp32 numbers to
                         1e-4 (configs/train_config.json:54 "mixed_precision": false); slower, same interface, same state_dict."""

    def __init__(self, num_commands=4, dropout=0.0, compute_dtype="bf16"):
        super().__init__()
        if num_commands != 4:
            raise ValueError("cilrs_b200 supports the reference's num_commands=4 only")
        if compute_dtype not in ("bf16", "fp32"):
            raise ValueError("compute_dtype must be 'bf16' or 'fp32'")
        self.compute_dtype = compute_dtype
        self.num_commands = num_commands
        self.dropout = float(dropout)
        ve = _Holder()
        mods = [_Conv(3, 64, 7), _BN(64), _Slot(), _Slot(), _layer(64, 64, 3), _layer(64, 128, 4), _layer(128, 256, 6),
                _layer(256, 512, 3), _Slot(), _Slot()]
        for i, mod in enumerate(mods):
            ve.add_module(str(i), mod)
        self.visual_encoder = ve
        self.speed_encoder = _mlp([("lin", 1, 128), None, None, ("lin", 128, 128), None])
        self.control_branches = nn.ModuleList(
            [_mlp([("lin", 640, 256), None, None, ("lin", 256, 256), None, None, ("lin", 256, 3)]) for _ in range(num_commands)])
        self.speed_predictor = _mlp([("lin", 512, 256), None, None, ("lin", 256, 256), None, ("lin", 256, 1)])

        # ---- flat arenas (layout comes from the C side: single source of truth) ----
        n = 256
        off = (ctypes.c_longlong * n)()
        siz = (ctypes.c_longlong * n)()
        tot = ctypes.c_longlong()
        buf = ctypes.c_longlong()
        nbn = ctypes.c_int()
        cnt = _lib.lib().cilrs_model_param_layout(off, siz, n, ctypes.byref(tot), ctypes.byref(buf), ctypes.byref(nbn))
        self._offsets = [off[i] for i in range(cnt)]
        self._sizes = [siz[i] for i in range(cnt)]
        self._total = tot.value
        self._buf_total = buf.value
        self._num_bn = nbn.value
        plist = list(self.parameters())
        if len(plist) != cnt or any(p.numel() != s for p, s in zip(plist, self._sizes)):
            raise RuntimeError("cilrs_b200: parameter layout mismatch between the Python module and the C plan")
        self._flat = None
        self._flat_grad = None
        self._flat_buf = None
        self._flat_nbt = None
        self._handle = None
        self._workspace = None
        self._max_batch = 0
        self._fwd_gen = 0
        self._last_mode = MODE_INFER
        self._last_dropout = 0.0
        self._wkey = None         # parameter version the packed bf16 operands were made from
        self._fkey = None         # (parameter, buffer) version the folded eval-mode BN vectors were made from
        self._extra_w = 0         # bumped by code that writes the parameter arena through raw pointers (FusedAdam)
        self._extra_b = 0         # same for the BN running statistics (train-mode forward)
        # dropout mask stream: follows torch.manual_seed and differs per data-parallel rank (identical masks on every rank
        # for the same local sample index would correlate the replicas)
        rank = 0
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            rank = torch.distributed.get_rank()
        self._seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + 0x5EED + (rank << 40)) & 0xFFFFFFFFFFFFFFFF
        self._plan_gen = 0        # bumped whenever the C plan / workspace is rebuilt: FusedTrainer / InferenceSession check it
        self._plist = None
        self._flatten()

    # ------------------------------------------------------------------------------------------
    # arenas
    # ------------------------------------------------------------------------------------------
    def _bn_modules(self):
        return [m for m in self.modules() if isinstance(m, _BN)]

    def _flatten(self):
        """(Re)build the flat parameter / buffer arenas on the parameters' current device and re-point every
        parameter and buffer at its view."""
        plist = list(self.parameters())
        dev = plist[0].device
        flat = torch.zeros(self._total, dtype=torch.float32, device=dev)
        for p, o, s in zip(plist, self._offsets, self._sizes):
            flat[o:o + s].copy_(p.detach().reshape(-1).float())
            p.data = flat[o:o + s].view(p.shape)
            p.grad = None
        bns = self._bn_modules()
        fbuf = torch.zeros(self._buf_total, dtype=torch.float32, device=dev)
        nbt = torch.zeros(len(bns), dtype=torch.long, device=dev)
        o = 0
        for i, bn in enumerate(bns):
            c = bn.running_mean.numel()
            fbuf[o:o + c].copy_(bn.running_mean.float())
            fbuf[o + c:o + 2 * c].copy_(bn.running_var.float())
            nbt[i] = bn.num_batches_tracked.to(dev)
            bn._buffers["running_mean"] = fbuf[o:o + c]
            bn._buffers["running_var"] = fbuf[o + c:o + 2 * c]
            bn._buffers["num_batches_tracked"] = nbt[i]
            o += 2 * c
        self._flat, self._flat_buf, self._flat_nbt = flat, fbuf, nbt
        self._flat_grad = None
        self._plist = plist
        self._destroy_handle()

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._flatten()
        return out

    def _destroy_handle(self):
        if getattr(self, "_handle", None) is not None:
            if self.compute_dtype == "fp32":
                _lib.lib().cilrs_model32_destroy(self._handle)
            else:
                _lib.lib().cilrs_model_destroy(self._handle)
        self._handle = None
        self._workspace = None
        self._max_batch = 0
        self._wkey = None
        self._fkey = None
        self._plan_gen = getattr(self, "_plan_gen", 0) + 1

    def __del__(self):
        try:
            self._destroy_handle()
        except Exception:  # pragma: no cover
            pass

    def _views(self, flat):
        return [flat[o:o + s].view(p.shape) for p, o, s in zip(self.parameters(), self._offsets, self._sizes)]

    def flat_parameters(self):
        """The fp32 arena all parameters are views of (padding between tensors is zero)."""
        return self._flat

    def flat_gradients(self):
        """The fp32 gradient arena (same layout); allocated on first use."""
        if self._flat_grad is None:
            self._flat_grad = torch.zeros_like(self._flat)
            self._bind()
        return self._flat_grad

    def _fresh_grad_arena(self):
        """Zeroed arena for one autograd backward. If parameter .grad tensors still alias the arena (the caller is
        accumulating gradients across backward calls), a new arena is used so they are not clobbered."""
        g = self.flat_gradients()
        p0 = next(self.parameters())
        if p0.grad is not None and p0.grad.data_ptr() == g.data_ptr():
            self._flat_grad = torch.zeros_like(self._flat)
            self._bind()
            return self._flat_grad
        g.zero_()
        return g

    # ------------------------------------------------------------------------------------------
    # C plan
    # ------------------------------------------------------------------------------------------
    def _ensure(self, batch):
        if not self._flat.is_cuda:
            raise RuntimeError("cilrs_b200.CILRS runs on CUDA only (no CPU fallback): call .to('cuda')")
        if self._handle is None or batch > self._max_batch:
            self._destroy_handle()
            lib = _lib.lib()
            fp32 = self.compute_dtype == "fp32"
            ws_bytes = lib.cilrs_model32_workspace_bytes if fp32 else lib.cilrs_model_workspace_bytes
            ws_bytes.restype = ctypes.c_size_t
            create = lib.cilrs_model32_create if fp32 else lib.cilrs_model_create
            nbytes = ws_bytes(int(batch))
            with torch.cuda.device(self._flat.device):
                self._workspace = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self._flat.device)
                base = (self._workspace.data_ptr() + 1023) // 1024 * 1024
                h = ctypes.c_void_p()
                st = create(ctypes.byref(h), int(batch), ctypes.c_void_p(base), ctypes.c_size_t(nbytes), _lib.stream_ptr())
            if st != 0:
                raise RuntimeError("cilrs_model_create failed: %s" % lib.cilrs_status_string(st).decode())
            self._handle = h
            self._max_batch = int(batch)
            self._bind()

    def _bind(self):
        if self._handle is not None:
            _lib.call("cilrs_model32_bind" if self.compute_dtype == "fp32" else "cilrs_model_bind", self._handle, self._flat,
                      self._flat_grad, self._flat_buf, self._flat_nbt)

    def _param_version(self):
        # `p.data = flat[...]` (see _flatten) leaves every parameter with its OWN version counter: in-place updates made through
        # the parameters (torch.optim.Adam.step(), load_state_dict(), p.mul_()) bump p._version, not flat._version
        return sum(p._version for p in self._plist) + self._flat._version

    def _refresh_if_needed(self, infer):
        if self.compute_dtype == "fp32":
            return  # the fp32 plan re-derives its operands from the masters in every forward
        wkey = (self._param_version(), self._extra_w)
        what = 0
        if wkey != self._wkey:
            what |= 1
        fkey = (wkey, self._flat_buf._version, self._extra_b)
        if infer and fkey != self._fkey:
            what |= 2
        if what:
            _lib.call("cilrs_model_refresh", self._handle, what, _lib.stream_ptr())
            self._wkey = wkey
            if what & 2:
                self._fkey = fkey

    def mark_parameters_changed(self, repacked=False):
        """For code that updates the parameter arena through raw pointers (FusedAdam). `repacked=True` says the bf16
        operands were already refreshed on the stream (so the next forward need not do it again)."""
        self._extra_w += 1
        if repacked:
            self._wkey = (self._param_version(), self._extra_w)

    def _check_inputs(self, image, speed, command):
        if image.dim() != 4 or tuple(image.shape[1:]) != (3, IMG_H, IMG_W):
            raise ValueError("cilrs_b200: image must be [B,3,88,200] (got %s)" % (tuple(image.shape),))
        b = image.shape[0]
        if speed.shape != (b,) or command.shape != (b,):
            raise ValueError("cilrs_b200: speed and command must be [B]")
        if not (image.is_cuda and speed.is_cuda and command.is_cuda):
            raise RuntimeError("cilrs_b200: inputs must be CUDA tensors (no CPU fallback)")
        if command.dtype != torch.long:
            raise TypeError("cilrs_b200: command must be int64 (torch.long), as in the reference")

    def _launch_forward(self, image, speed, command, keep, s2d=None):
        b = speed.shape[0]
        self._ensure(b)
        need_keep = bool(keep)
        if need_keep:
            self.flat_gradients()  # the gradient arena must exist (and be bound) before a forward that will be back-propagated
        mode = MODE_TRAIN if self.training else (MODE_FROZEN if need_keep else MODE_INFER)
        self._refresh_if_needed(infer=(mode == MODE_INFER))
        dropout = self.dropout if self.training else 0.0
        self._seed = (self._seed * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        controls = torch.empty(b, 3, dtype=torch.float32, device=speed.device)
        pred_speed = torch.empty(b, dtype=torch.float32, device=speed.device)
        img = None if s2d is not None else image.contiguous().float()
        if self.compute_dtype == "fp32":
            if img is None:
                raise RuntimeError("cilrs_b200: the fp32 mode takes the normalised image tensor, not the bf16 space-to-depth buffer")
            _lib.call("cilrs_model32_forward", self._handle, b, mode, img, speed.contiguous().float(), command.contiguous(),
                      controls, pred_speed, int(self.training), int(need_keep), ctypes.c_float(dropout),
                      ctypes.c_ulonglong(self._seed), _lib.stream_ptr())
        else:
            _lib.call("cilrs_model_forward", self._handle, b, mode, img, s2d, speed.contiguous().float(), command.contiguous(),
                      controls, pred_speed, int(self.training), int(need_keep), ctypes.c_float(dropout),
                      ctypes.c_ulonglong(self._seed), _lib.stream_ptr())
        if self.training:
            self._extra_b += 1  # running statistics were updated through raw pointers
        self._fwd_gen += 1
        self._last_mode = mode
        self._last_dropout = dropout
        return controls, pred_speed

    def forward(self, image, speed, command):
        self._check_inputs(image, speed, command)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return _Function.apply(self, image, speed, command, *self.parameters())
        return self._launch_forward(image, speed, command, keep=False)

    def _workspace_view(self, ptr, nbytes):
        start = ptr - self._workspace.data_ptr()
        if start < 0 or start + nbytes > self._workspace.numel():
            raise RuntimeError("cilrs_b200: pointer outside the model workspace")
        return self._workspace[start:start + nbytes]

    def error_flag(self):
        """int32[1] device tensor the kernels set to 1 when a command outside {0,1,2,3} was seen (the reference's gather
        would raise, model/autonomous_drive.py:395-398; here the command is clamped and the flag raised)."""
        self._ensure(max(1, self._max_batch))
        lib = _lib.lib()
        fn = lib.cilrs_model32_error_flag if self.compute_dtype == "fp32" else lib.cilrs_model_error_flag
        fn.restype = ctypes.c_void_p
        return self._workspace_view(fn(self._handle), 4).view(torch.int32)

    def check_errors(self):
        """Synchronising check of the device error flag; raises like the reference's out-of-range gather would."""
        if self._handle is None:
            return
        flag = self.error_flag()
        if int(flag.item()) != 0:
            flag.zero_()
            raise IndexError("cilrs_b200: a command index outside [0, %d) reached the model" % self.num_commands)

    def _bf16_plan_only(self, what):
        if self.compute_dtype != "bf16":
            raise RuntimeError("cilrs_b200: %s is a hook of the bf16 plan" % what)

    def debug_backward(self, batch, hi, lo, g_out):
        """Test hook (cilrs_model_debug_backward): backward of blocks hi..max(lo,0) (+ the stem when lo < 0) from `g_out`, a bf16
        padded-flat gradient w.r.t. block hi's output; returns the bf16 padded-flat gradient w.r.t. the input of block max(lo,0)."""
        self._bf16_plan_only("debug_backward")
        lib = _lib.lib()
        _lib.call("cilrs_model_debug_backward", self._handle, int(batch), self._last_mode, int(hi), int(lo), g_out.contiguous(),
                  _lib.stream_ptr())
        lib.cilrs_model_debug_gradient.restype = ctypes.c_void_p
        ptr = lib.cilrs_model_debug_gradient(self._handle)
        first = max(lo, 0)
        if hi < 0:
            return None  # stem only: the gradient w.r.t. the image is not computed (the input needs none)
        dims = (ctypes.c_int * 5)()
        lib.cilrs_model_debug_activation.restype = ctypes.c_void_p
        lib.cilrs_model_debug_activation(self._handle, int(first), dims)   # activation `first` is the input of block `first`
        h, w, c, hp, wp = dims[0], dims[1], dims[2], dims[3], dims[4]
        n = batch * hp * wp * c * 2
        return self._workspace_view(ptr, n).view(torch.bfloat16).view(batch, hp, wp, c)

    def debug_heads_saved(self, which, batch):
        """Test hook: fp32 [batch, width] head activation kept by the last forward (post-ReLU, post-Dropout); see the header."""
        self._bf16_plan_only("debug_heads_saved")
        lib = _lib.lib()
        lib.cilrs_model_debug_heads_saved.restype = ctypes.c_void_p
        width = ctypes.c_int()
        ptr = lib.cilrs_model_debug_heads_saved(self._handle, int(which), ctypes.byref(width))
        if not ptr:
            raise ValueError("no such head activation")
        return self._workspace_view(ptr, batch * width.value * 4).view(torch.float32).view(batch, width.value)

    def input_s2d_buffer(self, batch):
        """bf16 [batch,47,103,16] view of the plan's conv1 input: the preprocessing kernel can write frames there directly."""
        if self.compute_dtype == "fp32":
            raise RuntimeError("cilrs_b200: FusedTrainer / InferenceSession drive the bf16 plan; use the module interface in fp32 mode")
        self._ensure(batch)
        lib = _lib.lib()
        lib.cilrs_model_input_s2d.restype = ctypes.c_void_p
        ptr = lib.cilrs_model_input_s2d(self._handle)
        start = ptr - self._workspace.data_ptr()
        n = batch * 47 * 103 * 16 * 2
        return self._workspace[start:start + n].view(torch.bfloat16).view(batch, 47, 103, 16)

    def debug_activation(self, which, batch):
        """Test hook: bf16 [batch,H,W,C] activation of the last forward (0 = max-pool out, 1..16 = block outputs).
        Inside the plan the tensors are in the padded-flat layout [batch,H+1,W+1,C]; the real pixels are returned."""
        self._bf16_plan_only("debug_activation")
        lib = _lib.lib()
        lib.cilrs_model_debug_activation.restype = ctypes.c_void_p
        dims = (ctypes.c_int * 5)()
        ptr = lib.cilrs_model_debug_activation(self._handle, int(which), dims)
        if not ptr:
            raise ValueError("no such activation")
        h, w, c, hp, wp = dims[0], dims[1], dims[2], dims[3], dims[4]
        start = ptr - self._workspace.data_ptr()
        n = batch * hp * wp * c * 2
        return self._workspace[start:start + n].view(torch.bfloat16).view(batch, hp, wp, c)[:, :h, :w]
