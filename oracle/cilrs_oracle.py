"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's CILRS hot path. Nothing in the product package imports this module; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may.

What it restates (citations into /root/reference):
  * preprocess_image                     model/autonomous_drive.py:897-902   -> preprocess_np / preprocess_c
  * CILRS.__init__/forward               model/autonomous_drive.py:361-399 == notebook/notebook.ipynb:440-477
      (torchvision.models.resnet34 topology: torchvision/models/resnet.py BasicBlock :59-105, ResNet :166-290)
  * CILRSLoss.forward (L1x(5,1,1)+0.5 MSE) notebook/notebook.ipynb:504-527, and the README/config recipe
      MSE + 0.05 MSE                      configs/train_config.json:30-32, README.md:104-105
  * optim.Adam(lr, weight_decay) step    notebook/notebook.ipynb:533-534,555 (torch/optim/adam.py, L2-coupled decay)
  * predict_controls speed normalisation  model/autonomous_drive.py:908-920

The arithmetic itself lives in un-vendored third-party dependencies (requirements.txt:1-4: torch>=1.13.0,
torchvision>=0.14.0, opencv-python==4.6.0.66, numpy==1.21.6); the model oracle therefore is a *functional* torch
restatement (conv2d / batch_norm / linear on plain tensors keyed by the reference's state_dict names) that can run in
fp32 or fp64 on the CPU. It is pinned two ways (see tests/test_oracle_*.py):
  1. against the reference classes themselves, AST-extracted from /root/reference in the build container;
  2. against golden vectors those classes produced (tests/golden/*.npz, generator committed next to them) — the
     reference ships no tests or fixtures of its own ("parity unpinned" by the reference).
"""
import ctypes
import math
import os

import numpy as np
import torch
import torch.nn.functional as F

IMG_MEAN = (0.485, 0.456, 0.406)
IMG_STD = (0.229, 0.224, 0.225)
IMG_WIDTH, IMG_HEIGHT = 200, 88
SPEED_NORM_FACTOR = 90.0
STAGES = ((4, 64, 3), (5, 128, 4), (6, 256, 6), (7, 512, 3))  # (visual_encoder index, channels, blocks)

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------------------------------
# preprocessing
# --------------------------------------------------------------------------------------------------
def _axis_coef(dsize, ssize):
    scale = float(ssize) / float(dsize)
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    lo = s < 0
    s[lo] = 0
    f[lo] = 0.0
    hi = s >= ssize - 1
    s[hi] = ssize - 1
    f[hi] = 0.0
    s1 = np.minimum(s + 1, ssize - 1)
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048.0)).astype(np.int32)
    a1 = np.rint(f * np.float32(2048.0)).astype(np.int32)
    return s, s1, a0, a1


def resize_u8_np(frames, dst_hw=(IMG_HEIGHT, IMG_WIDTH)):
    """cv2.resize(img, (w, h)) INTER_LINEAR on uint8 [B,H,W,C] (model/autonomous_drive.py:898): OpenCV's 11-bit
    fixed-point bilinear, restated with numpy integer arithmetic."""
    frames = np.asarray(frames)
    assert frames.dtype == np.uint8 and frames.ndim == 4
    dh, dw = dst_hw
    _, sh, sw, _ = frames.shape
    x0, x1, a0, a1 = _axis_coef(dw, sw)
    y0, y1, b0, b1 = _axis_coef(dh, sh)
    src = frames.astype(np.int32)
    r0 = src[:, y0]  # [B, dh, sw, C]
    r1 = src[:, y1]
    h0 = r0[:, :, x0] * a0[None, None, :, None] + r0[:, :, x1] * a1[None, None, :, None]
    h1 = r1[:, :, x0] * a0[None, None, :, None] + r1[:, :, x1] * a1[None, None, :, None]
    v = (((b0[None, :, None, None] * (h0 >> 4)) >> 16) + ((b1[None, :, None, None] * (h1 >> 4)) >> 16) + 2) >> 2
    return v.astype(np.uint8)


def normalise_np(small_u8):
    """/255 -> CHW -> (x-mean)/std in float32, in the reference's operation order (autonomous_drive.py:899-901)."""
    x = small_u8.astype(np.float32) / np.float32(255.0)
    x = np.transpose(x, (0, 3, 1, 2))
    mean = np.asarray(IMG_MEAN, dtype=np.float32)[None, :, None, None]
    std = np.asarray(IMG_STD, dtype=np.float32)[None, :, None, None]
    return ((x - mean) / std).astype(np.float32)


def preprocess_np(frames_rgb_u8):
    small = resize_u8_np(frames_rgb_u8[..., :3])
    return small, normalise_np(small)


_clib = None


def _c():
    global _clib
    if _clib is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
        _clib = ctypes.CDLL(path)
        _clib.oracle_preprocess.restype = ctypes.c_int
    return _clib


def preprocess_c(frames_u8, reverse=False, dst_hw=(IMG_HEIGHT, IMG_WIDTH)):
    """Same as preprocess_np through the plain-C restatement (oracle/resize_oracle.c)."""
    frames_u8 = np.ascontiguousarray(frames_u8)
    b, sh, sw, sc = frames_u8.shape
    dh, dw = dst_hw
    u8 = np.empty((b, dh, dw, 3), dtype=np.uint8)
    f32 = np.empty((b, 3, dh, dw), dtype=np.float32)
    rc = _c().oracle_preprocess(frames_u8.ctypes.data_as(ctypes.c_void_p), b, sh, sw, sc, int(reverse), dh, dw,
                                u8.ctypes.data_as(ctypes.c_void_p), f32.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return u8, f32


# --------------------------------------------------------------------------------------------------
# state_dict layout (250 keys) and a deterministic synthetic fill
# --------------------------------------------------------------------------------------------------
def state_dict_spec(num_commands=4):
    """[(key, shape, dtype_kind)] in the reference's state_dict order; kind in {'w','bn_w','bn_b','rm','rv','nbt','lin_w','lin_b'}."""
    spec = []

    def bn(prefix, c):
        spec.extend([(prefix + ".weight", (c,), "bn_w"), (prefix + ".bias", (c,), "bn_b"),
                     (prefix + ".running_mean", (c,), "rm"), (prefix + ".running_var", (c,), "rv"),
                     (prefix + ".num_batches_tracked", (), "nbt")])

    spec.append(("visual_encoder.0.weight", (64, 3, 7, 7), "w"))
    bn("visual_encoder.1", 64)
    cin = 64
    for idx, c, blocks in STAGES:
        for b in range(blocks):
            p = "visual_encoder.%d.%d" % (idx, b)
            spec.append((p + ".conv1.weight", (c, cin if b == 0 else c, 3, 3), "w"))
            bn(p + ".bn1", c)
            spec.append((p + ".conv2.weight", (c, c, 3, 3), "w"))
            bn(p + ".bn2", c)
            if b == 0 and cin != c:
                spec.append((p + ".downsample.0.weight", (c, cin, 1, 1), "w"))
                bn(p + ".downsample.1", c)
        cin = c

    def lin(prefix, o, i):
        spec.extend([(prefix + ".weight", (o, i), "lin_w"), (prefix + ".bias", (o,), "lin_b")])

    lin("speed_encoder.0", 128, 1)
    lin("speed_encoder.3", 128, 128)
    for k in range(num_commands):
        lin("control_branches.%d.0" % k, 256, 640)
        lin("control_branches.%d.3" % k, 256, 256)
        lin("control_branches.%d.6" % k, 3, 256)
    lin("speed_predictor.0", 256, 512)
    lin("speed_predictor.3", 256, 256)
    lin("speed_predictor.5", 1, 256)
    return spec


def synthetic_state_dict(seed=0, dtype=torch.float32, perturb_bn=True):
    """Deterministic weights with the reference initialisers' scales (kaiming-normal fan_out convs, U(+-1/sqrt(fan_in))
    linears — torchvision/models/resnet.py:208-213, torch/nn/modules/linear.py) but generated per key from `seed`, so the
    oracle, the reference class and the CUDA build can all be loaded with bit-identical fp32 values without shipping a
    checkpoint. perturb_bn draws non-trivial BN affine/running statistics so that eval-mode BN is actually exercised."""
    g = torch.Generator().manual_seed(1000003 * seed + 17)
    sd = {}
    for key, shape, kind in state_dict_spec():
        if kind == "w":
            fan_out = shape[0] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_out)
        elif kind == "lin_w":
            bound = 1.0 / math.sqrt(shape[1])
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "lin_b":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        elif kind == "bn_w":
            t = 1.0 + 0.2 * (torch.rand(shape, generator=g) - 0.5) if perturb_bn else torch.ones(shape)
        elif kind == "bn_b":
            t = 0.1 * (torch.rand(shape, generator=g) - 0.5) if perturb_bn else torch.zeros(shape)
        elif kind == "rm":
            t = 0.1 * torch.randn(shape, generator=g) if perturb_bn else torch.zeros(shape)
        elif kind == "rv":
            t = 0.5 + torch.rand(shape, generator=g) if perturb_bn else torch.ones(shape)
        else:
            sd[key] = torch.zeros((), dtype=torch.int64)
            continue
        sd[key] = t.to(dtype)
    return sd


def synthetic_batch(batch, seed=0, smooth=True):
    """Seeded synthetic inputs: raw uint8 frames [B,600,800,3], speed U[0,1), command randint(0,4),
    targets (steer U[-1,1), throttle/brake U[0,1))."""
    rng = np.random.default_rng(seed)
    if smooth:
        coarse = torch.from_numpy(rng.uniform(0, 255, size=(batch, 3, 12, 16)).astype(np.float32))
        up = F.interpolate(coarse, size=(600, 800), mode="bicubic", align_corners=False).clamp(0, 255)
        frames = up.permute(0, 2, 3, 1).round().to(torch.uint8).numpy()
    else:
        frames = rng.integers(0, 256, size=(batch, 600, 800, 3), dtype=np.uint8)
    speed = rng.uniform(0, 1, size=(batch,)).astype(np.float32)
    command = rng.integers(0, 4, size=(batch,)).astype(np.int64)
    targets = np.stack([rng.uniform(-1, 1, size=batch), rng.uniform(0, 1, size=batch), rng.uniform(0, 1, size=batch)],
                       axis=1).astype(np.float32)
    return frames, speed, command, targets


# --------------------------------------------------------------------------------------------------
# the model, functionally (state dict in, tensors out)
# --------------------------------------------------------------------------------------------------
def _bn(sd, prefix, x, training, momentum=0.1, eps=1e-5, update=None):
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training and update is not None:
        # F.batch_norm updates the running statistics in place; keep the caller's dict untouched and report them
        rm, rv = rm.clone(), rv.clone()
        y = F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], True, momentum, eps)
        update[prefix + ".running_mean"] = rm
        update[prefix + ".running_var"] = rv
        update[prefix + ".num_batches_tracked"] = sd[prefix + ".num_batches_tracked"] + 1
        return y
    return F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], training, momentum, eps)


def block_prefixes():
    """state_dict prefixes of the 16 BasicBlocks in forward order, with their stride."""
    return [("visual_encoder.%d.%d" % (idx, b), 2 if (b == 0 and idx != 4) else 1) for idx, c, blocks in STAGES for b in range(blocks)]


def basic_block(sd, p, x, stride, training=False, update=None, rnd=None):
    """torchvision BasicBlock.forward (torchvision/models/resnet.py:89-105) on state-dict entries with prefix `p`.
    rnd: optional rounding hook applied where the CUDA path stores a bf16 tensor (see forward_bf16emu)."""
    r = rnd if rnd is not None else (lambda t: t)
    out = r(F.conv2d(x, sd[p + ".conv1.weight"], stride=stride, padding=1))
    out = r(F.relu(_bn(sd, p + ".bn1", out, training, update=update)))
    out = r(F.conv2d(out, sd[p + ".conv2.weight"], stride=1, padding=1))
    out = _bn(sd, p + ".bn2", out, training, update=update)
    if (p + ".downsample.0.weight") in sd:
        idn = r(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride))
        idn = _bn(sd, p + ".downsample.1", idn, training, update=update)
    else:
        idn = x
    return r(F.relu(out + idn))


def stem(sd, image, training=False, update=None, rnd=None):
    """resnet34 conv1 + bn1 + relu + maxpool (model/autonomous_drive.py:367)."""
    r = rnd if rnd is not None else (lambda t: t)
    x = r(F.conv2d(image, sd["visual_encoder.0.weight"], stride=2, padding=3))
    x = F.relu(_bn(sd, "visual_encoder.1", x, training, update=update))
    return r(F.max_pool2d(x, kernel_size=3, stride=2, padding=1))


def visual_encoder(sd, image, training=False, update=None, taps=None, rnd=None):
    """resnet34 conv1..avgpool + Flatten (model/autonomous_drive.py:366-370)."""
    x = stem(sd, image, training, update, rnd)
    if taps is not None:
        taps["stem"] = x
    for p, stride in block_prefixes():
        x = basic_block(sd, p, x, stride, training, update, rnd)
        if taps is not None:
            taps[p] = x
    return x.mean(dim=(2, 3))


class _RoundBF16(torch.autograd.Function):
    """Round to bf16 (nearest even) in the forward AND the backward pass: the value and its gradient are both tensors the CUDA
    path stores in bf16."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def round_bf16(x):
    return _RoundBF16.apply(x)


def bf16_weights(sd):
    """Copy of `sd` whose CONVOLUTION weights are rounded to bf16 (the tensor-core operands; BN, heads and everything else
    stay as they are). Gradients w.r.t. the rounded weights are what the CUDA path's fp32 weight gradients approximate."""
    out = {}
    for k, v in sd.items():
        if v.dim() == 4:
            out[k] = v.detach().to(torch.bfloat16).to(v.dtype)
            if v.requires_grad:
                out[k].requires_grad_(True)
        else:
            out[k] = v
    return out


def forward_bf16emu(sd, image, speed, command, training=False, update=None, taps=None):
    """CILRS.forward in the precision of `sd` (fp64 for tests) with bf16 rounding inserted exactly where the CUDA path stores
    bf16 tensors: the input image, every raw convolution output, every post-BN/ReLU activation, the max-pool output and every
    block output — forward values and (through _RoundBF16.backward) the gradients that flow through the same tensors.
    `sd` should come from bf16_weights(). The heads, BN statistics, losses stay in full precision (fp32 in the CUDA path)."""
    return forward(sd, round_bf16(image), speed, command, training, update, taps=taps, rnd=round_bf16)


def forward(sd, image, speed, command, training=False, update=None, num_commands=4, taps=None, rnd=None):
    """CILRS.forward (model/autonomous_drive.py:389-399); dropout p = 0 (the parity configuration)."""
    vf = visual_encoder(sd, image, training, update, taps, rnd)
    s = F.relu(F.linear(speed.unsqueeze(1), sd["speed_encoder.0.weight"], sd["speed_encoder.0.bias"]))
    s = F.relu(F.linear(s, sd["speed_encoder.3.weight"], sd["speed_encoder.3.bias"]))
    combined = torch.cat([vf, s], dim=1)
    p = F.relu(F.linear(vf, sd["speed_predictor.0.weight"], sd["speed_predictor.0.bias"]))
    p = F.relu(F.linear(p, sd["speed_predictor.3.weight"], sd["speed_predictor.3.bias"]))
    pred_speed = F.linear(p, sd["speed_predictor.5.weight"], sd["speed_predictor.5.bias"]).squeeze(1)
    outs = []
    for k in range(num_commands):
        pre = "control_branches.%d" % k
        h = F.relu(F.linear(combined, sd[pre + ".0.weight"], sd[pre + ".0.bias"]))
        h = F.relu(F.linear(h, sd[pre + ".3.weight"], sd[pre + ".3.bias"]))
        outs.append(F.linear(h, sd[pre + ".6.weight"], sd[pre + ".6.bias"]))
    allo = torch.stack(outs, dim=0)
    idx = command.unsqueeze(0).unsqueeze(2).expand(1, image.size(0), 3)
    controls = allo.gather(0, idx).squeeze(0)
    return controls, pred_speed


# --------------------------------------------------------------------------------------------------
# losses and optimiser
# --------------------------------------------------------------------------------------------------
def loss_l1(pred_controls, target_controls, pred_speed, target_speed, steer_w=5.0, throttle_w=1.0, brake_w=1.0, speed_w=0.5):
    """CILRSLoss.forward (notebook/notebook.ipynb:514-527). Returns (total, dict of the 6 components)."""
    steer = (pred_controls[:, 0] - target_controls[:, 0]).abs().mean()
    throttle = (pred_controls[:, 1] - target_controls[:, 1]).abs().mean()
    brake = (pred_controls[:, 2] - target_controls[:, 2]).abs().mean()
    control = steer_w * steer + throttle_w * throttle + brake_w * brake
    speed = ((pred_speed - target_speed) ** 2).mean()
    total = control + speed_w * speed
    return total, {"total": total, "control": control, "steer": steer, "throttle": throttle, "brake": brake, "speed": speed}


def loss_mse(pred_controls, target_controls, pred_speed, target_speed, speed_w=0.05):
    """README/config recipe (configs/train_config.json:30-32): MSE(controls) + 0.05 MSE(speed) — the BASELINE recipe."""
    control = ((pred_controls - target_controls) ** 2).mean()
    speed = ((pred_speed - target_speed) ** 2).mean()
    total = control + speed_w * speed
    return total, {"total": total, "control": control, "speed": speed}


def adam_step(p, g, m, v, step, lr=2e-4, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-4):
    """One torch.optim.Adam step (notebook/notebook.ipynb:533-534; torch/optim/adam.py `_single_tensor_adam`):
    L2-coupled decay, bias-corrected, eps added outside the sqrt. Operates on numpy or torch arrays, returns new (p, m, v)."""
    g = g + weight_decay * p
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v ** 0.5) / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def normalise_speed(speed_kmh):
    """predict_controls (model/autonomous_drive.py:910): min(speed / 90, 1.0)."""
    return min(speed_kmh / SPEED_NORM_FACTOR, 1.0)


def params_in_order(sd):
    """Parameter tensors (no buffers) in named_parameters() order == state_dict order minus BN buffers."""
    return [(k, sd[k]) for k, _, kind in state_dict_spec() if kind not in ("rm", "rv", "nbt")]
