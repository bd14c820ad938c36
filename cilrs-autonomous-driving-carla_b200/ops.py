"""Thin tensor-level wrappers over the single-kernel C-ABI entry points (used by tests and tools; the model path
uses the whole-network plan in model.py). All tensors must be CUDA tensors; nothing here has a CPU path."""
import torch

from . import _lib
from ._lib import ConvDesc

EPI_STATS, EPI_SCALE_BIAS, EPI_RESIDUAL, EPI_RELU = 1, 2, 4, 8
EPI_MASK, EPI_BNBWD, EPI_BNBWD2, EPI_DEFER = 16, 32, 64, 128


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("cilrs_b200 ops run on CUDA tensors only (no CPU fallback)")


def conv_desc(batch, in_h, in_w, in_c, out_c, k, stride):
    return ConvDesc(batch, in_h, in_w, in_c, out_c, k, k, stride, 1 if k == 3 else 0)


def out_hw(d):
    return ((d.in_h + 2 * d.pad - d.kh) // d.stride + 1, (d.in_w + 2 * d.pad - d.kw) // d.stride + 1)


def pack_weight(d, w_oihw):
    """fp32 OIHW -> (bf16 [tap][O][I] for fprop, bf16 [tap][I][O] for dgrad)."""
    _need_cuda(w_oihw)
    n = d.kh * d.kw * d.in_c * d.out_c
    wf = torch.empty(n, dtype=torch.bfloat16, device=w_oihw.device)
    wd = torch.empty(n, dtype=torch.bfloat16, device=w_oihw.device)
    _lib.call("cilrs_conv_pack_weight", d, w_oihw.contiguous(), wf, wd, _lib.stream_ptr())
    return wf, wd


def conv_fprop(d, x, wf, scale=None, bias=None, residual=None, stats=False, relu=False):
    """x: bf16 NHWC [B,H,W,Cin] -> bf16 NHWC [B,OH,OW,Cout] (+ stats partials [tiles,2,Cout] if stats)."""
    _need_cuda(x, wf)
    oh, ow = out_hw(d)
    y = torch.empty(d.batch, oh, ow, d.out_c, dtype=torch.bfloat16, device=x.device)
    flags = 0
    st = None
    if stats:
        flags |= EPI_STATS
        st = torch.zeros(_lib.query("cilrs_conv_stats_tiles", d), 2, d.out_c, dtype=torch.float32, device=x.device)
    if scale is not None:
        flags |= EPI_SCALE_BIAS
    if residual is not None:
        flags |= EPI_RESIDUAL
    if relu:
        flags |= EPI_RELU
    _lib.call("cilrs_conv_fprop", d, x, wf, y, scale, bias, residual, st, flags, _lib.stream_ptr())
    return (y, st) if stats else y


def conv_dgrad(d, dy, wd, residual=None):
    _need_cuda(dy, wd)
    dx = torch.zeros(d.batch, d.in_h, d.in_w, d.in_c, dtype=torch.bfloat16, device=dy.device)
    _lib.call("cilrs_conv_dgrad", d, dy, wd, dx, residual, _lib.stream_ptr())
    return dx


def conv_wgrad(d, dy, x):
    _need_cuda(dy, x)
    dw = torch.zeros(d.out_c, d.in_c, d.kh, d.kw, dtype=torch.float32, device=dy.device)
    _lib.call("cilrs_conv_wgrad", d, dy, x, dw, _lib.stream_ptr())
    return dw


def to_padded(x):
    """dense NHWC [B,H,W,C] -> padded-flat layout [B,H+1,W+1,C] (last row / column of every image zero)."""
    b, h, w, c = x.shape
    out = torch.zeros(b, h + 1, w + 1, c, dtype=x.dtype, device=x.device)
    out[:, :h, :w] = x
    return out


def from_padded(xp, h, w):
    return xp[:, :h, :w].contiguous()


def relu_bits(act_pad):
    """[B,Hp,Wp,C] activation -> uint8 [B,Hp,Wp,C/8], bit k of byte j = (act[..., 8j+k] > 0) (what cilrs_bn_apply emits)."""
    b, hp, wp, c = act_pad.shape
    m = (act_pad.float() > 0).view(b, hp, wp, c // 8, 8).to(torch.int32)
    w = (2 ** torch.arange(8, device=act_pad.device, dtype=torch.int32)).view(1, 1, 1, 1, 8)
    return (m * w).sum(-1).to(torch.uint8).contiguous()


def conv_flat(x_pad, w_pack, out_c, dgrad=False, scale=None, bias=None, residual=None, mask=None, relu=False,
              bn=None, bnbwd=None, bnbwd2=None, mask_bits=None, defer_sums=None):
    """3x3 stride-1 conv (fprop or dgrad) on padded-flat tensors. x_pad [B,H+1,W+1,Cin] bf16 -> [B,H+1,W+1,out_c].
    bn = dict(gamma, beta, running_mean, running_var, nbt, update) -> also returns vec [4,out_c] (fused train-mode BN
    statistics + finalize). bnbwd = dict(y, vec, dgamma, dbeta) -> also returns bred [2,out_c] (fused BN-backward reduce).
    defer_sums = "stats" | "bnbwd" (with bnbwd=dict(y=...), bnbwd2=dict(y=...)): deferred finalize - only the raw per-channel
    fp64 sums [3,out_c] are produced (CILRS_EPI_DEFER) and returned instead of vec / bred."""
    _need_cuda(x_pad, w_pack)
    b, hp, wp, cin = x_pad.shape
    dev = x_pad.device
    y = torch.full((b, hp, wp, out_c), float("nan"), dtype=torch.bfloat16, device=dev)
    a = _lib.FlatConvArgs()
    a.batch, a.H, a.W, a.in_c, a.out_c, a.dgrad = b, hp - 1, wp - 1, cin, out_c, int(dgrad)
    keep = [x_pad, w_pack, y]
    ptr = lambda t: None if t is None else t.data_ptr()
    a.x, a.w, a.y = ptr(x_pad), ptr(w_pack), ptr(y)
    flags = 0
    if scale is not None:
        flags |= EPI_SCALE_BIAS
        a.scale, a.bias = ptr(scale), ptr(bias)
    if residual is not None:
        flags |= EPI_RESIDUAL
        a.residual = ptr(residual)
    if mask is not None:
        flags |= EPI_MASK
        a.mask = ptr(mask)
        a.mask_bits = ptr(mask_bits)
    if relu:
        flags |= EPI_RELU
    extra = []
    if defer_sums is not None:
        acc = torch.zeros(3, out_c, dtype=torch.float64, device=dev)
        a.partials_ws = ptr(acc)
        keep.append(acc)
        flags |= EPI_DEFER | (EPI_STATS if defer_sums == "stats" else EPI_BNBWD)
        if defer_sums != "stats":
            a.y1 = ptr(bnbwd["y"])
            if bnbwd2 is not None:
                flags |= EPI_BNBWD2
                a.y2 = ptr(bnbwd2["y"])
        a.flags = flags
        _lib.call("cilrs_conv_flat", a, _lib.stream_ptr())
        return y, acc
    if bn is not None or bnbwd is not None:
        ws = torch.zeros(_lib.query("cilrs_conv_flat_workspace_floats", out_c), dtype=torch.float32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        a.partials_ws, a.counter_ws = ptr(ws), ptr(cnt)
        keep += [ws, cnt]
    if bn is not None:
        flags |= EPI_STATS
        vec = torch.empty(4, out_c, dtype=torch.float32, device=dev)
        a.gamma, a.beta, a.running_mean, a.running_var = ptr(bn["gamma"]), ptr(bn["beta"]), ptr(bn["running_mean"]), ptr(bn["running_var"])
        a.num_batches_tracked = ptr(bn.get("nbt"))
        a.vec, a.momentum, a.eps, a.update_running = ptr(vec), 0.1, 1e-5, int(bn.get("update", 1))
        extra.append(vec)
    if bnbwd is not None:
        flags |= EPI_BNBWD
        bred = torch.empty(2, out_c, dtype=torch.float32, device=dev)
        a.y1, a.vec1, a.bred1, a.dgamma1, a.dbeta1 = ptr(bnbwd["y"]), ptr(bnbwd["vec"]), ptr(bred), ptr(bnbwd.get("dgamma")), ptr(bnbwd.get("dbeta"))
        extra.append(bred)
        if bnbwd2 is not None:
            flags |= EPI_BNBWD2
            bred2 = torch.empty(2, out_c, dtype=torch.float32, device=dev)
            a.y2, a.vec2, a.bred2, a.dgamma2, a.dbeta2 = ptr(bnbwd2["y"]), ptr(bnbwd2["vec"]), ptr(bred2), ptr(bnbwd2.get("dgamma")), ptr(bnbwd2.get("dbeta"))
            extra.append(bred2)
    a.flags = flags
    _lib.call("cilrs_conv_flat", a, _lib.stream_ptr())
    return (y, *extra) if extra else y


def wgrad_flat(dy_pad, x_pad):
    _need_cuda(dy_pad, x_pad)
    b, hp, wp, cout = dy_pad.shape
    cin = x_pad.shape[3]
    dw = torch.zeros(cout, cin, 3, 3, dtype=torch.float32, device=dy_pad.device)
    ws = torch.zeros(_lib.query("cilrs_wgrad_flat_workspace_bytes") // 4, dtype=torch.float32, device=dy_pad.device)
    _lib.call("cilrs_wgrad_flat", b, hp - 1, wp - 1, cin, cout, dy_pad, x_pad, dw, ws, _lib.stream_ptr())
    return dw


def image_to_s2d(image):
    _need_cuda(image)
    b = image.shape[0]
    out = torch.empty(b, 47, 103, 16, dtype=torch.bfloat16, device=image.device)
    _lib.call("cilrs_image_to_s2d", image.contiguous(), b, out, _lib.stream_ptr())
    return out


def stem_pack_weight(w):
    wp = torch.empty(4 * 64 * 64, dtype=torch.bfloat16, device=w.device)
    _lib.call("cilrs_stem_pack_weight", w.contiguous(), wp, _lib.stream_ptr())
    return wp


def stem_fprop(x_s2d, wp, scale=None, bias=None, stats=False, relu=False):
    b = x_s2d.shape[0]
    y = torch.empty(b, 44, 100, 64, dtype=torch.bfloat16, device=x_s2d.device)
    flags = 0
    st = None
    if stats:
        flags |= EPI_STATS
        st = torch.zeros(_lib.query("cilrs_stem_stats_tiles", b), 2, 64, dtype=torch.float32, device=x_s2d.device)
    if scale is not None:
        flags |= EPI_SCALE_BIAS
    if relu:
        flags |= EPI_RELU
    _lib.call("cilrs_stem_fprop", b, x_s2d, wp, y, scale, bias, st, flags, _lib.stream_ptr())
    return (y, st) if stats else y


def stem_wgrad(dy, x_s2d):
    dw = torch.zeros(64, 3, 7, 7, dtype=torch.float32, device=dy.device)
    _lib.call("cilrs_stem_wgrad", dy.shape[0], dy, x_s2d, dw, _lib.stream_ptr())
    return dw


def preprocess(frames_u8, reverse=False, dst_hw=(88, 200), want_u8=False, want_f32=True, want_s2d=False):
    """uint8 [B,H,W,C] -> dict(u8=[B,h,w,3] uint8, f32=[B,3,h,w] float32, s2d=[B,47,103,16] bf16)."""
    _need_cuda(frames_u8)
    if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4:
        raise ValueError("frames must be uint8 [B,H,W,C]")
    frames_u8 = frames_u8.contiguous()
    b, h, w, c = frames_u8.shape
    dh, dw = dst_hw
    dev = frames_u8.device
    u8 = torch.empty(b, dh, dw, 3, dtype=torch.uint8, device=dev) if want_u8 else None
    f32 = torch.empty(b, 3, dh, dw, dtype=torch.float32, device=dev) if want_f32 else None
    s2d = torch.empty(b, 47, 103, 16, dtype=torch.bfloat16, device=dev) if want_s2d else None
    _lib.call("cilrs_preprocess_u8", frames_u8, b, h, w, c, int(reverse), dh, dw, u8, f32, s2d, _lib.stream_ptr())
    return {"u8": u8, "f32": f32, "s2d": s2d}
