"""GPU parity of the tcgen05 implicit-GEMM convolution kernels (fprop / dgrad / wgrad / stem) through the C-ABI.

The checker is a plain fp32 torch convolution of the same bf16-rounded operands (reference call site:
torchvision resnet34 convs used by CILRS.visual_encoder, model/autonomous_drive.py:365-369). Tolerance: the kernel
accumulates bf16 products in fp32, so it must match the fp32 result to accumulation-order noise; outputs stored in
bf16 add one rounding (2^-9 relative).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from cilrs_b200 import ops
    return ops


def _ref_setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _mk(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale)


def _nhwc_bf16(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _report(name, got, ref, tol):
    got = got.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-20
    rel = err.max().item() / denom
    if rel > tol:
        bad = (err > tol * denom)
        idx = bad.nonzero()
        print(f"[{name}] rel={rel:.3e} tol={tol:.1e} bad={bad.sum().item()}/{bad.numel()} shape={tuple(got.shape)}")
        print("  first bad idx:", idx[:8].tolist())
        for d in range(got.dim()):
            dims = [i for i in range(got.dim()) if i != d]
            cnt = bad.sum(dim=dims)
            nz = cnt.nonzero().flatten().tolist()
            print(f"  dim{d}: bad positions {nz[:40]}{'...' if len(nz) > 40 else ''}")
        print("  got:", got.flatten()[:8].tolist(), " ref:", ref.flatten()[:8].tolist())
    assert rel <= tol, f"{name}: rel err {rel:.3e} > {tol:.1e}"
    return rel


FPROP_CASES = [
    # batch, H, W, Cin, Cout, k, stride
    (2, 8, 16, 64, 64, 1, 1),
    (5, 22, 50, 64, 64, 3, 1),
    (3, 11, 25, 128, 128, 3, 1),
    (9, 6, 13, 256, 256, 3, 1),
    (7, 3, 7, 512, 512, 3, 1),
    (5, 22, 50, 64, 128, 3, 2),
    (5, 22, 50, 64, 128, 1, 2),
    (4, 11, 25, 128, 256, 3, 2),
    (13, 6, 13, 256, 512, 3, 2),
    (1, 22, 50, 64, 64, 3, 1),
    (1, 3, 7, 512, 512, 3, 1),
    (40, 11, 25, 128, 128, 3, 1),
]


@pytest.mark.parametrize("case", FPROP_CASES)
def test_fprop_matches_fp32_conv(case):
    ops = _ops()
    _ref_setup()
    b, h, w, ci, co, k, s = case
    d = ops.conv_desc(b, h, w, ci, co, k, s)
    x = _mk((b, ci, h, w), 1).to(torch.bfloat16)
    wt = (_mk((co, ci, k, k), 2) * (2.0 / (ci * k * k)) ** 0.5)
    wf, _ = ops.pack_weight(d, wt)
    y = ops.conv_fprop(d, _nhwc_bf16(x.float()), wf)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float(), wt.to(torch.bfloat16).float(), stride=s, padding=d.pad).permute(0, 2, 3, 1)
    _report(f"fprop{case}", y, ref, 1.2e-2)


def test_fprop_stats_and_fused_epilogue():
    ops = _ops()
    _ref_setup()
    b, h, w, ci, co, k, s = 6, 11, 25, 128, 128, 3, 1
    d = ops.conv_desc(b, h, w, ci, co, k, s)
    x = _mk((b, ci, h, w), 3).to(torch.bfloat16)
    wt = _mk((co, ci, k, k), 4) * (2.0 / (ci * 9)) ** 0.5
    wf, _ = ops.pack_weight(d, wt)
    xn = _nhwc_bf16(x.float())
    y, st = ops.conv_fprop(d, xn, wf, stats=True)
    torch.cuda.synchronize()
    yf = y.float()
    ssum = st[:, 0].sum(0)
    ssq = st[:, 1].sum(0)
    _report("stats.sum", ssum, yf.sum((0, 1, 2)), 1e-3)
    _report("stats.sumsq", ssq, (yf * yf).sum((0, 1, 2)), 1e-3)
    scale = _mk((co,), 5).abs() + 0.5
    bias = _mk((co,), 6)
    res = _mk((b, h, w, co), 7).to(torch.bfloat16)
    y2 = ops.conv_fprop(d, xn, wf, scale=scale, bias=bias, residual=res, relu=True)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float(), wt.to(torch.bfloat16).float(), stride=s, padding=1).permute(0, 2, 3, 1)
    ref = torch.relu(ref * scale + bias + res.float())
    _report("fused-epilogue", y2, ref, 1.2e-2)


@pytest.mark.parametrize("batch", [1, 5, 12])
def test_stem_fprop(batch):
    ops = _ops()
    _ref_setup()
    img = _mk((batch, 3, 88, 200), 8)
    wt = _mk((64, 3, 7, 7), 9) * (2.0 / 147) ** 0.5
    xs = ops.image_to_s2d(img)
    wp = ops.stem_pack_weight(wt)
    y = ops.stem_fprop(xs, wp)
    torch.cuda.synchronize()
    ref = F.conv2d(img.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float(), stride=2, padding=3).permute(0, 2, 3, 1)
    _report(f"stem{batch}", y, ref, 1.2e-2)


DGRAD_CASES = [
    (5, 22, 50, 64, 64, 3, 1),
    (3, 11, 25, 128, 128, 3, 1),
    (7, 3, 7, 512, 512, 3, 1),
    (5, 22, 50, 64, 128, 3, 2),
    (4, 11, 25, 128, 256, 3, 2),
    (5, 6, 13, 256, 512, 3, 2),
    (5, 22, 50, 64, 128, 1, 2),
    (3, 11, 25, 128, 256, 1, 2),
]


@pytest.mark.parametrize("case", DGRAD_CASES)
def test_dgrad(case):
    ops = _ops()
    _ref_setup()
    b, h, w, ci, co, k, s = case
    d = ops.conv_desc(b, h, w, ci, co, k, s)
    oh, ow = ops.out_hw(d)
    dy = _mk((b, co, oh, ow), 10).to(torch.bfloat16)
    wt = _mk((co, ci, k, k), 11) * (2.0 / (co * k * k)) ** 0.5
    _, wd = ops.pack_weight(d, wt)
    dx = ops.conv_dgrad(d, _nhwc_bf16(dy.float()), wd)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_input((b, ci, h, w), wt.to(torch.bfloat16).float(), dy.float(), stride=s, padding=d.pad)
    _report(f"dgrad{case}", dx, ref.permute(0, 2, 3, 1), 1.2e-2)


WGRAD_CASES = [
    (5, 22, 50, 64, 64, 3, 1),
    (12, 11, 25, 128, 128, 3, 1),
    (20, 6, 13, 256, 256, 3, 1),
    (9, 3, 7, 512, 512, 3, 1),
    (5, 22, 50, 64, 128, 3, 2),
    (4, 11, 25, 128, 256, 3, 2),
    (5, 22, 50, 64, 128, 1, 2),
]


@pytest.mark.parametrize("case", WGRAD_CASES)
def test_wgrad(case):
    ops = _ops()
    _ref_setup()
    b, h, w, ci, co, k, s = case
    d = ops.conv_desc(b, h, w, ci, co, k, s)
    oh, ow = ops.out_hw(d)
    x = _mk((b, ci, h, w), 12).to(torch.bfloat16)
    dy = (_mk((b, co, oh, ow), 13) * 0.1).to(torch.bfloat16)
    dw = ops.conv_wgrad(d, _nhwc_bf16(dy.float()), _nhwc_bf16(x.float()))
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x.float(), (co, ci, k, k), dy.float(), stride=s, padding=d.pad)
    _report(f"wgrad{case}", dw, ref, 2e-3)


@pytest.mark.parametrize("batch", [2, 7])
def test_stem_wgrad(batch):
    ops = _ops()
    _ref_setup()
    img = _mk((batch, 3, 88, 200), 14)
    dy = (_mk((batch, 64, 44, 100), 15) * 0.1).to(torch.bfloat16)
    xs = ops.image_to_s2d(img)
    dw = ops.stem_wgrad(_nhwc_bf16(dy.float()), xs)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(img.to(torch.bfloat16).float(), (64, 3, 7, 7), dy.float(), stride=2, padding=3)
    _report(f"stem_wgrad{batch}", dw, ref, 2e-3)
