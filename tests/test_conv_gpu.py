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


# ---------------------------------------------------------------------------------------------------------------
# padded-flat kernels (3x3 stride-1): fprop / dgrad / wgrad, fused BatchNorm statistics + finalize, fused ReLU mask +
# BatchNorm-backward reductions. Same checker (fp32 torch conv of the same bf16-rounded operands).
# ---------------------------------------------------------------------------------------------------------------
FLAT_CASES = [
    # batch, H, W, Cin, Cout
    (5, 22, 50, 64, 64),
    (3, 11, 25, 128, 128),
    (9, 6, 13, 256, 256),
    (7, 3, 7, 512, 512),
    (1, 22, 50, 64, 64),
    (1, 3, 7, 512, 512),
    (40, 11, 25, 128, 128),
    (128, 22, 50, 64, 64),
    (128, 6, 13, 256, 256),
    (4, 9, 10, 64, 128),
]


def _pads_are_zero(yp, h, w):
    return float(yp[:, h:, :].float().abs().max()) == 0.0 and float(yp[:, :, w:].float().abs().max()) == 0.0


@pytest.mark.parametrize("case", FLAT_CASES)
def test_flat_fprop(case):
    ops = _ops()
    _ref_setup()
    b, h, w, ci, co = case
    d = ops.conv_desc(b, h, w, ci, co, 3, 1)
    x = _mk((b, ci, h, w), 21).to(torch.bfloat16)
    wt = _mk((co, ci, 3, 3), 22) * (2.0 / (ci * 9)) ** 0.5
    wf, _ = ops.pack_weight(d, wt)
    yp = ops.conv_flat(ops.to_padded(_nhwc_bf16(x.float())), wf, co)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float(), wt.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 1)
    _report(f"flat_fprop{case}", ops.from_padded(yp, h, w), ref, 1.2e-2)
    assert _pads_are_zero(yp, h, w)


@pytest.mark.parametrize("case", FLAT_CASES[:7])
def test_flat_dgrad(case):
    ops = _ops()
    _ref_setup()
    b, h, w, ci, co = case
    d = ops.conv_desc(b, h, w, ci, co, 3, 1)
    dy = _mk((b, co, h, w), 23).to(torch.bfloat16)
    wt = _mk((co, ci, 3, 3), 24) * (2.0 / (co * 9)) ** 0.5
    _, wd = ops.pack_weight(d, wt)
    dxp = ops.conv_flat(ops.to_padded(_nhwc_bf16(dy.float())), wd, ci, dgrad=True)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_input((b, ci, h, w), wt.to(torch.bfloat16).float(), dy.float(), padding=1)
    _report(f"flat_dgrad{case}", ops.from_padded(dxp, h, w), ref.permute(0, 2, 3, 1), 1.2e-2)
    assert _pads_are_zero(dxp, h, w)


@pytest.mark.parametrize("case", [(5, 22, 50, 64, 64), (12, 11, 25, 128, 128), (20, 6, 13, 256, 256), (9, 3, 7, 512, 512),
                                  (128, 11, 25, 128, 128), (3, 9, 10, 64, 128)])
def test_flat_wgrad(case):
    ops = _ops()
    _ref_setup()
    from cilrs_b200 import _lib
    b, h, w, ci, co = case
    x = _mk((b, ci, h, w), 25).to(torch.bfloat16)
    dy = (_mk((b, co, h, w), 26) * 0.1).to(torch.bfloat16)
    ref = torch.nn.grad.conv2d_weight(x.float(), (co, ci, 3, 3), dy.float(), padding=1)
    # the default kernel (one CTA per filter row) and the opt-in cluster-of-three kernel with multicast loads
    for cluster in (0, 1):
        prev = _lib.lib().cilrs_set_wgrad_cluster(cluster)
        try:
            dw = ops.wgrad_flat(ops.to_padded(_nhwc_bf16(dy.float())), ops.to_padded(_nhwc_bf16(x.float())))
            torch.cuda.synchronize()
        finally:
            _lib.lib().cilrs_set_wgrad_cluster(prev)
        _report(f"flat_wgrad{case} cluster={cluster}", dw, ref, 2e-3)


def test_flat_inference_epilogue():
    ops = _ops()
    _ref_setup()
    b, h, w, c = 6, 11, 25, 128
    d = ops.conv_desc(b, h, w, c, c, 3, 1)
    x = _mk((b, c, h, w), 27).to(torch.bfloat16)
    wt = _mk((c, c, 3, 3), 28) * (2.0 / (c * 9)) ** 0.5
    wf, _ = ops.pack_weight(d, wt)
    scale = _mk((c,), 29).abs() + 0.5
    bias = _mk((c,), 30)
    res = _mk((b, h, w, c), 31).to(torch.bfloat16)
    yp = ops.conv_flat(ops.to_padded(_nhwc_bf16(x.float())), wf, c, scale=scale, bias=bias, residual=ops.to_padded(res), relu=True)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float(), wt.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 1)
    ref = torch.relu(ref * scale + bias + res.float())
    _report("flat fused-epilogue", ops.from_padded(yp, h, w), ref, 1.2e-2)
    assert _pads_are_zero(yp, h, w)


@pytest.mark.parametrize("case", [(6, 11, 25, 128), (128, 22, 50, 64), (33, 6, 13, 256), (128, 3, 7, 512)])
def test_flat_fprop_fused_bn_statistics(case):
    """train-mode BatchNorm statistics + finalize inside the conv kernel == torch batch_norm statistics of its output"""
    ops = _ops()
    _ref_setup()
    b, h, w, c = case
    d = ops.conv_desc(b, h, w, c, c, 3, 1)
    x = _mk((b, c, h, w), 32).to(torch.bfloat16)
    wt = _mk((c, c, 3, 3), 33) * (2.0 / (c * 9)) ** 0.5
    wf, _ = ops.pack_weight(d, wt)
    gamma = 1 + 0.3 * _mk((c,), 34)
    beta = 0.2 * _mk((c,), 35)
    rm, rv = _mk((c,), 36), _mk((c,), 37).abs() + 0.5
    rm0, rv0 = rm.clone(), rv.clone()
    nbt = torch.zeros(1, dtype=torch.long, device="cuda")
    for rep in range(2):  # twice: the kernel must leave its counter / workspace reusable
        yp, vec = ops.conv_flat(ops.to_padded(_nhwc_bf16(x.float())), wf, c,
                                bn=dict(gamma=gamma, beta=beta, running_mean=rm, running_var=rv, nbt=nbt, update=1 if rep == 0 else 0))
    torch.cuda.synchronize()
    y = ops.from_padded(yp, h, w).float()
    mean = y.mean((0, 1, 2))
    var = y.var((0, 1, 2), unbiased=False)
    rstd = 1.0 / torch.sqrt(var + 1e-5)
    _report("bn mean", vec[2], mean, 2e-4)
    _report("bn rstd", vec[3], rstd, 1e-4)
    _report("bn scale", vec[0], gamma * rstd, 1e-4)
    _report("bn shift", vec[1], beta - mean * gamma * rstd, 2e-4)
    n = b * h * w
    _report("running_mean", rm, 0.9 * rm0 + 0.1 * mean, 1e-4)
    _report("running_var", rv, 0.9 * rv0 + 0.1 * var * n / (n - 1), 1e-4)
    assert int(nbt) == 1


@pytest.mark.parametrize("case", [(6, 11, 25, 128, False), (128, 22, 50, 64, True), (20, 6, 13, 256, True), (128, 3, 7, 512, False)])
def test_flat_dgrad_fused_relu_mask_and_bn_backward_reduce(case):
    """dz = (dgrad + residual) * (act > 0) and the BatchNorm-backward reductions sum(dz), sum(dz * xhat) in one kernel"""
    ops = _ops()
    _ref_setup()
    b, h, w, c, two = case
    d = ops.conv_desc(b, h, w, c, c, 3, 1)
    dy = _mk((b, c, h, w), 40).to(torch.bfloat16)
    wt = _mk((c, c, 3, 3), 41) * (2.0 / (c * 9)) ** 0.5
    _, wd = ops.pack_weight(d, wt)
    res = _mk((b, h, w, c), 42).to(torch.bfloat16)
    act = torch.relu(_mk((b, h, w, c), 43)).to(torch.bfloat16)
    y1 = _mk((b, h, w, c), 44).to(torch.bfloat16)
    y2 = _mk((b, h, w, c), 45).to(torch.bfloat16)

    def mkvec(y):
        yf = y.float()
        mean = yf.mean((0, 1, 2))
        rstd = 1.0 / torch.sqrt(yf.var((0, 1, 2), unbiased=False) + 1e-5)
        return torch.stack([torch.ones_like(mean), torch.zeros_like(mean), mean, rstd]).contiguous()

    v1, v2 = mkvec(y1), mkvec(y2)
    dg1, db1 = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    dg2, db2 = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    out = ops.conv_flat(ops.to_padded(_nhwc_bf16(dy.float())), wd, c, dgrad=True, residual=ops.to_padded(res), mask=ops.to_padded(act),
                        bnbwd=dict(y=ops.to_padded(y1), vec=v1, dgamma=dg1, dbeta=db1),
                        bnbwd2=dict(y=ops.to_padded(y2), vec=v2, dgamma=dg2, dbeta=db2) if two else None)
    torch.cuda.synchronize()
    dzp, bred1 = out[0], out[1]
    ref = torch.nn.grad.conv2d_input((b, c, h, w), wt.to(torch.bfloat16).float(), dy.float(), padding=1).permute(0, 2, 3, 1)
    ref = (ref + res.float()) * (act.float() > 0)
    dz = ops.from_padded(dzp, h, w)
    _report("fused dz", dz, ref, 1.2e-2)
    assert _pads_are_zero(dzp, h, w)
    dzf = dz.float()  # the reductions are defined on the stored (bf16) dz
    xhat1 = (y1.float() - v1[2]) * v1[3]
    _report("bsum", bred1[0], dzf.sum((0, 1, 2)), 2e-4)
    _report("bdot", bred1[1], (dzf * xhat1).sum((0, 1, 2)), 2e-4)
    _report("dgamma", dg1, (dzf * xhat1).sum((0, 1, 2)), 2e-4)
    _report("dbeta", db1, dzf.sum((0, 1, 2)), 2e-4)
    if two:
        bred2 = out[2]
        xhat2 = (y2.float() - v2[2]) * v2[3]
        _report("bsum2", bred2[0], dzf.sum((0, 1, 2)), 2e-4)
        _report("bdot2", bred2[1], (dzf * xhat2).sum((0, 1, 2)), 2e-4)
        _report("dgamma2", dg2, (dzf * xhat2).sum((0, 1, 2)), 2e-4)


@pytest.mark.parametrize("case", [(9, 11, 25, 128), (128, 22, 50, 64), (40, 3, 7, 512)])
def test_flat_deferred_finalize_sums(case):
    """CILRS_EPI_DEFER: the kernel only adds the per-channel fp64 sums (what the network plan uses; the BatchNorm apply
    kernels finalize them) - forward: sum y, sum y^2; backward: sum dz, sum dz*y1, sum dz*y2 of the stored bf16 values"""
    ops = _ops()
    _ref_setup()
    b, h, w, c = case
    d = ops.conv_desc(b, h, w, c, c, 3, 1)
    x = _mk((b, c, h, w), 60).to(torch.bfloat16)
    wt = _mk((c, c, 3, 3), 61) * (2.0 / (c * 9)) ** 0.5
    wf, wd = ops.pack_weight(d, wt)
    xp = ops.to_padded(_nhwc_bf16(x.float()))
    yp, acc = ops.conv_flat(xp, wf, c, defer_sums="stats")
    yp_ref = ops.conv_flat(xp, wf, c)
    torch.cuda.synchronize()
    assert torch.equal(yp, yp_ref)
    y = ops.from_padded(yp, h, w).double()
    _report("sum y", acc[0], y.sum((0, 1, 2)), 1e-5)
    _report("sum y^2", acc[1], (y * y).sum((0, 1, 2)), 1e-5)
    res = ops.to_padded(_mk((b, h, w, c), 62).to(torch.bfloat16))
    act = ops.to_padded(torch.relu(_mk((b, h, w, c), 63)).to(torch.bfloat16))
    y1 = _mk((b, h, w, c), 64).to(torch.bfloat16)
    y2 = _mk((b, h, w, c), 65).to(torch.bfloat16)
    dzp, acc = ops.conv_flat(xp, wd, c, dgrad=True, residual=res, mask=act, mask_bits=ops.relu_bits(act),
                             bnbwd=dict(y=ops.to_padded(y1)), bnbwd2=dict(y=ops.to_padded(y2)), defer_sums="bnbwd")
    torch.cuda.synchronize()
    assert _pads_are_zero(dzp, h, w)
    dz = ops.from_padded(dzp, h, w).double()
    _report("sum dz", acc[0], dz.sum((0, 1, 2)), 1e-5)
    _report("sum dz*y1", acc[1], (dz * y1.double()).sum((0, 1, 2)), 1e-5)
    _report("sum dz*y2", acc[2], (dz * y2.double()).sum((0, 1, 2)), 1e-5)


@pytest.mark.parametrize("case", [(128, 11, 25, 128, "1,128,0,3,1"), (37, 22, 50, 64, "2,64,1,9,1"), (128, 6, 13, 256, "1,256,0,1,1"),
                                  (5, 3, 7, 512, "1,128,0,3,1"), (64, 3, 7, 512, "2,256,0,1,1")])
def test_flat_cta_pair_kernel_is_bit_identical_to_the_single_cta_kernel(case, monkeypatch):
    """conv_flat_kernel<MT, PAIR>: clusters of two CTAs sharing tcgen05.mma.cta_group::2 (M = 256, weight tiles split between
    the CTAs) accumulate in the same order as the single-CTA kernel: outputs equal bit for bit, fused sums to fp32 rounding.
    Tile shapes are forced through CILRS_FLAT_SHAPE = "mt,block_n,resident,tap_group,pair"."""
    ops = _ops()
    _ref_setup()
    b, h, w, c, shape = case
    d = ops.conv_desc(b, h, w, c, c, 3, 1)
    x = ops.to_padded(_nhwc_bf16(_mk((b, c, h, w), 70)))
    wf, wd = ops.pack_weight(d, _mk((c, c, 3, 3), 71) * (2.0 / (c * 9)) ** 0.5)
    act = ops.to_padded(torch.relu(_mk((b, h, w, c), 72)).to(torch.bfloat16))
    y1 = ops.to_padded(_mk((b, h, w, c), 73).to(torch.bfloat16))
    res = ops.to_padded(_mk((b, h, w, c), 74).to(torch.bfloat16))
    bits = ops.relu_bits(act)

    def run():
        f, sf = ops.conv_flat(x, wf, c, defer_sums="stats")
        g, sg = ops.conv_flat(x, wd, c, dgrad=True, residual=res, mask=act, mask_bits=bits, bnbwd=dict(y=y1), defer_sums="bnbwd")
        torch.cuda.synchronize()
        return f, sf, g, sg

    monkeypatch.setenv("CILRS_FLAT_SHAPE", "1,64,0,3,0")
    ref = run()
    monkeypatch.setenv("CILRS_FLAT_SHAPE", shape)
    got = run()
    assert torch.equal(got[0], ref[0]) and torch.equal(got[2], ref[2])
    _report("pair: forward sums", got[1], ref[1].double(), 1e-5)
    _report("pair: backward sums", got[3], ref[3].double(), 1e-5)


def test_flat_dgrad_relu_mask_as_bit_tensor_equals_the_bf16_mask():
    """the ReLU mask read as one bit per element (written by cilrs_bn_apply) gives the same dz as the bf16 activation"""
    ops = _ops()
    from cilrs_b200 import _lib
    import ctypes
    _ref_setup()
    b, h, w, c = 7, 11, 25, 128
    d = ops.conv_desc(b, h, w, c, c, 3, 1)
    dy = _mk((b, c, h, w), 50).to(torch.bfloat16)
    wt = _mk((c, c, 3, 3), 51) * (2.0 / (c * 9)) ** 0.5
    _, wd = ops.pack_weight(d, wt)
    ypre = ops.to_padded(_mk((b, h, w, c), 52).to(torch.bfloat16))
    vec = torch.stack([torch.ones(c), torch.zeros(c), torch.zeros(c), torch.ones(c)]).cuda().contiguous()
    act = torch.empty_like(ypre)
    bits = torch.full((b, h + 1, w + 1, c // 8), 255, dtype=torch.uint8, device="cuda")
    _lib.call("cilrs_bn_apply", ypre, vec, None, None, None, act, ctypes.c_longlong(ypre.numel()), c, 1, h, w, bits, _lib.stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(bits, ops.relu_bits(act))
    dyp = ops.to_padded(_nhwc_bf16(dy.float()))
    a = ops.conv_flat(dyp, wd, c, dgrad=True, mask=act)
    bb = ops.conv_flat(dyp, wd, c, dgrad=True, mask=act, mask_bits=bits)
    torch.cuda.synchronize()
    assert torch.equal(a, bb)
