"""Unit parity of the BatchNorm / ReLU / max-pool kernels (forward and backward) and of the fp32 heads, through the C-ABI.

Checker: fp32/fp64 torch autograd of the same operation on the same (bf16-valued) inputs. These kernels are exercised
in isolation with tight tolerances because in the whole model a single flipped ReLU (bf16 noise upstream) dominates any
max-norm comparison (see tests/test_model_gpu.py).
"""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _conv_stats(B, H, W, C, seed):
    """a raw conv output y (bf16 NHWC) with its stats partials, produced by the real fprop kernel"""
    from cilrs_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(B, H, W, C, generator=g, device="cuda").to(torch.bfloat16)
    w = torch.randn(C, C, 3, 3, generator=g, device="cuda") * (2.0 / (9 * C)) ** 0.5
    d = ops.conv_desc(B, H, W, C, C, 3, 1)
    wf, _ = ops.pack_weight(d, w)
    y, st = ops.conv_fprop(d, x, wf, stats=True)
    return y, st


def _finalize(st, C, count, gamma, beta, rm, rv, nbt, training, update=1):
    from cilrs_b200 import _lib
    vec = torch.empty(4, C, device="cuda")
    _lib.call("cilrs_bn_finalize", st, st.shape[0] if st is not None else 0, C, ctypes.c_double(count), gamma, beta, rm, rv, nbt,
              ctypes.c_float(0.1), ctypes.c_float(1e-5), int(training), update, vec, _lib.stream_ptr())
    return vec


@pytest.mark.parametrize("shape", [(6, 11, 25, 128), (3, 22, 50, 64), (9, 3, 7, 512)])
def test_bn_train_forward_and_running_stats(shape):
    from cilrs_b200 import _lib
    B, H, W, C = shape
    y, st = _conv_stats(B, H, W, C, 1)
    g = torch.Generator(device="cuda").manual_seed(2)
    gamma = 1 + 0.3 * torch.randn(C, generator=g, device="cuda")
    beta = 0.2 * torch.randn(C, generator=g, device="cuda")
    rm, rv = torch.randn(C, generator=g, device="cuda"), torch.rand(C, generator=g, device="cuda") + 0.5
    nbt = torch.zeros(1, dtype=torch.long, device="cuda")
    rm_ref, rv_ref = rm.clone(), rv.clone()
    res = torch.randn(B, H, W, C, generator=g, device="cuda").to(torch.bfloat16)
    vec = _finalize(st, C, B * H * W, gamma, beta, rm, rv, nbt, True)
    out = torch.empty_like(y)
    _lib.call("cilrs_bn_apply", y, vec, res, None, None, out, ctypes.c_longlong(y.numel()), C, 1, 0, 0, None, _lib.stream_ptr())
    torch.cuda.synchronize()
    yf = y.float().permute(0, 3, 1, 2)
    ref = F.batch_norm(yf, rm_ref, rv_ref, gamma, beta, True, 0.1, 1e-5)
    ref = torch.relu(ref + res.float().permute(0, 3, 1, 2))
    assert _rel(out.float().permute(0, 3, 1, 2), ref) <= 6e-3
    assert _rel(rm, rm_ref) <= 1e-5 and _rel(rv, rv_ref) <= 1e-5 and int(nbt) == 1
    # frozen (eval) statistics + second BN'd input (downsample branch)
    vec_e = _finalize(None, C, 1.0, gamma, beta, rm, rv, None, False, 0)
    out2 = torch.empty_like(y)
    _lib.call("cilrs_bn_apply", y, vec_e, None, res, vec, out2, ctypes.c_longlong(y.numel()), C, 0, 0, 0, None, _lib.stream_ptr())
    torch.cuda.synchronize()
    ref2 = F.batch_norm(yf, rm, rv, gamma, beta, False, 0.1, 1e-5) + F.batch_norm(res.float().permute(0, 3, 1, 2), None, None, gamma, beta, True, 0.1, 1e-5) * 0
    # second input normalised with `vec` (batch statistics of y): restate directly
    sc, sh = vec[0], vec[1]
    ref2 = F.batch_norm(yf, rm, rv, gamma, beta, False, 0.1, 1e-5) + (res.float() * sc + sh).permute(0, 3, 1, 2)
    assert _rel(out2.float().permute(0, 3, 1, 2), ref2) <= 6e-3


@pytest.mark.parametrize("frozen", [0, 1])
@pytest.mark.parametrize("shape", [(6, 11, 25, 128), (5, 22, 50, 64), (9, 3, 7, 512)])
def test_bn_relu_backward(shape, frozen):
    from cilrs_b200 import _lib
    B, H, W, C = shape
    y, st = _conv_stats(B, H, W, C, 3)
    gen = torch.Generator(device="cuda").manual_seed(4)
    gamma = 1 + 0.3 * torch.randn(C, generator=gen, device="cuda")
    beta = 0.2 * torch.randn(C, generator=gen, device="cuda")
    rm, rv = 0.1 * torch.randn(C, generator=gen, device="cuda"), torch.rand(C, generator=gen, device="cuda") + 0.5
    vec = _finalize(st, C, B * H * W, gamma, beta, rm.clone(), rv.clone(), None, not frozen, 0)
    act = torch.empty_like(y)
    _lib.call("cilrs_bn_apply", y, vec, None, None, None, act, ctypes.c_longlong(y.numel()), C, 1, 0, 0, None, _lib.stream_ptr())
    gup = torch.randn(B, H, W, C, generator=gen, device="cuda").to(torch.bfloat16)
    dy, dz = torch.empty_like(y), torch.empty_like(y)
    dgamma, dbeta = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    nws = _lib.lib().cilrs_bn_backward_workspace_floats
    nws.restype = ctypes.c_size_t
    ws = torch.zeros(nws(C), device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.call("cilrs_bn_backward", gup, act, y, vec, gamma, ctypes.c_longlong(y.numel()), C, ctypes.c_double(B * H * W), frozen, dy, dz,
              dgamma, dbeta, ws, cnt, None, 0, 0, 0, 0, _lib.stream_ptr())
    torch.cuda.synchronize()
    x = y.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    out = torch.relu(F.batch_norm(x, rm.clone(), rv.clone(), gm, bt, not frozen, 0.1, 1e-5))
    out.backward(gup.float().permute(0, 3, 1, 2))
    assert _rel(dy.float().permute(0, 3, 1, 2), x.grad) <= 8e-3
    assert _rel(dgamma, gm.grad) <= 2e-3 and _rel(dbeta, bt.grad) <= 2e-3
    assert int(cnt) == 0  # the reduction counter is reset for the next launch / graph replay
    mask = (act.float() > 0)
    assert torch.equal(dz.float(), gup.float() * mask)


@pytest.mark.parametrize("B", [1, 5])
def test_stem_bn_relu_maxpool_forward_backward(B):
    from cilrs_b200 import _lib, ops
    gen = torch.Generator(device="cuda").manual_seed(6)
    img = torch.randn(B, 3, 88, 200, generator=gen, device="cuda")
    w = torch.randn(64, 3, 7, 7, generator=gen, device="cuda") * (2.0 / 147) ** 0.5
    y, st = ops.stem_fprop(ops.image_to_s2d(img), ops.stem_pack_weight(w), stats=True)
    C = 64
    gamma = 1 + 0.3 * torch.randn(C, generator=gen, device="cuda")
    beta = 0.2 * torch.randn(C, generator=gen, device="cuda")
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    vec = _finalize(st, C, B * 4400, gamma, beta, rm, rv, None, True, 0)
    pooled = torch.empty(B, 22, 50, C, dtype=torch.bfloat16, device="cuda")
    arg = torch.empty(B, 22, 50, C, dtype=torch.uint8, device="cuda")
    _lib.call("cilrs_bn_relu_maxpool", y, vec, pooled, arg, B, 44, 100, C, 0, _lib.stream_ptr())
    gp = torch.randn(B, 22, 50, C, generator=gen, device="cuda").to(torch.bfloat16)
    dy = torch.empty_like(y)
    dgamma, dbeta = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    nws = _lib.lib().cilrs_bn_backward_workspace_floats
    nws.restype = ctypes.c_size_t
    ws = torch.zeros(nws(C), device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.call("cilrs_bn_backward", gp, None, y, vec, gamma, ctypes.c_longlong(y.numel()), C, ctypes.c_double(B * 4400), 0, dy, None, dgamma,
              dbeta, ws, cnt, arg, 44, 100, 0, 0, _lib.stream_ptr())
    torch.cuda.synchronize()
    x = y.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    a = torch.relu(F.batch_norm(x, None, None, gm, bt, True, 0.1, 1e-5))
    a_r = a + (a.to(torch.bfloat16).float() - a).detach()  # the pool compares bf16-rounded activations (straight-through)
    po = F.max_pool2d(a_r, 3, 2, 1)
    assert _rel(pooled.float().permute(0, 3, 1, 2), po) <= 8e-3
    po.backward(gp.float().permute(0, 3, 1, 2))
    assert _l2(dy.float().permute(0, 3, 1, 2), x.grad) <= 1.5e-2
    assert _rel(dgamma, gm.grad) <= 1e-2 and _rel(dbeta, bt.grad) <= 1e-2


@pytest.mark.parametrize("B,padded", [(1, 0), (5, 1), (24, 1)])
def test_stem_pool_with_selected_values_and_fused_backward(B, padded):
    """The training plan's form of bn1 -> relu -> maxpool and its backward (csrc/stem_pool.cuh): the pool also keeps the raw conv1
    value at every arg-max, the BatchNorm-backward sums run over the pool outputs, one pass routes + applies. Checked against
    torch autograd and against the two-pass kernels behind cilrs_bn_backward (an independent implementation of the same maths)."""
    from cilrs_b200 import _lib, ops
    gen = torch.Generator(device="cuda").manual_seed(16 + B)
    img = torch.randn(B, 3, 88, 200, generator=gen, device="cuda")
    w = torch.randn(64, 3, 7, 7, generator=gen, device="cuda") * (2.0 / 147) ** 0.5
    y, st = ops.stem_fprop(ops.image_to_s2d(img), ops.stem_pack_weight(w), stats=True)
    C = 64
    gamma = 1 + 0.3 * torch.randn(C, generator=gen, device="cuda")
    gamma[::7] *= -1                                    # negative scales: the pool must not assume a monotone map
    beta = 0.2 * torch.randn(C, generator=gen, device="cuda")
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    vec = _finalize(st, C, B * 4400, gamma, beta, rm, rv, None, True, 0)
    sp = _lib.stream_ptr()
    OHp, OWp = (23, 51) if padded else (22, 50)
    pooled = torch.zeros(B, OHp, OWp, C, dtype=torch.bfloat16, device="cuda")
    ysel = torch.full((B, OHp, OWp, C), float("nan"), dtype=torch.bfloat16, device="cuda")
    arg = torch.empty(B, 22, 50, C, dtype=torch.uint8, device="cuda")
    _lib.call("cilrs_bn_relu_maxpool_sel", y, vec, pooled, arg, ysel, B, 44, 100, C, padded, sp)
    # reference pool: bf16-rounded activations, first maximum in scan order
    # (the kernel's y * scale + shift is one fused multiply-add: exact in fp64, then rounded to fp32, then to bf16)
    a = torch.relu((y.double() * vec[0].double() + vec[1].double()).float()).to(torch.bfloat16)
    win = F.unfold(F.pad(a.float().permute(0, 3, 1, 2), (1, 1, 1, 1), value=float("-inf")), 3, stride=2).view(B, C, 9, 22, 50)
    ref_val, ref_idx = win.max(dim=2)                   # torch.max returns the first maximal index on CUDA for exact ties? make it explicit:
    first = (win == ref_val.unsqueeze(2)).float().argmax(dim=2)
    assert torch.equal(pooled[:, :22, :50].float().permute(0, 3, 1, 2), ref_val)
    assert torch.equal(arg.permute(0, 3, 1, 2).long(), first)
    if padded:
        assert float(pooled[:, 22:].float().abs().max()) == 0.0 and float(pooled[:, :, 50:].float().abs().max()) == 0.0
    # ysel = the raw conv output at the arg-max position
    ywin = F.unfold(F.pad(y.float().permute(0, 3, 1, 2), (1, 1, 1, 1)), 3, stride=2).view(B, C, 9, 22, 50)
    want_sel = ywin.gather(2, first.unsqueeze(2)).squeeze(2)
    assert torch.equal(ysel[:, :22, :50].float().permute(0, 3, 1, 2), want_sel)

    gp = torch.randn(B, 22, 50, C, generator=gen, device="cuda").to(torch.bfloat16)
    g_in = torch.full((B, OHp, OWp, C), 7.0, dtype=torch.bfloat16, device="cuda")   # garbage in the gradient's padding: never read
    g_in[:, :22, :50] = gp
    nws = _lib.lib().cilrs_bn_backward_workspace_floats
    nws.restype = ctypes.c_size_t
    outs = []
    for fused in (True, False):
        dy = torch.full_like(y, float("nan"))
        dgamma, dbeta = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        ws = torch.zeros(nws(C), device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        if fused:
            _lib.call("cilrs_stem_bn_backward", g_in.clone(), pooled, ysel, arg, y, vec, gamma, B, 44, 100, C, padded, 0, dy, dgamma, dbeta,
                      ws, cnt, sp)
        else:
            _lib.call("cilrs_bn_backward", gp, None, y, vec, gamma, ctypes.c_longlong(y.numel()), C, ctypes.c_double(B * 4400), 0, dy, None,
                      dgamma, dbeta, ws, cnt, arg, 44, 100, 0, 0, sp)
        torch.cuda.synchronize()
        assert int(cnt) == 0 and float(ws[:4 * C].abs().max()) == 0.0     # accumulators left ready for the next launch
        outs.append((dy, dgamma, dbeta))
    (dy_f, dg_f, db_f), (dy_t, dg_t, db_t) = outs
    assert _rel(dg_f, dg_t) <= 1e-5 and _rel(db_f, db_t) <= 1e-5            # same sums, taken over a quarter of the elements
    assert _rel(dy_f.float(), dy_t.float()) <= 8e-3 and _l2(dy_f.float(), dy_t.float()) <= 3e-3   # bf16 ulps of the outputs
    x = y.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    act = torch.relu(F.batch_norm(x, None, None, gm, bt, True, 0.1, 1e-5))
    a_r = act + (act.to(torch.bfloat16).float() - act).detach()
    po = F.max_pool2d(a_r, 3, 2, 1)
    po.backward(gp.float().permute(0, 3, 1, 2))
    assert _l2(dy_f.float().permute(0, 3, 1, 2), x.grad) <= 1.5e-2
    assert _rel(dg_f, gm.grad) <= 1e-2 and _rel(db_f, bt.grad) <= 1e-2
    # frozen statistics: dy = gamma * rstd * dz only
    dy0 = torch.empty_like(y)
    ws = torch.zeros(nws(C), device="cuda")
    _lib.call("cilrs_stem_bn_backward", g_in.clone(), pooled, ysel, arg, y, vec, gamma, B, 44, 100, C, padded, 1, dy0, None, None, ws,
              torch.zeros(1, dtype=torch.int32, device="cuda"), sp)
    dz = torch.zeros(B, C, 46, 102, device="cuda")
    gm_ = (gp.float() * (pooled[:, :22, :50].float() > 0)).permute(0, 3, 1, 2)
    for code in range(9):
        r, s_ = divmod(code, 3)
        dz[:, :, r:r + 44:2, s_:s_ + 100:2] += gm_ * (first == code)
    want0 = dz[:, :, 1:45, 1:101] * (gamma * vec[3]).view(1, C, 1, 1)
    assert _rel(dy0.float().permute(0, 3, 1, 2), want0) <= 8e-3


@pytest.mark.parametrize("shape", [(6, 11, 25, 128), (3, 22, 50, 64), (9, 3, 7, 512)])
def test_padded_flat_layout_bn_apply_and_backward_equal_the_dense_kernels(shape):
    """the same kernels on the padded-flat layout [B,H+1,W+1,C]: identical results on the real pixels, exact zeros on the
    padding pixels, and stale values in the padding of the incoming gradient are never read"""
    from cilrs_b200 import _lib, ops
    B, H, W, C = shape
    y, st = _conv_stats(B, H, W, C, 11)
    gen = torch.Generator(device="cuda").manual_seed(12)
    gamma = 1 + 0.3 * torch.randn(C, generator=gen, device="cuda")
    beta = 0.2 * torch.randn(C, generator=gen, device="cuda")
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    vec = _finalize(st, C, B * H * W, gamma, beta, rm, rv, None, True, 0)
    res = torch.randn(B, H, W, C, generator=gen, device="cuda").to(torch.bfloat16)
    out_d = torch.empty_like(y)
    _lib.call("cilrs_bn_apply", y, vec, res, None, None, out_d, ctypes.c_longlong(y.numel()), C, 1, 0, 0, None, _lib.stream_ptr())
    yp, resp = ops.to_padded(y), ops.to_padded(res)
    out_p = torch.full_like(yp, float("nan"))
    _lib.call("cilrs_bn_apply", yp, vec, resp, None, None, out_p, ctypes.c_longlong(yp.numel()), C, 1, H, W, None, _lib.stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(ops.from_padded(out_p, H, W), out_d)
    assert float(out_p[:, H:].float().abs().max()) == 0.0 and float(out_p[:, :, W:].float().abs().max()) == 0.0
    # backward: dense vs padded (gradient padding filled with garbage on purpose)
    gup = torch.randn(B, H, W, C, generator=gen, device="cuda").to(torch.bfloat16)
    nws = _lib.lib().cilrs_bn_backward_workspace_floats
    nws.restype = ctypes.c_size_t
    outs = []
    for padded in (False, True):
        g_in = gup
        a_in, y_in = out_d, y
        if padded:
            g_in = torch.full((B, H + 1, W + 1, C), 7.0, dtype=torch.bfloat16, device="cuda")
            g_in[:, :H, :W] = gup
            a_in, y_in = ops.to_padded(out_d), yp
        dy, dz = torch.full_like(y_in, float("nan")), torch.full_like(y_in, float("nan"))
        dgamma, dbeta = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        ws = torch.zeros(nws(C), device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.call("cilrs_bn_backward", g_in, a_in, y_in, vec, gamma, ctypes.c_longlong(y_in.numel()), C, ctypes.c_double(B * H * W), 0, dy, dz,
                  dgamma, dbeta, ws, cnt, None, 0, 0, H if padded else 0, W if padded else 0, _lib.stream_ptr())
        torch.cuda.synchronize()
        outs.append((dy, dz, dgamma, dbeta))
    (dy_d, dz_d, dg_d, db_d), (dy_p, dz_p, dg_p, db_p) = outs
    assert torch.equal(ops.from_padded(dz_p, H, W), dz_d)
    assert float(dy_p[:, H:].float().abs().max()) == 0.0 and float(dy_p[:, :, W:].float().abs().max()) == 0.0
    assert float(dz_p[:, H:].float().abs().max()) == 0.0 and float(dz_p[:, :, W:].float().abs().max()) == 0.0
    assert _rel(dg_p, dg_d) <= 1e-4 and _rel(db_p, db_d) <= 1e-4      # different partial order
    assert _rel(ops.from_padded(dy_p, H, W).float(), dy_d.float()) <= 8e-3


@pytest.mark.parametrize("B", [1, 4, 37])
def test_heads_forward_backward_fp32_exact(B):
    """heads on given fp32 features vs the fp64 oracle: 1e-5 (they are fp32 CUDA-core kernels)"""
    from cilrs_b200 import _lib
    from cilrs_b200.model import CILRS
    from oracle import cilrs_oracle as O
    sd = O.synthetic_state_dict(3)
    m = CILRS().to("cuda")
    m.load_state_dict(sd)
    m._ensure(B)
    grads = m.flat_gradients()
    grads.zero_()
    gen = torch.Generator().manual_seed(B)
    feat = torch.randn(B, 512, generator=gen).abs()
    speed = torch.rand(B, generator=gen)
    command = torch.randint(0, 4, (B,), generator=gen)
    dctrl, dspd = torch.randn(B, 3, generator=gen), torch.randn(B, generator=gen)
    controls = torch.empty(B, 3, device="cuda")
    ps = torch.empty(B, device="cuda")
    dfeat = torch.empty(B, 512, device="cuda")
    sp_d, cm_d = speed.cuda(), command.cuda()
    _lib.call("cilrs_model_heads_forward", m._handle, B, feat.cuda(), sp_d, cm_d, controls, ps, 1, ctypes.c_float(0.0),
              ctypes.c_ulonglong(1), _lib.stream_ptr())
    _lib.call("cilrs_model_heads_backward", m._handle, B, dctrl.cuda(), dspd.cuda(), sp_d, cm_d, ctypes.c_float(0.0), dfeat,
              _lib.stream_ptr())
    torch.cuda.synchronize()
    # oracle: the head part of O.forward on the same features
    sd64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()
            if v.is_floating_point() and not k.startswith("visual_encoder")}
    f64 = feat.double().clone().requires_grad_(True)
    s = F.relu(F.linear(speed.double().unsqueeze(1), sd64["speed_encoder.0.weight"], sd64["speed_encoder.0.bias"]))
    s = F.relu(F.linear(s, sd64["speed_encoder.3.weight"], sd64["speed_encoder.3.bias"]))
    comb = torch.cat([f64, s], 1)
    p = F.relu(F.linear(f64, sd64["speed_predictor.0.weight"], sd64["speed_predictor.0.bias"]))
    p = F.relu(F.linear(p, sd64["speed_predictor.3.weight"], sd64["speed_predictor.3.bias"]))
    ps64 = F.linear(p, sd64["speed_predictor.5.weight"], sd64["speed_predictor.5.bias"]).squeeze(1)
    outs = []
    for k in range(4):
        pre = "control_branches.%d" % k
        h = F.relu(F.linear(comb, sd64[pre + ".0.weight"], sd64[pre + ".0.bias"]))
        h = F.relu(F.linear(h, sd64[pre + ".3.weight"], sd64[pre + ".3.bias"]))
        outs.append(F.linear(h, sd64[pre + ".6.weight"], sd64[pre + ".6.bias"]))
    c64 = torch.stack(outs, 0).gather(0, command.view(1, B, 1).expand(1, B, 3)).squeeze(0)
    ((c64 * dctrl.double()).sum() + (ps64 * dspd.double()).sum()).backward()
    assert _rel(controls, c64) <= 1e-5 and _rel(ps, ps64) <= 1e-5
    assert _rel(dfeat, f64.grad) <= 1e-5
    views = dict(zip([n for n, _ in m.named_parameters()], m._views(grads)))
    for k, v in sd64.items():
        ref = v.grad if v.grad is not None else torch.zeros_like(v)
        assert _rel(views[k], ref) <= 2e-5 or float(ref.abs().max()) == 0.0 and float(views[k].abs().max()) == 0.0, k


def test_heads_dropout_statistics():
    """dropout p>0 cannot match torch's RNG stream: check what nn.Dropout guarantees instead - every kept activation
    is scaled by exactly 1/(1-p), the keep-rate is 1-p within 5 sigma at every dropout site (speed_encoder.2,
    control_branches.k.2 / .5, speed_predictor.2; model/autonomous_drive.py:371-387), the mask is a function of the seed."""
    from cilrs_b200 import _lib
    from cilrs_b200.model import CILRS
    m = CILRS(dropout=0.5).to("cuda")
    B = 256
    m._ensure(B)
    feat = torch.rand(B, 512, device="cuda")
    speed = torch.rand(B, device="cuda")
    cmd = torch.randint(0, 4, (B,), device="cuda")
    m.flat_gradients()

    def run(p, seed):
        c, ps = torch.empty(B, 3, device="cuda"), torch.empty(B, device="cuda")
        _lib.call("cilrs_model_heads_forward", m._handle, B, feat, speed, cmd, c, ps, 1, ctypes.c_float(p), ctypes.c_ulonglong(seed),
                  _lib.stream_ptr())
        torch.cuda.synchronize()
        return c, ps, [m.debug_heads_saved(w, B).clone() for w in range(6)]

    c_ref, p_ref, sv0 = run(0.0, 7)
    c0, p0, sv1 = run(0.5, 7)
    # sites whose INPUT does not depend on another dropout: speed_encoder.0 (which 0) and speed_predictor.0 (which 4):
    # kept values are exactly 2x the p = 0 activation, dropped ones exactly 0
    for w in (0, 4):
        a0, a1 = sv0[w], sv1[w]
        active = a0 > 0
        kept = a1 != 0
        assert not bool((kept & ~active).any())
        assert torch.equal(a1[kept], 2.0 * a0[kept])
        n = int(active.sum())
        rate = float(kept.sum()) / n
        assert abs(rate - 0.5) <= 5 * 0.5 / n ** 0.5, (w, rate, n)
    # the other two sites (branch layers, which 2 and 3) see a dropout-perturbed input: keep-rate among their own active units
    # cannot be read off directly, but the zero fraction must rise to about 1 - 0.5 * (active fraction without dropout)
    for w in (2, 3):
        act0 = float((sv0[w] > 0).float().mean())
        act1 = float((sv1[w] != 0).float().mean())
        assert abs(act1 - 0.5 * act0) <= 0.08 * act0, (w, act0, act1)
    # layers without dropout behind them (speed_encoder.3, speed_predictor.3) are never zeroed beyond ReLU
    assert float((sv1[5] != 0).float().mean()) > 0.5 * float((sv0[5] > 0).float().mean())
    assert not torch.equal(c0, c_ref)
    c1, _, _ = run(0.5, 8)
    assert not torch.equal(c0, c1)          # different seed -> different mask
    c2, _, _ = run(0.5, 7)
    assert torch.equal(c0, c2)              # same seed -> same mask
