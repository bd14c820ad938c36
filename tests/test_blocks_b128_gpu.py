"""Parity at the BENCHMARKED configuration (batch 128): every BasicBlock and the stem, train-mode forward AND backward, against
the fp64 oracle on identical bf16 inputs.

Why per block: at random init the 36-BatchNorm train-mode backward of the whole network amplifies rounding ~1e5x (SURVEY
§7.3-H1: the reference's own fp32 is 2e-3..8e-3 from fp64; measured on B200: our whole-model trunk gradient is 0.53-0.58 from
the bf16-EMULATING fp64 oracle and 0.66-0.73 from plain fp64), so a whole-model gradient bound cannot tell a correct bf16
implementation from a broken one. One block is well conditioned: with the same bf16 input activation and the same output
gradient a correct implementation must agree to bf16 rounding, and a wrong BN-backward, mask, deferred finalize, fused dgrad
reduction or stride-2 parity launch shows up as an O(1) error. The tile shapes the cost model picks (pair, MT, block_n,
tap_group) depend on the batch, hence B = 128, the bench batch.

Two references per block, both the oracle's `basic_block` in fp64:
  * bf16-emulating (`rnd=round_bf16`: the block's raw conv outputs, post-BN/ReLU activation and output - and the gradients
    flowing through them - are rounded exactly where the CUDA path stores bf16): what remains is accumulation order; bar 2e-2.
  * plain fp64: measured 2-5e-2 on every tensor and every block. That floor is ReLU sign flips - bf16 storage moves ~0.15 % of
    the pre-activations across zero, and with an i.i.d. random output gradient each flip costs a full-size gradient element
    (sqrt(1.5e-3) = 4e-2); reported, and bounded at 8e-2.

The block under test runs through the real plan (`cilrs_model_forward`, then `cilrs_model_debug_backward(hi=lo=block)`): the
same launches, buffers and fused epilogues as a training step. The checker is the oracle's `basic_block` / `stem` in fp64
(torch fp64 on the GPU only to keep the suite short; it is the checker, never the product).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
B = 128
BAR = 2e-2   # BASELINE.json north_star, bf16 mode: 2e-2 relative


def _O():
    from oracle import cilrs_oracle as O
    return O


def _rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-300))


def _rel_max(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


@pytest.fixture(scope="module", params=["separate-bn", "fused-bn"])
def net(request):
    """One train-mode forward at B=128 (activations, ReLU bits and BatchNorm vectors of every layer are then in the plan).
    Every test below runs twice: with the separate bn_apply / bn_bwd_apply launches (the default) and with the grid-synchronous
    BatchNorm fused into the flat convolutions (conv_params.h: CF_FUSE)."""
    from cilrs_b200 import _lib
    from cilrs_b200.model import CILRS
    prev = _lib.lib().cilrs_set_bn_fusion(1 if request.param == "fused-bn" else 0)
    request.addfinalizer(lambda: _lib.lib().cilrs_set_bn_fusion(prev))
    O = _O()
    sd = O.synthetic_state_dict(0)
    m = CILRS(num_commands=4, dropout=0.0)
    m.load_state_dict(sd, strict=True)
    m = m.to("cuda").train()
    g = torch.Generator().manual_seed(77)
    coarse = torch.randn(B, 3, 11, 25, generator=g)
    image = torch.nn.functional.interpolate(coarse, size=(88, 200), mode="bicubic", align_corners=False).contiguous()
    speed = torch.rand(B, generator=g)
    command = torch.randint(0, 4, (B,), generator=g)
    controls, pred_speed = m(image.cuda(), speed.cuda(), command.cuda())   # grad enabled: activations are kept
    torch.cuda.synchronize()
    sd64 = {k: (v.double().cuda() if v.is_floating_point() else v.cuda()) for k, v in sd.items()}
    return m, sd64, image


def _leaf(sd64, keys):
    out = dict(sd64)
    for k in keys:
        v = sd64[k]
        if v.dim() == 4:   # conv weights: the tensor-core operand is the bf16 rounding of the fp32 master
            v = v.float().to(torch.bfloat16).double()
        out[k] = v.clone().requires_grad_(True)
    return out


def _param(m, name):
    return dict(m.named_parameters())[name]


@pytest.mark.parametrize("bi", list(range(16)))
def test_basic_block_train_forward_backward_b128(net, bi):
    from cilrs_b200 import ops
    O = _O()
    m, sd64, _ = net
    prefix, stride = O.block_prefixes()[bi]
    x = m.debug_activation(bi, B).clone()            # [B,H,W,C] bf16: the block's input as the plan holds it
    out = m.debug_activation(bi + 1, B).clone()
    keys = [k for k in sd64 if k.startswith(prefix + ".") and sd64[k].is_floating_point() and "running" not in k]
    sd = _leaf(sd64, keys)
    # ---- backward from a random output gradient (bf16-exact values so both sides start from the same tensor) ----
    gen = torch.Generator(device="cuda").manual_seed(1000 + bi)
    g = (torch.randn(out.shape, generator=gen, device="cuda") * 0.05).to(torch.bfloat16)
    m.flat_gradients().zero_()
    dx = m.debug_backward(B, bi, bi, ops.to_padded(g))
    torch.cuda.synchronize()
    dx = dx[:, :x.shape[1], :x.shape[2]].double().permute(0, 3, 1, 2)
    views = dict(zip([n for n, _ in m.named_parameters()], m._views(m.flat_gradients())))
    report = {}
    for label, rnd in (("emu", O.round_bf16), ("fp64", None)):
        sd = _leaf(sd64, keys)
        x64 = x.double().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
        out64 = O.basic_block(sd, prefix, x64, stride, training=True, rnd=rnd)
        (out64 * g.double().permute(0, 3, 1, 2)).sum().backward()
        ref_dx = x64.grad
        if bi > 0:   # the fused dgrad epilogue also applies the ReLU mask of the tensor it produces the gradient of
            ref_dx = ref_dx * (x64 > 0)
        rows = [("forward out", _rel_max(out.double().permute(0, 3, 1, 2), out64)), ("dx", _rel_l2(dx, ref_dx))]
        for k in keys:
            rows.append((k[len(prefix) + 1:], _rel_l2(views[k], sd[k].grad)))
        report[label] = rows
        print("block %2d (%s, stride %d) vs %-4s: %s" % (bi, prefix, stride, label, ", ".join("%s %.2e" % r for r in rows)))
    bad = [r for r in report["emu"] if not r[1] <= BAR] + [r for r in report["fp64"] if not r[1] <= 8e-2]
    assert not bad, bad


def test_stem_train_forward_backward_b128(net):
    from cilrs_b200 import ops
    O = _O()
    m, sd64, image = net
    keys = ["visual_encoder.0.weight", "visual_encoder.1.weight", "visual_encoder.1.bias"]
    img64 = image.cuda().to(torch.bfloat16).double()   # conv1's operand is the bf16 space-to-depth copy of the image
    y = m.debug_activation(17, B).double().permute(0, 3, 1, 2)
    pool = m.debug_activation(0, B)
    gen = torch.Generator(device="cuda").manual_seed(4242)
    g = (torch.randn(pool.shape, generator=gen, device="cuda") * 0.05).to(torch.bfloat16)
    m.flat_gradients().zero_()
    m.debug_backward(B, -1, -1, ops.to_padded(g))
    torch.cuda.synchronize()
    views = dict(zip([n for n, _ in m.named_parameters()], m._views(m.flat_gradients())))
    report = {}
    for label, rnd in (("emu", O.round_bf16), ("fp64", None)):
        sd = _leaf(sd64, keys)
        y64 = torch.nn.functional.conv2d(img64, sd["visual_encoder.0.weight"].detach(), stride=2, padding=3)
        pool64 = O.stem(sd, img64, training=True, rnd=rnd)
        (pool64 * g.double().permute(0, 3, 1, 2)).sum().backward()
        rows = [("conv1 out", _rel_max(y, y64)), ("pool out", _rel_max(pool.double().permute(0, 3, 1, 2), pool64))]
        rows += [(k, _rel_l2(views[k], sd[k].grad)) for k in keys]
        report[label] = rows
        print("stem vs %-4s: %s" % (label, ", ".join("%s %.2e" % r for r in rows)))
    bad = [r for r in report["emu"] if not r[1] <= BAR] + [r for r in report["fp64"] if not r[1] <= 8e-2]
    assert not bad, bad


def _chain(net, hi, lo):
    """debug backward of blocks hi..lo against the bf16-emulating fp64 oracle: (forward err, global param-grad err, dx err, worst)"""
    from cilrs_b200 import ops
    O = _O()
    m, sd64, _ = net
    pre = O.block_prefixes()
    keys = [k for k in sd64 if any(k.startswith(pre[b][0] + ".") for b in range(lo, hi + 1)) and sd64[k].is_floating_point()
            and "running" not in k]
    sd = _leaf(sd64, keys)
    x = m.debug_activation(lo, B).clone()
    out = m.debug_activation(hi + 1, B).clone()
    x64 = x.double().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    h = x64
    for b in range(lo, hi + 1):
        h = O.basic_block(sd, pre[b][0], h, pre[b][1], training=True, rnd=O.round_bf16)
    e_fwd = _rel_max(out.double().permute(0, 3, 1, 2), h)
    gen = torch.Generator(device="cuda").manual_seed(99 + 16 * hi + lo)
    g = (torch.randn(out.shape, generator=gen, device="cuda") * 0.05).to(torch.bfloat16)
    (h * g.double().permute(0, 3, 1, 2)).sum().backward()
    m.flat_gradients().zero_()
    dx = m.debug_backward(B, hi, lo, ops.to_padded(g))
    torch.cuda.synchronize()
    dx = dx[:, :x.shape[1], :x.shape[2]].double().permute(0, 3, 1, 2)
    views = dict(zip([n for n, _ in m.named_parameters()], m._views(m.flat_gradients())))
    flat_got = torch.cat([views[k].double().reshape(-1) for k in keys])
    flat_ref = torch.cat([sd[k].grad.reshape(-1) for k in keys])
    ref_dx = x64.grad * (x64 > 0) if lo > 0 else x64.grad
    per = sorted(((_rel_l2(views[k], sd[k].grad), k) for k in keys), reverse=True)
    return e_fwd, _rel_l2(flat_got, flat_ref), _rel_l2(dx, ref_dx), per[:3]


@pytest.mark.parametrize("lo", list(range(15)))
def test_adjacent_block_pair_backward_b128(net, lo):
    """Blocks lo+1 and lo in one backward: here the BatchNorm-backward reductions of block lo's output come from the dgrad
    epilogue of block lo+1 (CF_BNBWD, and CF_BNBWD2 when block lo has a downsample branch) as raw fp64 sums that bn_bwd_apply
    finalizes in its prologue - the path a training step takes at every block boundary - or, below a stride-2 block, from the
    four-parity conv_gemm_multi launch + stand-alone reduce. Measured on B200 vs the bf16-emulating oracle: <= 2.5e-2."""
    e_fwd, e_glob, e_dx, worst = _chain(net, lo + 1, lo)
    print("blocks %d..%d: forward %.2e, global param-grad %.2e, dx %.2e, worst tensors %s" % (lo, lo + 1, e_fwd, e_glob, e_dx, worst))
    assert e_fwd <= BAR and e_glob <= 4e-2 and e_dx <= 4e-2 and worst[0][0] <= 6e-2, (e_fwd, e_glob, e_dx, worst)


@pytest.mark.parametrize("hi,lo,bar", [(2, 0, 8e-2), (6, 3, 0.12), (12, 7, 0.2), (15, 13, 0.12), (15, 0, 0.75)])
def test_layer_group_backward_b128(net, hi, lo, bar):
    """Whole layer groups (layer1..layer4, each with its stride-2 / downsample head where it has one) and the whole 16-block
    trunk in ONE debug backward, against the bf16-emulating fp64 oracle. Every train-mode block re-amplifies the perturbation it
    receives (~1.5x per block): measured on B200 4.2e-2 (3 blocks), 6.7e-2 (4), 0.115 (6), 6.1e-2 (layer4's 3) and 0.47 for all
    16 - the whole-model figure of ~0.5 is this growth, not an error of any one kernel (single blocks: 1e-3..1e-2). The bars
    are ~1.8x the measured values; a wiring error between blocks (wrong buffer, wrong mask) measures >= 1.0 (seen during
    development). The 16-block bar has little discriminating power and is kept as a record."""
    e_fwd, e_glob, e_dx, worst = _chain(net, hi, lo)
    print("blocks %d..%d: forward %.2e, global param-grad %.2e, dx %.2e, worst tensors %s" % (lo, hi, e_fwd, e_glob, e_dx, worst))
    assert e_fwd <= (BAR if hi - lo < 8 else 1e-1) and e_glob <= bar and e_dx <= bar, (e_fwd, e_glob, e_dx)


@pytest.mark.parametrize("B2", [16, 128, 160])
def test_fused_batchnorm_equals_the_unfused_kernels(B2):
    """The grid-synchronous BatchNorm inside the flat convolutions (CF_FUSE) against the separate bn_apply / bn_bwd_apply
    launches it replaces: same model, same batch, train-mode forward + backward with the fusion on and off. Both derive the
    per-channel constants with the same code and apply them to the same bf16-rounded tensors, so activations, running
    statistics and gradients agree to the last bit or two of bf16 (a different FMA contraction is the only freedom). B = 160:
    layer1 does not fit tensor memory (10 tiles per CTA pair > 8 accumulator sets) and stays unfused while layers 2-4 fuse."""
    from cilrs_b200 import _lib
    from cilrs_b200.model import CILRS
    O = _O()
    sd = O.synthetic_state_dict(0)
    g = torch.Generator().manual_seed(5)
    coarse = torch.randn(B2, 3, 11, 25, generator=g)
    image = torch.nn.functional.interpolate(coarse, size=(88, 200), mode="bicubic", align_corners=False).contiguous().cuda()
    speed = torch.rand(B2, generator=g).cuda()
    command = torch.randint(0, 4, (B2,), generator=g).cuda()
    targets = torch.rand(B2, 3, generator=g).cuda()
    lib = _lib.lib()
    results = []
    prev = lib.cilrs_set_bn_fusion(1)
    try:
        for fuse in (1, 0):
            lib.cilrs_set_bn_fusion(fuse)
            m = CILRS(num_commands=4, dropout=0.0)
            m.load_state_dict(sd, strict=True)
            m = m.to("cuda").train()
            c, p = m(image, speed, command)
            loss, _ = O.loss_mse(c, targets, p, speed)
            loss.backward()
            torch.cuda.synchronize()
            acts = [m.debug_activation(i, B2).float().clone() for i in (1, 4, 8, 16)]
            results.append((c.detach().clone(), m.flat_gradients().clone(), m._flat_buf.clone(), acts))
    finally:
        lib.cilrs_set_bn_fusion(prev)
    (c1, g1, b1, a1), (c0, g0, b0, a0) = results
    for x, y in zip(a1, a0):
        assert _rel_max(x, y) <= 1e-2      # (one bf16 ulp on isolated elements at most)
    assert _rel_max(b1, b0) <= 1e-6        # running statistics: same sums, same finalize
    assert _rel_max(c1, c0) <= 1e-2
    e = _rel_l2(g1, g0)
    print("B=%d fused vs unfused BatchNorm: controls %.2e, global gradient %.2e" % (B2, _rel_max(c1, c0), e))
    assert e <= 2e-2
