"""Per-layer timing of the padded-flat conv kernels at the BASELINE batch (CUDA events, L2 flushed between reps).
Arguments are pre-built so only the kernel launch is inside the timed region."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cilrs_b200 import ops, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def timeit_hot(fn, reps=20):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


layers = [("layer1 3x3 64->64", 22, 50, 64, 6), ("layer2 3x3 128->128", 11, 25, 128, 7), ("layer3 3x3 256->256", 6, 13, 256, 11),
          ("layer4 3x3 512->512", 3, 7, 512, 5)]
tot = {"fprop": 0.0, "dgrad": 0.0, "wgrad": 0.0}
print("B=%d (cold L2 / back-to-back)" % B)
P = lambda t: t.data_ptr()
for name, h, w, c, count in layers:
    d = ops.conv_desc(B, h, w, c, c, 3, 1)
    x = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
    dy = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
    act = ops.to_padded(torch.relu(torch.randn(B, h, w, c, device="cuda")).to(torch.bfloat16))
    y1 = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
    res = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
    wt = torch.randn(c, c, 3, 3, device="cuda") * 0.05
    wf, wd = ops.pack_weight(d, wt)
    out = torch.empty_like(x)
    ws = torch.zeros(_lib.query("cilrs_conv_flat_workspace_floats", c), device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    gamma, beta, rm, rv = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    vec = torch.zeros(4, c, device="cuda"); vec[3] = 1
    bred = torch.zeros(2, c, device="cuda")
    dg, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    dw = torch.zeros(c, c, 3, 3, device="cuda")
    wsw = torch.empty(_lib.query("cilrs_wgrad_flat_workspace_bytes") // 4, device="cuda")
    a = _lib.FlatConvArgs()
    a.batch, a.H, a.W, a.in_c, a.out_c, a.dgrad, a.flags = B, h, w, c, c, 0, ops.EPI_STATS
    a.x, a.w, a.y = P(x), P(wf), P(out)
    a.gamma, a.beta, a.running_mean, a.running_var, a.vec = P(gamma), P(beta), P(rm), P(rv), P(vec)
    a.momentum, a.eps, a.update_running = 0.1, 1e-5, 1
    a.partials_ws, a.counter_ws = P(ws), P(cnt)
    g = _lib.FlatConvArgs()
    g.batch, g.H, g.W, g.in_c, g.out_c, g.dgrad = B, h, w, c, c, 1
    g.flags = ops.EPI_RESIDUAL | ops.EPI_MASK | ops.EPI_BNBWD
    bits = ops.relu_bits(act)
    g.x, g.w, g.y, g.residual, g.mask, g.mask_bits = P(dy), P(wd), P(out), P(res), P(act), P(bits)
    g.y1, g.vec1, g.bred1, g.dgamma1, g.dbeta1 = P(y1), P(vec), P(bred), P(dg), P(db)
    g.partials_ws, g.counter_ws = P(ws), P(cnt)
    p = _lib.FlatConvArgs()
    p.batch, p.H, p.W, p.in_c, p.out_c, p.dgrad, p.flags = B, h, w, c, c, 0, 0
    p.x, p.w, p.y = P(x), P(wf), P(out)
    sp = _lib.stream_ptr()
    f_f = lambda: _lib.call("cilrs_conv_flat", a, sp)
    f_p = lambda: _lib.call("cilrs_conv_flat", p, sp)
    f_d = lambda: _lib.call("cilrs_conv_flat", g, sp)
    f_w = lambda: _lib.call("cilrs_wgrad_flat", B, h, w, c, c, dy, x, dw, wsw, sp)
    flops = 2.0 * B * h * w * c * c * 9
    t_p, t_f, t_d, t_w = timeit(f_p), timeit(f_f), timeit(f_d), timeit(f_w)
    h_p, h_f, h_d, h_w = timeit_hot(f_p), timeit_hot(f_f), timeit_hot(f_d), timeit_hot(f_w)
    tf = lambda t: flops / t / 1e9
    print("%-22s plain %6.1f/%6.1f us %5.0f TF | fprop+bn %6.1f/%6.1f us %5.0f TF | dgrad+bnbwd %6.1f/%6.1f us %5.0f TF | wgrad %6.1f/%6.1f us %5.0f TF (x%d)" %
          (name, t_p * 1e3, h_p * 1e3, tf(h_p), t_f * 1e3, h_f * 1e3, tf(h_f), t_d * 1e3, h_d * 1e3, tf(h_d), t_w * 1e3, h_w * 1e3, tf(h_w), count))
    tot["fprop"] += h_f * count; tot["dgrad"] += h_d * count; tot["wgrad"] += h_w * count
print("stride-1 3x3 totals per step (ms, back-to-back):", tot, "sum %.3f" % sum(tot.values()))
