"""Tile-shape sweep of the padded-flat conv kernel (CILRS_FLAT_SHAPE override), back-to-back timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cilrs_b200 import ops, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mode = sys.argv[2] if len(sys.argv) > 2 else "plain"


def timeit_hot(fn, reps=20):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


layers = [("layer1", 22, 50, 64), ("layer2", 11, 25, 128), ("layer3", 6, 13, 256), ("layer4", 3, 7, 512)]
P = lambda t: t.data_ptr()
for name, h, w, c in layers:
    d = ops.conv_desc(B, h, w, c, c, 3, 1)
    x = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
    wt = torch.randn(c, c, 3, 3, device="cuda") * 0.05
    wf, wd = ops.pack_weight(d, wt)
    out = torch.empty_like(x)
    ws = torch.zeros(_lib.query("cilrs_conv_flat_workspace_floats", c), device="cuda")
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    gamma, beta, rm, rv = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    vec = torch.zeros(4, c, device="cuda")
    a = _lib.FlatConvArgs()
    a.batch, a.H, a.W, a.in_c, a.out_c, a.dgrad, a.flags = B, h, w, c, c, 0, ((ops.EPI_STATS | ops.EPI_DEFER) if mode == "bn" else 0)
    a.x, a.w, a.y = P(x), P(wf), P(out)
    if mode == "dgrad":  # what the network's backward launches: residual + ReLU mask bits + BN-backward sums, deferred finalize
        act = ops.to_padded(torch.relu(torch.randn(B, h, w, c, device="cuda")).to(torch.bfloat16))
        y1 = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
        resid = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
        bits = ops.relu_bits(act)
        a.dgrad, a.flags = 1, ops.EPI_RESIDUAL | ops.EPI_MASK | ops.EPI_BNBWD | ops.EPI_DEFER
        a.w, a.residual, a.mask, a.mask_bits, a.y1 = P(wd), P(resid), P(act), P(bits), P(y1)
    a.gamma, a.beta, a.running_mean, a.running_var, a.vec = P(gamma), P(beta), P(rm), P(rv), P(vec)
    a.momentum, a.eps, a.update_running = 0.1, 1e-5, 1
    a.partials_ws, a.counter_ws = P(ws), P(cnt)
    sp = _lib.stream_ptr()
    flops = 2.0 * B * h * w * c * c * 9
    res = []
    for mt in (1, 2, 4):
        for bn in (64, 128, 256):
            if c % bn or mt * bn > 512:
                continue
            for r, G in ((1, 9), (0, 3), (0, 2), (0, 1)):
              for pr in (0, 1):
                os.environ["CILRS_FLAT_SHAPE"] = "%d,%d,%d,%d,%d" % (mt, bn, r, G, pr)
                os.environ["CILRS_FLAT_DEBUG"] = "1"
                try:
                    t = timeit_hot(lambda: _lib.call("cilrs_conv_flat", a, sp))
                except RuntimeError as e:
                    continue
                res.append((t, mt, bn, r, G, pr))
    del os.environ["CILRS_FLAT_SHAPE"]
    del os.environ["CILRS_FLAT_DEBUG"]
    t_auto = timeit_hot(lambda: _lib.call("cilrs_conv_flat", a, sp))
    res.sort()
    print("%s %s auto %.1f us (%.0f TF) | best:" % (name, mode, t_auto * 1e3, flops / t_auto / 1e9),
          "  ".join("%.1fus %smt%d bn%d r%d G%d" % (t * 1e3, "PAIR " if pr else "", mt, bn, r, G) for t, mt, bn, r, G, pr in res[:12]))
