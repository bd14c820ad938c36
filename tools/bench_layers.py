"""Per-layer timing of the tcgen05 conv kernels at the BASELINE batch (CUDA events, L2 flushed between reps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cilrs_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


layers = [("layer1 3x3 64->64", 22, 50, 64, 64, 3, 1, 6), ("layer2.0 3x3/2 64->128", 22, 50, 64, 128, 3, 2, 1),
          ("layer2 ds 1x1/2", 22, 50, 64, 128, 1, 2, 1), ("layer2 3x3 128->128", 11, 25, 128, 128, 3, 1, 7),
          ("layer3.0 3x3/2", 11, 25, 128, 256, 3, 2, 1), ("layer3 3x3 256->256", 6, 13, 256, 256, 3, 1, 11),
          ("layer4.0 3x3/2", 6, 13, 256, 512, 3, 2, 1), ("layer4 3x3 512->512", 3, 7, 512, 512, 3, 1, 5)]
tot = {"fprop": 0.0, "dgrad": 0.0, "wgrad": 0.0}
print("B=%d" % B)
for name, h, w, ci, co, k, s, count in layers:
    d = ops.conv_desc(B, h, w, ci, co, k, s)
    oh, ow = ops.out_hw(d)
    x = torch.randn(B, h, w, ci, device="cuda").to(torch.bfloat16)
    dy = torch.randn(B, oh, ow, co, device="cuda").to(torch.bfloat16)
    wt = torch.randn(co, ci, k, k, device="cuda") * 0.05
    wf, wd = ops.pack_weight(d, wt)
    flops = 2.0 * B * oh * ow * co * ci * k * k
    t_f = timeit(lambda: ops.conv_fprop(d, x, wf, stats=True))
    t_d = timeit(lambda: ops.conv_dgrad(d, dy, wd))
    t_w = timeit(lambda: ops.conv_wgrad(d, dy, x))
    print("%-26s fprop %7.3f ms %6.1f TF | dgrad %7.3f ms %6.1f TF | wgrad %7.3f ms %6.1f TF  (x%d)" %
          (name, t_f, flops / t_f / 1e9, t_d, flops / t_d / 1e9, t_w, flops / t_w / 1e9, count))
    tot["fprop"] += t_f * count; tot["dgrad"] += t_d * count; tot["wgrad"] += t_w * count
img = torch.randn(B, 3, 88, 200, device="cuda")
xs = ops.image_to_s2d(img)
wp = ops.stem_pack_weight(torch.randn(64, 3, 7, 7, device="cuda") * 0.05)
dy = torch.randn(B, 44, 100, 64, device="cuda").to(torch.bfloat16)
flops = 2.0 * B * 44 * 100 * 64 * 147
t_f = timeit(lambda: ops.stem_fprop(xs, wp, stats=True))
t_w = timeit(lambda: ops.stem_wgrad(dy, xs))
print("%-26s fprop %7.3f ms %6.1f TF |                          | wgrad %7.3f ms %6.1f TF" % ("stem 7x7/2", t_f, flops / t_f / 1e9, t_w, flops / t_w / 1e9))
tot["fprop"] += t_f; tot["wgrad"] += t_w
print("conv totals per step (ms):", tot, "sum %.3f" % sum(tot.values()))
print("conv-only roofline fraction at this sum: %.3f of 1645.6 TF" % (B * 8.304955392e9 / (sum(tot.values()) * 1e-3) / 1645.6e12))
