"""Back-to-back timing of the BatchNorm elementwise kernels per layer geometry (C-ABI entry points, non-deferred)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cilrs_b200 import ops, _lib

B = 128
def timeit(fn, reps=20):
    """device time per launch: `reps` launches captured in one CUDA graph (no host launch overhead in the number)"""
    fn(_lib.stream_ptr()); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        sp = _lib.stream_ptr()
        for _ in range(reps): fn(sp)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for h, w, c in [(22, 50, 64), (11, 25, 128), (6, 13, 256), (3, 7, 512)]:
    x = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
    res = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
    out = torch.empty_like(x)
    bits = torch.empty(B, h + 1, w + 1, c // 8, dtype=torch.uint8, device="cuda")
    vec = torch.stack([torch.ones(c), torch.zeros(c), torch.zeros(c), torch.ones(c)]).cuda().contiguous()
    n = ctypes.c_longlong(x.numel())
    t1 = timeit(lambda sp: _lib.call("cilrs_bn_apply", x, vec, None, None, None, out, n, c, 1, h, w, bits, sp))
    t2 = timeit(lambda sp: _lib.call("cilrs_bn_apply", x, vec, res, None, None, out, n, c, 1, h, w, bits, sp))
    mb = x.numel() * 2 / 1e6
    print("C=%d %dx%d  tensor %.1f MB | bn_apply %.1f us (%.0f GB/s) | +residual %.1f us (%.0f GB/s)" %
          (c, h, w, mb, t1, 2.06 * mb / t1 * 1e3, t2, 3.06 * mb / t2 * 1e3))
