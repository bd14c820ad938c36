"""Fused Adam variants timed alone (CUDA events, L2 flushed): by-value entry, device-hyper entry with / without the fused
zero_grad, bf16 gradient source. One JSON line per variant."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cilrs_b200 import _lib

n = 22421504
sp = _lib.stream_ptr()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
p = torch.randn(n, device="cuda"); g = torch.randn(n, device="cuda") * 1e-3
g16 = g.to(torch.bfloat16)
mm, vv = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
hyper = torch.tensor([2e-4, 0.9, 0.999, 1e-8, 1e-4, 1.0, 0, 0], device="cuda")
step = torch.zeros(1, dtype=torch.long, device="cuda")
L = ctypes.c_longlong(n)
F = ctypes.c_float


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        g.normal_(0, 1e-3)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


variants = {
    "by_value": lambda: _lib.call("cilrs_adam_step", p, g, mm, vv, L, F(2e-4), F(0.9), F(0.999), F(1e-8), F(1e-4), ctypes.c_longlong(3), None, F(1.0), None, sp),
    "by_value_stepdev": lambda: _lib.call("cilrs_adam_step", p, g, mm, vv, L, F(2e-4), F(0.9), F(0.999), F(1e-8), F(1e-4), ctypes.c_longlong(0), step, F(1.0), None, sp),
    "ex": lambda: _lib.call("cilrs_adam_step_ex", p, g, None, mm, vv, L, hyper, step, None, 0, sp),
    "ex_zero": lambda: _lib.call("cilrs_adam_step_ex", p, g, None, mm, vv, L, hyper, step, None, 1, sp),
    "ex_bf16": lambda: _lib.call("cilrs_adam_step_ex", p, g, g16, mm, vv, L, hyper, step, None, 0, sp),
    "to_bf16_zero": lambda: _lib.call("cilrs_grad_to_bf16", g, g16, L, 1, sp),
}
for name, fn in variants.items():
    print(json.dumps({"variant": name, "ms": timeit(fn)}))
