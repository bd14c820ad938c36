"""K0 alone on configs[2]: 1024 frames 600x800x3 u8 -> 88x200 fp32 NCHW (one launch per variant), for ncu and for timing."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cilrs_b200 import _lib
B = 1024
sp = _lib.stream_ptr()
frames = torch.randint(0, 256, (B, 600, 800, 3), dtype=torch.uint8, device="cuda")
f32 = torch.empty(B, 3, 88, 200, device="cuda")
s2d = torch.empty(B, 47, 103, 16, dtype=torch.bfloat16, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, reps=10):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
ms = t(lambda: _lib.call("cilrs_preprocess_u8", frames, B, 600, 800, 3, 0, 88, 200, None, f32, None, sp))
ms2 = t(lambda: _lib.call("cilrs_preprocess_u8", frames, B, 600, 800, 3, 0, 88, 200, None, None, s2d, sp))
print(json.dumps({"f32_ms": ms, "hbm_frac_f32": B * 633600 / (ms * 1e-3) / 1e9 / 6543.1, "s2d_ms": ms2}))
