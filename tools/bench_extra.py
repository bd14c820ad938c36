"""Secondary measurements of the hot path on one B200 (BASELINE.json configs 2 and 4, and the HBM-bound kernels):
  C3  frame preprocessing, 1024 uint8 600x800x3 frames -> 88x200 fp32 NCHW + conv1-ready bf16   (K0, HBM roofline)
  C5  batched inference, 512 frames per GPU (= 4096 / 8), mixed commands, eval mode              (tensor roofline)
  O1  fused Adam over the 22.4 M-parameter arena                                                  (HBM roofline)
  K3  heads forward / backward at batch 128
CUDA events, inputs resident in HBM, L2 flushed between repetitions (a 256 MB memset). Prints one JSON object per line."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cilrs_b200 import _lib, ops
from cilrs_b200.model import CILRS
from cilrs_b200.optim import FusedAdam

HBM = 6543.1
try:
    HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=10, cold=True):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if cold:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


only = sys.argv[1] if len(sys.argv) > 1 else "all"
sp = _lib.stream_ptr()
if only in ("all", "pre"):
    B = 1024
    frames = torch.randint(0, 256, (B, 600, 800, 3), dtype=torch.uint8, device="cuda")
    f32 = torch.empty(B, 3, 88, 200, device="cuda")
    s2d = torch.empty(B, 47, 103, 16, dtype=torch.bfloat16, device="cuda")
    t = timeit(lambda: _lib.call("cilrs_preprocess_u8", frames, B, 600, 800, 3, 0, 88, 200, None, f32, None, sp))
    t2 = timeit(lambda: _lib.call("cilrs_preprocess_u8", frames, B, 600, 800, 3, 0, 88, 200, None, None, s2d, sp))
    alg = 633600.0
    print(json.dumps({"bench": "C3 preprocess 1024 x (600x800x3 u8 -> 88x200 f32 NCHW)", "ms": t * 1e3, "frames_per_s": B / t,
                      "algorithmic_GBps": B * alg / t / 1e9, "hbm_frac": B * alg / t / 1e9 / HBM, "hbm_peak_GBps": HBM,
                      "bytes_per_frame": alg, "bf16_s2d_output_ms": t2 * 1e3, "bf16_s2d_frames_per_s": B / t2}))
    del frames, f32, s2d
if only in ("all", "infer"):
    B = 512
    torch.manual_seed(0)
    m = CILRS(num_commands=4, dropout=0.0).cuda().eval()
    img = torch.randn(B, 3, 88, 200, device="cuda")
    speed = torch.rand(B, device="cuda")
    cmd = torch.randint(0, 4, (B,), device="cuda")
    with torch.no_grad():
        t = timeit(lambda: m(img, speed, cmd), reps=10, cold=False)
    fl = 2795915264.0
    print(json.dumps({"bench": "C5 batched inference, 512 frames per GPU, eval mode (fused BN)", "ms": t * 1e3, "frames_per_s": B / t,
                      "conv_TFLOPs": B * fl / t / 1e12}))
    del m
if only in ("all", "adam"):
    n = 22421504
    p = torch.randn(n, device="cuda").requires_grad_(True)
    p.grad = torch.randn(n, device="cuda") * 1e-3
    mm, vv = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    t = timeit(lambda: _lib.call("cilrs_adam_step", p.data, p.grad, mm, vv, ctypes.c_longlong(n), ctypes.c_float(2e-4), ctypes.c_float(0.9),
                                 ctypes.c_float(0.999), ctypes.c_float(1e-8), ctypes.c_float(1e-4), ctypes.c_longlong(3), None,
                                 ctypes.c_float(1.0), None, sp))
    print(json.dumps({"bench": "O1 fused Adam, 22.4 M parameters (28 B/param)", "ms": t * 1e3, "GBps": n * 28 / t / 1e9,
                      "hbm_frac": n * 28 / t / 1e9 / HBM}))
if only in ("all", "heads"):
    B = 128
    m = CILRS(num_commands=4, dropout=0.0).cuda().train()
    m._ensure(B)
    feat = torch.randn(B, 512, device="cuda")
    speed = torch.rand(B, device="cuda")
    cmd = torch.randint(0, 4, (B,), device="cuda")
    c, ps = torch.empty(B, 3, device="cuda"), torch.empty(B, device="cuda")
    dc, dsp, df = torch.randn(B, 3, device="cuda"), torch.randn(B, device="cuda"), torch.empty(B, 512, device="cuda")
    m._refresh_if_needed(infer=False)
    m.flat_gradients()
    t = timeit(lambda: _lib.call("cilrs_model_heads_forward", m._handle, B, feat, speed, cmd, c, ps, 1, ctypes.c_float(0.0),
                                 ctypes.c_ulonglong(1), sp), cold=False)
    t2 = timeit(lambda: _lib.call("cilrs_model_heads_backward", m._handle, B, dc, dsp, speed, cmd, ctypes.c_float(0.0), df, sp), cold=False)
    print(json.dumps({"bench": "K3 heads at batch 128 (fp32 CUDA cores)", "forward_ms": t * 1e3, "backward_incl_wgrad_ms": t2 * 1e3}))
