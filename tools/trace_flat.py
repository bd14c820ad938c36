"""Event timeline of CTA 0 of the padded-flat conv kernel (needs tools/libcilrs_trace.so: CILRS_B200_LIB=tools/libcilrs_trace.so)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cilrs_b200 import ops, _lib

li = int(sys.argv[1]); B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
mode = sys.argv[3] if len(sys.argv) > 3 else "plain"
h, w, c = [(22, 50, 64), (11, 25, 128), (6, 13, 256), (3, 7, 512)][li - 1]
P = lambda t: t.data_ptr()
d = ops.conv_desc(B, h, w, c, c, 3, 1)
x = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
wf, wd = ops.pack_weight(d, torch.randn(c, c, 3, 3, device="cuda") * 0.05)
out = torch.empty_like(x)
a = _lib.FlatConvArgs()
a.batch, a.H, a.W, a.in_c, a.out_c, a.dgrad, a.flags = B, h, w, c, c, 0, 0
a.x, a.w, a.y = P(x), P(wf), P(out)
ws = torch.zeros(_lib.query("cilrs_conv_flat_workspace_floats", c), device="cuda")
cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
gamma, beta, rm, rv = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
vec = torch.zeros(4, c, device="cuda"); vec[3] = 1
a.partials_ws, a.counter_ws = P(ws), P(cnt)
if mode == "dgrad":
    dy = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
    act = ops.to_padded(torch.relu(torch.randn(B, h, w, c, device="cuda")).to(torch.bfloat16))
    y1 = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
    res = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
    bred = torch.zeros(2, c, device="cuda"); dg = torch.zeros(c, device="cuda"); db = torch.zeros(c, device="cuda")
    a.dgrad, a.flags = 1, ops.EPI_RESIDUAL | ops.EPI_MASK | ops.EPI_BNBWD
    bits = ops.relu_bits(act)
    a.x, a.w, a.y, a.residual, a.mask, a.mask_bits = P(dy), P(wd), P(out), P(res), P(act), P(bits)
    a.y1, a.vec1, a.bred1, a.dgamma1, a.dbeta1 = P(y1), P(vec), P(bred), P(dg), P(db)
if mode == "bn":
    a.flags = ops.EPI_STATS
    a.gamma, a.beta, a.running_mean, a.running_var, a.vec = P(gamma), P(beta), P(rm), P(rv), P(vec)
    a.momentum, a.eps, a.update_running = 0.1, 1e-5, 1
sp = _lib.stream_ptr()
buf = (ctypes.c_ulonglong * (3 * 2048))()
cnt = (ctypes.c_int * 3)()
for it in range(3):
    _lib.call("cilrs_conv_flat", a, sp)
    _lib.lib().cilrs_conv_flat_trace(buf, cnt)
ev = []
for role in range(3):
    for i in range(cnt[role]):
        v = buf[role * 2048 + i]
        ev.append((v >> 16, role, v & 0xFFFF))
ev.sort()
t0 = ev[0][0]
names = {0: "TMA", 1: "MMA", 2: "EPI"}
print("layer%d B=%d events=%d span=%d clk" % (li, B, len(ev), ev[-1][0] - t0))
last = {0: t0, 1: t0, 2: t0}
for t, role, code in ev[:int(os.environ.get("TRACE_MAX", "400"))]:
    kind = {0x100: "A", 0x200: "B", 0x300: "acc-free", 0x400: "tile-commit", 0x500: "acc-full", 0x600: "tile-epi-done", 0x700: "stats"}[code & 0xF00]
    print("%8d  (+%6d)  %s %s %d" % (t - t0, t - last[role], names[role], kind, code & 0xFF))
    last[role] = t
