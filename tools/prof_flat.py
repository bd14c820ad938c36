"""Runs one padded-flat conv configuration a few times (driver for ncu). usage: prof_flat.py <layer 1..4> <plain|bn|dgrad|wgrad> [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cilrs_b200 import ops, _lib

li = int(sys.argv[1]); mode = sys.argv[2]; B = int(sys.argv[3]) if len(sys.argv) > 3 else 128
h, w, c = [(22, 50, 64), (11, 25, 128), (6, 13, 256), (3, 7, 512)][li - 1]
P = lambda t: t.data_ptr()
d = ops.conv_desc(B, h, w, c, c, 3, 1)
x = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
dy = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
act = ops.to_padded(torch.relu(torch.randn(B, h, w, c, device="cuda")).to(torch.bfloat16))
y1 = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
res = ops.to_padded(torch.randn(B, h, w, c, device="cuda").to(torch.bfloat16))
wf, wd = ops.pack_weight(d, torch.randn(c, c, 3, 3, device="cuda") * 0.05)
out = torch.empty_like(x)
ws = torch.zeros(_lib.query("cilrs_conv_flat_workspace_floats", c), device="cuda")
cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
gamma, beta, rm, rv = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
vec = torch.zeros(4, c, device="cuda"); vec[3] = 1
bred = torch.zeros(2, c, device="cuda"); dg = torch.zeros(c, device="cuda"); db = torch.zeros(c, device="cuda")
dw = torch.zeros(c, c, 3, 3, device="cuda")
wsw = torch.empty(_lib.query("cilrs_wgrad_flat_workspace_bytes") // 4, device="cuda")
a = _lib.FlatConvArgs()
a.batch, a.H, a.W, a.in_c, a.out_c = B, h, w, c, c
a.partials_ws, a.counter_ws = P(ws), P(cnt)
if mode in ("plain", "bn"):
    a.dgrad, a.flags = 0, (ops.EPI_STATS if mode == "bn" else 0)
    a.x, a.w, a.y = P(x), P(wf), P(out)
    a.gamma, a.beta, a.running_mean, a.running_var, a.vec = P(gamma), P(beta), P(rm), P(rv), P(vec)
    a.momentum, a.eps, a.update_running = 0.1, 1e-5, 1
elif mode == "dgrad":
    a.dgrad, a.flags = 1, ops.EPI_RESIDUAL | ops.EPI_MASK | ops.EPI_BNBWD
    bits = ops.relu_bits(act)
    a.x, a.w, a.y, a.residual, a.mask, a.mask_bits = P(dy), P(wd), P(out), P(res), P(act), P(bits)
    a.y1, a.vec1, a.bred1, a.dgamma1, a.dbeta1 = P(y1), P(vec), P(bred), P(dg), P(db)
sp = _lib.stream_ptr()
for _ in range(6):
    if mode == "wgrad":
        _lib.call("cilrs_wgrad_flat", B, h, w, c, c, dy, x, dw, wsw, sp)
    else:
        _lib.call("cilrs_conv_flat", a, sp)
torch.cuda.synchronize()
print("ok")
