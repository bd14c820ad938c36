"""Kernel timeline of the captured training step (CUPTI through torch.profiler): which kernels overlap, where the SMs idle.
   python tools/step_timeline.py [out.csv]      (GPU box only; B = 128, same trainer configuration as bench.py)"""
import os, sys, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import cilrs_b200  # noqa: F401
from cilrs_b200.model import CILRS
from cilrs_b200.train import FusedTrainer

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r2_timeline.csv"
dev = torch.device("cuda", 0)
sd, _ = bench.own_initial_state_dict()
model = CILRS(num_commands=4, dropout=0.0)
model.load_state_dict(sd, strict=True)
model = model.to(dev)
tr = FusedTrainer(model, bench.BATCH, lr=2e-4, weight_decay=1e-4, loss="mse", speed_w=0.05, frames="u8", use_graph=True)
devb = [tuple(t.to(dev) for t in b) for b in bench.synthetic_host_batches(bench.BATCH, 2, 100)]
for i in range(6):
    tr.load_batch(*devb[i % 2]); tr.step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        tr.load_batch(*devb[i % 2]); tr.step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
with open(out, "w") as f:
    f.write("start_us,dur_us,stream,name\n")
    for e in ev:
        name = re.sub(r"\(.*", "", e.name).replace("cilrs::", "")
        stream = getattr(e, "stream", None)
        f.write("%.2f,%.2f,%s,%s\n" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, stream, name))
print("events", len(ev), "->", out)
