"""Three eager training steps at the benchmark configuration (B = 128, u8 frames, MSE recipe) - the program the ncu captures under
profiles/r02_* profile. python tools/prof_step.py [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import cilrs_b200  # noqa: F401
from cilrs_b200.model import CILRS
from cilrs_b200.train import FusedTrainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
sd, _ = bench.own_initial_state_dict()
model = CILRS(num_commands=4, dropout=0.0)
model.load_state_dict(sd, strict=True)
model = model.to(dev)
tr = FusedTrainer(model, bench.BATCH, lr=2e-4, weight_decay=1e-4, loss="mse", speed_w=0.05, frames="u8", use_graph=False)
devb = [tuple(t.to(dev) for t in b) for b in bench.synthetic_host_batches(bench.BATCH, 2, 100)]
for i in range(steps):
    tr.load_batch(*devb[i % 2]); tr.step()
torch.cuda.synchronize()
print("loss", tr.read_loss()["total"])
