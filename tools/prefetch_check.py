"""Run-to-run determinism of the trainer and prefetch vs direct loading (GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cilrs_b200  # noqa
from cilrs_b200.model import CILRS
from cilrs_b200.train import FusedTrainer
torch.manual_seed(0)
sd = {k: v.clone() for k, v in CILRS(num_commands=4, dropout=0.0).state_dict().items()}   # one random initialisation for every run
g = torch.Generator().manual_seed(21)
host = []
for _ in range(4):
    img = torch.randint(0, 256, (8, 88, 200, 3), generator=g, dtype=torch.uint8)
    host.append(tuple(t.pin_memory() for t in (img, torch.rand(8, generator=g), torch.randint(0, 4, (8,), generator=g), torch.rand(8, 3, generator=g))))
for use_graph in (True, False):
    for mode in ("direct", "direct", "prefetch", "prefetch", "direct_sync"):
        m = CILRS(num_commands=4, dropout=0.0)
        m.load_state_dict(sd, strict=True)
        m = m.cuda().train()
        tr = FusedTrainer(m, 8, lr=1e-3, weight_decay=1e-4, eps=1e-3, use_graph=use_graph, frames="u8")
        losses = []
        if mode == "prefetch":
            tr.prefetch_batch(*host[0])
        for i in range(4):
            if mode == "prefetch":
                tr.load_prefetched()
                if i + 1 < 4:
                    tr.prefetch_batch(*host[i + 1])
            else:
                tr.load_batch(*host[i])
            if mode == "direct_sync":
                torch.cuda.synchronize()
            tr.step()
            losses.append(tr.read_loss()["total"])
        print(use_graph, mode, ["%.7f" % v for v in losses], float(m.flat_parameters().double().sum()))
