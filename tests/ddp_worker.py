"""torchrun worker of tests/test_ddp_multigpu.py (NCCL, one rank per GPU): for every allreduce schedule x gradient exchange
dtype x (async parts) x (CUDA graph) combination FusedTrainer offers:
 (1) the exchanged gradient equals the sum over ranks of the local gradients (all_gather'ed): exactly for the fp32 exchange
     (the weight gradients are deterministic), to bf16 rounding for the bf16 exchange;
 (2) after 3 steps every rank holds bit-identical parameters and BN statistics stay per rank (plain-DDP semantics)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from cilrs_b200.ddp import SCHEDULES  # noqa: E402
from cilrs_b200.model import CILRS  # noqa: E402
from cilrs_b200.train import FusedTrainer  # noqa: E402

B = 16
g = torch.Generator().manual_seed(100 + rank)
frames = torch.randint(0, 256, (B, 88, 200, 3), generator=g, dtype=torch.uint8).cuda()
speed, cmd, tgt = torch.rand(B, generator=g).cuda(), torch.randint(0, 4, (B,), generator=g).cuda(), torch.rand(B, 3, generator=g).cuda()
configs = [(s, c, a, gr) for s in sorted(SCHEDULES) for c in ("bf16", "fp32") for a in (False, True) for gr in (False,)]
configs += [("two", "bf16", False, True), ("all", "fp32", True, True)]
for sched, comm, asy, graph in configs:
    torch.manual_seed(1234 + rank)                      # different init per rank on purpose: the trainer's broadcast must fix it
    model = CILRS().cuda()
    tr = FusedTrainer(model, B, frames="u8", overlap_allreduce=sched, grad_comm=comm, async_parts=asy, use_graph=False)
    p0 = model.flat_parameters().clone()
    chk = [torch.zeros_like(p0) for _ in range(world)]
    dist.all_gather(chk, p0)
    assert all(torch.equal(chk[0], c) for c in chk), "broadcast_parameters failed"
    # local gradient of this rank's shard through the plain (non-exchanging) path of the same kernels
    tr.load_batch(frames, speed, cmd, tgt)
    w, tr.world = tr.world, 1
    tr.skip_optimizer = True
    nbt0 = model._flat_nbt.clone(); buf0 = model._flat_buf.clone()
    tr._device_step()
    torch.cuda.synchronize()
    local_grad = model.flat_gradients().clone()
    model.flat_gradients().zero_()
    model._flat_nbt.copy_(nbt0); model._flat_buf.copy_(buf0)
    tr.world = w
    allg = [torch.zeros_like(local_grad) for _ in range(world)]
    dist.all_gather(allg, local_grad)
    expect = torch.stack(allg).sum(0)
    tr._device_step()                                   # with the exchange, still without the optimizer
    torch.cuda.synchronize()
    got = tr.g16.float() if comm == "bf16" else model.flat_gradients()
    err = float((got - expect).abs().max() / expect.abs().max())
    if comm == "fp32":
        assert torch.equal(got, expect) or err < 1e-6, (sched, comm, asy, err)
    else:
        assert err < 1e-2, (sched, comm, asy, err)
        assert float(model.flat_gradients().abs().max()) == 0.0   # the conversion left the fp32 arena zeroed
    model.flat_gradients().zero_()
    model._flat_nbt.copy_(nbt0); model._flat_buf.copy_(buf0)
    tr.skip_optimizer = False
    if graph:
        tr = FusedTrainer(model, B, frames="u8", overlap_allreduce=sched, grad_comm=comm, async_parts=asy, use_graph=True)
        assert tr.graph is not None, tr.graph_error
        tr.load_batch(frames, speed, cmd, tgt)
    for it in range(3):
        loss6 = tr.step()
    torch.cuda.synchronize()
    pf = model.flat_parameters()
    chk = [torch.zeros_like(pf) for _ in range(world)]
    dist.all_gather(chk, pf)
    same = all(torch.equal(chk[0], c) for c in chk)
    bufs = [torch.zeros_like(model._flat_buf) for _ in range(world)]
    dist.all_gather(bufs, model._flat_buf)
    per_rank_bn = not torch.equal(bufs[0], bufs[-1])
    if rank == 0:
        print("ddp world=%d schedule=%-5s comm=%s async=%d graph=%d: exchanged-vs-gathered rel err %.2e, params identical %s, BN stats per rank %s, loss %.5f"
              % (world, sched, comm, asy, graph, err, same, per_rank_bn, float(loss6[0])), flush=True)
    assert same and per_rank_bn and not torch.equal(pf, p0)
    tr.close()
    del tr, model
dist.barrier()
if rank == 0:
    print("DDP_WORKER_OK", flush=True)
dist.destroy_process_group()
