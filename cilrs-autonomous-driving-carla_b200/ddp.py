"""Data-parallel plumbing (new: the reference is single-GPU, notebook/notebook.ipynb:479).

The batch is sharded across ranks (one process per GPU); parameters, Adam state and BN running statistics are replicated
(BN batch statistics stay per rank, plain-DDP semantics). The only exchange per step is a sum-allreduce of the flat gradient
arena, issued in five ranges that the backward pass completes back to front (heads+layer4, layer3, layer2, layer1, stem) so
that each range's NCCL allreduce overlaps the backward kernels of the next; the 1/world average is folded into the fused
Adam's grad_scale.
"""
import torch
import torch.distributed as dist

from . import _lib


# which backward parts (0 = heads + layer4, 1 = layer3, 2 = layer2, 3 = layer1, 4 = stem; completed in this order) share one
# allreduce: a group's collective starts when its LAST part is complete and covers [lo(last part), hi(first part))
SCHEDULES = {
    "all": ((0,), (1,), (2,), (3,), (4,)),
    "two": ((0,), (1,), (2, 3, 4)),
    "first": ((0,), (1, 2, 3, 4)),
    "tail": ((0, 1, 2, 3, 4),),
}


def schedule_ranges(part_ranges, schedule):
    """[(lo, hi)] element range each group of `schedule` exchanges (contiguous because the parts complete back to front)."""
    return [(part_ranges[g[-1]][0], part_ranges[g[0]][1]) for g in schedule]


def backward_part_ranges(model):
    """[(lo, hi)] element ranges of the flat gradient arena completed by backward part 0..4."""
    lib = _lib.lib()
    nt = len(model._offsets)
    bounds = [nt] + [lib.cilrs_model_backward_part_first_tensor(p) for p in range(5)]
    out = []
    for p in range(5):
        lo = model._offsets[bounds[p + 1]]
        hi = model._total if p == 0 else model._offsets[bounds[p]]
        out.append((lo, hi))
    return out


def allreduce_ranges(flat_grad, ranges, group=None):
    """Start an asynchronous sum-allreduce per range; returns the work handles (wait() before the optimizer)."""
    return [dist.all_reduce(flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True) for lo, hi in ranges]


def broadcast_parameters(model, src=0, group=None):
    """Make every rank start from rank `src`'s parameters and BN buffers."""
    dist.broadcast(model.flat_parameters(), src, group=group)
    dist.broadcast(model._flat_buf, src, group=group)
    dist.broadcast(model._flat_nbt, src, group=group)
    model.mark_parameters_changed()
