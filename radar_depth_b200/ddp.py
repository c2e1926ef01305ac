"""Batch-sharded data parallelism: one process per GPU, identical replicas, BatchNorm statistics and
loss means local to the rank (torch DDP semantics; the reference has no multi-GPU mode), and ONE exchange per step:
an all-reduce (SUM, then /world) of the flat gradient arena over NCCL / NVLink 5.  The arena is a single contiguous
fp32 bucket (58.8 MB for latefusion, 2 x 58.8 MB for multistage), so there is exactly one collective per engine."""
from __future__ import annotations

import torch
import torch.distributed as dist

from .optim import _engines


def broadcast_parameters(model: torch.nn.Module, src: int = 0) -> None:
    """Rank ``src``'s weights and BN buffers to every rank (call once after construction)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src)


def allreduce_gradients(model: torch.nn.Module) -> int:
    """Average gradients across ranks.  Returns the number of collectives issued."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    world = dist.get_world_size()
    n = 0
    owned = set()
    for eng in _engines(model):
        dist.all_reduce(eng.gflat, op=dist.ReduceOp.SUM)
        eng.gflat.div_(world)
        n += 1
        for _, p in eng.module.named_parameters():
            owned.add(id(p))
    loose = [p.grad for p in model.parameters() if id(p) not in owned and p.grad is not None]
    if loose:
        flat = torch.cat([g.reshape(-1) for g in loose])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        off = 0
        for g in loose:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n += 1
    return n
