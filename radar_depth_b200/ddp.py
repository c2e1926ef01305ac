"""Batch-sharded data parallelism: one process per GPU, identical replicas, BatchNorm statistics and
loss means local to the rank (torch DDP semantics; the reference has no multi-GPU mode), and ONE exchange per step: the
all-reduce (SUM, averaged by 1/world) of the flat gradient arena over NCCL / NVLink 5 (58.8 MB for latefusion, 2 x 58.8 MB
for multistage).

Two ways to issue it:
  * ``allreduce_gradients(model)`` after ``loss.backward()``: one collective per engine arena, fully exposed;
  * ``enable_overlap(model)`` once, then the same call: the engine runs its backward in three segments
    (engine._grad_buckets: head+decoder+fusion | RGB layer4 | the rest) and starts the all-reduce of each bucket as soon as
    its segment has been enqueued, so NCCL works beside the remaining backward kernels; ``allreduce_gradients`` then only
    waits for the outstanding collectives (and reduces whatever was not covered: parameters outside the engines, or a
    backward that accumulated into existing gradients).
Passing the FusedSGD optimizer lets the 1/world averaging ride in the update kernel (rd_sgd_scaled) instead of a pass over
the arena."""
from __future__ import annotations

import torch
import torch.distributed as dist

from .optim import _engines


def _active() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def broadcast_parameters(model: torch.nn.Module, src: int = 0) -> None:
    """Rank ``src``'s weights and BN buffers to every rank (call once after construction)."""
    if not _active():
        return
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src)


def _bucket_hook(eng, k: int) -> None:
    """Called by the engine after backward segment k has been enqueued (engine.backward)."""
    if not _active():
        return
    works = eng.__dict__.setdefault("_rd_works", [])
    for a, b in eng.grad_ranges[k]:
        if b > a:
            works.append(dist.all_reduce(eng.gflat[a:b], op=dist.ReduceOp.SUM, async_op=True))
    if k == 2:
        eng._rd_reduced = True


def enable_overlap(model: torch.nn.Module, on: bool = True) -> None:
    """Bucketed all-reduce overlapped with the backward pass, for every engine-backed network inside ``model``."""
    for m in model.modules():
        if hasattr(m, "_get_engine"):
            m._rd_grad_hook = _bucket_hook if on else None


def allreduce_gradients(model: torch.nn.Module, optimizer=None) -> int:
    """Average gradients across ranks.  Returns the number of collectives issued for this step.  With ``optimizer`` (a
    FusedSGD) the arenas are left holding the SUM and the optimizer applies 1/world in its update kernel."""
    if not _active():
        return 0
    world = dist.get_world_size()
    n = 0
    owned = set()
    fold = optimizer is not None and hasattr(optimizer, "grad_scale")
    for eng in _engines(model):
        works = eng.__dict__.get("_rd_works", [])
        if getattr(eng, "_rd_reduced", False):
            for w in works:
                w.wait()                       # current stream waits for NCCL's stream
            n += len(works)
        else:
            dist.all_reduce(eng.gflat, op=dist.ReduceOp.SUM)
            n += 1
        eng._rd_works, eng._rd_reduced = [], False
        if not fold:
            eng.gflat.div_(world)
        for _, p in eng.module.named_parameters():
            owned.add(id(p))
    loose = [p.grad for p in model.parameters() if id(p) not in owned and p.grad is not None]
    if loose:
        flat = torch.cat([g.reshape(-1) for g in loose])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if not fold:
            flat.div_(world)
        off = 0
        for g in loose:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n += 1
    if fold:
        optimizer.grad_scale = 1.0 / world
    return n
