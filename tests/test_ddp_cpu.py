"""Multi-GPU host logic on CPU (gloo, world_size 2): the gradient exchange is ONE all-reduce per engine arena plus
one for parameters living outside an arena (w_stage1/w_stage2), averaged over ranks; weights/buffers are broadcast
from rank 0.  The device kernels are not involved: the arenas are stubbed with CPU tensors."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _ArenaStub:
    """What radar_depth_b200.engine.LatefusionEngine exposes to ddp.py / optim.py."""

    def __init__(self, module, n):
        self.module = module
        self.flat = torch.zeros(n)
        self.gflat = torch.zeros(n)


class _Stage(torch.nn.Module):
    def __init__(self, n):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(n))
        self.register_buffer("running", torch.zeros(3))
        self._engine = _ArenaStub(self, n)
        self.w.data = self._engine.flat[:n]


class _Model(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.stage1, self.stage2 = _Stage(10), _Stage(6)
        self.w_stage1 = torch.nn.Parameter(torch.tensor(1.0))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from radar_depth_b200 import ddp
    torch.manual_seed(100 + rank)
    m = _Model()
    with torch.no_grad():
        m.stage1._engine.flat.normal_()
        m.stage2._engine.flat.normal_()
        m.stage1.running.fill_(float(rank + 1))
        m.w_stage1.fill_(float(rank + 5))
    ddp.broadcast_parameters(m)
    w_after = torch.cat([m.stage1.w.detach(), m.stage2.w.detach(), m.w_stage1.detach().reshape(1), m.stage1.running])
    m.stage1._engine.gflat.fill_(float(rank + 1))
    m.stage2._engine.gflat.copy_(torch.arange(6.0) * (rank + 1))
    m.stage1.w.grad = m.stage1._engine.gflat[:10]
    m.stage2.w.grad = m.stage2._engine.gflat[:6]
    m.w_stage1.grad = torch.tensor(float(10 * (rank + 1)))
    n = ddp.allreduce_gradients(m)
    q.put((rank, n, w_after, m.stage1.w.grad.clone(), m.stage2.w.grad.clone(), m.w_stage1.grad.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_and_broadcast_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, n0, w0, g1a, g2a, gsa), (_, n1, w1, g1b, g2b, gsb) = res
    assert n0 == n1 == 3                                   # two arenas + one bucket of loose parameters
    assert torch.equal(w0, w1)                              # rank 0's weights and buffers everywhere
    assert float(w0[16]) == 5.0 and float(w0[17]) == 1.0
    for ga, gb in ((g1a, g1b), (g2a, g2b), (gsa, gsb)):
        assert torch.equal(ga, gb)
    assert torch.allclose(g1a, torch.full((10,), 1.5))      # mean of 1 and 2
    assert torch.allclose(g2a, torch.arange(6.0) * 1.5)
    assert float(gsa) == 15.0


def _worker_overlap(rank, world, port, q):
    """The bucketed path: the engine calls the hook after each of its three backward segments (stubbed here: the segments
    are just the moments the ranges of the arena become final); allreduce_gradients then only waits, and with the optimizer
    passed the arenas keep the SUM while the optimizer's grad_scale becomes 1/world."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from radar_depth_b200 import ddp, optim
    m = _Model()
    for st, n in ((m.stage1, 10), (m.stage2, 6)):
        st._get_engine = lambda: None                       # marks an engine-backed network for enable_overlap
        st._engine.grad_ranges = [[(n - 3, n)], [(2, n - 3)], [(0, 2)]]
    ddp.enable_overlap(m)
    m.stage1._engine.gflat.copy_(torch.arange(10.0) + 100 * rank)
    m.stage2._engine.gflat.copy_(torch.arange(6.0) * (rank + 1))
    m.stage1.w.grad = m.stage1._engine.gflat[:10]
    m.stage2.w.grad = m.stage2._engine.gflat[:6]
    m.w_stage1.grad = torch.tensor(float(10 * (rank + 1)))
    for st in (m.stage2, m.stage1):                         # backward order: stage 2 first
        for k in range(3):
            st._rd_grad_hook(st._engine, k)
    opt = optim.FusedSGD(m)
    n = ddp.allreduce_gradients(m, opt)
    first = (n, opt.grad_scale, m.stage1._engine.gflat.clone(), m.stage2._engine.gflat.clone(), m.w_stage1.grad.clone())
    # a backward that did not go through the hook (gradient accumulation): whole-arena fallback, averaged in place
    m.stage1._engine.gflat.fill_(float(rank + 1))
    m.stage2._engine.gflat.fill_(float(rank + 1))
    n2 = ddp.allreduce_gradients(m)
    q.put((rank, first, n2, m.stage1._engine.gflat.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_overlap_path_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_overlap, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, (n, scale, g1, g2, gs), n2, g1b in res:
        assert n == 3 + 3 + 1                               # three buckets per arena + the loose parameters
        assert scale == 0.5
        assert torch.equal(g1, 2 * torch.arange(10.0) + 100)        # SUM over the two ranks, averaging left to the optimizer
        assert torch.equal(g2, torch.arange(6.0) * 3)
        assert float(gs) == 30.0
        assert n2 == 3 and torch.allclose(g1b, torch.full((10,), 1.5))


def test_single_process_is_a_no_op():
    from radar_depth_b200 import ddp
    m = _Model()
    assert ddp.allreduce_gradients(m) == 0
    ddp.broadcast_parameters(m)
