"""Host-side bookkeeping that runs every step, on CPU: the engine's cached parameter / gradient-view lists (adoption
check, .grad binding with torch's accumulation semantics) and FusedSGD's cached engine / loose-parameter discovery (the
rd_sgd kernel is replaced by the same rule in torch; the kernel itself is pinned in tests/test_elementwise_gpu.py)."""
import pytest
import torch

from radar_depth_b200 import _lib, optim
from radar_depth_b200.engine import LatefusionEngine
from radar_depth_b200.model.models import ResNet_latefusion


@pytest.fixture()
def eng():
    m = ResNet_latefusion(18, "upproj", (64, 96), 4, pretrained=False)
    e = LatefusionEngine(m, 4, (64, 96), _lib.RD_F32)
    e.adopt("cpu")
    return e


def test_adoption_is_tracked_through_cached_owner_lists(eng):
    m = eng.module
    assert eng.params_adopted()
    names = [n for n, _ in m.named_parameters()]
    assert [e[2] for e in eng._plist] == [p for _, p in m.named_parameters()]           # same objects, same order
    for (n, (off, shape)), e in zip(eng.offs.items(), eng._plist):
        assert e[3] == off and e[5] == shape and e[2].data_ptr() == eng.flat.data_ptr() + 4 * off, n
    # in-place updates (optimizer steps, load_state_dict) keep the adoption
    with torch.no_grad():
        m.conv3.weight.add_(1.0)
    m.load_state_dict(m.state_dict())
    assert eng.params_adopted()
    # storage moved away from the arena -> not adopted
    keep = m.layer1[0].conv1.weight.data
    m.layer1[0].conv1.weight.data = keep.clone()
    assert not eng.params_adopted()
    m.layer1[0].conv1.weight.data = keep
    assert eng.params_adopted()
    # parameter OBJECT replaced (e.g. load_state_dict(assign=True)) -> not adopted
    old = m.bn1.weight
    m.bn1.weight = torch.nn.Parameter(old.detach().clone())
    assert not eng.params_adopted()
    m.bn1._parameters["weight"] = old
    assert eng.params_adopted()
    # a parameter ADDED inside the module is caught by the periodic full walk (every 128th check)
    m.register_parameter("extra", torch.nn.Parameter(torch.zeros(3)))
    assert not all(eng.params_adopted() for _ in range(130))
    del m._parameters["extra"]
    assert names == [n for n, _ in m.named_parameters()]
    eng.adopt("cpu")                                   # re-adoption rebuilds every cached list
    assert eng.params_adopted() and len(eng._gviews) == len(names)


def test_grad_views_bind_rebind_and_accumulate(eng):
    m = eng.module
    assert not eng.grads_bound()                        # adopt() leaves .grad = None
    eng.bind_grads()
    assert eng.grads_bound()
    for e, v in zip(eng._plist, eng._gviews):
        p, off, n, shape = e[2], e[3], e[4], e[5]
        assert p.grad is v and tuple(v.shape) == shape and v.data_ptr() == eng.gflat.data_ptr() + 4 * off and v.numel() == n
    eng.gflat.fill_(2.0)                                # the views alias the arena
    assert float(m.conv3.weight.grad.sum()) == 2.0 * m.conv3.weight.numel()
    views = [p.grad for p in m.parameters()]
    eng.bind_grads()                                    # idempotent: the same view objects stay
    assert all(a is b for a, b in zip(views, (p.grad for p in m.parameters())))
    # optimizer.zero_grad(set_to_none=False): zeroed in place, still bound -> the next backward accumulates into zeros
    for p in m.parameters():
        p.grad.zero_()
    assert eng.grads_bound() and float(eng.gflat.abs().sum()) == 0.0
    # an equivalent view made by someone else (same arena slot) counts as bound and is left alone
    e0 = eng._plist[0]
    other = eng.gflat[e0[3]:e0[3] + e0[4]].view(e0[5])
    e0[2].grad = other
    assert eng.grads_bound()
    eng.bind_grads()
    assert e0[2].grad is other
    # optimizer.zero_grad() (set_to_none) on ONE parameter -> not bound (the engine then zeroes the arena and re-binds)
    m.bn1.bias.grad = None
    assert not eng.grads_bound()
    eng.bind_grads()
    idx = next(i for i, e in enumerate(eng._plist) if e[2] is m.bn1.bias)
    assert eng.grads_bound() and m.bn1.bias.grad is eng._gviews[idx]
    # a foreign tensor as .grad is replaced by the arena view
    m.bn1.bias.grad = torch.ones_like(m.bn1.bias)
    assert not eng.grads_bound()
    eng.bind_grads()
    assert eng.grads_bound()


class _Stage(torch.nn.Module):
    def __init__(self, n):
        super().__init__()
        self.w = torch.nn.Parameter(torch.randn(n))

        class _E:
            pass
        self._engine = _E()
        self._engine.module = self
        self._engine.flat = torch.zeros(n)
        self._engine.gflat = torch.zeros(n)
        with torch.no_grad():
            self._engine.flat.copy_(self.w)
        self.w.data = self._engine.flat[:n]


class _Two(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.stage1, self.stage2 = _Stage(10), _Stage(6)
        self.w_stage1 = torch.nn.Parameter(torch.tensor(1.0))          # main.py:166-172 registers these on the model
        self.w_stage2 = torch.nn.Parameter(torch.tensor(1.0))


def test_fused_sgd_matches_torch_sgd_with_cached_discovery(monkeypatch):
    calls = []

    def fake_call(name, p, g, mom, n, lr, mo, wd, first, stream):      # rd_sgd's rule (torch.optim.SGD, dampening 0)
        assert name == "rd_sgd" and n == p.numel()
        calls.append(p)
        d = g + wd * p
        if first:
            mom.copy_(d)
        else:
            mom.mul_(mo).add_(d)
        p.sub_(lr * mom)

    monkeypatch.setattr(optim._lib, "call", fake_call)
    monkeypatch.setattr(optim, "ptr", lambda t: t)
    monkeypatch.setattr(optim, "stream_ptr", lambda: 0)
    torch.manual_seed(0)
    m = _Two()
    ref = _Two()
    ref.load_state_dict(m.state_dict())
    opt = optim.FusedSGD(m, lr=0.01, momentum=0.9, weight_decay=1e-4)
    opt_ref = torch.optim.SGD(ref.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    for it in range(3):
        gs = [torch.randn_like(p) for p in m.parameters()]
        opt.zero_grad()
        assert all(p.grad is None for p in m.parameters())
        opt_ref.zero_grad()
        m.stage1._engine.gflat.copy_(gs[2])            # parameters(): w_stage1, w_stage2, stage1.w, stage2.w
        m.stage2._engine.gflat.copy_(gs[3])
        m.w_stage1.grad, m.w_stage2.grad = gs[0].clone(), gs[1].clone()
        for p, g in zip(ref.parameters(), gs):
            p.grad = g.clone()
        opt.step()
        opt_ref.step()
        for (k, a), b in zip(m.named_parameters(), ref.parameters()):
            torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7, msg=k)
    assert len(calls) == 6 and opt._loose_params == [m.w_stage1, m.w_stage2]
    # a replaced engine (precision switch / re-adoption) is picked up at the next step
    mods = opt._eng_modules
    new = _Stage(10)._engine
    new.module = m.stage1
    with torch.no_grad():
        new.flat.copy_(m.stage1.w)
    m.stage1.w.data = new.flat[:10]
    m.stage1._engine = new
    new.gflat.fill_(1.0)
    m.stage2._engine.gflat.zero_()
    opt.zero_grad()
    opt.step()
    assert calls[-2] is new.flat and opt._eng_modules == mods and id(new.flat) in opt._mom
    # learning-rate schedule through param_groups (utils.py:85-89)
    opt.param_groups[0]["lr"] = 0.0
    before = m.stage1.w.detach().clone()
    opt.step()
    assert torch.equal(m.stage1.w.detach(), before)


def test_fused_sgd_before_the_first_forward_fails_loudly():
    m = torch.nn.Linear(2, 2)
    with pytest.raises(_lib.RdError):
        optim.FusedSGD(m).step()
