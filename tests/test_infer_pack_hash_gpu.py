"""The inference program re-packs the bf16 weight tiles only when the content hash of the fp32 parameter arena changed
(rd_weights_hash / rd_pack_weights_if).  Every way of changing the weights must be seen -- optimizer-style in-place
updates, load_state_dict, writes through .data (which no host-side version counter notices) -- and a program that re-packs
unconditionally in between must not leave a stale hash behind."""
import pytest
import torch

from radar_depth_b200.model.models import ResNet_latefusion

pytestmark = pytest.mark.gpu


def _model():
    torch.manual_seed(3)
    m = ResNet_latefusion(18, "upproj", (64, 96), 4, pretrained=False).cuda().eval()
    m.precision = "bf16"
    return m


def _fresh_output(sd, x):
    """The same weights in a new model (its first forward packs unconditionally)."""
    m = _model()
    m.load_state_dict(sd)
    with torch.no_grad():
        return m(x).clone()


def test_repack_follows_every_kind_of_weight_change():
    m = _model()
    x = torch.rand(1, 4, 64, 96, device="cuda")
    with torch.no_grad():
        outs = [m(x).clone() for _ in range(4)]                 # eager, capture, replays
    assert all(torch.equal(outs[0], o) for o in outs[1:])
    eng = m._engine
    assert eng.fwd_infer[0].name == "weights_hash" and eng.fwd_infer[1].name == "pack_weights_if"
    assert int(eng._wdirty.item()) == 0                         # the last forward did NOT re-pack

    def check(tag):
        with torch.no_grad():
            got = m(x).clone()
        ref = _fresh_output({k: v.clone() for k, v in m.state_dict().items()}, x)
        assert torch.equal(got, ref), tag
        return got

    with torch.no_grad():
        m.layer2[0].conv1.weight.mul_(1.5)                      # in-place (what an optimizer does)
    a = check("in-place update")
    assert not torch.equal(a, outs[0])
    m.conv3.weight.data.mul_(-1.0)                              # through .data: invisible to tensor version counters
    b = check(".data update")
    assert not torch.equal(a, b)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    sd["decoder.layer1.upper_branch.conv1.weight"] *= 0.5
    m.load_state_dict(sd)
    check("load_state_dict")
    # a single changed element in the middle of the arena
    with torch.no_grad():
        m.layer3[1].conv2.weight.view(-1)[12345] += 1.0
    check("one element")
    with torch.no_grad():
        again = m(x).clone()
    assert int(eng._wdirty.item()) == 0 and torch.equal(again, check("unchanged"))


def test_unconditional_repack_in_between_does_not_leave_a_stale_hash():
    m = _model()
    x = torch.rand(1, 4, 64, 96, device="cuda")
    w0 = {k: v.clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        for _ in range(3):
            y0 = m(x).clone()                                   # hash state = hash(W0), packed tiles = pack(W0)
    with torch.no_grad():
        m.layer1[0].conv1.weight.mul_(2.0)                      # W1
    m(x).sum().backward()                                       # eval program WITH grad: packs W1 unconditionally
    m.load_state_dict(w0)                                       # back to W0: the arena hashes to the stored value again
    with torch.no_grad():
        y = m(x).clone()
    assert torch.equal(y, y0)
