"""Drop-in boundary checks that need no GPU: state_dict layout vs the reference's (pinned through the oracle's
entry table, which tests/test_oracle_golden.py checks against the live reference), constructor error conventions
(SURVEY.md 8b), and that the C-ABI library loads and exports every symbol include/radar_depth_b200.h declares."""
import ctypes
import os
import re

import pytest
import torch

from oracle import torch_oracle as O
from radar_depth_b200 import _lib
from radar_depth_b200.model.models import ResNet_latefusion, choose_decoder, Decoder

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("cin", [4, 5])
def test_state_dict_keys_shapes_order_match_reference(cin):
    m = ResNet_latefusion(18, "upproj", (64, 96), cin, pretrained=False)
    sd = m.state_dict()
    ent = O.latefusion_entries(cin)
    assert list(sd.keys()) == list(ent.keys())
    assert len(sd) == 325
    for k, (shape, kind) in ent.items():
        assert tuple(sd[k].shape) == tuple(shape), k
        assert sd[k].dtype == (torch.int64 if kind == "bn_count" else torch.float32), k
    m.load_state_dict(O.synth_state_dict(ent), strict=True)


def test_constructor_error_conventions():
    with pytest.raises(RuntimeError):
        ResNet_latefusion(19, "upproj", (64, 96), 4, pretrained=False)       # models.py:522-523
    with pytest.raises(AssertionError):
        ResNet_latefusion(18, "upproj", (64, 96), 3, pretrained=False)       # models.py:535
    with pytest.raises(AssertionError):
        choose_decoder("nonsense", 256)                                       # models.py:230
    with pytest.raises(NotImplementedError):
        ResNet_latefusion(50, "upproj", (64, 96), 4, pretrained=False)       # outside the hot path
    assert "upproj" in Decoder.names                                          # utils.py:11,23


def test_forward_on_cpu_fails_loudly():
    m = ResNet_latefusion(18, "upproj", (64, 96), 4, pretrained=False)
    with pytest.raises(_lib.RdError):
        m(torch.zeros(1, 4, 64, 96))


def test_init_distributions_follow_reference():
    torch.manual_seed(0)
    m = ResNet_latefusion(18, "upproj", (64, 96), 4, pretrained=False)
    w = m.conv2.weight                                    # weights_init: N(0, sqrt(2/(k*k*Cout)))  models.py:30-36
    assert abs(float(w.std()) - (2.0 / (1 * 1 * 256)) ** 0.5) < 5e-3
    w = m.decoder.layer1.upper_branch.conv1.weight
    assert abs(float(w.std()) - (2.0 / (25 * 128)) ** 0.5) < 1e-3
    w = m.layer2_depth[0].conv1.weight                    # kaiming fan_out relu
    assert abs(float(w.std()) - (2.0 / (9 * 32)) ** 0.5) < 1e-2
    assert float(m.bn_fusion.weight.min()) == 1.0 and float(m.bn1_depth.bias.abs().max()) == 0.0


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "radar_depth_b200.h")).read()
    declared = set(re.findall(r"\b(rd_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(_lib.EXPORTS) == declared
    assert lib.rd_version() >= 1
    assert lib.rd_sizeof(0) == ctypes.sizeof(_lib.ConvParams) and lib.rd_sizeof(1) == ctypes.sizeof(_lib.WgradParams)
