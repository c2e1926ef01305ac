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


def test_multistage_state_dict_and_error_conventions(tmp_path, monkeypatch):
    from radar_depth_b200.model import multistage_model as mm
    with pytest.raises(RuntimeError):
        mm.ResNet_multistage(20, "upproj", (64, 96), pretrained=False)          # multistage_model.py:24-25
    monkeypatch.setattr(mm.cfg, "PROJECT_ROOT", str(tmp_path))
    with pytest.raises(ValueError):
        mm.ResNet_multistage(18, "upproj", (64, 96), pretrained=True)           # multistage_model.py:37-39
    m = mm.ResNet_multistage(18, "upproj", (64, 96), pretrained=False)
    m.register_parameter("w_stage1", torch.nn.Parameter(torch.tensor(1.0)))     # main.py:166-172
    m.register_parameter("w_stage2", torch.nn.Parameter(torch.tensor(1.0)))
    ent = O.multistage_entries()
    assert list(m.state_dict().keys()) == list(ent.keys()) and len(ent) == 652
    for k, (shape, _) in ent.items():
        assert tuple(m.state_dict()[k].shape) == tuple(shape), k
    # "init from latefusion weights" (BASELINE.json configs[3]): stage1 == checkpoint, stage2 differs only in conv1_depth
    sd = O.synth_state_dict(O.latefusion_entries(4))
    (tmp_path / "pretrained").mkdir()
    torch.save({"model_state_dict": sd}, str(tmp_path / "pretrained" / "resnet18_latefusion.pth.tar"))
    m2 = mm.ResNet_multistage(18, "upproj", (64, 96), pretrained=True)
    for k, v in sd.items():
        assert torch.equal(m2.stage1.state_dict()[k], v), k
        if k != "conv1_depth.weight":
            assert torch.equal(m2.stage2.state_dict()[k], v), k
    assert m2.stage2.conv1_depth.weight.shape == (16, 2, 7, 7)
    d = dict(sd)
    out = m2.filter_state_dict(d, m2.stage2.state_dict())
    assert out is d and "conv1_depth.weight" not in d                            # mutates its argument (multistage_model.py:58-59)
