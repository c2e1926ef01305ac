"""Imports the reference's UNMODIFIED main.py (only where /root/reference exists) with the binding of INTEGRATION.md
section 2 applied: the four hot-path names main.py imports at lines 16-31 (ResNet_latefusion, ResNet_multistage,
MaskedL1Loss, SmoothnessLoss) are re-pointed at radar_depth_b200's classes; everything else in main.py -- create_model,
the optimizer construction, train(), validate(), checkpointing through utils.save_checkpoint -- runs as written.

TEST INFRASTRUCTURE.  main.py cannot be imported as is in this image: it parses sys.argv at import (main.py:35) and pulls
in tensorboardX, matplotlib, h5py and the nuScenes dataset stack, none of which are installed; they are stubbed here
(SURVEY.md 8c lists the same stubs).  No reference source is copied or edited.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF = os.environ.get("RADAR_DEPTH_REFERENCE", "/root/reference")


def reference_present() -> bool:
    return os.path.isfile(os.path.join(REF, "main.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Writer:
    """tensorboardX.SummaryWriter stand-in: records nothing."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return lambda *a, **k: None


def import_main(argv):
    """Returns the reference's ``main`` module with the B200 binding in place.  ``argv``: the command line main.py parses."""
    if not reference_present():
        raise FileNotFoundError(REF)
    import torchvision.models as tvm
    if "tensorboardX" not in sys.modules:
        _stub("tensorboardX", SummaryWriter=_Writer)
    if "matplotlib" not in sys.modules:
        plt = _stub("matplotlib.pyplot", cm=types.SimpleNamespace(viridis=lambda x: x))
        _stub("matplotlib", pyplot=plt)
    if "h5py" not in sys.modules:
        _stub("h5py", File=None)
    if "attrdict" not in sys.modules:
        class AttrDict(dict):
            __getattr__ = dict.__getitem__
        _stub("attrdict", AttrDict=AttrDict)
    # the dataset stack needs h5py / nuscenes-devkit / scipy.misc.imresize: main.py only needs the class name at import
    _stub("dataset.nuscenes_dataset_torch_new", nuscenes_dataset_torch=object)
    orig = tvm.resnet18
    if not getattr(orig, "_rd_nodl", False):               # no network: never download ImageNet weights
        def resnet18(pretrained=False, **k):
            return orig(weights=None)
        resnet18._rd_nodl = True
        tvm.__dict__["resnet18"] = resnet18
        tvm.resnet18 = resnet18
    if REF not in sys.path:
        sys.path.insert(0, REF)
    old_argv = sys.argv
    sys.argv = ["main.py"] + list(argv)
    try:
        for name in ("main", "utils"):
            sys.modules.pop(name, None)
        main = importlib.import_module("main")
    finally:
        sys.argv = old_argv
    # ---- the binding of INTEGRATION.md section 2
    from radar_depth_b200.model.models import ResNet_latefusion
    from radar_depth_b200.model.multistage_model import ResNet_multistage
    from radar_depth_b200.evaluation.criteria_new import MaskedL1Loss, SmoothnessLoss
    main.ResNet_latefusion = ResNet_latefusion
    main.ResNet_multistage = ResNet_multistage
    main.MaskedL1Loss = MaskedL1Loss
    main.SmoothnessLoss = SmoothnessLoss
    return main
