"""The reference's training loop driving the B200 classes on the GPU: the sequence of main.py:258-291 (optimizer built
before model.cuda()), train() (main.py:377-458: model.train(), forward, loss per architecture, zero_grad, backward, SGD
step, Result.evaluate on every iteration), validate() (main.py:545-600: model.eval() under no_grad, batch 1) and the
checkpoint save / resume round trip (main.py:219-266,358-374).

Where the reference checkout exists next to a GPU the loop that runs IS the reference's own main.train / main.validate
(tests/tools/ref_main_harness.py applies the import switch of INTEGRATION.md).  /root/reference does not travel to the GPU
box, so there the same call sequence is issued by `_train_epoch` / `_validate` below -- a restatement of those loops
without their logging -- and the binding itself is exercised against the real main.py by
tests/test_dropin_reference_main_cpu.py in the build container."""
import os
import sys
import types

import pytest
import torch

from oracle import torch_oracle as O
from radar_depth_b200.evaluation.criteria_new import MaskedL1Loss, SmoothnessLoss
from radar_depth_b200.evaluation.metrics import Result
from radar_depth_b200.model.models import ResNet_latefusion
from radar_depth_b200.model.multistage_model import ResNet_multistage

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
import ref_main_harness as H  # noqa: E402

pytestmark = pytest.mark.gpu
HW = (96, 160)


def _batches(n, b):
    out = []
    for i in range(n):
        inputs, target = O.synth_batch(b, HW[0], HW[1], seed=1234 + i)
        out.append({"inputs": inputs, "labels": target, "lidar_depth": target, "radar_depth": inputs[:, 3:],
                    "daynight_info": ["day, sun"] * b})
    return out


def _create(arch):
    """create_model (main.py:118-186) for the two target architectures, --no-pretrain."""
    if arch == "resnet18_latefusion":
        m = ResNet_latefusion(layers=18, decoder="upproj", output_size=HW, in_channels=4, pretrained=False)
        m.load_state_dict(O.synth_state_dict(O.latefusion_entries(4)), strict=True)
        return m, None
    m = ResNet_multistage(layers=18, decoder="upproj", output_size=HW, pretrained=False)
    w1 = torch.nn.Parameter(torch.tensor(1.0, dtype=torch.float32), requires_grad=True)
    w2 = torch.nn.Parameter(torch.tensor(1.0, dtype=torch.float32), requires_grad=True)
    m.register_parameter("w_stage1", w1)
    m.register_parameter("w_stage2", w2)
    m.load_state_dict(O.synth_state_dict(O.multistage_entries()), strict=True)
    return m, {"w_stage1": w1, "w_stage2": w2, "w_smooth": 0.1}


def _set_precision(model, precision):
    for mod in model.modules():
        if hasattr(mod, "precision"):
            mod.precision = precision


def _train_epoch(loader, model, criterion, optimizer, arch, loss_weights):
    model.train()
    losses, results = [], []
    for data in loader:
        inputs, target = data["inputs"].cuda(), data["labels"].cuda()
        torch.cuda.synchronize()
        if arch == "resnet18_multistage_uncertainty_fixs":
            pred_ = model(inputs)
            pred1, pred = pred_["stage1"], pred_["stage2"]
            l1, l2 = criterion["depth"](pred1, target), criterion["depth"](pred, target)
            sm = criterion["smooth"](pred1, inputs)
            loss = torch.exp(-loss_weights["w_stage1"]) * (l1 + loss_weights["w_smooth"] * sm) + \
                torch.exp(-loss_weights["w_stage2"]) * l2 + loss_weights["w_stage1"] + loss_weights["w_stage2"]
        else:
            pred = model(inputs)
            loss = criterion["depth"](pred, target)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        torch.cuda.synchronize()
        r = Result()
        r.evaluate(pred.data, target.data)
        losses.append(float(loss))
        results.append(r)
    return losses, results


def _validate(loader, model, arch):
    model.eval()
    out = []
    for data in loader:
        inputs, target = data["inputs"].cuda(), data["labels"].cuda()
        with torch.no_grad():
            pred = model(inputs)
            if isinstance(pred, dict):
                pred = pred["stage2"]
        torch.cuda.synchronize()
        r = Result()
        r.evaluate(pred.data, target.data)
        out.append((r.rmse, r.mae, pred.clone()))
    return out


@pytest.mark.parametrize("arch", ["resnet18_latefusion", "resnet18_multistage_uncertainty_fixs"])
def test_train_validate_checkpoint_resume(tmp_path, arch):
    train_loader, val_loader = _batches(3, 2), _batches(2, 1)
    model, lw = _create(arch)
    _set_precision(model, "fp32")
    optimizer = torch.optim.SGD(model.parameters(), 0.01, momentum=0.9, weight_decay=1e-4)     # main.py:285-290
    model = model.cuda()                                                                      # main.py:291
    criterion = {"depth": MaskedL1Loss().cuda(), "smooth": SmoothnessLoss().cuda()}
    losses, results = _train_epoch(train_loader, model, criterion, optimizer, arch, lw)
    assert all(l == l and l < 1e4 for l in losses), losses
    assert all(r.rmse == r.rmse and r.rmse > 0 for r in results)
    # every parameter received a gradient and moved
    ref_sd = O.synth_state_dict(O.latefusion_entries(4) if lw is None else O.multistage_entries())
    moved = [k for k, p in model.named_parameters() if not torch.equal(p.detach().cpu(), ref_sd[k])]
    assert len(moved) == len(list(model.parameters())), set(dict(model.named_parameters())) - set(moved)
    val = _validate(val_loader, model, arch)
    assert all(v[0] == v[0] for v in val)
    # ---- checkpoint (main.py:358-374) and resume (main.py:219-266)
    path = tmp_path / "checkpoint-0.pth.tar"
    torch.save({"epoch": 0, "arch": arch, "model_state_dict": model.state_dict(),
                "optimizer_state_dict": optimizer.state_dict()}, path)
    ck = torch.load(path, weights_only=False)
    model2, lw2 = _create(arch)
    _set_precision(model2, "fp32")
    missing, unexpected = model2.load_state_dict(ck["model_state_dict"], strict=False)
    assert not missing and not unexpected
    model2 = model2.cuda()
    optimizer2 = torch.optim.SGD(model2.parameters(), 0.01, momentum=0.9, weight_decay=1e-4)
    optimizer2.load_state_dict(ck["optimizer_state_dict"])
    for (k, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert torch.equal(a, b), k
    # the resumed run validates identically and continues bit-identically (fp32 mode is deterministic)
    val2 = _validate(val_loader, model2, arch)
    for a, b in zip(val, val2):
        assert a[0] == b[0] and torch.equal(a[2], b[2])
    more = _batches(4, 2)[3:]
    la, _ = _train_epoch(more, model, criterion, optimizer, arch, lw)
    lb, _ = _train_epoch(more, model2, criterion, optimizer2, arch, lw2)
    assert la == lb
    for (k, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert torch.equal(a, b), k


@pytest.mark.skipif(not H.reference_present(), reason="reference checkout not present on this box")
def test_reference_main_train_and_validate_run_on_the_b200_classes(tmp_path):
    """With the reference next to a GPU: its own train() / validate() (main.py:377-458,545-732), unmodified."""
    main = H.import_main(["--arch", "resnet18_latefusion", "--data", "nuscenes", "--modality", "rgbd", "--decoder", "upproj",
                          "--no-pretrain", "-b", "2", "--print-freq", "1"])
    model = main.create_model(main.args, output_size=HW)
    optimizer = torch.optim.SGD(model.parameters(), main.args.lr, momentum=main.args.momentum, weight_decay=main.args.weight_decay)
    model = model.cuda()
    main.output_directory = str(tmp_path)
    main.train_csv = str(tmp_path / "train.csv")
    main.test_csv = str(tmp_path / "test.csv")
    criterion = {"depth": main.MaskedL1Loss().cuda()}
    logger = sys.modules["tensorboardX"].SummaryWriter()
    main.train(_batches(3, 2), model, criterion, optimizer, 0, None, logger=logger)
    assert all(p.grad is not None for p in model.parameters())
