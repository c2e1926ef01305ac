"""The reference's own main.py (unmodified, /root/reference) driving the B200 classes through the binding of
INTEGRATION.md section 2: create_model for both target architectures, the optimizer built BEFORE model.cuda() like
main.py:285-291, and the checkpoint round trip through the reference's utils.save_checkpoint / load_state_dict(strict=False)
(main.py:219-266,358-374).  Runs where the reference is present (the build container); the training loop itself needs
a GPU and is covered by tests/test_train_loop_gpu.py."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
import ref_main_harness as H  # noqa: E402

pytestmark = pytest.mark.skipif(not H.reference_present(), reason="reference checkout not present")

ARGV = ["--data", "nuscenes", "--modality", "rgbd", "--decoder", "upproj", "--no-pretrain", "-b", "2", "--sparsifier", "radar"]


@pytest.mark.parametrize("arch,nkeys", [("resnet18_latefusion", 325), ("resnet18_multistage_uncertainty_fixs", 652)])
def test_reference_main_builds_and_checkpoints_the_b200_model(tmp_path, arch, nkeys):
    main = H.import_main(["--arch", arch] + ARGV)
    from radar_depth_b200.model.models import ResNet_latefusion
    from radar_depth_b200.model.multistage_model import ResNet_multistage
    assert main.args.arch == arch and main.args.pretrained is False
    made = main.create_model(main.args, output_size=(64, 96))
    if arch.endswith("fixs"):
        model, loss_weights = made
        assert isinstance(model, ResNet_multistage)
        assert loss_weights["w_smooth"] == 0.1 and loss_weights["w_stage1"] is model.w_stage1
    else:
        model = made
        assert isinstance(model, ResNet_latefusion)
    assert len(model.state_dict()) == nkeys
    # main.py:285-290: the optimizer is built on the CPU parameters
    opt = torch.optim.SGD(model.parameters(), main.args.lr, momentum=main.args.momentum, weight_decay=main.args.weight_decay)
    crit = main.MaskedL1Loss()
    assert type(crit).__module__.startswith("radar_depth_b200")
    # main.py:358-374 -> :219-266: save through the reference's own utils, resume with strict=False
    import utils as ref_utils
    state = {"args": main.args, "epoch": 0, "arch": main.args.arch, "model_state_dict": model.state_dict(),
             "optimizer_state_dict": opt.state_dict()}
    ref_utils.save_checkpoint(state, True, 0, str(tmp_path))
    assert os.path.isfile(tmp_path / "model_best.pth.tar")
    ck = torch.load(tmp_path / "checkpoint-0.pth.tar", weights_only=False)
    made2 = main.create_model(ck["args"], output_size=(64, 96))
    model2 = made2[0] if isinstance(made2, tuple) else made2
    missing, unexpected = model2.load_state_dict(ck["model_state_dict"], strict=False)
    assert not missing and not unexpected
    for (k, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert torch.equal(a, b), k
    opt2 = torch.optim.SGD(model2.parameters(), ck["args"].lr, momentum=ck["args"].momentum, weight_decay=ck["args"].weight_decay)
    opt2.load_state_dict(ck["optimizer_state_dict"])
    # utils.adjust_learning_rate (utils.py:85-89) drives param_groups as usual
    ref_utils.adjust_learning_rate(opt2, 5, ck["args"].lr)
    assert abs(opt2.param_groups[0]["lr"] - 0.001) < 1e-12
    # there is no CPU fallback: the product path fails loudly without a CUDA device
    if not torch.cuda.is_available():
        from radar_depth_b200 import _lib
        with pytest.raises(_lib.RdError):
            model(torch.zeros(1, 4, 64, 96))
