"""The other constructors on the same kernels: ResNet (reference models.py:233-303, one encoder over all
input channels) and the UpConv / DeConv decoders (models.py:135-176) under either encoder, on the GPU against the CPU oracle
and the goldens of the REAL reference (tests/golden/{resnet,latefusion}_*_b2_64x96.npz, oracle/gen_golden.py::run_variant).
Tolerances as in tests/test_model_gpu.py (fp32 parity mode: outputs 1e-3, losses 1e-4)."""
import os

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from radar_depth_b200.evaluation.criteria_new import MaskedL1Loss
from radar_depth_b200.model.models import ResNet, ResNet_latefusion
from test_model_gpu import _check_grads, _rel      # tests/ is on sys.path (pytest prepend mode)

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VARIANTS = [("resnet_rgbd_upproj_b2_64x96", "resnet", 4, "upproj"), ("resnet_rgb_deconv3_b2_64x96", "resnet", 3, "deconv3"),
            ("resnet_rgbd_upconv_b2_64x96", "resnet", 4, "upconv"), ("latefusion_deconv2_b2_64x96", "latefusion", 4, "deconv2"),
            ("latefusion_upconv_b2_64x96", "latefusion", 4, "upconv")]


def _make(arch, cin, decoder, hw, precision):
    cls = ResNet if arch == "resnet" else ResNet_latefusion
    m = cls(18, decoder, hw, cin, pretrained=False)
    sd = O.synth_state_dict(O.resnet_entries(cin, decoder) if arch == "resnet" else O.latefusion_entries(cin, decoder))
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train()
    m.precision = precision
    return m, sd


@pytest.mark.parametrize("name,arch,cin,decoder", VARIANTS)
def test_variant_train_step_parity_fp32_mode(name, arch, cin, decoder):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m, sd = _make(arch, cin, decoder, (64, 96), "fp32")
    inputs, target = O.synth_batch(2, 64, 96)
    inputs = inputs[:, :cin].contiguous()
    ref = O.train_step(sd, inputs, target, arch, dtype=torch.float64, decoder=decoder)
    pred = m(inputs.cuda())
    loss = MaskedL1Loss()(pred, target.cuda())
    loss.backward()
    assert _rel(pred, ref["pred"]) < 1e-3
    assert _rel(pred[..., ::2, ::2], torch.from_numpy(g["pred"])) < 1e-3               # the real reference's output
    assert abs(float(loss) - float(ref["loss"])) <= 1e-4 * abs(float(ref["loss"]))
    assert abs(float(loss) - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    # tiny maps (see tests/test_model_gpu.py): measured worst cosine 0.979 (bn1_depth.bias, 16 values behind two ReLU masks)
    _check_grads(m, ref, rel_tol=0.25, cos_tol=0.97)
    for k, v in ref["new_buffers"].items():
        got = m.state_dict()[k]
        if k.endswith("num_batches_tracked"):
            assert int(got) == int(v), k
        else:
            np.testing.assert_allclose(got.cpu().numpy(), v.float().numpy(), rtol=2e-4, atol=1e-5, err_msg=k)


@pytest.mark.parametrize("arch,cin,decoder,hw", [("resnet", 4, "upconv", (90, 160)), ("latefusion", 4, "deconv3", (90, 160)),
                                                ("resnet", 1, "deconv2", (64, 96))])
def test_variant_odd_sizes_and_eval_fp32_mode(arch, cin, decoder, hw):
    """Odd intermediate sizes (90x160 -> 3x5 at 1/32) and the eval path (running statistics, no_grad)."""
    m, sd = _make(arch, cin, decoder, hw, "fp32")
    inputs, target = O.synth_batch(2, hw[0], hw[1])
    inputs = inputs[:, -cin:].contiguous() if cin == 1 else inputs[:, :cin].contiguous()
    ref = O.train_step(sd, inputs, target, arch, dtype=torch.float64, decoder=decoder)
    pred = m(inputs.cuda())
    loss = MaskedL1Loss()(pred, target.cuda())
    loss.backward()
    assert _rel(pred, ref["pred"]) < 1e-3
    assert abs(float(loss) - float(ref["loss"])) <= 1e-4 * abs(float(ref["loss"]))
    _check_grads(m, ref, rel_tol=0.3, cos_tol=0.97)
    m.load_state_dict(sd, strict=True)
    m.eval()
    with torch.no_grad():
        pe = m(inputs.cuda())
    fwd = O.resnet_forward if arch == "resnet" else O.latefusion_forward
    with torch.no_grad():
        re = fwd({k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()}, inputs.double(), hw, False, None,
                 decoder=decoder)
    assert _rel(pe, re) < 1e-3


def test_variant_bf16_mode_within_autocast_level():
    m, sd = _make("resnet", 4, "upconv", (352, 1216), "bf16")
    inputs, target = O.synth_batch(2, 352, 1216)
    ref = O.train_step(sd, inputs, target, "resnet", dtype=torch.float32, decoder="upconv")
    pred = m(inputs.cuda())
    loss = MaskedL1Loss()(pred, target.cuda())
    loss.backward()
    assert _rel(pred, ref["pred"]) < 0.25
    assert abs(float(loss) - float(ref["loss"])) <= 2e-2 * abs(float(ref["loss"]))
    assert _rel(m.conv3.weight.grad, ref["grads"]["conv3.weight"]) < 0.1
