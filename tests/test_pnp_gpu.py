"""Graph cut at the bottleneck on the GPU: ResNet_latefusion.pnp_forward_front / pnp_forward_rear (reference
models.py:669-707) against goldens of the REAL reference (tests/golden/pnp_*.npz, oracle/gen_golden.py::run_pnp), the
gradient of the loss w.r.t. the bottleneck feature (what a PnP-Depth loop differentiates), and eval-mode backward
(BatchNorm on running statistics) of the un-cut model against the oracle.

Tolerances: forward 1e-3 relative / loss 1e-4 (north_star) in the fp32 parity mode.  Gradients at 64x96 are taken through
ReLU masks and the sign() of the L1 loss on 2x3-pixel bottleneck maps, see the docstring of tests/test_model_gpu.py for
why they are compared by rel-L2 + cosine with the bounds measured there."""
import os

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from radar_depth_b200 import _lib
from radar_depth_b200.evaluation.criteria_new import MaskedL1Loss
from test_model_gpu import _build, _check_grads, _inputs, _rel      # tests/ is on sys.path (pytest prepend mode)

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cos(a, b):
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize("name,training", [("pnp_train_b2_64x96", True), ("pnp_eval_b2_64x96", False)])
def test_front_and_rear_match_reference_golden_fp32_mode(name, training):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    m, _ = _build(4, (64, 96), "fp32", training=training)
    inputs, target = _inputs(2, 64, 96, 4)
    with torch.no_grad():
        feat = m.pnp_forward_front(inputs.cuda())
    assert feat.shape == (2, 256, 2, 3) and feat.dtype == torch.float32
    r_feat = _rel(feat, torch.from_numpy(g["feature"]))
    f = torch.from_numpy(g["feature"]).cuda().requires_grad_(True)
    pred = m.pnp_forward_rear(f)
    loss = MaskedL1Loss()(pred, target.cuda())
    loss.backward()
    r_pred = _rel(pred, torch.from_numpy(g["pred"]))
    dref = torch.from_numpy(g["dfeature"])
    r_grad, c_grad = _rel(f.grad, dref), _cos(f.grad, dref)
    print(f"[pnp {name}] feature rel {r_feat:.3e}  pred rel {r_pred:.3e}  loss {float(loss):.6f} vs {float(g['loss']):.6f}  "
          f"dfeature rel {r_grad:.3e} cos {c_grad:.6f}")
    assert r_feat < 1e-3
    assert r_pred < 1e-3
    assert abs(float(loss) - float(g["loss"])) <= 1e-4 * max(1.0, abs(float(g["loss"])))
    assert f.grad.shape == f.shape
    assert r_grad < 5e-2 and c_grad > 0.999        # measured on B200: 1.1e-2 / 0.99994 (train), 2.1e-3 / 0.999998 (eval)
    # no parameter gradient is produced (or disturbed) on the cut path
    assert all(p.grad is None or float(p.grad.abs().max()) == 0.0 for p in m.parameters())


@pytest.mark.parametrize("precision,training,tol", [("fp32", True, 1e-3), ("fp32", False, 1e-6), ("bf16", True, 0.25), ("bf16", False, 1e-6)])
def test_rear_of_front_is_the_uncut_forward(precision, training, tol):
    """rear(front(x)) == forward(x) on the same engine.  Measured on B200: bit-identical in eval mode (both precisions);
    in training mode 8e-5 (fp32) / 4e-2 (bf16: the imported feature is rounded to bf16 once more; bound = the bf16-mode
    bound of smoke(), the level at which the reference's own autocast run differs from its fp32 run) -- the batch
    statistics are summed with atomics whose order varies from launch to launch (fp32 shared-memory atomics per CTA in
    the conv epilogue, rd_conv_fprop.cuh:429-435, then fp64 L2 atomics across CTAs), and the variance E[z^2] - mean^2 of
    the 12-sample BatchNorm channels of a 64x96 image amplifies the last-bit differences; the bar is north_star's 1e-3."""
    m, _ = _build(4, (64, 96), precision, training=training)
    x = _inputs(2, 64, 96, 4)[0].cuda()
    with torch.no_grad():
        full = m(x)
        bufs = {k: v.clone() for k, v in m.state_dict().items() if "running" in k or "num_batches" in k}
        m.load_state_dict(O.synth_state_dict(O.latefusion_entries(4)), strict=True)     # undo the running-stat update
        cut = m.pnp_forward_rear(m.pnp_forward_front(x))
    r = _rel(cut, full)
    print(f"[pnp rear(front)] {precision} training={training}: rel {r:.3e}")
    assert r < tol
    if training:                       # each BatchNorm was updated exactly once by front + rear, as by the un-cut forward
        for k, v in m.state_dict().items():
            if k in bufs:
                # (bf16 mode: the decoder sees the feature rounded to bf16 once more, so its batch statistics move a little)
                # fp32 mode: run-to-run noise of the atomically summed statistics, measured > 1e-4 relative on layer4's
                # 12-sample channels; a BatchNorm updated twice (or not at all) would be off by ~0.1 * (batch - running)
                rt, at = (1e-3, 1e-4) if precision == "fp32" else (3e-2, 3e-3)
                assert torch.allclose(v.float(), bufs[k].float(), rtol=rt, atol=at), k


def test_rear_rejects_a_feature_of_the_wrong_shape():
    m, _ = _build(4, (64, 96), "fp32", training=False)
    with torch.no_grad():
        with pytest.raises(_lib.RdError):
            m.pnp_forward_rear(torch.zeros(2, 128, 2, 3, device="cuda"))
        with pytest.raises(_lib.RdError):
            m.pnp_forward_rear(torch.zeros(2, 256, 5, 7, device="cuda"))       # not the 1/32 map of a 64x96 image
        with pytest.raises(_lib.RdError):
            m.pnp_forward_rear(torch.zeros(2, 256, 2, 3))                       # CPU tensor: no fallback


def test_eval_mode_backward_uses_running_statistics_fp32_mode():
    """model.eval() + loss.backward(): BatchNorm is the running-statistics affine map in forward AND backward
    (torch semantics; the oracle runs F.batch_norm(training=False) under autograd)."""
    m, sd = _build(4, (64, 96), "fp32", training=False)
    inputs, target = _inputs(2, 64, 96, 4)
    ref = O.train_step(sd, inputs, target, "latefusion", training=False, dtype=torch.float64)
    pred = m(inputs.cuda())
    loss = MaskedL1Loss()(pred, target.cuda())
    loss.backward()
    assert _rel(pred, ref["pred"]) < 1e-3
    assert abs(float(loss) - float(ref["loss"])) <= 1e-4 * max(1.0, abs(float(ref["loss"])))
    _check_grads(m, ref, rel_tol=0.25, cos_tol=0.98)
    # the batch counts are back: a training step right after gives the training-mode gradients
    m.train()
    m.zero_grad(set_to_none=True)
    ref_t = O.train_step(sd, inputs, target, "latefusion", training=True, dtype=torch.float64)
    loss_t = MaskedL1Loss()(m(inputs.cuda()), target.cuda())
    loss_t.backward()
    assert abs(float(loss_t) - float(ref_t["loss"])) <= 1e-4 * max(1.0, abs(float(ref_t["loss"])))
    _check_grads(m, ref_t, rel_tol=0.25, cos_tol=0.98)
