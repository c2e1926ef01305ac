"""Directional-derivative check of the engine's backward against its own forward (no oracle involved).  Kept in a file
that sorts after every other GPU test: it is a finite-difference check through ReLU / max-pool / |.| kinks and must never
gate the parity suites under ``pytest -x``."""
import pytest
import torch

from radar_depth_b200.evaluation.criteria_new import MaskedL1Loss
from test_model_gpu import _build, _inputs      # tests/ is on sys.path (pytest prepend mode)

pytestmark = pytest.mark.gpu


def test_backward_is_the_derivative_of_forward_fp32_mode():
    """Directional derivatives: (L(w + eps d) - L(w - eps d)) / 2 eps == <grad, d> using only the engine itself
    (the loss is re-evaluated in fp64 from the fp32 prediction so that eps can be small)."""
    m, sd = _build(4, (96, 160), "fp32")
    inputs, target = _inputs(2, 96, 160, 4)
    x, t = inputs.cuda(), target.cuda()
    valid = t > 0

    def loss64(pred):
        return float((t.double() - pred.double())[valid].abs().mean())

    loss = MaskedL1Loss()(m(x), t)
    loss.backward()
    params = dict(m.named_parameters())
    grads = {k: p.grad.detach().clone() for k, p in params.items()}
    groups = {"all": list(params), "stems": ["conv1.weight", "conv1_depth.weight", "bn1.weight", "bn1_depth.bias"],
              "rgb_encoder": [k for k in params if k.startswith("layer") and "_depth" not in k],
              "depth_encoder": [k for k in params if "_depth" in k and k.startswith("layer")],
              "fusion": ["conv_fusion.weight", "bn_fusion.weight", "conv2.weight", "bn2.bias"],
              "decoder": [k for k in params if k.startswith("decoder")], "head": ["conv3.weight"]}
    report = {}
    for gname, keys in groups.items():
        # direction = the gradient itself, rescaled per tensor to the weight's magnitude: every term of <grad, d> is
        # positive, so there is no cancellation that would amplify the (ReLU-mask) noise of individual tensors
        dirs = {k: grads[k] * (params[k].detach().abs().mean().clamp_min(1e-3) / grads[k].abs().mean().clamp_min(1e-20))
                for k in keys}
        analytic = sum(float((grads[k].double() * dirs[k].double()).sum()) for k in keys)
        nums = []
        steps = (4e-5, 2e-5, 1e-5)
        for eps in steps:
            vals = []
            with torch.no_grad():
                for sgn in (+1, -1):
                    for k in keys:
                        params[k].add_(dirs[k], alpha=sgn * eps)
                    vals.append(loss64(m(x)))
                    for k in keys:
                        params[k].add_(dirs[k], alpha=-sgn * eps)
            nums.append((vals[0] - vals[1]) / (2 * eps))
        # the loss is piecewise linear in the weights (ReLU / max-pool / |.| kinks): the central difference has an
        # error linear in eps.  Least-squares line through three step sizes, intercept at eps -> 0 (a two-point
        # extrapolation triples the fp32 rounding noise of the forward; the forward itself is deterministic in this mode)
        n = len(steps)
        mx, my = sum(steps) / n, sum(nums) / n
        slope = sum((e - mx) * (v - my) for e, v in zip(steps, nums)) / sum((e - mx) ** 2 for e in steps)
        extrap = my - slope * mx
        resid = max(abs(v - (extrap + slope * e)) for e, v in zip(steps, nums))
        report[gname] = (extrap, analytic, resid, *nums)
    print("[directional]", {k: tuple(round(x, 5) for x in v) for k, v in report.items()})
    # tolerance: 3 % plus the scatter of the three finite differences around their own line (measured forward noise)
    bad = {k: v for k, v in report.items() if abs(v[0] - v[1]) > 3e-2 * abs(v[1]) + 3.0 * v[2]}
    assert not bad, (bad, report)


