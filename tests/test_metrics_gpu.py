"""rd_depth_metrics / radar_depth_b200.evaluation.metrics against the oracle and the reference's golden values
(evaluation/metrics.py:34-58, 91-140)."""
import os

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O

pytestmark = pytest.mark.gpu
FIELDS = ("mse", "rmse", "mae", "lg10", "absrel", "delta1", "delta2", "delta3", "irmse", "imae")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_result_matches_reference_golden_and_oracle():
    from radar_depth_b200.evaluation.metrics import Result, Result_multidist
    g = np.load(os.path.join(GOLDEN, "metrics.npz"))
    out, tgt = torch.from_numpy(g["output"]).cuda(), torch.from_numpy(g["target"]).cuda()
    r = Result()
    r.evaluate(out, tgt)
    # element math is fp32 on both sides; the reference averages in fp32, the kernel sums in fp64
    np.testing.assert_allclose([getattr(r, k) for k in FIELDS], g["result"], rtol=2e-5, atol=1e-7)
    assert r.data_time == 0 and r.gpu_time == 0
    md = Result_multidist()
    md.evaluate(out, tgt)
    got = np.array([[getattr(x, k) for k in FIELDS] for x in md.result_lst])
    np.testing.assert_allclose(got, g["multidist"], rtol=2e-5, atol=1e-7, equal_nan=True)
    assert md.valid_label == [int(v) for v in g["valid_label"]]


@pytest.mark.parametrize("shape,p_valid", [((16, 1, 352, 1216), 0.05), ((1, 1, 7, 13), 0.5), ((2, 1, 24, 40), 0.0)])
def test_result_matches_oracle_at_other_sizes(shape, p_valid):
    from radar_depth_b200.evaluation.metrics import Result
    g = torch.Generator().manual_seed(3)
    tgt = torch.rand(*shape, generator=g) * 79 + 1
    tgt[torch.rand(*shape, generator=g) >= p_valid] = 0
    out = torch.rand(*shape, generator=g) * 60 + 0.5
    ref = O.depth_metrics(out, tgt)
    r = Result()
    r.evaluate(out.cuda(), tgt.cuda())
    np.testing.assert_allclose([getattr(r, k) for k in FIELDS], [ref[k] for k in FIELDS], rtol=2e-5, atol=1e-7, equal_nan=True)


def test_cpu_tensors_are_refused():
    from radar_depth_b200 import _lib
    from radar_depth_b200.evaluation.metrics import Result
    with pytest.raises(_lib.RdError):
        Result().evaluate(torch.ones(1, 1, 4, 4), torch.ones(1, 1, 4, 4))
