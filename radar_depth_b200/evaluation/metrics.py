"""On-device evaluation metrics: the reference's ``evaluation/metrics.py`` (``Result``: metrics.py:12-58,
``Result_multidist``: metrics.py:61-140) behind the same class names, attributes and ``evaluate(output, target)`` call,
computed by ONE masked multi-reduction kernel (``rd_depth_metrics``) and one 80-byte device->host copy per call instead
of ~10 boolean gathers + host synchronisations (the reference runs this every training iteration, main.py:450-458).
There is no CPU fallback: CPU tensors raise ``RdError``."""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _lib

_SLOTS = 10


def _sums(output: torch.Tensor, target: torch.Tensor, lo: float, hi: float) -> np.ndarray:
    if not (output.is_cuda and target.is_cuda):
        raise _lib.RdError("radar_depth_b200 metrics run on a CUDA (sm_100a) device only; there is no CPU fallback")
    assert output.shape == target.shape, (output.shape, target.shape)
    o = output.detach().float().contiguous()
    t = target.detach().float().contiguous()
    acc = torch.empty(_SLOTS, dtype=torch.float64, device=o.device)
    _lib.call("rd_depth_metrics", o.data_ptr(), t.data_ptr(), o.numel(), float(lo), float(hi), acc.data_ptr(),
              torch.cuda.current_stream().cuda_stream)
    return acc.cpu().numpy()


def _fill(res: "Result", s: np.ndarray) -> None:
    n = s[0]
    mean = (lambda v: float(v / n)) if n > 0 else (lambda v: float("nan"))     # mean of an empty selection is NaN in torch
    res.mse = mean(s[1])
    res.rmse = math.sqrt(res.mse) if res.mse == res.mse else float("nan")
    res.mae = mean(s[2])
    res.lg10 = mean(s[3])
    res.absrel = mean(s[4])
    res.delta1, res.delta2, res.delta3 = mean(s[5]), mean(s[6]), mean(s[7])
    imse = mean(s[8])
    res.irmse = math.sqrt(imse) if imse == imse else float("nan")
    res.imae = mean(s[9])


class Result(object):
    """metrics.py:12-58."""

    def __init__(self):
        self.irmse, self.imae = 0, 0
        self.mse, self.rmse, self.mae = 0, 0, 0
        self.absrel, self.lg10 = 0, 0
        self.delta1, self.delta2, self.delta3 = 0, 0, 0
        self.data_time, self.gpu_time = 0, 0

    def set_to_worst(self):
        self.irmse, self.imae = np.inf, np.inf
        self.mse, self.rmse, self.mae = np.inf, np.inf, np.inf
        self.absrel, self.lg10 = np.inf, np.inf
        self.delta1, self.delta2, self.delta3 = 0, 0, 0
        self.data_time, self.gpu_time = 0, 0

    def update(self, irmse, imae, mse, rmse, mae, absrel, lg10, delta1, delta2, delta3, gpu_time, data_time):
        self.irmse, self.imae = irmse, imae
        self.mse, self.rmse, self.mae = mse, rmse, mae
        self.absrel, self.lg10 = absrel, lg10
        self.delta1, self.delta2, self.delta3 = delta1, delta2, delta3
        self.data_time, self.gpu_time = data_time, gpu_time

    def evaluate(self, output, target):
        _fill(self, _sums(output, target, 0.0, float("inf")))       # valid = target > 0 (metrics.py:35)
        self.data_time = 0
        self.gpu_time = 0


class Result_multidist(object):
    """metrics.py:61-140: the same statistics per distance interval (both interval ends inclusive, metrics.py:110)."""

    def __init__(self):
        self.dist_interval = [10., 20., 30., 40., 50., 60., 70., 80., 90., 100.]
        self.result_lst = [Result() for _ in range(len(self.dist_interval))]
        self.valid_label = [1 for _ in range(len(self.dist_interval))]

    def set_to_worst(self):
        for res in self.result_lst:
            res.set_to_worst()

    def update(self, result):
        # (the reference's version, metrics.py:76-87, iterates the object itself and reads a non-existent ``log10``
        # attribute, so it cannot run; this is what it evidently means)
        assert isinstance(result, Result_multidist)
        for idx, res in enumerate(result.result_lst):
            self.result_lst[idx].update(res.irmse, res.imae, res.mse, res.rmse, res.mae, res.absrel, res.lg10,
                                        res.delta1, res.delta2, res.delta3, res.gpu_time, res.data_time)

    def evaluate(self, output, target):
        for idx, interval in enumerate(self.dist_interval):
            if idx == 0:
                dist_min, dist_max = 0., interval
            elif idx == len(self.dist_interval) - 1:
                dist_min, dist_max = self.dist_interval[idx - 1], np.inf
            else:
                dist_min, dist_max = self.dist_interval[idx - 1], interval
            s = _sums(output, target, dist_min, dist_max)
            if s[0] == 0:
                self.valid_label[idx] = 0
            _fill(self.result_lst[idx], s)
