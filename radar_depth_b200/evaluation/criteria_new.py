"""Drop-in for the hot-path losses of the reference's evaluation/criteria_new.py on sm_100a kernels.

MaskedL1Loss (criteria_new.py:44-54) is computed WITHOUT the reference's boolean gather (``diff[valid_mask]``
forces a device->host sync for the dynamic shape): one reduction kernel accumulates sum|target-pred| and the valid
count in fp64 on the device, a one-thread kernel divides; the backward kernel writes -sign(target-pred)/count.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib, determinism
from ..ops import ptr, stream_ptr


class _MaskedL1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        pred_c = pred.detach().float().contiguous()
        tgt_c = target.detach().float().contiguous()
        acc = torch.empty(2, dtype=torch.float64, device=pred.device)
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        with determinism.mode(determinism.loss_scratch(pred.device)):     # fixed-order loss sums: reproducible run to run
            _lib.call("rd_l1_fwd", ptr(pred_c), ptr(tgt_c), pred_c.numel(), ptr(acc), ptr(loss), stream_ptr())
        ctx.save_for_backward(pred_c, tgt_c, acc)
        return loss

    @staticmethod
    def backward(ctx, gout):
        pred_c, tgt_c, acc = ctx.saved_tensors
        gpred = torch.empty_like(pred_c)
        g = gout.detach().float().contiguous()
        _lib.call("rd_l1_bwd", ptr(pred_c), ptr(tgt_c), pred_c.numel(), ptr(acc), ptr(g), ptr(gpred), 0, stream_ptr())
        return gpred, None


class MaskedL1Loss(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, pred, target):
        assert pred.dim() == target.dim(), "inconsistent dimensions"      # criteria_new.py:49
        if not pred.is_cuda:
            raise _lib.RdError("radar_depth_b200 losses run on a CUDA (sm_100a) device only; there is no CPU fallback")
        self.loss = _MaskedL1Fn.apply(pred, target)
        return self.loss


class _SmoothnessFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, image):
        pred_c = pred.detach().float().contiguous()
        img_c = image.detach().float().contiguous()
        B, _, H, W = pred_c.shape
        scratch = torch.empty(2 * B + 2, dtype=torch.float64, device=pred.device)
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        with determinism.mode(determinism.loss_scratch(pred.device)):
            _lib.call("rd_smoothness_fwd", ptr(pred_c), ptr(img_c), B, img_c.shape[1], H, W, ptr(scratch), ptr(loss), stream_ptr())
        ctx.save_for_backward(pred_c, img_c, scratch)
        return loss

    @staticmethod
    def backward(ctx, gout):
        pred_c, img_c, scratch = ctx.saved_tensors
        B, _, H, W = pred_c.shape
        gpred = torch.empty_like(pred_c)
        g = gout.detach().float().contiguous()
        with determinism.mode(determinism.loss_scratch(pred_c.device)):
            _lib.call("rd_smoothness_bwd", ptr(pred_c), ptr(img_c), B, img_c.shape[1], H, W, ptr(scratch), ptr(g), ptr(gpred), 0,
                      stream_ptr())
        return gpred, None


class SmoothnessLoss(nn.Module):
    """Edge-aware first-difference smoothness of the mean-normalised depth (criteria_new.py:8-28).  ``image`` is
    whatever the caller passes -- main.py:422 passes the full 4-channel network input, which is preserved."""

    def __init__(self):
        super().__init__()

    def forward(self, pred_depth, image):
        if not pred_depth.is_cuda:
            raise _lib.RdError("radar_depth_b200 losses run on a CUDA (sm_100a) device only; there is no CPU fallback")
        assert pred_depth.dim() == 4 and pred_depth.shape[1] == 1 and image.shape[-2:] == pred_depth.shape[-2:]
        return _SmoothnessFn.apply(pred_depth, image)
