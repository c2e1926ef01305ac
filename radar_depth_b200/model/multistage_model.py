"""Drop-in for the reference's model/multistage_model.py: the two-stage refinement network.

ResNet_multistage(layers, decoder, output_size, pretrained=True) (multistage_model.py:22-83):
  depth1 = stage1(x);  radar_f, mask = Filter_layer(x[:, 3:], depth1);  depth2 = stage2(cat(rgb, radar_f, depth1))
Both stages are ResNet_latefusion graphs on the sm_100a engine (stage 2 with a 2-channel depth stem, in_channels=5);
depth1 is NOT detached (multistage_model.py:75), so the stage-2 loss back-propagates into stage 1 through the stem
data-gradient of stage 2's fifth input channel.  The SID filter is one elementwise kernel (no gradient flows through
the mask, exactly as ``<=`` in the reference).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import _lib
from ..config.config_nuscenes import config_nuscenes as cfg
from ..ops import ptr, stream_ptr
from .models import ResNet_latefusion, _LAYER_CHOICES


class ResNet_latefusion2(ResNet_latefusion):
    """multistage_model.py:123-276: the latefusion graph whose depth stem takes ``in_channels - 3`` channels."""


class Filter_layer(nn.Module):
    """multistage_model.py:87-119: keep a radar return if |depth - radar| <= 5 * (18/5)^(depth/100)."""

    def __init__(self):
        super().__init__()
        self.alpha = torch.tensor(5.0)
        self.beta = torch.tensor(18.0)
        self.K = torch.tensor(100.0)

    def forward(self, sparse_depth, dense_depth):
        if not sparse_depth.is_cuda:
            raise _lib.RdError("radar_depth_b200 runs on a CUDA (sm_100a) device only; there is no CPU fallback")
        r = sparse_depth.detach().float().contiguous()
        d = dense_depth.detach().float().contiguous()
        radar_f, mask = torch.empty_like(r), torch.empty_like(r)
        _lib.call("rd_sid_filter", ptr(r), ptr(d), r.numel(), ptr(radar_f), ptr(mask), stream_ptr())
        return radar_f, mask


class ResNet_multistage(nn.Module):
    def __init__(self, layers, decoder, output_size, pretrained=True):
        if layers not in _LAYER_CHOICES:
            raise RuntimeError("Only 18, 34, 50, 101, and 152 layer model are defined for ResNet. Got {}".format(layers))
        super().__init__()
        # The reference always asks torchvision for ImageNet weights here (multistage_model.py:29-30) and then
        # overwrites every tensor from the latefusion checkpoint when pretrained=True.  Without network access that
        # download is skipped; the result after the checkpoint load is identical.
        self.stage1 = ResNet_latefusion2(layers, decoder, output_size, in_channels=4, pretrained=False)
        self.stage2 = ResNet_latefusion2(layers, decoder, output_size, in_channels=5, pretrained=False)
        self.filter_layer = Filter_layer()
        if pretrained is True:
            path = os.path.join(cfg.PROJECT_ROOT, "pretrained/resnet18_latefusion.pth.tar")
            if not os.path.exists(path):
                raise ValueError("[Error] Can't find pretrained latefusion model. "
                                 "Please follow the instructions in README.md to download the weights!")
            checkpoint = torch.load(path, map_location="cpu", weights_only=False)
            weights = checkpoint["model_state_dict"]
            self.stage1.load_state_dict(weights)
            self.stage2.load_state_dict(self.filter_state_dict(weights, self.stage2.state_dict()), strict=False)

    def filter_state_dict(self, pretrain_dict, target_dict):
        """Drops (IN PLACE, like multistage_model.py:51-61) the entries whose shape differs from the target's --
        only conv1_depth.weight (16x1x7x7 vs 16x2x7x7) for a latefusion checkpoint."""
        for key in [k for k, v in pretrain_dict.items() if target_dict[k].shape != v.shape]:
            pretrain_dict.pop(key)
        return pretrain_dict

    def forward(self, x):
        x_img = x[:, :3, :, :]
        x_d = x[:, 3:, :, :]
        depth_stage1 = self.stage1(x)
        x_d_filtered, mask = self.filter_layer(x_d, depth_stage1)
        # multistage_model.py:78-79: stage2(cat(x_img, x_d_filtered, depth_stage1)); the three sources are packed by one
        # kernel (rd_input_pack_parts) and the gradient of depth_stage1 comes back from rd_input_grad_channel
        if x.requires_grad and torch.is_grad_enabled():
            # a caller that differentiates w.r.t. the network INPUT (saliency, adversarial probes): the general path, whose
            # autograd node returns the gradient of all five stage-2 input channels
            depth_stage2 = self.stage2(torch.cat((x_img.float(), x_d_filtered, depth_stage1), dim=1))
        else:
            depth_stage2 = self.stage2.forward_parts(x_img, x_d_filtered, depth_stage1)
        return {"stage1": depth_stage1, "stage2": depth_stage2, "mask": mask, "radar_filtered": x_d_filtered}
