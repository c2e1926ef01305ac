"""Module-level execution of the reference graph on stock torch.nn layers (cuDNN / ATen) -- TEST AND BENCH INFRASTRUCTURE ONLY.

``latefusion_forward(m, x)`` walks the sub-modules of a model object that carries the reference's attribute names
(`/root/reference/model/models.py:539-594`): either the reference's own ``ResNet_latefusion`` or this repo's parameter-holder
class (whose children are ordinary ``nn.Conv2d`` / ``nn.BatchNorm2d`` objects), calling the children exactly in the order of
the reference's ``forward`` (models.py:627-664), its ``BasicBlock.forward`` (models.py:96-112), ``UpProjModule.forward``
(models.py:201-208) and ``Unpool.forward`` (models.py:26-27: a depthwise ``F.conv_transpose2d`` with the [[1,0],[0,0]] kernel).
So the GPU speed bar of SURVEY.md 8(d) ("reference modules as written on cuDNN") is measured through ``nn.BatchNorm2d``,
``nn.Conv2d`` and ``F.conv_transpose2d`` themselves, not through the functional restatement in torch_oracle.py, and it runs on
the GPU box where /root/reference does not exist.  tests/test_oracle_golden.py pins this walk to the goldens of the real
reference.  Nothing under radar_depth_b200/ imports this file.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _block(blk, x):                                   # models.py:96-112
    idt = x
    y = blk.relu(blk.bn1(blk.conv1(x)))
    y = blk.bn2(blk.conv2(y))
    if blk.downsample is not None:
        idt = blk.downsample[1](blk.downsample[0](x))
    return blk.relu(y + idt)


def _stage(seq, x):
    for blk in seq:
        x = _block(blk, x)
    return x


_UNPOOL_KERNELS = {}


def _unpool(x):                                       # models.py:13-27
    key = (x.shape[1], x.device, x.dtype)
    w = _UNPOOL_KERNELS.get(key)
    if w is None:
        w = torch.zeros(x.shape[1], 1, 2, 2, device=x.device, dtype=x.dtype)
        w[:, :, 0, 0] = 1
        _UNPOOL_KERNELS[key] = w
    return F.conv_transpose2d(x, w, stride=2, groups=x.shape[1])


def _upproj(mod, x):                                  # models.py:201-208
    u = _unpool(x)
    ub, bb = mod.upper_branch, mod.bottom_branch
    a = ub.batchnorm2(ub.conv2(F.relu(ub.batchnorm1(ub.conv1(u)))))
    b = bb.batchnorm(bb.conv(u))
    return F.relu(a + b)


def latefusion_forward(m, x):                         # models.py:627-664 / multistage_model.py:232-276
    rgb, d = x[:, :3], x[:, 3:]
    rgb = F.max_pool2d(F.relu(m.bn1(m.conv1(rgb))), 3, 2, 1)
    for name in ("layer1", "layer2", "layer3", "layer4"):
        rgb = _stage(getattr(m, name), rgb)
    d = F.max_pool2d(F.leaky_relu(m.bn1_depth(m.conv1_depth(d)), 0.2), 3, 2, 1)
    for name in ("layer1_depth", "layer2_depth", "layer3_depth", "layer4_depth"):
        d = _stage(getattr(m, name), d)
    y = m.bn_fusion(m.conv_fusion(torch.cat((rgb, d), dim=1)))
    y = m.bn2(m.conv2(y))
    for name in ("layer1", "layer2", "layer3", "layer4"):
        y = _upproj(getattr(m.decoder, name), y)
    y = m.conv3(y)
    return F.interpolate(y, size=tuple(m.output_size), mode="bilinear", align_corners=True)


def sid_filter(radar, depth):                         # multistage_model.py:87-119
    alpha, beta, K = torch.tensor(5.0), torch.tensor(18.0), torch.tensor(100.0)
    thr = torch.exp(((depth * torch.log(beta / alpha).to(depth)) / K.to(depth)) + torch.log(alpha).to(depth))
    mask = ((depth - radar).abs() <= thr).to(radar.dtype)
    return radar * mask, mask


def multistage_forward(m, x):                         # multistage_model.py:63-83
    d1 = latefusion_forward(m.stage1, x)
    rf, mask = sid_filter(x[:, 3:], d1)
    d2 = latefusion_forward(m.stage2, torch.cat((x[:, :3], rf, d1), dim=1))
    return {"stage1": d1, "stage2": d2, "mask": mask, "radar_filtered": rf}


def masked_l1(pred, target):                          # criteria_new.py:44-54
    valid = (target > 0).detach()
    return (target - pred)[valid].abs().mean()


def smoothness(pred, image):                          # criteria_new.py:8-28
    dn = pred / (pred.mean(dim=(2, 3), keepdim=True) + 1e-7)
    gx = (dn[:, :, :, :-1] - dn[:, :, :, 1:]).abs()
    gy = (dn[:, :, :-1, :] - dn[:, :, 1:, :]).abs()
    wx = torch.exp(-(image[:, :, :, :-1] - image[:, :, :, 1:]).abs().mean(1, keepdim=True))
    wy = torch.exp(-(image[:, :, :-1, :] - image[:, :, 1:, :]).abs().mean(1, keepdim=True))
    return (gx * wx).mean() + (gy * wy).mean()
