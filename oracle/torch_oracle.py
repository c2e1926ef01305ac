"""CPU restatement of the radar_depth hot path as pure functions over a state_dict.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the reference
file:line (relative to /root/reference) whose behaviour it restates.  The arithmetic
primitives (conv2d, max_pool2d, interpolate) are torch's, exactly as in the reference,
whose own arithmetic lives in third-party torch/torchvision (pinned torch==1.3.1,
torchvision==0.4.2 in requirements.txt:16,19; torch 2.11 / torchvision 0.26 here).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5        # nn.BatchNorm2d default, models.py:540
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------
# state_dict layout (names, shapes, registration order)  -- models.py:539-594
# --------------------------------------------------------------------------------------
def _bn_entries(out: "OrderedDict[str, Tuple[tuple, str]]", name: str, c: int) -> None:
    out[name + ".weight"] = ((c,), "bn_weight")
    out[name + ".bias"] = ((c,), "bn_bias")
    out[name + ".running_mean"] = ((c,), "bn_mean")
    out[name + ".running_var"] = ((c,), "bn_var")
    out[name + ".num_batches_tracked"] = ((), "bn_count")


def _encoder_entries(out, suffix: str, widths, first_in: int) -> None:
    """layer{1..4}{suffix}: two BasicBlocks each (models.py:75-112, 597-625; torchvision resnet18)."""
    cin = first_in
    for li, c in enumerate(widths, start=1):
        for bi in range(2):
            p = f"layer{li}{suffix}.{bi}"
            out[p + ".conv1.weight"] = ((c, cin if bi == 0 else c, 3, 3), "conv")
            _bn_entries(out, p + ".bn1", c)
            out[p + ".conv2.weight"] = ((c, c, 3, 3), "conv")
            _bn_entries(out, p + ".bn2", c)
            if bi == 0 and (li > 1):
                out[p + ".downsample.0.weight"] = ((c, cin, 1, 1), "conv")
                _bn_entries(out, p + ".downsample.1", c)
        cin = c


def _decoder_entries(e, decoder: str, c: int = 256) -> None:
    """choose_decoder (models.py:219-230): UpProj (models.py:178-216), UpConv (models.py:158-176), DeConv (models.py:135-156)."""
    for li in range(1, 5):
        p = f"decoder.layer{li}"
        if decoder == "upproj":
            e[p + ".upper_branch.conv1.weight"] = ((c // 2, c, 5, 5), "conv")
            _bn_entries(e, p + ".upper_branch.batchnorm1", c // 2)
            e[p + ".upper_branch.conv2.weight"] = ((c // 2, c // 2, 3, 3), "conv")
            _bn_entries(e, p + ".upper_branch.batchnorm2", c // 2)
            e[p + ".bottom_branch.conv.weight"] = ((c // 2, c, 5, 5), "conv")
            _bn_entries(e, p + ".bottom_branch.batchnorm", c // 2)
        elif decoder == "upconv":
            e[p + ".conv.weight"] = ((c // 2, c, 5, 5), "conv")
            _bn_entries(e, p + ".batchnorm", c // 2)
        elif decoder in ("deconv2", "deconv3"):
            k = int(decoder[6])
            e[p + f".deconv{k}.weight"] = ((c, c // 2, k, k), "convT")      # nn.ConvTranspose2d: [in, out, k, k]
            _bn_entries(e, p + ".batchnorm", c // 2)
        else:
            raise ValueError(decoder)
        c //= 2


def resnet_entries(in_channels: int = 3, decoder: str = "upproj") -> "OrderedDict[str, Tuple[tuple, str]]":
    """Names/shapes of ResNet(18, decoder) in registration order (models.py:233-281): one encoder over all input channels."""
    e: "OrderedDict[str, Tuple[tuple, str]]" = OrderedDict()
    e["conv1.weight"] = ((64, in_channels, 7, 7), "conv")
    _bn_entries(e, "bn1", 64)
    _encoder_entries(e, "", (64, 128, 256, 512), 64)
    e["conv2.weight"] = ((256, 512, 1, 1), "conv")
    _bn_entries(e, "bn2", 256)
    _decoder_entries(e, decoder)
    e["conv3.weight"] = ((1, 16, 3, 3), "conv")
    return e


def latefusion_entries(in_channels: int = 4, decoder: str = "upproj") -> "OrderedDict[str, Tuple[tuple, str]]":
    """Names/shapes of ResNet_latefusion(18, decoder) in registration order (models.py:539-588)."""
    assert in_channels > 3                                   # models.py:535
    e: "OrderedDict[str, Tuple[tuple, str]]" = OrderedDict()
    e["conv1.weight"] = ((64, 3, 7, 7), "conv")
    _bn_entries(e, "bn1", 64)
    _encoder_entries(e, "", (64, 128, 256, 512), 64)
    e["conv1_depth.weight"] = ((16, in_channels - 3, 7, 7), "conv")   # multistage_model.py:163-164
    _bn_entries(e, "bn1_depth", 16)
    _encoder_entries(e, "_depth", (16, 32, 64, 128), 16)
    e["conv_fusion.weight"] = ((512, 640, 1, 1), "conv")
    _bn_entries(e, "bn_fusion", 512)
    e["conv2.weight"] = ((256, 512, 1, 1), "conv")
    _bn_entries(e, "bn2", 256)
    _decoder_entries(e, decoder)
    e["conv3.weight"] = ((1, 16, 3, 3), "conv")
    return e


def multistage_entries() -> "OrderedDict[str, Tuple[tuple, str]]":
    """ResNet_multistage + fixs uncertainty scalars (multistage_model.py:29-31, main.py:166-172)."""
    e: "OrderedDict[str, Tuple[tuple, str]]" = OrderedDict()
    # parameters registered on the top module precede sub-module entries in state_dict()
    e["w_stage1"] = ((), "scalar")
    e["w_stage2"] = ((), "scalar")
    for k, v in latefusion_entries(4).items():
        e["stage1." + k] = v
    for k, v in latefusion_entries(5).items():
        e["stage2." + k] = v
    return e


def synth_state_dict(entries, seed: int = 7, dtype=torch.float32) -> "OrderedDict[str, Tensor]":
    """Deterministic, construction-order-independent weights (one generator per key).

    BN affine/running values are deliberately non-trivial so parity tests exercise them.
    """
    sd: "OrderedDict[str, Tensor]" = OrderedDict()
    for i, (name, (shape, kind)) in enumerate(entries.items()):
        g = torch.Generator().manual_seed(seed * 100003 + i)
        if kind == "conv":
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif kind == "convT":
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / (shape[0] * shape[2] * shape[3] / 4.0))
        elif kind == "bn_weight":
            t = torch.rand(shape, generator=g) * 0.8 + 0.6
        elif kind == "bn_bias":
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_mean":
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_var":
            t = torch.rand(shape, generator=g) * 0.8 + 0.6
        elif kind == "bn_count":
            sd[name] = torch.tensor(3, dtype=torch.int64)
            continue
        elif kind == "scalar":
            t = torch.tensor(1.0)                            # main.py:166-167
        else:
            raise KeyError(kind)
        sd[name] = t.to(dtype)
    return sd


# --------------------------------------------------------------------------------------
# synthetic batch (SURVEY.md section 8d recipe)
# --------------------------------------------------------------------------------------
def synth_batch(b: int, h: int, w: int, p_lidar: float = 0.05, seed: int = 1234,
                radar_per_image: float = 100.0) -> Tuple[Tensor, Tensor]:
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(b, 3, h, w, generator=g)
    mk = torch.rand(b, 1, h, w, generator=g) < radar_per_image / (h * w)
    radar = torch.zeros(b, 1, h, w)
    radar[mk] = torch.rand(int(mk.sum()), generator=g) * 79 + 1
    inputs = torch.cat((rgb, radar), dim=1)
    tm = torch.rand(b, 1, h, w, generator=g) < p_lidar
    target = torch.zeros(b, 1, h, w)
    target[tm] = torch.rand(int(tm.sum()), generator=g) * 79 + 1
    return inputs, target


# --------------------------------------------------------------------------------------
# layer primitives
# --------------------------------------------------------------------------------------
def _conv(x: Tensor, sd, name: str, stride: int = 1, pad: int = 0) -> Tensor:
    return F.conv2d(x, sd[name + ".weight"], None, stride, pad)   # every hot-path conv is bias=False


def _bn(x: Tensor, sd, name: str, training: bool, new_buffers: Optional[dict]) -> Tensor:
    """nn.BatchNorm2d: biased var to normalise, unbiased var into running_var (Appendix B)."""
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    if training:
        n = x.numel() // x.shape[1]
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if new_buffers is not None:
            with torch.no_grad():
                rm, rv = sd[name + ".running_mean"], sd[name + ".running_var"]
                new_buffers[name + ".running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean
                new_buffers[name + ".running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * (n / max(n - 1, 1))
                new_buffers[name + ".num_batches_tracked"] = sd[name + ".num_batches_tracked"] + 1
    else:
        mean, var = sd[name + ".running_mean"], sd[name + ".running_var"]
    scale = w * torch.rsqrt(var + BN_EPS)
    shift = b - mean * scale
    return x * scale[None, :, None, None] + shift[None, :, None, None]


def _basic_block(x, sd, p, stride, has_ds, training, nb):
    """models.py:96-112 (identical math to torchvision BasicBlock)."""
    out = F.relu(_bn(_conv(x, sd, p + ".conv1", stride, 1), sd, p + ".bn1", training, nb))
    out = _bn(_conv(out, sd, p + ".conv2", 1, 1), sd, p + ".bn2", training, nb)
    if has_ds:
        x = _bn(_conv(x, sd, p + ".downsample.0", stride, 0), sd, p + ".downsample.1", training, nb)
    return F.relu(out + x)


def _encoder(x, sd, suffix, training, nb):
    for li in range(1, 5):
        x = _basic_block(x, sd, f"layer{li}{suffix}.0", 1 if li == 1 else 2, li > 1, training, nb)
        x = _basic_block(x, sd, f"layer{li}{suffix}.1", 1, False, training, nb)
    return x


def unpool(x: Tensor) -> Tensor:
    """models.py:13-27: out[2i,2j] = x[i,j], zero elsewhere, size exactly 2H x 2W."""
    b, c, h, w = x.shape
    out = x.new_zeros(b, c, 2 * h, 2 * w)
    out[:, :, ::2, ::2] = x
    return out


def _upproj(x, sd, p, training, nb):
    """models.py:181-209."""
    u = unpool(x)
    x1 = F.relu(_bn(_conv(u, sd, p + ".upper_branch.conv1", 1, 2), sd, p + ".upper_branch.batchnorm1", training, nb))
    x1 = _bn(_conv(x1, sd, p + ".upper_branch.conv2", 1, 1), sd, p + ".upper_branch.batchnorm2", training, nb)
    x2 = _bn(_conv(u, sd, p + ".bottom_branch.conv", 1, 2), sd, p + ".bottom_branch.batchnorm", training, nb)
    return F.relu(x1 + x2)


def _upconv(x, sd, p, training, nb):
    """UpConv.upconv_module (models.py:160-169): unpool -> 5x5 conv -> BN -> ReLU."""
    return F.relu(_bn(_conv(unpool(x), sd, p + ".conv", 1, 2), sd, p + ".batchnorm", training, nb))


def _deconv(x, sd, p, k, training, nb):
    """DeConv.convt (models.py:140-151): ConvTranspose2d(k, stride 2, padding (k-1)//2, output_padding k%2) -> BN -> ReLU."""
    y = F.conv_transpose2d(x, sd[p + f".deconv{k}.weight"], None, 2, (k - 1) // 2, k % 2)
    return F.relu(_bn(y, sd, p + ".batchnorm", training, nb))


def _decoder(f, sd, decoder, training, nb):
    for li in range(1, 5):
        p = f"decoder.layer{li}"
        if decoder == "upproj":
            f = _upproj(f, sd, p, training, nb)
        elif decoder == "upconv":
            f = _upconv(f, sd, p, training, nb)
        elif decoder in ("deconv2", "deconv3"):
            f = _deconv(f, sd, p, int(decoder[6]), training, nb)
        else:
            raise ValueError(decoder)
    return f


def resnet_forward(sd, x: Tensor, output_size, training: bool = True, new_buffers: Optional[dict] = None,
                   decoder: str = "upproj") -> Tensor:
    """ResNet.forward (models.py:283-303): single encoder over all input channels, conv2/bn2, decoder, head, bilinear."""
    nb = new_buffers
    y = F.relu(_bn(_conv(x, sd, "conv1", 2, 3), sd, "bn1", training, nb))
    y = F.max_pool2d(y, 3, 2, 1)
    y = _encoder(y, sd, "", training, nb)
    y = _bn(_conv(y, sd, "conv2"), sd, "bn2", training, nb)
    y = _decoder(y, sd, decoder, training, nb)
    y = _conv(y, sd, "conv3", 1, 1)
    return F.interpolate(y, size=tuple(output_size), mode="bilinear", align_corners=True)


def _scoped(sd, new_buffers, prefix):
    if prefix:
        return _PrefixView(sd, prefix), (_PrefixSink(new_buffers, prefix) if new_buffers is not None else None)
    return sd, new_buffers


def latefusion_front(sd, x: Tensor, training: bool = True, new_buffers: Optional[dict] = None, prefix: str = "") -> Tensor:
    """ResNet_latefusion.pnp_forward_front (models.py:669-700): both encoders, the concat and the two 1x1 conv+BN
    pairs; returns bn2's output (256 channels at 1/32 resolution)."""
    assert x.shape[1] >= 4                                   # multistage_model.py:233
    sd, nb = _scoped(sd, new_buffers, prefix)
    xi, xd = x[:, :3], x[:, 3:]
    xi = F.relu(_bn(_conv(xi, sd, "conv1", 2, 3), sd, "bn1", training, nb))
    xi = F.max_pool2d(xi, 3, 2, 1)
    xi = _encoder(xi, sd, "", training, nb)
    xd = F.leaky_relu(_bn(_conv(xd, sd, "conv1_depth", 2, 3), sd, "bn1_depth", training, nb), 0.2)
    xd = F.max_pool2d(xd, 3, 2, 1)
    xd = _encoder(xd, sd, "_depth", training, nb)
    f = torch.cat((xi, xd), dim=1)
    f = _bn(_conv(f, sd, "conv_fusion"), sd, "bn_fusion", training, nb)     # no activation, models.py:652-657
    return _bn(_conv(f, sd, "conv2"), sd, "bn2", training, nb)


def latefusion_rear(sd, f: Tensor, output_size, training: bool = True, new_buffers: Optional[dict] = None,
                    prefix: str = "", decoder: str = "upproj") -> Tensor:
    """ResNet_latefusion.pnp_forward_rear (models.py:702-707): decoder, 3x3 head, bilinear resize."""
    sd, nb = _scoped(sd, new_buffers, prefix)
    f = _decoder(f, sd, decoder, training, nb)
    f = _conv(f, sd, "conv3", 1, 1)
    return F.interpolate(f, size=tuple(output_size), mode="bilinear", align_corners=True)


def latefusion_forward(sd, x: Tensor, output_size, training: bool = True,
                       new_buffers: Optional[dict] = None, prefix: str = "", decoder: str = "upproj") -> Tensor:
    """ResNet_latefusion.forward (models.py:627-664) / ResNet_latefusion2.forward
    (multistage_model.py:232-276) = rear(front(x))."""
    f = latefusion_front(sd, x, training, new_buffers, prefix)
    return latefusion_rear(sd, f, output_size, training, new_buffers, prefix, decoder)


class _PrefixView:
    def __init__(self, sd, prefix):
        self.sd, self.prefix = sd, prefix

    def __getitem__(self, k):
        return self.sd[self.prefix + k]


class _PrefixSink:
    def __init__(self, d, prefix):
        self.d, self.prefix = d, prefix

    def __setitem__(self, k, v):
        self.d[self.prefix + k] = v


def filter_layer(sparse_depth: Tensor, dense_depth: Tensor) -> Tuple[Tensor, Tensor]:
    """Filter_layer.forward (multistage_model.py:87-119): thr = exp(d*ln(18/5)/100 + ln 5)."""
    thr = torch.exp(dense_depth * math.log(18.0 / 5.0) / 100.0 + math.log(5.0))
    mask = (torch.abs(dense_depth - sparse_depth) <= thr).to(sparse_depth.dtype)
    return sparse_depth * mask, mask


def multistage_forward(sd, x: Tensor, output_size, training: bool = True,
                       new_buffers: Optional[dict] = None) -> Dict[str, Tensor]:
    """ResNet_multistage.forward (multistage_model.py:63-83); depth1 is NOT detached."""
    d1 = latefusion_forward(sd, x, output_size, training, new_buffers, prefix="stage1.")
    radar_f, mask = filter_layer(x[:, 3:], d1)
    x2 = torch.cat((x[:, :3], radar_f, d1), dim=1)
    d2 = latefusion_forward(sd, x2, output_size, training, new_buffers, prefix="stage2.")
    return {"stage1": d1, "stage2": d2, "mask": mask, "radar_filtered": radar_f}


# --------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------
def masked_l1(pred: Tensor, target: Tensor) -> Tensor:
    """MaskedL1Loss.forward (criteria_new.py:48-54)."""
    assert pred.dim() == target.dim(), "inconsistent dimensions"
    valid = target > 0
    return (target - pred)[valid].abs().mean()


def smoothness(pred_depth: Tensor, image: Tensor) -> Tensor:
    """SmoothnessLoss.forward (criteria_new.py:12-28)."""
    d = pred_depth / (pred_depth.mean(2, True).mean(3, True) + 1e-7)
    gx = (d[:, :, :, :-1] - d[:, :, :, 1:]).abs()
    gy = (d[:, :, :-1, :] - d[:, :, 1:, :]).abs()
    ix = (image[:, :, :, :-1] - image[:, :, :, 1:]).abs().mean(1, keepdim=True)
    iy = (image[:, :, :-1, :] - image[:, :, 1:, :]).abs().mean(1, keepdim=True)
    return (gx * torch.exp(-ix)).mean() + (gy * torch.exp(-iy)).mean()


def fixs_loss(out: Dict[str, Tensor], inputs: Tensor, target: Tensor, w1: Tensor, w2: Tensor,
              w_smooth: float = 0.1) -> Tensor:
    """main.py:416-429 (resnet18_multistage_uncertainty_fixs); smoothness sees all 4 input channels."""
    l1 = masked_l1(out["stage1"], target)
    l2 = masked_l1(out["stage2"], target)
    s = smoothness(out["stage1"], inputs)
    return torch.exp(-w1) * (l1 + w_smooth * s) + torch.exp(-w2) * l2 + w1 + w2


# --------------------------------------------------------------------------------------
# one training step (main.py:416-445), functional
# --------------------------------------------------------------------------------------
def _leaf_params(sd, dtype):
    params = OrderedDict()
    work = OrderedDict()
    for k, v in sd.items():
        if v.dtype.is_floating_point and not (k.endswith("running_mean") or k.endswith("running_var")):
            p = v.detach().to(dtype).clone().requires_grad_(True)
            params[k] = p
            work[k] = p
        elif v.dtype.is_floating_point:
            work[k] = v.detach().to(dtype)
        else:
            work[k] = v
    return params, work


def train_step(sd, inputs: Tensor, target: Tensor, arch: str = "latefusion", training: bool = True,
               dtype=torch.float32, decoder: str = "upproj"):
    """fwd + loss + bwd.  Returns dict(pred|preds, loss, grads, new_buffers)."""
    params, work = _leaf_params(sd, dtype)
    inputs, target = inputs.to(dtype), target.to(dtype)
    nb: dict = {}
    size = inputs.shape[-2:]
    if arch == "latefusion":
        pred = latefusion_forward(work, inputs, size, training, nb, decoder=decoder)
        loss = masked_l1(pred, target)
        outs = {"pred": pred.detach()}
    elif arch == "resnet":
        pred = resnet_forward(work, inputs, size, training, nb, decoder=decoder)
        loss = masked_l1(pred, target)
        outs = {"pred": pred.detach()}
    elif arch == "multistage_fixs":
        out = multistage_forward(work, inputs, size, training, nb)
        loss = fixs_loss(out, inputs, target, work["w_stage1"], work["w_stage2"])
        outs = {k: v.detach() for k, v in out.items()}
    else:
        raise ValueError(arch)
    loss.backward()
    grads = OrderedDict((k, p.grad) for k, p in params.items())
    return {**outs, "loss": loss.detach(), "grads": grads, "new_buffers": nb}


def sgd_step(sd, grads, momentum_buf: Optional[dict], lr=0.01, momentum=0.9, weight_decay=1e-4):
    """torch.optim.SGD as configured at main.py:285-290 (dampening 0, nesterov False)."""
    new_sd = OrderedDict(sd)
    new_buf = {}
    for k, g in grads.items():
        if g is None:
            continue
        p = sd[k]
        d = g.to(p.dtype) + weight_decay * p
        buf = d.clone() if momentum_buf is None or k not in momentum_buf else momentum * momentum_buf[k] + d
        new_buf[k] = buf
        new_sd[k] = p - lr * buf
    return new_sd, new_buf


def depth_metrics(output: Tensor, target: Tensor, lo: float = 0.0, hi: float = float("inf")) -> Dict[str, float]:
    """evaluation/metrics.py:34-58 (Result.evaluate) and :91-140 (one interval of Result_multidist.evaluate):
    statistics over the pixels with target > 0 and lo <= target <= hi.  Means of an empty selection are NaN."""
    import math
    m = (target > 0) & (target >= lo) & (target <= hi)
    o, t = output[m], target[m]
    d = (o - t).abs()
    ratio = torch.max(o / t, t / o)
    di = (1 / o - 1 / t).abs()
    mse = float((d ** 2).mean())
    return dict(mse=mse, rmse=math.sqrt(mse) if mse == mse else float("nan"), mae=float(d.mean()),
                lg10=float((torch.log(o) / math.log(10) - torch.log(t) / math.log(10)).abs().mean()),
                absrel=float((d / t).mean()), delta1=float((ratio < 1.25).float().mean()),
                delta2=float((ratio < 1.25 ** 2).float().mean()), delta3=float((ratio < 1.25 ** 3).float().mean()),
                irmse=math.sqrt(float((di ** 2).mean())) if o.numel() else float("nan"), imae=float(di.mean()),
                count=int(m.sum()))
