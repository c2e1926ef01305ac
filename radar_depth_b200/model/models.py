"""Drop-in for the hot-path classes of the reference's model/models.py.

``ResNet_latefusion(layers, decoder, output_size, in_channels=4, pretrained=True)`` keeps the reference's
constructor signature, attribute names, parameter/buffer names, shapes and registration order
(models.py:519-594 -> 325 state_dict entries), so ``main.py``'s ``create_model`` / ``load_state_dict`` /
``state_dict`` / optimizer code works unchanged.  The sub-modules below are PARAMETER HOLDERS ONLY: their own
``forward`` is never called.  ``forward(x)`` runs the whole encoder-decoder through the sm_100a kernels
(radar_depth_b200.engine) as one autograd node; ``loss.backward()`` fills ``.grad`` of every parameter.

north_star names layers=18 and decoder='upproj'; the other decoders (`upconv`, `deconv2`, `deconv3`, models.py:135-176) and
the single-encoder `ResNet` (models.py:233-303) run on the same convolution programs.  Deeper encoders
(34/50/101/152) raise NotImplementedError after the reference's own argument errors, which are mirrored.
There is no CPU / PyTorch fallback: without a CUDA device and the built library, forward raises.
"""
from __future__ import annotations

import collections
import math
import os

import torch
import torch.nn as nn

from .. import _lib
from ..engine import LatefusionEngine

_LAYER_CHOICES = (18, 34, 50, 101, 152)          # models.py:522
_DECODER_NAMES = ("deconv2", "deconv3", "upconv", "upproj")


def _precision_from_env() -> str:
    return os.environ.get("RADAR_DEPTH_B200_PRECISION", "bf16").lower()


# ---------------------------------------------------------------------------------------------- init (models.py:30-72)
def weights_init(m):
    """N(0, sqrt(2/(kh*kw*Cout))) for convs, BN -> (1, 0).  Same distributions as models.py:30-44."""
    if isinstance(m, nn.Conv2d):
        n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
        with torch.no_grad():
            m.weight.normal_(0.0, math.sqrt(2.0 / n))
            if m.bias is not None:
                m.bias.zero_()
    elif isinstance(m, nn.BatchNorm2d):
        with torch.no_grad():
            m.weight.fill_(1.0)
            m.bias.zero_()


def _kaiming(m, nonlinearity):
    if isinstance(m, nn.Conv2d):
        nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity=nonlinearity)
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
        nn.init.ones_(m.weight)
        nn.init.zeros_(m.bias)


def weights_init_kaiming(m):          # models.py:47-58
    _kaiming(m, "relu")


def weights_init_kaiming_leaky(m):    # models.py:61-72
    _kaiming(m, "leaky_relu")


# ---------------------------------------------------------------------------------------------- parameter holders
class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError(f"{type(self).__name__} only holds parameters; the owning model's forward runs the "
                           "fused sm_100a kernels")


class BasicBlock(_Holder):
    """conv3x3-bn-relu-conv3x3-bn (+downsample) + residual + relu (models.py:75-112, identical to torchvision's)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=64, dilation=1, norm_layer=None):
        super().__init__()
        norm_layer = norm_layer or nn.BatchNorm2d
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = norm_layer(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = norm_layer(planes)
        self.downsample = downsample
        self.stride = stride


def _make_layer(inplanes, planes, blocks, stride):
    downsample = None
    if stride != 1 or inplanes != planes:
        downsample = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
    layers = [BasicBlock(inplanes, planes, stride, downsample)]
    for _ in range(1, blocks):
        layers.append(BasicBlock(planes, planes))
    seq = nn.Sequential(*layers)
    for m in seq.modules():           # models.py:621-623 (and torchvision's own init for the RGB branch)
        weights_init_kaiming(m)
    return seq


class Unpool(_Holder):
    """Zero-stuffing x2 upsampling (models.py:13-27).  Never executed on its own here: it is folded into the
    4-phase sub-pixel form of the UpProj 5x5 convolutions (convplan.gconv_upproj)."""

    def __init__(self, num_channels, stride=2):
        super().__init__()
        self.num_channels, self.stride = num_channels, stride


class Decoder(_Holder):
    names = list(_DECODER_NAMES)      # utils.py:11,23 reads Decoder.names


class UpProj(Decoder):
    class UpProjModule(_Holder):      # models.py:181-209
        def __init__(self, in_channels):
            super().__init__()
            out_channels = in_channels // 2
            self.unpool = Unpool(in_channels)
            self.upper_branch = nn.Sequential(collections.OrderedDict([
                ("conv1", nn.Conv2d(in_channels, out_channels, 5, 1, 2, bias=False)),
                ("batchnorm1", nn.BatchNorm2d(out_channels)),
                ("relu", nn.ReLU()),
                ("conv2", nn.Conv2d(out_channels, out_channels, 3, 1, 1, bias=False)),
                ("batchnorm2", nn.BatchNorm2d(out_channels)),
            ]))
            self.bottom_branch = nn.Sequential(collections.OrderedDict([
                ("conv", nn.Conv2d(in_channels, out_channels, 5, 1, 2, bias=False)),
                ("batchnorm", nn.BatchNorm2d(out_channels)),
            ]))
            self.relu = nn.ReLU()

    def __init__(self, in_channels):
        super().__init__()
        self.layer1 = self.UpProjModule(in_channels)
        self.layer2 = self.UpProjModule(in_channels // 2)
        self.layer3 = self.UpProjModule(in_channels // 4)
        self.layer4 = self.UpProjModule(in_channels // 8)


class DeConv(Decoder):
    """Four ConvTranspose2d(k, stride 2) + BN + ReLU stages (models.py:135-156): on the engine each is the 4-phase
    data-gradient form of a stride-2 convolution (convplan.gconv_deconv)."""

    def __init__(self, in_channels, kernel_size):
        assert kernel_size >= 2, "kernel_size out of range: {}".format(kernel_size)
        super().__init__()
        if kernel_size not in (2, 3):
            raise NotImplementedError("only deconv2 / deconv3 (Decoder.names) are built")

        def convt(c):
            padding = (kernel_size - 1) // 2
            output_padding = kernel_size % 2
            assert -2 - 2 * padding + kernel_size + output_padding == 0, "deconv parameters incorrect"
            return nn.Sequential(collections.OrderedDict([
                ("deconv{}".format(kernel_size), nn.ConvTranspose2d(c, c // 2, kernel_size, 2, padding, output_padding, bias=False)),
                ("batchnorm", nn.BatchNorm2d(c // 2)),
                ("relu", nn.ReLU(inplace=True)),
            ]))
        self.layer1 = convt(in_channels)
        self.layer2 = convt(in_channels // 2)
        self.layer3 = convt(in_channels // 4)
        self.layer4 = convt(in_channels // 8)


class UpConv(Decoder):
    """Four unpool -> 5x5 conv -> BN -> ReLU stages (models.py:158-176): 4-phase sub-pixel programs (convplan.gconv_upconv)."""

    def upconv_module(self, in_channels):
        return nn.Sequential(collections.OrderedDict([
            ("unpool", Unpool(in_channels)),
            ("conv", nn.Conv2d(in_channels, in_channels // 2, 5, 1, 2, bias=False)),
            ("batchnorm", nn.BatchNorm2d(in_channels // 2)),
            ("relu", nn.ReLU()),
        ]))

    def __init__(self, in_channels):
        super().__init__()
        self.layer1 = self.upconv_module(in_channels)
        self.layer2 = self.upconv_module(in_channels // 2)
        self.layer3 = self.upconv_module(in_channels // 4)
        self.layer4 = self.upconv_module(in_channels // 8)


def choose_decoder(decoder, in_channels):
    """models.py:219-230."""
    if decoder[:6] == "deconv":
        assert len(decoder) == 7
        return DeConv(in_channels, int(decoder[6]))
    elif decoder == "upproj":
        return UpProj(in_channels)
    elif decoder == "upconv":
        return UpConv(in_channels)
    assert False, "invalid option for decoder: {}".format(decoder)


# ---------------------------------------------------------------------------------------------- autograd node
class _LatefusionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, anchor, module):
        eng = module._get_engine()
        pred = eng.forward(x, module.training)
        module._fwd_serial += 1
        ctx.module, ctx.serial = module, module._fwd_serial
        ctx.need_dx = x.requires_grad
        ctx.training = module.training
        return pred.clone()

    @staticmethod
    def backward(ctx, dpred):
        module = ctx.module
        if ctx.serial != module._fwd_serial:
            raise RuntimeError("radar_depth_b200: the activations of this forward pass were overwritten by a later "
                               "forward of the same model; call backward before the next forward")
        eng = module._engine
        accumulate = eng.grads_bound()
        eng.backward(dpred.contiguous(), accumulate, ctx.training)
        eng.bind_grads()
        dx = eng.input_grad() if ctx.need_dx else None
        return dx, None, None


class _LatefusionPartsFn(torch.autograd.Function):
    """forward() on an input given as channel groups (rgb, radar, stage-1 depth): the stage-2 call of ResNet_multistage
    without torch.cat (multistage_model.py:78).  Only the last group (ONE channel) may carry a gradient."""

    @staticmethod
    def forward(ctx, a, b, c, anchor, module):
        eng = module._get_engine()
        pred = eng.forward((a, b, c), module.training)
        module._fwd_serial += 1
        ctx.module, ctx.serial = module, module._fwd_serial
        ctx.need_dc = c.requires_grad
        ctx.c_index = int(a.shape[1] + b.shape[1])
        ctx.training = module.training
        return pred.clone()

    @staticmethod
    def backward(ctx, dpred):
        module = ctx.module
        if ctx.serial != module._fwd_serial:
            raise RuntimeError("radar_depth_b200: the activations of this forward pass were overwritten by a later "
                               "forward of the same model; call backward before the next forward")
        eng = module._engine
        accumulate = eng.grads_bound()
        eng.backward(dpred.contiguous(), accumulate, ctx.training)
        eng.bind_grads()
        dc = eng.input_grad_channel(ctx.c_index) if ctx.need_dc else None
        return None, None, dc, None, None


class _RearFn(torch.autograd.Function):
    """pnp_forward_rear with a gradient w.r.t. its input feature (decoder parameters get no gradient on this path)."""

    @staticmethod
    def forward(ctx, feat, module):
        eng = module._get_engine()
        pred = eng.forward_rear(feat.detach().float().contiguous(), module.training, module._image_hw())
        module._fwd_serial += 1
        ctx.module, ctx.serial, ctx.training = module, module._fwd_serial, module.training
        return pred.clone()

    @staticmethod
    def backward(ctx, dpred):
        module = ctx.module
        if ctx.serial != module._fwd_serial:
            raise RuntimeError("radar_depth_b200: the activations of this rear pass were overwritten by a later forward "
                               "of the same model; call backward before the next forward")
        return module._engine.backward_rear(dpred.contiguous(), ctx.training).clone(), None


class ResNet_latefusion(nn.Module):
    def __init__(self, layers, decoder, output_size, in_channels=4, pretrained=True):
        if layers not in _LAYER_CHOICES:
            raise RuntimeError("Only 18, 34, 50, 101, and 152 layer model are defined for ResNet. Got {}".format(layers))
        super().__init__()
        if layers != 18:
            raise NotImplementedError("only the ResNet-18 encoder is built for B200 (north_star: resnet18_latefusion)")
        assert in_channels > 3                       # models.py:535
        self.output_size = output_size
        self.in_channels = in_channels
        self.decoder_name = decoder
        self._arch = "latefusion"
        # ---- RGB branch (models.py:539-551)
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = _make_layer(64, 64, 2, 1)
        self.layer2 = _make_layer(64, 128, 2, 2)
        self.layer3 = _make_layer(128, 256, 2, 2)
        self.layer4 = _make_layer(256, 512, 2, 2)
        # ---- depth branch (models.py:559-569; multistage_model.py:163-164 for in_channels=5)
        self.conv1_depth = nn.Conv2d(in_channels - 3, 16, 7, 2, 3, bias=False)
        self.bn1_depth = nn.BatchNorm2d(16)
        # reference quirk kept (models.py:561-562): the leaky-Kaiming init lands on the RGB stem, the depth stem
        # keeps PyTorch's default init
        weights_init_kaiming_leaky(self.conv1)
        weights_init_kaiming(self.bn1)
        self.relu_depth = nn.LeakyReLU(0.2, inplace=True)
        self.maxpool_depth = nn.MaxPool2d(3, 2, 1)
        self.layer1_depth = _make_layer(16, 16, 2, 1)
        self.layer2_depth = _make_layer(16, 32, 2, 2)
        self.layer3_depth = _make_layer(32, 64, 2, 2)
        self.layer4_depth = _make_layer(64, 128, 2, 2)
        # ---- fusion, decoder, head (models.py:573-594)
        self.conv_fusion = nn.Conv2d(640, 512, 1, bias=False)
        self.bn_fusion = nn.BatchNorm2d(512)
        self.conv2 = nn.Conv2d(512, 256, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(256)
        self.decoder = choose_decoder(decoder, 256)
        self.conv3 = nn.Conv2d(16, 1, 3, 1, 1, bias=False)
        self.bilinear = nn.Upsample(size=self.output_size, mode="bilinear", align_corners=True)
        self.conv2.apply(weights_init)
        self.bn2.apply(weights_init)
        self.decoder.apply(weights_init)
        self.conv3.apply(weights_init)
        if pretrained:
            self._load_torchvision_rgb_layers()
        self.precision = _precision_from_env()       # "bf16" (throughput) | "fp32" (parity: 3-term bf16 split)
        self._engine = None
        self._anchor = None
        self._fwd_serial = 0

    # ImageNet weights for layer1-4 exactly like the reference's torchvision resnet18(pretrained=True) (models.py:526)
    def _load_torchvision_rgb_layers(self):
        import torchvision
        tv = torchvision.models.resnet18(weights=torchvision.models.ResNet18_Weights.IMAGENET1K_V1)
        for name in ("layer1", "layer2", "layer3", "layer4"):
            getattr(self, name).load_state_dict(getattr(tv, name).state_dict())

    def _get_engine(self) -> LatefusionEngine:
        want = _lib.RD_F32 if self.precision in ("fp32", "f32", "parity") else _lib.RD_BF16
        if self._engine is None or self._engine.act_dtype != want or tuple(self._engine.output_size) != tuple(self.output_size):
            self._engine = LatefusionEngine(self, self.in_channels, self.output_size, want, arch=self._arch,
                                            decoder=self.decoder_name)
        return self._engine

    def forward(self, x):
        if self._arch == "latefusion":
            assert x.shape[1] >= 4                   # multistage_model.py:233
        if not x.is_cuda:
            raise _lib.RdError("radar_depth_b200 runs on a CUDA (sm_100a) device only; there is no CPU fallback")
        x = x.float().contiguous()
        if torch.is_grad_enabled():
            if self._anchor is None or self._anchor.device != x.device:
                self._anchor = torch.zeros(1, device=x.device, requires_grad=True)
            return _LatefusionFn.apply(x, self._anchor, self)
        eng = self._get_engine()
        self._fwd_serial += 1                        # the saved activations of an earlier grad-enabled forward are gone:
        return eng.forward(x, self.training, inference=True).clone()  # its backward() must raise, not differentiate the wrong pass

    def forward_parts(self, rgb, radar, depth):
        """forward(torch.cat((rgb, radar, depth), 1)) without the concatenation: every part is fp32 NCHW with dense channel
        planes (a channel slice of a contiguous tensor qualifies); ``depth`` is one channel and may require grad, the other
        parts are treated as constants (what ResNet_multistage passes, multistage_model.py:73-79)."""
        assert depth.shape[1] == 1 and not rgb.requires_grad and not radar.requires_grad
        parts = []
        for t in (rgb, radar, depth):
            if not t.is_cuda:
                raise _lib.RdError("radar_depth_b200 runs on a CUDA (sm_100a) device only; there is no CPU fallback")
            t = t.float()
            dense = t.stride(3) == 1 and t.stride(2) == t.shape[3] and t.stride(1) == t.shape[2] * t.shape[3]
            parts.append(t if dense else t.contiguous())
        if torch.is_grad_enabled():
            if self._anchor is None or self._anchor.device != depth.device:
                self._anchor = torch.zeros(1, device=depth.device, requires_grad=True)
            return _LatefusionPartsFn.apply(parts[0], parts[1], parts[2], self._anchor, self)
        eng = self._get_engine()
        self._fwd_serial += 1
        return eng.forward(tuple(parts), self.training, inference=True).clone()

    # API surface of models.py:669-707 (PnP-Depth refinement); main.py never calls them.  front = encoder
    # + fusion 1x1s up to bn2's output, rear = decoder + head + bilinear.  rear(front(x)) == forward(x).  The usual PnP
    # loop differentiates the rear with respect to the FEATURE, which is what the rear's autograd node provides;
    # parameter gradients are only produced by the un-cut forward().
    def pnp_forward_front(self, x):
        assert x.shape[1] >= 4
        if not x.is_cuda:
            raise _lib.RdError("radar_depth_b200 runs on a CUDA (sm_100a) device only; there is no CPU fallback")
        self._fwd_serial += 1                        # the activations of an earlier forward() are overwritten
        return self._get_engine().forward_front(x.float().contiguous(), self.training,
                                                inference=not torch.is_grad_enabled()).clone()

    def pnp_forward_rear(self, x):
        if not x.is_cuda:
            raise _lib.RdError("radar_depth_b200 runs on a CUDA (sm_100a) device only; there is no CPU fallback")
        if torch.is_grad_enabled() and x.requires_grad:
            return _RearFn.apply(x, self)
        self._fwd_serial += 1
        return self._get_engine().forward_rear(x.detach().float().contiguous(), self.training, self._image_hw(),
                                               inference=not torch.is_grad_enabled()).clone()

    def _image_hw(self):
        eng = self._engine
        return (eng.cfg["H"], eng.cfg["W"]) if eng is not None and eng.cfg is not None else None


class ResNet(ResNet_latefusion):
    """models.py:233-303: ONE ResNet-18 encoder over all input channels (rgb: 3, rgbd: 4, d: 1), conv2/bn2, decoder, head.
    Same engine and kernels as the late-fusion network (no depth branch, no fusion convolution); parameter / buffer names
    and registration order are the reference's.  The parent class only lends forward() and the engine plumbing."""

    def __init__(self, layers, decoder, output_size, in_channels=3, pretrained=True):
        if layers not in _LAYER_CHOICES:
            raise RuntimeError("Only 18, 34, 50, 101, and 152 layer model are defined for ResNet. Got {}".format(layers))
        nn.Module.__init__(self)
        if layers != 18:
            raise NotImplementedError("only the ResNet-18 encoder is built for B200")
        if not 1 <= in_channels <= 4:
            raise NotImplementedError("the B200 stem packs at most 4 input channels")
        self.output_size = output_size
        self.in_channels = in_channels
        self.decoder_name = decoder
        self._arch = "resnet"
        self.conv1 = nn.Conv2d(in_channels, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        if in_channels != 3:                          # models.py:243-247 (3 channels: torchvision's own conv1 / bn1)
            weights_init(self.conv1)
            weights_init(self.bn1)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = _make_layer(64, 64, 2, 1)
        self.layer2 = _make_layer(64, 128, 2, 2)
        self.layer3 = _make_layer(128, 256, 2, 2)
        self.layer4 = _make_layer(256, 512, 2, 2)
        self.conv2 = nn.Conv2d(512, 256, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(256)
        self.decoder = choose_decoder(decoder, 256)
        self.conv3 = nn.Conv2d(16, 1, 3, 1, 1, bias=False)
        self.bilinear = nn.Upsample(size=self.output_size, mode="bilinear", align_corners=True)
        self.conv2.apply(weights_init)
        self.bn2.apply(weights_init)
        self.decoder.apply(weights_init)
        self.conv3.apply(weights_init)
        if pretrained:
            import torchvision
            tv = torchvision.models.resnet18(weights=torchvision.models.ResNet18_Weights.IMAGENET1K_V1)
            names = ("layer1", "layer2", "layer3", "layer4") + (("conv1", "bn1") if in_channels == 3 else ())
            for name in names:
                getattr(self, name).load_state_dict(getattr(tv, name).state_dict())
        self.precision = _precision_from_env()
        self._engine = None
        self._anchor = None
        self._fwd_serial = 0
