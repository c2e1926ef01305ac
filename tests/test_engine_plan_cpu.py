"""The whole launch program of the engine can be planned without a GPU (buffers on the CPU, no launches): every
convolution of ResNet_latefusion gets a feasible forward / data-gradient / weight-gradient program for the shapes the
tests and BASELINE.json configs use, in both precisions."""
import pytest

from radar_depth_b200 import _lib
from radar_depth_b200.engine import LatefusionEngine
from radar_depth_b200.model.models import ResNet_latefusion


@pytest.mark.parametrize("cin,hw,b,act", [(4, (64, 96), 2, _lib.RD_F32), (5, (64, 96), 2, _lib.RD_F32), (4, (90, 160), 2, _lib.RD_F32),
                                          (4, (64, 96), 1, _lib.RD_BF16), (4, (352, 1216), 2, _lib.RD_BF16),
                                          (5, (128, 192), 2, _lib.RD_BF16)])
def test_engine_plans_every_program(cin, hw, b, act):
    m = ResNet_latefusion(18, "upproj", hw, cin, pretrained=False)
    eng = LatefusionEngine(m, cin, hw, act)
    eng.adopt("cpu")
    eng.configure(b, *hw)
    assert len(eng.fwd) == 74 and len(eng.fwd_eval) == 75      # training: BN finalisation rides in the conv tails; eval: one batched launch
    assert len(eng.bwd) == (175 if cin > 4 else 174)
    names = [L.name for L in eng.fwd]
    assert names[0] == "pack_weights" and names[-1] == "bilinear"
    assert sum(n.startswith("conv_f:") for n in names) == 49          # 55 reference convs: two stems fused (-1), 4x2 5x5s fused (-4), conv3 in the head kernel (-1)
    assert sum(L.name.startswith("wgrad:") for L in eng.bwd) == 49
    # every trainable tensor is reachable from the scatter table or written directly (BN affine, conv3)
    covered = (eng.unpack_idx >= 0).sum().item()
    n_conv = sum(p.numel() for n, p in m.named_parameters() if p.dim() == 4 and n != "conv3.weight")
    assert covered == n_conv
