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


def test_graph_cut_programs_partition_the_forward():
    """pnp_forward_front / pnp_forward_rear (reference models.py:669-707): front + rear cover the forward program exactly
    once (plus the feature export / import at the cut), in both BatchNorm modes; the rear's backward is the prefix of the
    backward program that ends with the data gradient of decoder.layer1's 5x5 pair."""
    m = ResNet_latefusion(18, "upproj", (64, 96), 4, pretrained=False)
    eng = LatefusionEngine(m, 4, (64, 96), _lib.RD_F32)
    eng.adopt("cpu")
    eng.configure(2, 64, 96)
    f, r = [L.name for L in eng.front], [L.name for L in eng.rear]
    assert f[-1] == "feature_export" and f[-2] == "conv_f:conv2"
    assert r[:2] == ["pack_weights", "feature_import"] and r[2] == "conv_f:decoder.layer1.up5x5" and r[-1] == "bilinear"
    assert f[:-1] + r[2:] == [L.name for L in eng.fwd]
    fe, re_ = [L.name for L in eng.front_eval], [L.name for L in eng.rear_eval]
    assert fe[-1] == "feature_export" and fe[-2] == "conv_f:conv2(eval)"
    assert re_[:3] == ["pack_weights", "bn_fin_eval_all", "feature_import"]
    assert fe[:-1] + re_[3:] == [L.name for L in eng.fwd_eval]
    rb = [L.name for L in eng.rear_bwd]
    assert rb[0] == "bilinear_bwd" and rb[-2] == "conv_d:decoder.layer1.up5x5" and rb[-1] == "feature_export(grad)"
    assert rb[:-1] == [L.name for L in eng.bwd[:len(rb) - 1]]
    assert eng.bneck.shape == (2, 256, 2, 3)


def test_eval_mode_backward_switches_every_batchnorm_job():
    """Eval-mode BatchNorm backward = the training formula with an infinite batch count (the two batch-mean terms
    vanish); every one of the 62 BatchNorm layers has exactly one backward job and the switch is reversible."""
    m = ResNet_latefusion(18, "upproj", (64, 96), 4, pretrained=False)
    eng = LatefusionEngine(m, 4, (64, 96), _lib.RD_F32)
    eng.adopt("cpu")
    eng.configure(2, 64, 96)
    n_bn = sum(1 for mod in m.modules() if mod.__class__.__name__ == "BatchNorm2d")
    assert len(eng._bwd_jobs) == n_bn
    counts = [float(j.count) for j, _ in eng._bwd_jobs]
    assert all(c == c0 and c > 0 and c != float("inf") for c, (_, c0) in zip(counts, eng._bwd_jobs))
    eng._bn_backward_mode(False)
    assert all(float(j.count) == float("inf") for j, _ in eng._bwd_jobs)
    eng._bn_backward_mode(True)
    assert [float(j.count) for j, _ in eng._bwd_jobs] == counts


@pytest.mark.parametrize("cin,b", [(4, 16), (5, 8)])
def test_encoder_lanes_own_disjoint_sm_sets(cin, b):
    """Between the stem and the fusion convolution the depth encoder (lane 1) and the RGB encoder (lane 0) run at the same
    time; every launch of a lane is planned for that lane's CTA budget, the budgets add up to the GPU, and the launches
    outside the two-lane region may use all of it.  (A conv CTA owns its SM: > half the shared memory + all TMEM columns.)"""
    from radar_depth_b200 import convplan as cp
    hw = (352, 1216)
    m = ResNet_latefusion(18, "upproj", hw, cin, pretrained=False)
    eng = LatefusionEngine(m, cin, hw, _lib.RD_BF16)
    eng.det = False
    eng.adopt("cpu")
    eng.configure(b, *hw)
    assert eng._par and eng._depth_sms == 20
    budgets = {rec["wargs"][3] for rec in eng.convs}
    assert budgets == {cp.NUM_SMS, cp.NUM_SMS - 20, 20}
    for rec in eng.convs:
        sms = rec["wargs"][3]
        plans = [rec["fplan"]] + ([rec["dplan"]] if rec["dplan"] is not None else [])
        for pl in plans:
            ctas = min(int(pl.params.max_ctas), pl.ntiles) * int(pl.params.nblk)
            assert ctas <= sms, (rec["name"], ctas, sms)
        for key in ("wplan", "wplan_bn"):
            w = rec.get(key)
            if w is not None:
                q = w.params
                assert int(q.max_ctas) * int(q.ncob) * int(q.ncib) * int(q.ntg) <= sms, (rec["name"], key, sms)
    # the launch programs: one fork and one join per direction, lane-1 launches only between them
    for prog in (eng.fwd, eng.bwd):
        syncs = [L.sync for L in prog if L.sync]
        assert syncs == ["fork", "join"], syncs
        inside = False
        for L in prog:
            if L.sync == "fork":
                inside = True
            elif L.sync == "join":
                inside = False
            assert L.lane == 0 or inside, L.name
