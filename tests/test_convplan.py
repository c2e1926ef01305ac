"""Host-side planning logic (no GPU): the GConv descriptions reproduce torch's convolutions exactly, their
transposes reproduce autograd's data gradients, and the weight-gradient reference reproduces autograd's
weight gradients.  These are the semantics the tcgen05 programs are derived from."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from radar_depth_b200 import convplan as cp
from radar_depth_b200 import _lib


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("k,stride,pad,H,W", [(3, 1, 1, 9, 13), (3, 2, 1, 10, 14), (3, 2, 1, 9, 13), (1, 2, 0, 9, 12),
                                              (1, 1, 0, 5, 7), (5, 1, 2, 8, 9), (7, 2, 3, 12, 16)])
def test_standard_conv_forward_dgrad_wgrad(k, stride, pad, H, W):
    torch.manual_seed(0)
    Cin, Cout, B = 16, 32, 2
    w = torch.randn(Cout, Cin, k, k, dtype=torch.float64, requires_grad=True)
    x = torch.randn(B, Cin, H, W, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x, w, None, stride, pad)
    dy = torch.randn_like(y)
    y.backward(dy)
    g = cp.gconv_standard(0, Cout, Cin, k, stride, pad)
    wflat = w.detach().reshape(-1)
    out = cp.gconv_reference(g, _nhwc(x.detach()), wflat, y.shape[-2:])
    assert torch.allclose(_nchw(out), y.detach(), atol=1e-10)
    dx = cp.gconv_reference(g.transposed(), _nhwc(dy), wflat, (H, W))
    assert torch.allclose(_nchw(dx), x.grad, atol=1e-10)
    dw = cp.gconv_wgrad_reference(g, _nhwc(x.detach()), _nhwc(dy), wflat.numel())
    assert torch.allclose(dw.reshape(w.shape), w.grad, atol=1e-9)


@pytest.mark.parametrize("cin_depth,H,W", [(1, 16, 24), (2, 16, 24), (1, 18, 22)])
def test_stem_matches_two_7x7_stride2_convs(cin_depth, H, W):
    torch.manual_seed(1)
    B, C = 2, 3 + cin_depth
    w_rgb = torch.randn(64, 3, 7, 7, dtype=torch.float64, requires_grad=True)
    w_d = torch.randn(16, cin_depth, 7, 7, dtype=torch.float64, requires_grad=True)
    x = torch.randn(B, C, H, W, dtype=torch.float64, requires_grad=True)
    y = torch.cat([F.conv2d(x[:, :3], w_rgb, None, 2, 3), F.conv2d(x[:, 3:], w_d, None, 2, 3)], 1)
    dy = torch.randn_like(y)
    y.backward(dy)
    wflat = torch.cat([w_rgb.detach().reshape(-1), w_d.detach().reshape(-1)])
    g = cp.gconv_stem(0, w_rgb.numel(), cin_depth)
    Cs = g.Cx // 4
    # space-to-depth input, channel = parity*Cs + c (what rd_input_pack writes)
    xs = torch.zeros(B, H // 2, W // 2, g.Cx, dtype=torch.float64)
    for py in range(2):
        for px in range(2):
            xs[..., (py * 2 + px) * Cs:(py * 2 + px) * Cs + C] = x.detach()[:, :, py::2, px::2].permute(0, 2, 3, 1)
    out = cp.gconv_reference(g, xs, wflat, y.shape[-2:])
    assert torch.allclose(_nchw(out), y.detach(), atol=1e-10)
    dw = cp.gconv_wgrad_reference(g, xs, _nhwc(dy), wflat.numel())
    assert torch.allclose(dw[:w_rgb.numel()].reshape(w_rgb.shape), w_rgb.grad, atol=1e-9)
    assert torch.allclose(dw[w_rgb.numel():].reshape(w_d.shape), w_d.grad, atol=1e-9)
    # data gradient in space-to-depth form
    dxs = cp.gconv_reference(g.transposed(), _nhwc(dy), wflat, (H // 2, W // 2))
    for py in range(2):
        for px in range(2):
            got = dxs[..., (py * 2 + px) * Cs:(py * 2 + px) * Cs + C].permute(0, 3, 1, 2)
            assert torch.allclose(got, x.grad[:, :, py::2, px::2], atol=1e-10)


@pytest.mark.parametrize("H,W", [(5, 7), (4, 6)])
def test_upproj_subpixel_equals_5x5_on_unpooled(H, W):
    torch.manual_seed(2)
    B, Cin, Cout = 2, 32, 16
    wu = torch.randn(Cout, Cin, 5, 5, dtype=torch.float64, requires_grad=True)
    wb = torch.randn(Cout, Cin, 5, 5, dtype=torch.float64, requires_grad=True)
    x = torch.randn(B, Cin, H, W, dtype=torch.float64, requires_grad=True)
    u = torch.zeros(B, Cin, 2 * H, 2 * W, dtype=torch.float64)
    u = u.clone()
    u[:, :, ::2, ::2] = x            # Unpool, models.py:13-27
    y = torch.cat([F.conv2d(u, wu, None, 1, 2), F.conv2d(u, wb, None, 1, 2)], 1)
    dy = torch.randn_like(y)
    y.backward(dy)
    wflat = torch.cat([wu.detach().reshape(-1), wb.detach().reshape(-1)])
    g = cp.gconv_upproj(0, wu.numel(), Cin, Cout)
    assert len(g.taps) == 25
    out = cp.gconv_reference(g, _nhwc(x.detach()), wflat, (2 * H, 2 * W))
    assert torch.allclose(_nchw(out), y.detach(), atol=1e-10)
    dx = cp.gconv_reference(g.transposed(), _nhwc(dy), wflat, (H, W))
    assert torch.allclose(_nchw(dx), x.grad, atol=1e-10)
    dw = cp.gconv_wgrad_reference(g, _nhwc(x.detach()), _nhwc(dy), wflat.numel())
    assert torch.allclose(dw[:wu.numel()].reshape(wu.shape), wu.grad, atol=1e-9)
    assert torch.allclose(dw[wu.numel():].reshape(wb.shape), wb.grad, atol=1e-9)


@pytest.mark.parametrize("dtype", [_lib.RD_BF16, _lib.RD_F32])
def test_plans_are_self_consistent(dtype):
    """Every hot-path conv shape at 352x1216 plans within TMEM / smem limits."""
    shapes = [("l1", cp.gconv_standard(0, 64, 64, 3, 1, 1), (88, 304), (88, 304)),
              ("l2.0", cp.gconv_standard(0, 128, 64, 3, 2, 1), (88, 304), (44, 152)),
              ("l4", cp.gconv_standard(0, 512, 512, 3, 1, 1), (11, 38), (11, 38)),
              ("ds", cp.gconv_standard(0, 256, 128, 1, 2, 0), (44, 152), (22, 76)),
              ("fusion", cp.gconv_standard(0, 512, 640, 1, 1, 0), (11, 38), (11, 38)),
              ("d16", cp.gconv_standard(0, 16, 16, 3, 1, 1), (88, 304), (88, 304)),
              ("stem", cp.gconv_stem(0, 64 * 3 * 49, 1), (176, 608), (176, 608)),
              ("up1", cp.gconv_upproj(0, 128 * 256 * 25, 256, 128), (11, 38), (22, 76)),
              ("up4", cp.gconv_upproj(0, 16 * 32 * 25, 32, 16), (88, 304), (176, 608))]
    for name, g, src, dst in shapes:
        for gg, s, d in ((g, src, dst), (g.transposed(), dst, src)):
            pl = cp.plan_fprop(gg, 2, s, d, dtype)
            p = pl.params
            assert p.P * p.MB * p.N <= 512, name
            assert cp.FPROP_HEADER + p.IS * p.istage_bytes + p.WS * p.wstage_bytes <= cp.SMEM_BUDGET, name
            assert pl.pack_idx.size == pl.wpk_elems
            mx = max(p.taps[i].a_shift for i in range(p.ntaps))
            assert mx + p.MB * 128 <= p.S * p.S * p.plane_slots, name
        wp = cp.plan_wgrad(g, 2, src, dst, dtype)
        q = wp.params
        assert q.tg_size * q.Nc <= 512 and cp.WGRAD_HEADER + q.NS * q.stage_bytes + wp.info['geo']['pad'] <= cp.SMEM_BUDGET, name
        assert len(set(wp.scatter[0].tolist())) == wp.scatter[0].size       # each parameter has one source


def test_wgrad_planner_folds_tap_rows_of_16_channel_sources():
    """rd_wgrad_params.fold_rows/fold_len: 16-channel stride-1 sources whose taps form full rows of <= 4 adjacent taps
    (3x3 convs of the depth branch / decoder layer 4, the fused 4x4 stem) get one N=32 UMMA per tap row and chunk."""
    g = cp.gconv_standard(0, 16, 16, 3, 1, 1)
    w = cp.plan_wgrad(g, 2, (24, 40), (24, 40), gcopy=False)
    assert (w.params.fold_rows, w.params.fold_len) == (3, 3) and w.params.Nc == 16 and w.params.ntg == 1
    taps = [w.params.taps[i] for i in range(w.params.ntaps)]
    for r in range(3):
        for i in range(3):
            assert taps[3 * r + i].x_shift == taps[3 * r].x_shift + i and taps[3 * r + i].g_off == taps[3 * r].g_off
        # the junk 4th tap of a row still reads inside the staged source plane
        assert taps[3 * r].x_shift + 3 + w.params.KS <= w.params.x_plane_slots
    s = cp.gconv_stem(0, 10 ** 6, 1)
    ws = cp.plan_wgrad(s, 2, (32, 48), (32, 48))
    assert (ws.params.fold_rows, ws.params.fold_len) == (4, 4)
    assert ws.params.gcopies == 0                       # 80 output channels: no room for a second copy in M
    # wider sources, strided sources and the fp32 parity mode are not folded
    assert cp.plan_wgrad(cp.gconv_standard(0, 32, 32, 3, 1, 1), 2, (24, 40), (24, 40)).params.fold_len == 0
    assert cp.plan_wgrad(cp.gconv_standard(0, 32, 16, 3, 2, 1), 2, (24, 40), (12, 20)).params.fold_len == 0
    assert cp.plan_wgrad(g, 2, (24, 40), (24, 40), _lib.RD_F32).params.fold_len == 0


def test_wgrad_planner_stacks_gradient_copies_in_m():
    """rd_wgrad_params.gcopies: stride-1 convolutions with Cout <= 64 stage the gradient tile R = min(128/Cout, rows) times,
    copy r shifted r rows down, so that one UMMA covers R taps; every tap is produced exactly once."""
    for co, ci, R, njobs, fold in ((64, 64, 2, 6, 0), (32, 32, 3, 3, 0), (16, 16, 3, 1, 3), (16, 32, 3, 3, 0)):
        w = cp.plan_wgrad(cp.gconv_standard(0, co, ci, 3, 1, 1), 2, (24, 40), (24, 40), use_tuned=False)
        q = w.params
        assert (q.gcopies, q.njobs, q.fold_len) == (R, njobs, fold), (co, ci, w.info)
        assert q.gcopies * q.Mc <= 128 and q.tile_oy == R - 1 and q.tiles_y * q.Ht >= 24 + q.tile_oy
        assert q.KS == q.Ht * q.Wl                                       # copies are TMA boxes
        assert q.g_bytes >= R * (q.Mc // 8) * q.KS * 16
        produced = []
        for j in range(q.njobs):
            sy, sx = divmod(q.taps[j].x_shift, q.Wl)                     # source shift of copy 0, relative to (sy_min, sx_min)
            for r in range(R):
                t = q.job_tap[j][r]
                if t < 0:
                    assert sy - r < 0                                    # only rows above the first tap row are dropped
                    continue
                taps_of = [t + i for i in range(fold)] if fold else [t]
                assert t == (sy - r) * 3 + (0 if fold else sx)           # tap (sy - r, sx): copy r is the tile r rows further down
                produced += taps_of
        assert sorted(produced) == list(range(9)), (co, ci, produced)
    # strided convolutions, 1x1 convolutions, wide layers, the fp32 parity mode and RD_WGRAD_GCOPY=0 keep one accumulator per tap
    assert cp.plan_wgrad(cp.gconv_standard(0, 32, 16, 3, 2, 1), 2, (24, 40), (12, 20)).params.gcopies == 0
    assert cp.plan_wgrad(cp.gconv_standard(0, 64, 64, 1, 1, 0), 2, (24, 40), (24, 40)).params.gcopies == 0
    assert cp.plan_wgrad(cp.gconv_standard(0, 128, 128, 3, 1, 1), 2, (24, 40), (24, 40)).params.gcopies == 0
    assert cp.plan_wgrad(cp.gconv_standard(0, 64, 64, 3, 1, 1), 2, (24, 40), (24, 40), _lib.RD_F32).params.gcopies == 0
    assert cp.plan_wgrad(cp.gconv_standard(0, 64, 64, 3, 1, 1), 2, (24, 40), (24, 40), gcopy=False).params.gcopies == 0


def test_planner_stages_one_parity_plane_for_1x1_stride2():
    g = cp.gconv_standard(0, 128, 64, 1, 2, 0)
    assert cp.plan_fprop(g, 2, (24, 40), (12, 20)).params.src_planes == 1
    assert cp.plan_wgrad(g, 2, (24, 40), (12, 20)).params.x_planes == 1
    g3 = cp.gconv_standard(0, 128, 64, 3, 2, 1)
    assert cp.plan_fprop(g3, 2, (24, 40), (12, 20)).params.src_planes == 0
    assert cp.plan_wgrad(g3, 2, (24, 40), (12, 20)).params.x_planes == 0


def test_wgrad_planner_prefers_tiles_that_are_one_dense_tma_box():
    """KS == Ht*Wl (no tail in a chunk plane) is what lets rd_conv_wgrad stage the gradient tile with one TMA box."""
    for co, ci, hw in ((128, 128, (44, 152)), (256, 256, (22, 76)), (512, 512, (11, 38))):
        w = cp.plan_wgrad(cp.gconv_standard(0, co, ci, 3, 1, 1), 16, hw, hw, use_tuned=False)
        assert w.params.KS == w.params.Ht * w.params.Wl and w.params.KS % 16 == 0, (co, w.info)


def test_tuned_lookup_is_keyed_by_the_sm_budget_of_the_launch(monkeypatch):
    """The measured tile table holds entries per SM budget (the encoder chains run side by side on disjoint SM sets):
    a launch gets the entry measured with exactly its budget, the whole-GPU entry when its budget is at least half the
    GPU, and NO entry (cost model) when it owns a few SMs only and was never measured there."""
    tab = {"k": {"Ht": 1}, "k|sm128": {"Ht": 2}, "k|bn": {"Ht": 3}, "k|sm128|bn": {"Ht": 4}, "j": {"Ht": 5}}
    monkeypatch.setattr(cp, "_TUNED", tab, raising=False)
    monkeypatch.setattr(cp, "tuned_table", lambda: tab)
    assert cp.tuned_lookup("k", cp.NUM_SMS) == {"Ht": 1}
    assert cp.tuned_lookup("k", 128) == {"Ht": 2}
    assert cp.tuned_lookup("k", 128, "|bn") == {"Ht": 4}
    assert cp.tuned_lookup("k", cp.NUM_SMS, "|bn") == {"Ht": 3}
    assert cp.tuned_lookup("j", 128) == {"Ht": 5}            # no entry for this budget: whole-GPU entry (>= half the GPU)
    assert cp.tuned_lookup("j", 20) is None                  # a small lane without its own entry: cost model
    assert cp.tuned_lookup("missing", 128) is None


def test_wgrad_plans_with_and_without_the_source_transform_share_the_dw_layout():
    """engine.emit_wgrad swaps in the '|bn' blocking for launches that transform their source tile: the weight-gradient
    buffer layout (taps x Cout x Cin) and the scatter table must not depend on the blocking."""
    g = cp.gconv_standard(0, 128, 128, 3, 1, 1)
    a = cp.plan_wgrad(g, 16, (44, 152), (44, 152), use_tuned=False, nc=128, ks_target=256)
    b = cp.plan_wgrad(g, 16, (44, 152), (44, 152), use_tuned=False, nc=64, ks_target=512, sm_budget=128)
    assert a.dw_elems == b.dw_elems
    assert all((x == y).all() for x, y in zip(a.scatter, b.scatter))
    assert b.params.max_ctas * b.params.ncob * b.params.ncib * b.params.ntg <= 128


def test_compact_pack_table_reproduces_the_index_table():
    """rd_pack_weights_g8's (base, stride) groups + fallback rows must describe exactly the gather of the per-element table:
    checked on real plans (standard, stride-2 data gradient, stem with zero-padded taps, fused UpProj) in both precisions."""
    tables = []
    for act in (_lib.RD_BF16, _lib.RD_F32):
        for g, s_hw, d_hw in ((cp.gconv_standard(0, 64, 64, 3, 1, 1), (20, 37), (20, 37)),
                              (cp.gconv_standard(0, 128, 64, 3, 2, 1).transposed(), (11, 19), (22, 38)),
                              (cp.gconv_stem(0, 64 * 3 * 49, 1), (24, 40), (24, 40)),
                              (cp.gconv_upproj(0, 16 * 32 * 25, 32, 16), (20, 33), (40, 66))):
            tables.append(cp.plan_fprop(g, 2, s_hw, d_hw, act, use_tuned=False).pack_idx)
    idx = np.concatenate(tables)
    grp, fb = cp.compact_pack_table(idx)
    assert grp.shape == (idx.size // 8, 2) and fb.shape[1] == 8
    rebuilt = np.empty((grp.shape[0], 8), dtype=np.int64)
    ap = grp[:, 1] >= 0
    zero = ap & (grp[:, 0] < 0)
    flag = grp[:, 0].astype(np.int64) & 0x40000000
    base = grp[:, 0].astype(np.int64) & 0x3FFFFFFF
    rebuilt[ap] = (base[ap, None] + np.arange(8)[None, :] * grp[ap, 1].astype(np.int64)[:, None]) | flag[ap, None]
    rebuilt[zero] = -1
    rebuilt[~ap] = fb[grp[~ap, 0]]
    assert (rebuilt.reshape(-1) == idx.astype(np.int64)).all()
    assert (~ap).mean() < 0.2 and ap.mean() > 0.8            # the compact form is the common case
