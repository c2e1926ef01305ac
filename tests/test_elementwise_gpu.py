"""GPU parity of the HBM-bound kernels (through the C ABI) against plain torch fp32/fp64 references of the same ops."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from radar_depth_b200 import _lib, ops
from radar_depth_b200._lib import View, call
from radar_depth_b200.ops import ptr, stream_ptr, view

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False          # the torch references below must be true fp32
torch.backends.cuda.matmul.allow_tf32 = False
NULLV = View(None, 0, 0)
ACTS = [("bf16", _lib.RD_BF16, torch.bfloat16), ("f32", _lib.RD_F32, torch.float32)]


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def tol(act):
    return dict(rtol=2e-2, atol=2e-2) if act == _lib.RD_BF16 else dict(rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("name,act,td", ACTS)
@pytest.mark.parametrize("C,H,W", [(4, 12, 16), (5, 13, 17)])
def test_input_pack(name, act, td, C, H, W):
    x = torch.randn(2, C, H, W, device="cuda")
    Cs = 4 if C <= 4 else 8
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    out = torch.full((2, H2, W2, 4 * Cs), float("nan"), device="cuda", dtype=td)
    call("rd_input_pack", ptr(x), ptr(out), 2, C, H, W, Cs, act, stream_ptr())
    ref = torch.zeros(2, H2, W2, 4 * Cs, device="cuda")
    for py in range(2):
        for px in range(2):
            sub = x[:, :, py::2, px::2]
            ref[:, :sub.shape[2], :sub.shape[3], (py * 2 + px) * Cs:(py * 2 + px) * Cs + C] = sub.permute(0, 2, 3, 1)
    torch.testing.assert_close(out.float(), ref.to(td).float(), rtol=0, atol=0)


@pytest.mark.parametrize("name,act,td", ACTS)
@pytest.mark.parametrize("H,W", [(12, 16), (13, 17)])
def test_input_pack_parts_equals_input_pack_of_the_concatenation(name, act, td, H, W):
    """multistage_model.py:78: cat((x[:, :3], radar_filtered, depth_stage1), 1) packed straight from its three sources (a
    channel slice of a 4-channel tensor, two 1-channel tensors), and the gradient slice that flows back into depth_stage1."""
    import ctypes as C
    x4 = torch.randn(2, 4, H, W, device="cuda")
    parts = [x4[:, :3], torch.randn(2, 1, H, W, device="cuda"), torch.randn(2, 1, H, W, device="cuda")]
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    ref = torch.full((2, H2, W2, 32), float("nan"), device="cuda", dtype=td)
    call("rd_input_pack", ptr(torch.cat(parts, 1).contiguous()), ptr(ref), 2, 5, H, W, 8, act, stream_ptr())
    planes, strides = [], []
    for t in parts:
        for c in range(t.shape[1]):
            planes.append(t.data_ptr() + 4 * c * t.stride(1))
            strides.append(t.stride(0))
    out = torch.full((2, H2, W2, 32), float("nan"), device="cuda", dtype=td)
    call("rd_input_pack_parts", (C.c_void_p * 5)(*planes), (C.c_longlong * 5)(*strides), ptr(out), 2, 5, H, W, 8, act, stream_ptr())
    torch.testing.assert_close(out.float(), ref.float(), rtol=0, atol=0)
    # channel 4 of a gradient in the same space-to-depth layout
    g = torch.randn(2, H2, W2, 32, device="cuda").to(td)
    dc = torch.empty(2, 1, H, W, device="cuda")
    call("rd_input_grad_channel", ptr(g), ptr(dc), 2, H, W, 8, 4, act, stream_ptr())
    full = g.float().view(2, H2, W2, 2, 2, 8).permute(0, 5, 1, 3, 2, 4).reshape(2, 8, 2 * H2, 2 * W2)[:, 4:5, :H, :W]
    torch.testing.assert_close(dc, full.contiguous(), rtol=0, atol=0)


def test_bn_finalize_train_and_eval_match_torch_batchnorm():
    C, n = 48, 2 * 7 * 9
    x = torch.randn(2, C, 7, 9, device="cuda", dtype=torch.float64) * 2 + 1
    bn = torch.nn.BatchNorm2d(C).cuda().double()
    bn.weight.data.uniform_(0.5, 1.5)
    bn.bias.data.normal_()
    bn.running_mean.normal_()
    bn.running_var.uniform_(0.5, 2.0)
    rm, rv = bn.running_mean.clone().float(), bn.running_var.clone().float()
    nbt = torch.tensor(3, device="cuda", dtype=torch.int64)
    y_ref = bn(x)
    s = torch.stack([x.sum(dim=(0, 2, 3)), (x * x).sum(dim=(0, 2, 3))]).contiguous()
    gamma, beta = bn.weight.detach().float().contiguous(), bn.bias.detach().float().contiguous()
    vec = torch.zeros(4, C, device="cuda")
    call("rd_bn_finalize", ptr(s[0]), ptr(s[1]), float(n), ptr(gamma), ptr(beta), ptr(rm), ptr(rv), ptr(nbt), C, 1, 0.1, 1e-5,
         ptr(vec[0]), ptr(vec[1]), ptr(vec[2]), ptr(vec[3]), stream_ptr())
    y = x.float() * vec[0][None, :, None, None] + vec[1][None, :, None, None]
    torch.testing.assert_close(y, y_ref.float(), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(rm, bn.running_mean.float(), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(rv, bn.running_var.float(), rtol=1e-6, atol=1e-6)
    assert int(nbt) == 4
    bn.eval()
    y_ref = bn(x)
    call("rd_bn_finalize", None, None, float(n), ptr(gamma), ptr(beta), ptr(rm), ptr(rv), None, C, 0, 0.1, 1e-5,
         ptr(vec[0]), ptr(vec[1]), ptr(vec[2]), ptr(vec[3]), stream_ptr())
    y = x.float() * vec[0][None, :, None, None] + vec[1][None, :, None, None]
    torch.testing.assert_close(y, y_ref.float(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name,act,td", ACTS)
@pytest.mark.parametrize("with_ds", [False, True])
def test_residual_join_forward_backward_vs_autograd(name, act, td, with_ds):
    """bn_add_act + join_bwd + bn_bwd_finalize + bn_bwd_apply == autograd of relu(bn(z) + identity)."""
    torch.manual_seed(0)
    B, C, H, W = 2, 32, 6, 10
    n = B * H * W
    z = torch.randn(B, H, W, C, device="cuda").to(td)
    idt = torch.randn(B, H, W, C, device="cuda").to(td)
    dout = torch.randn(B, H, W, C, device="cuda").to(td)
    bn_a, bn_b = torch.nn.BatchNorm2d(C).cuda(), torch.nn.BatchNorm2d(C).cuda()
    for bn in (bn_a, bn_b):
        bn.weight.data.uniform_(0.5, 1.5)
        bn.bias.data.normal_(0, 0.3)
    zr = nchw(z.float()).requires_grad_(True)
    ir = nchw(idt.float()).requires_grad_(True)
    out_ref = F.relu(bn_a(zr) + (bn_b(ir) if with_ds else ir))
    out_ref.backward(nchw(dout.float()))

    def finalize(t, bn):
        tf = t.double()
        s = torch.stack([tf.sum(dim=(0, 1, 2)), (tf * tf).sum(dim=(0, 1, 2))]).contiguous()
        vec = torch.zeros(7, C, device="cuda")
        rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
        call("rd_bn_finalize", ptr(s[0]), ptr(s[1]), float(n), ptr(bn.weight.data), ptr(bn.bias.data), ptr(rm), ptr(rv), None, C, 1,
             0.1, 1e-5, ptr(vec[0]), ptr(vec[1]), ptr(vec[2]), ptr(vec[3]), stream_ptr())
        return vec

    va = finalize(z, bn_a)
    vb = finalize(idt, bn_b) if with_ds else None
    out = torch.empty_like(z)
    call("rd_bn_add_act", view(z), ptr(va[0]), ptr(va[1]), view(idt), ptr(vb[0]) if with_ds else None, ptr(vb[1]) if with_ds else None,
         view(out), n, C, 0.0, act, stream_ptr())
    torch.testing.assert_close(out.float(), nhwc(out_ref.detach()), **tol(act))
    g = torch.empty_like(z)
    st = torch.zeros(3, C, device="cuda", dtype=torch.float64)
    # the BatchNorm-backward finalisation rides in the tail of join_bwd (rd_bn_tail, last block): it must give exactly
    # what the stand-alone rd_bn_bwd_finalize computes from the same statistics
    import ctypes
    fused = torch.zeros(5, C, device="cuda")                  # dgamma dbeta A B C
    counter = torch.zeros(1, device="cuda", dtype=torch.float64)
    tail = _lib.BnTail()
    tail.counter, tail.njobs = ptr(counter), 1
    j = tail.job[0]
    j.kind, j.C, j.count = 2, C, float(n)
    j.sum_a, j.sum_b, j.gamma = ptr(st[0]), ptr(st[1]), ptr(bn_a.weight.data)
    j.v0, j.v1, j.v2, j.v3 = ptr(va[2]), ptr(va[3]), ptr(fused[0]), ptr(fused[1])
    j.cA, j.cB, j.cC = ptr(fused[2]), ptr(fused[3]), ptr(fused[4])
    call("rd_join_bwd", view(dout), view(out), view(z), view(idt) if with_ds else NULLV, view(g), n, C, 0.0, ptr(st[0]), ptr(st[1]),
         ptr(st[2]), ctypes.byref(tail), act, stream_ptr())
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    call("rd_bn_bwd_finalize", ptr(st[0]), ptr(st[1]), float(n), ptr(bn_a.weight.data), ptr(va[2]), ptr(va[3]), C, 1, ptr(dg), ptr(db),
         ptr(va[4]), ptr(va[5]), ptr(va[6]), stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(fused[0], dg) and torch.equal(fused[1], db)
    assert torch.equal(fused[2], va[4]) and torch.equal(fused[3], va[5]) and torch.equal(fused[4], va[6])
    dz = torch.empty_like(z)
    call("rd_bn_bwd_apply", view(g), view(z), view(dz), ptr(va[4]), ptr(va[5]), ptr(va[6]), n, C, act, stream_ptr())
    t = dict(rtol=3e-2, atol=3e-2) if act == _lib.RD_BF16 else dict(rtol=1e-4, atol=1e-5)
    # elements whose pre-activation is within rounding of 0 may flip the mask in bf16: compare in L2
    rel = (dz.float() - nhwc(zr.grad)).norm() / nhwc(zr.grad).norm()
    assert rel < (3e-2 if act == _lib.RD_BF16 else 1e-4), rel
    torch.testing.assert_close(dg, bn_a.weight.grad, **t)
    torch.testing.assert_close(db, bn_a.bias.grad, **t)
    if with_ds:
        dg2, db2 = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        call("rd_bn_bwd_finalize", ptr(st[0]), ptr(st[2]), float(n), ptr(bn_b.weight.data), ptr(vb[2]), ptr(vb[3]), C, 1, ptr(dg2),
             ptr(db2), ptr(vb[4]), ptr(vb[5]), ptr(vb[6]), stream_ptr())
        did = torch.empty_like(z)
        call("rd_bn_bwd_apply", view(g), view(idt), view(did), ptr(vb[4]), ptr(vb[5]), ptr(vb[6]), n, C, act, stream_ptr())
        rel = (did.float() - nhwc(ir.grad)).norm() / nhwc(ir.grad).norm()
        assert rel < (3e-2 if act == _lib.RD_BF16 else 1e-4), rel
        torch.testing.assert_close(dg2, bn_b.weight.grad, **t)
    else:
        rel = (g.float() - nhwc(ir.grad)).norm() / nhwc(ir.grad).norm()
        assert rel < (3e-2 if act == _lib.RD_BF16 else 1e-5), rel


@pytest.mark.parametrize("name,act,td", ACTS)
@pytest.mark.parametrize("H,W,ties", [(12, 16, False), (11, 15, False), (12, 16, True)])
def test_maxpool_forward_backward_vs_autograd(name, act, td, H, W, ties):
    """Both stems at once: channels [0,64) ReLU, [64,80) LeakyReLU(0.2); arg-max tie rule = first max in window order."""
    torch.manual_seed(1)
    B, C, split = 2, 80, 64
    z = torch.randn(B, H, W, C, device="cuda")
    if ties:   # mostly-constant map like the depth stem on a ~empty radar image
        z = torch.zeros(B, H, W, C, device="cuda")
        z[:, 3, 4] = 1.0
        z[:, 7, 9] = -2.0
    z = z.to(td)
    sc = (torch.rand(C, device="cuda") + 0.5) * torch.where(torch.rand(C, device="cuda") < 0.2, -1.0, 1.0)
    sh = torch.randn(C, device="cuda") * 0.3
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    oa = torch.empty(B, Ho, Wo, split, device="cuda", dtype=td)
    ob = torch.empty(B, Ho, Wo, C - split, device="cuda", dtype=td)
    amax = torch.zeros(B, Ho, Wo, C, device="cuda", dtype=torch.uint8)
    zarg = torch.zeros(B, Ho, Wo, C, device="cuda", dtype=torch.bfloat16) if act == _lib.RD_BF16 else None
    call("rd_maxpool_fwd", view(z), ptr(sc), ptr(sh), B, H, W, C, split, 0.0, 0.2, view(oa), view(ob), ptr(amax), Ho, Wo, ptr(zarg), act,
         stream_ptr())
    zr = nchw(z.float()).requires_grad_(True)
    y = zr * sc[None, :, None, None] + sh[None, :, None, None]
    y = torch.cat([F.relu(y[:, :split]), F.leaky_relu(y[:, split:], 0.2)], 1)
    if act == _lib.RD_BF16:
        y = y + (y.detach().bfloat16().float() - y.detach())      # straight-through rounding, like a stored activation
    pooled = F.max_pool2d(y.cpu(), 3, 2, 1).cuda() if False else F.max_pool2d(y, 3, 2, 1)
    got = torch.cat([oa, ob], dim=-1).float()
    torch.testing.assert_close(got, nhwc(pooled.detach()), **tol(act))
    dpa = torch.randn(B, Ho, Wo, split, device="cuda").to(td)
    dpb = torch.randn(B, Ho, Wo, C - split, device="cuda").to(td)
    # CPU autograd defines the tie rule the oracle follows
    zc = nchw(z.float()).cpu().requires_grad_(True)
    yc = zc * sc.cpu()[None, :, None, None] + sh.cpu()[None, :, None, None]
    yc = torch.cat([F.relu(yc[:, :split]), F.leaky_relu(yc[:, split:], 0.2)], 1)
    if act == _lib.RD_BF16:
        yc = yc + (yc.detach().bfloat16().float() - yc.detach())
    F.max_pool2d(yc, 3, 2, 1).backward(nchw(torch.cat([dpa, dpb], -1).float()).cpu())
    g = torch.empty(B, H, W, C, device="cuda", dtype=td)
    st = torch.zeros(2, C, device="cuda", dtype=torch.float64)
    call("rd_maxpool_bwd", view(dpa), view(dpb), ptr(amax), view(z), ptr(sc), ptr(sh), B, H, W, C, split, 0.0, 0.2, Ho, Wo, view(g),
         ptr(st[0]), ptr(st[1]), None, act, stream_ptr())
    gref = nhwc(zc.grad).cuda() / sc          # gradient w.r.t. the BN output's pre-activation y... undo the affine
    torch.testing.assert_close(g.float(), gref, **(dict(rtol=2e-2, atol=2e-2) if act == _lib.RD_BF16 else dict(rtol=1e-5, atol=1e-5)))
    # the statistics are accumulated from the fp32 values before the store rounds them (bf16: ~2^-9 per element)
    st_tol = dict(rtol=1e-3, atol=1e-3) if act == _lib.RD_F32 else dict(rtol=2e-2, atol=8e-2)
    torch.testing.assert_close(st[0].float(), g.float().sum(dim=(0, 1, 2)), **st_tol)
    torch.testing.assert_close(st[1].float(), (g.float() * z.float()).sum(dim=(0, 1, 2)), **st_tol)
    if act != _lib.RD_BF16:
        return
    # ---- the two-pass backward (rd_maxpool_bwd_stats + rd_maxpool_bwd_apply) against the path above
    # zarg = z at the arg-max position of every window (per channel)
    code = amax.long()
    dy, dx = code // 3, code % 3
    oy = torch.arange(Ho, device="cuda")[None, :, None, None]
    ox = torch.arange(Wo, device="cuda")[None, None, :, None]
    iy, ix = (2 * oy - 1 + dy).clamp(0, H - 1), (2 * ox - 1 + dx).clamp(0, W - 1)
    bi = torch.arange(B, device="cuda")[:, None, None, None].expand_as(iy)
    ci = torch.arange(C, device="cuda")[None, None, None, :].expand_as(iy)
    assert torch.equal(zarg, z[bi, iy, ix, ci])
    st2 = torch.zeros(2, C, device="cuda", dtype=torch.float64)
    call("rd_maxpool_bwd_stats", view(dpa), view(dpb), ptr(zarg), ptr(sc), ptr(sh), B, Ho, Wo, C, split, 0.0, 0.2, ptr(st2[0]), ptr(st2[1]),
         None, stream_ptr())
    # same products as the stem-resolution sums, added in another order (and from un-rounded per-window gradients)
    torch.testing.assert_close(st2.float(), st.float(), rtol=2e-2, atol=8e-2)
    cA = torch.rand(C, device="cuda") + 0.5
    cB = torch.randn(C, device="cuda") * 0.1
    cC = torch.randn(C, device="cuda") * 0.1
    dz = torch.empty(B, H, W, C, device="cuda", dtype=td)
    call("rd_maxpool_bwd_apply", view(dpa), view(dpb), ptr(amax), view(z), ptr(sc), ptr(sh), ptr(cA), ptr(cB), ptr(cC), B, H, W, C, split,
         0.0, 0.2, Ho, Wo, view(dz), stream_ptr())
    want = cA * gref + cB * z.float() + cC           # gref: exact fp32 gradient w.r.t. the BatchNorm output (CPU autograd)
    torch.testing.assert_close(dz.float(), want, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("name,act,td", ACTS)
def test_head_conv_and_bilinear_vs_autograd(name, act, td):
    torch.manual_seed(2)
    B, H, W, OH, OW = 2, 11, 19, 23, 37
    x = torch.randn(B, H, W, 16, device="cuda").to(td)
    w = (torch.randn(1, 16, 3, 3, device="cuda") * 0.2).contiguous()
    xr = nchw(x.float()).requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    c3_ref = F.conv2d(xr, wr, None, 1, 1)
    pred_ref = F.interpolate(c3_ref, size=(OH, OW), mode="bilinear", align_corners=True)
    dpred = torch.randn(B, 1, OH, OW, device="cuda")
    pred_ref.backward(dpred)
    c3 = torch.empty(B, H, W, device="cuda")
    pred = torch.empty(B, 1, OH, OW, device="cuda")
    call("rd_head_conv_fwd", view(x), ptr(w), B, H, W, ptr(c3), act, stream_ptr())
    call("rd_bilinear_fwd", ptr(c3), B, H, W, ptr(pred), OH, OW, stream_ptr())
    torch.testing.assert_close(pred, pred_ref.detach(), rtol=1e-4, atol=1e-4)
    dc3 = torch.empty(B, H, W, device="cuda")
    call("rd_bilinear_bwd", ptr(dpred), B, H, W, ptr(dc3), OH, OW, stream_ptr())
    dx = torch.empty_like(x)
    dw = torch.zeros(144, device="cuda")
    call("rd_head_conv_bwd", ptr(dc3), view(x), ptr(w), B, H, W, view(dx), ptr(dw), act, stream_ptr())
    torch.testing.assert_close(dx.float(), nhwc(xr.grad), **(dict(rtol=2e-2, atol=2e-2) if act == _lib.RD_BF16 else dict(rtol=1e-4, atol=1e-5)))
    torch.testing.assert_close(dw.view(1, 16, 3, 3), wr.grad, rtol=1e-3, atol=1e-3)


@pytest.mark.parametrize("B,Hi,Wi,Ho,Wo", [(16, 176, 608, 352, 1216), (17, 225, 400, 450, 800)])
def test_bilinear_at_benchmark_batch_matches_interpolate(B, Hi, Wi, Ho, Wo):
    """Flat output indices of the headline configuration exceed the range in which a multiply-high division without a
    correction step is exact (n * d >= 2^32: the last column of the last ~114 rows of image 15 landed one row down).
    Forward and backward at B=16 352x1216 (and a W=800 output at B=17) against F.interpolate / autograd."""
    torch.manual_seed(7)
    c3 = torch.randn(B, 1, Hi, Wi, device="cuda").requires_grad_(True)
    ref = F.interpolate(c3, size=(Ho, Wo), mode="bilinear", align_corners=True)
    dpred = torch.randn(B, 1, Ho, Wo, device="cuda")
    ref.backward(dpred)
    pred = torch.full((B, 1, Ho, Wo), float("nan"), device="cuda")
    call("rd_bilinear_fwd", ptr(c3.detach()), B, Hi, Wi, ptr(pred), Ho, Wo, stream_ptr())
    torch.testing.assert_close(pred, ref.detach(), rtol=1e-4, atol=1e-4)
    assert torch.equal(pred[-1, 0, -3:, -1], pred[-1, 0, -3:, -1]) and not torch.isnan(pred).any()
    dc3 = torch.empty(B, Hi, Wi, device="cuda")
    call("rd_bilinear_bwd", ptr(dpred), B, Hi, Wi, ptr(dc3), Ho, Wo, stream_ptr())
    torch.testing.assert_close(dc3.view_as(c3), c3.grad, rtol=1e-4, atol=1e-4)


def test_loss_sums_are_deterministic():
    """MaskedL1 / Smoothness reduce in a fixed order (determinism.py): identical bits on repeated calls at full size."""
    from radar_depth_b200.evaluation.criteria_new import MaskedL1Loss, SmoothnessLoss
    torch.manual_seed(11)
    pred = torch.rand(8, 1, 352, 1216, device="cuda") * 60 + 0.5
    tgt = torch.rand(8, 1, 352, 1216, device="cuda") * 80
    tgt[torch.rand_like(tgt) < 0.95] = 0
    img = torch.rand(8, 4, 352, 1216, device="cuda")
    vals = []
    for _ in range(3):
        p = pred.clone().requires_grad_(True)
        l1, sm = MaskedL1Loss()(p, tgt), SmoothnessLoss()(p, img)
        (l1 + 0.1 * sm).backward()
        vals.append((float(l1), float(sm), p.grad.clone()))
    for v in vals[1:]:
        assert v[0] == vals[0][0] and v[1] == vals[0][1] and torch.equal(v[2], vals[0][2])
    d = (tgt - pred)[tgt > 0].abs().double().mean()
    assert abs(vals[0][0] - float(d)) <= 1e-6 * float(d)


def test_masked_l1_forward_backward_vs_reference_formula():
    torch.manual_seed(3)
    pred = (torch.rand(2, 1, 24, 40, device="cuda") * 30).requires_grad_(True)
    tgt = torch.rand(2, 1, 24, 40, device="cuda") * 50
    tgt[torch.rand_like(tgt) < 0.7] = 0
    from radar_depth_b200.evaluation.criteria_new import MaskedL1Loss
    crit = MaskedL1Loss()
    loss = crit(pred, tgt)
    (loss * 3.0).backward()
    p2 = pred.detach().clone().requires_grad_(True)
    ref = (tgt - p2)[tgt > 0].abs().mean()            # criteria_new.py:50-53
    (ref * 3.0).backward()
    assert abs(float(loss) - float(ref)) < 1e-5 and crit.loss is loss
    torch.testing.assert_close(pred.grad, p2.grad, rtol=1e-5, atol=1e-9)
    with pytest.raises(AssertionError):
        crit(pred[0], tgt)


@pytest.mark.parametrize("name,act,td", ACTS)
def test_feature_export_import_roundtrip(name, act, td):
    """rd_feature_export / rd_feature_import (graph cut of pnp_forward_front / rear): NHWC slice of a wider buffer <-> NCHW
    fp32, with and without the per-channel affine; bit-exact in fp32, one bf16 rounding in bf16."""
    torch.manual_seed(5)
    B, H, W, Cc, pitch, coff = 2, 3, 5, 24, 40, 8
    buf = torch.randn(B, H, W, pitch, device="cuda").to(td)
    sc, sh = torch.rand(Cc, device="cuda") + 0.5, torch.randn(Cc, device="cuda")
    out = torch.full((B, Cc, H, W), float("nan"), device="cuda")
    call("rd_feature_export", view(buf, coff), ptr(sc), ptr(sh), ptr(out), B, H, W, Cc, act, stream_ptr())
    sl = buf[..., coff:coff + Cc].float()
    ref = torch.addcmul(sh.view(1, 1, 1, -1), sl, sc.view(1, 1, 1, -1)).permute(0, 3, 1, 2)
    torch.testing.assert_close(out, ref, rtol=1e-6, atol=1e-6)
    call("rd_feature_export", view(buf, coff), None, None, ptr(out), B, H, W, Cc, act, stream_ptr())
    assert torch.equal(out, sl.permute(0, 3, 1, 2))
    x = torch.randn(B, Cc, H, W, device="cuda")
    dst = torch.full((B, H, W, pitch), 7.0, device="cuda").to(td)
    call("rd_feature_import", ptr(x), view(dst, coff), B, H, W, Cc, act, stream_ptr())
    assert torch.equal(dst[..., coff:coff + Cc].float(), x.permute(0, 2, 3, 1).to(td).float())
    assert bool((dst[..., :coff] == 7).all()) and bool((dst[..., coff + Cc:] == 7).all())      # neighbours untouched
    with pytest.raises(_lib.RdError):
        call("rd_feature_export", view(buf, coff), ptr(sc), None, ptr(out), B, H, W, Cc, act, stream_ptr())


def test_masked_l1_edge_cases_follow_the_reference():
    """criteria_new.py:50-53 is unguarded: no valid pixel -> mean of an empty selection = NaN (loss and gradient are not
    "fixed" here either); one valid pixel -> |t - p| and a single +-1 gradient; non-contiguous / fp64 inputs are accepted;
    pred == target on a valid pixel has sign 0."""
    from radar_depth_b200.evaluation.criteria_new import MaskedL1Loss
    crit = MaskedL1Loss()
    pred = torch.rand(1, 1, 8, 12, device="cuda").requires_grad_(True)
    loss = crit(pred, torch.zeros(1, 1, 8, 12, device="cuda"))
    assert torch.isnan(loss)
    tgt = torch.zeros(1, 1, 8, 12, device="cuda")
    tgt[0, 0, 3, 5] = 7.5
    pred = torch.full((1, 1, 8, 12), 2.0, device="cuda", requires_grad=True)
    loss = crit(pred, tgt)
    loss.backward()
    assert float(loss) == 5.5
    exp = torch.zeros_like(tgt)
    exp[0, 0, 3, 5] = -1.0
    assert torch.equal(pred.grad, exp)
    # exact hit on a valid pixel: sign(0) = 0 like torch's abs backward; the count still includes it
    tgt2 = tgt.clone()
    tgt2[0, 0, 0, 0] = 2.0
    pred2 = torch.full((1, 1, 8, 12), 2.0, device="cuda", requires_grad=True)
    loss2 = crit(pred2, tgt2)
    loss2.backward()
    assert float(loss2) == 2.75 and float(pred2.grad[0, 0, 0, 0]) == 0.0 and float(pred2.grad[0, 0, 3, 5]) == -0.5
    # a transposed (non-contiguous) fp64 pair gives the same number as the reference formula
    torch.manual_seed(8)
    p3 = (torch.rand(2, 1, 12, 8, device="cuda", dtype=torch.float64) * 20).transpose(2, 3)
    t3 = (torch.rand(2, 1, 12, 8, device="cuda", dtype=torch.float64) * 20).transpose(2, 3)
    t3 = torch.where(t3 > 10, t3, torch.zeros_like(t3))
    ref = (t3 - p3)[t3 > 0].abs().mean()
    assert abs(float(crit(p3, t3)) - float(ref)) < 1e-5


def test_sid_filter_matches_reference_formula():
    torch.manual_seed(4)
    d = torch.rand(2, 1, 20, 30, device="cuda") * 80
    r = torch.zeros_like(d)
    mk = torch.rand_like(d) < 0.3
    r[mk] = torch.rand(int(mk.sum()), device="cuda") * 80 + 1
    rf, mask = torch.empty_like(d), torch.empty_like(d)
    call("rd_sid_filter", ptr(r), ptr(d), d.numel(), ptr(rf), ptr(mask), stream_ptr())
    thr = torch.exp(d * math.log(18.0 / 5.0) / 100.0 + math.log(5.0))      # multistage_model.py:96-100
    mref = ((d - r).abs() <= thr).float()
    assert (mask != mref).float().mean() < 1e-3      # only exact-threshold ties may differ
    torch.testing.assert_close(rf, r * mask)


def test_pack_unpack_and_sgd():
    torch.manual_seed(5)
    src = torch.randn(1000, device="cuda")
    idx = torch.randint(-1, 1000, (4096,), device="cuda", dtype=torch.int32)
    lo = idx.clone()
    lo[idx >= 0] |= (1 << 30)
    out = ops.pack_weights(src, torch.cat([idx, lo]))
    hi_ref = torch.where(idx >= 0, src[idx.clamp(min=0).long()], torch.zeros((), device="cuda")).bfloat16()
    assert torch.equal(out[:4096], hi_ref)
    lo_ref = torch.where(idx >= 0, src[idx.clamp(min=0).long()] - hi_ref.float(), torch.zeros((), device="cuda")).bfloat16()
    assert torch.equal(out[4096:], lo_ref)
    p, g, m = torch.randn(777, device="cuda"), torch.randn(777, device="cuda"), torch.zeros(777, device="cuda")
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.SGD([pr], lr=0.01, momentum=0.9, weight_decay=1e-4)
    for step in range(3):
        pr.grad = g.clone()
        opt.step()
        call("rd_sgd", ptr(p), ptr(g), ptr(m), 777, 0.01, 0.9, 1e-4, 1 if step == 0 else 0, stream_ptr())
    torch.testing.assert_close(p, pr.detach(), rtol=1e-6, atol=1e-7)


def test_pack_weights_g8_equals_the_per_element_gather():
    """Compact (base, stride) table + fallback rows (convplan.compact_pack_table) against rd_pack_weights on a table with
    arithmetic-progression groups, all-invalid groups, groups with holes and lo-part (bit 30) groups; also the `dirty` switch."""
    import numpy as np
    from radar_depth_b200 import convplan as cp
    rng = np.random.default_rng(7)
    n_src = 5000
    rows = []
    for _ in range(300):
        kind = rng.integers(0, 4)
        if kind == 0:                                           # arithmetic progression (stride 0..9)
            b, st = int(rng.integers(0, n_src - 80)), int(rng.integers(0, 10))
            r = b + st * np.arange(8)
        elif kind == 1:
            r = np.full(8, -1)
        elif kind == 2:                                         # holes
            r = rng.integers(0, n_src, 8)
            r[rng.integers(0, 8, 3)] = -1
        else:                                                   # lo part of the split
            b, st = int(rng.integers(0, n_src - 80)), int(rng.integers(1, 10))
            r = (b + st * np.arange(8)) | (1 << 30)
        rows.append(r)
    idx_np = np.concatenate(rows).astype(np.int32)
    src = torch.randn(n_src, device="cuda")
    idx = torch.from_numpy(idx_np).cuda()
    ref = ops.pack_weights(src, idx)
    grp, fb = cp.compact_pack_table(idx_np)
    g_t, f_t = torch.from_numpy(grp).cuda(), torch.from_numpy(fb).cuda()
    out = torch.full_like(ref, float("nan"))
    call("rd_pack_weights_g8", ptr(src), ptr(g_t), ptr(f_t), ptr(out), grp.shape[0], None, stream_ptr())
    assert torch.equal(out.view(torch.int16), ref.view(torch.int16))
    dirty = torch.zeros(1, dtype=torch.int32, device="cuda")
    out2 = torch.zeros_like(ref)
    call("rd_pack_weights_g8", ptr(src), ptr(g_t), ptr(f_t), ptr(out2), grp.shape[0], ptr(dirty), stream_ptr())
    assert not out2.float().abs().sum().item()                  # *dirty == 0: nothing written
    dirty.fill_(1)
    call("rd_pack_weights_g8", ptr(src), ptr(g_t), ptr(f_t), ptr(out2), grp.shape[0], ptr(dirty), stream_ptr())
    assert torch.equal(out2.view(torch.int16), ref.view(torch.int16))
