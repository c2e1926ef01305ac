"""GPU parity of the tcgen05 convolution programs (through the C ABI) against the slow torch evaluation of the
same GConv descriptions (which tests/test_convplan.py pins to F.conv2d / autograd on CPU)."""
import numpy as np
import pytest
import torch

from radar_depth_b200 import _lib, convplan as cp, ops

pytestmark = pytest.mark.gpu

DT = {"bf16": _lib.RD_BF16, "f32": _lib.RD_F32}


def _mk(g, B, src_hw, seed, nparams, act):
    gen = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(B, src_hw[0], src_hw[1], g.Cx, generator=gen).cuda()
    w = (torch.randn(nparams, generator=gen) * 0.1).cuda()
    if act == _lib.RD_BF16:
        x = x.bfloat16()
    return x.contiguous(), w


def _nparams(g):
    return int(max(int(t.widx.max()) for t in g.taps)) + 1


CASES = [
    ("c3s1_64", lambda: cp.gconv_standard(0, 64, 64, 3, 1, 1), (20, 37), (20, 37)),
    ("c3s1_16", lambda: cp.gconv_standard(0, 16, 16, 3, 1, 1), (17, 50), (17, 50)),
    ("c3s2_128", lambda: cp.gconv_standard(0, 128, 64, 3, 2, 1), (22, 38), (11, 19)),
    ("c3s2_odd", lambda: cp.gconv_standard(0, 32, 16, 3, 2, 1), (23, 39), (12, 20)),
    ("c1s2", lambda: cp.gconv_standard(0, 128, 64, 1, 2, 0), (22, 38), (11, 19)),
    ("c1s1_640", lambda: cp.gconv_standard(0, 512, 640, 1, 1, 0), (11, 38), (11, 38)),
    ("c3s1_512", lambda: cp.gconv_standard(0, 512, 512, 3, 1, 1), (11, 38), (11, 38)),
    ("stem4", lambda: cp.gconv_stem(0, 64 * 3 * 49, 1), (24, 40), (24, 40)),
    ("stem5", lambda: cp.gconv_stem(0, 64 * 3 * 49, 2), (24, 40), (24, 40)),
    ("up_256", lambda: cp.gconv_upproj(0, 128 * 256 * 25, 256, 128), (11, 19), (22, 38)),
    ("up_32", lambda: cp.gconv_upproj(0, 16 * 32 * 25, 32, 16), (20, 33), (40, 66)),
]


def _tols(act):
    return (2e-2, 2e-2) if act == _lib.RD_BF16 else (3e-4, 3e-4)


def _close(got, ref, act, what):
    rtol, atol = _tols(act)
    got, ref = got.float(), ref.float()
    scale = ref.abs().max().item() + 1e-6
    err = (got - ref).abs().max().item()
    rel = ((got - ref).norm() / (ref.norm() + 1e-12)).item()
    assert rel < (8e-3 if act == _lib.RD_BF16 else 1e-4), (what, rel, err, scale)   # f32 mode = 3-term bf16 split, ~2^-16
    assert err <= atol * scale + rtol * scale, (what, rel, err, scale)


@pytest.mark.parametrize("act", ["bf16", "f32"])
@pytest.mark.parametrize("name,mk,src_hw,dst_hw", CASES, ids=[c[0] for c in CASES])
def test_fprop_and_dgrad(name, mk, src_hw, dst_hw, act):
    act = DT[act]
    g = mk()
    B = 2
    for gg, s_hw, d_hw, tag in ((g, src_hw, dst_hw, "fwd"), (g.transposed(), dst_hw, src_hw, "dgrad")):
        x, w = _mk(gg, B, s_hw, 11, _nparams(g), act)
        plan = cp.plan_fprop(gg, B, s_hw, d_hw, act)
        wpk = ops.pack_weights(w, torch.from_numpy(plan.pack_idx).cuda())
        out = torch.full((B, d_hw[0], d_hw[1], gg.N), float("nan"), device="cuda", dtype=ops.act_torch_dtype(act))
        stats = torch.zeros(2, gg.N, dtype=torch.float64, device="cuda")
        # BatchNorm finalisation fused into the kernel tail (last CTA): must equal rd_bn_finalize on the same sums
        Cn = gg.N
        gamma, beta = torch.rand(Cn, device="cuda") + 0.5, torch.randn(Cn, device="cuda")
        rm, rv = torch.randn(Cn, device="cuda"), torch.rand(Cn, device="cuda") + 0.5
        rm2, rv2 = rm.clone(), rv.clone()
        nbt, nbt2 = torch.zeros(1, dtype=torch.int64, device="cuda"), torch.zeros(1, dtype=torch.int64, device="cuda")
        fused, alone = torch.zeros(4, Cn, device="cuda"), torch.zeros(4, Cn, device="cuda")
        counter = torch.zeros(1, dtype=torch.float64, device="cuda")
        count = float(B * d_hw[0] * d_hw[1])
        tail = _lib.BnTail()
        tail.counter, tail.njobs = counter.data_ptr(), 1
        j = tail.job[0]
        j.kind, j.C, j.count, j.momentum, j.eps = 1, Cn, count, 0.1, 1e-5
        j.sum_a, j.sum_b, j.gamma, j.beta = stats[0].data_ptr(), stats[1].data_ptr(), gamma.data_ptr(), beta.data_ptr()
        j.running_mean, j.running_var, j.nbt = rm.data_ptr(), rv.data_ptr(), nbt.data_ptr()
        j.v0, j.v1, j.v2, j.v3 = (fused[i].data_ptr() for i in range(4))
        ops.conv_fprop(plan, ops.view(x), wpk, ops.view(out), stats=(stats, gg.N), tail=tail)
        assert ops.device_error() == 0
        _lib.call("rd_bn_finalize", stats[0].data_ptr(), stats[1].data_ptr(), count, gamma.data_ptr(), beta.data_ptr(),
                  rm2.data_ptr(), rv2.data_ptr(), nbt2.data_ptr(), Cn, 1, 0.1, 1e-5, *(alone[i].data_ptr() for i in range(4)),
                  ops.stream_ptr())
        torch.cuda.synchronize()
        assert torch.equal(fused, alone) and int(nbt) == 1 == int(nbt2), (name, tag)
        # running statistics: same formula, but the compiler may contract (1-m)*r + m*x differently in the two kernels
        torch.testing.assert_close(rm, rm2, rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(rv, rv2, rtol=1e-6, atol=1e-7)
        wref = w.bfloat16().float() if act == _lib.RD_BF16 else w
        ref = cp.gconv_reference(gg, x.float(), wref, d_hw)
        covered = torch.zeros(gg.OS, gg.OS, dtype=torch.bool)
        for t in gg.taps:
            covered[t.ph] = True
        if not covered.all():        # phases without taps are not written by the kernel
            for a in range(gg.OS):
                for b in range(gg.OS):
                    if not covered[a, b]:
                        out[:, a::gg.OS, b::gg.OS] = 0
        assert not torch.isnan(out.float()).any(), (name, tag)
        _close(out, ref, act, (name, tag))
        _close(stats[0].float(), ref.sum(dim=(0, 1, 2)), act, (name, tag, "sum"))
        _close(stats[1].float(), (ref * ref).sum(dim=(0, 1, 2)), act, (name, tag, "sumsq"))


@pytest.mark.parametrize("act", ["bf16", "f32"])
def test_fprop_fused_bn_act_load_and_gradient_epilogue(act):
    act = DT[act]
    g = cp.gconv_standard(0, 64, 32, 3, 1, 1)
    B, hw = 2, (19, 45)
    x, w = _mk(g, B, hw, 5, _nparams(g), act)
    gen = torch.Generator().manual_seed(3)
    sc = (torch.rand(g.Cx, generator=gen) + 0.5).cuda()
    sh = (torch.randn(g.Cx, generator=gen) * 0.3).cuda()
    plan = cp.plan_fprop(g, B, hw, hw, act)
    wpk = ops.pack_weights(w, torch.from_numpy(plan.pack_idx).cuda())
    td = ops.act_torch_dtype(act)
    out = torch.empty(B, hw[0], hw[1], g.N, device="cuda", dtype=td)
    ops.conv_fprop(plan, ops.view(x), wpk, ops.view(out), ld=(sc, sh, 0.2))
    y = x.float() * sc + sh
    y = torch.where(y > 0, y, 0.2 * y)
    if act == _lib.RD_BF16:
        y = y.bfloat16().float()
    wref = w.bfloat16().float() if act == _lib.RD_BF16 else w
    ref = cp.gconv_reference(g, y, wref, hw)
    _close(out, ref, act, "bn_act_load")
    # epilogue 1: g = (acc + addend) * act'(z*sc+sh), stats sum g, sum g*z
    z = torch.randn(B, hw[0], hw[1], g.N, generator=gen).cuda().to(td)
    add = torch.randn(B, hw[0], hw[1], g.N, generator=gen).cuda().to(td)
    esc = (torch.rand(g.N, generator=gen) + 0.5).cuda()
    esh = (torch.randn(g.N, generator=gen) * 0.3).cuda()
    stats = torch.zeros(2, g.N, dtype=torch.float64, device="cuda")
    out2 = torch.empty_like(out)
    ops.conv_fprop(plan, ops.view(x), wpk, ops.view(out2), epi=1, addend=ops.view(add), zsrc=ops.view(z),
                   ep=(esc, esh, 0.0), stats=(stats, g.N))
    assert ops.device_error() == 0
    raw = cp.gconv_reference(g, x.float(), wref, hw) + add.float()
    yz = z.float() * esc + esh
    gref = torch.where(yz > 0, raw, torch.zeros_like(raw))
    _close(out2, gref, act, "epi1")
    _close(stats[0].float(), gref.sum(dim=(0, 1, 2)), act, "epi1 sum g")
    _close(stats[1].float(), (gref * z.float()).sum(dim=(0, 1, 2)), act, "epi1 sum gz")


@pytest.mark.parametrize("act", ["bf16", "f32"])
@pytest.mark.parametrize("name,mk,src_hw,dst_hw", CASES, ids=[c[0] for c in CASES])
def test_wgrad(name, mk, src_hw, dst_hw, act):
    act = DT[act]
    g = mk()
    B = 2
    npar = _nparams(g)
    x, _ = _mk(g, B, src_hw, 21, npar, act)
    gen = torch.Generator().manual_seed(9)
    dy = torch.randn(B, dst_hw[0], dst_hw[1], g.N, generator=gen).cuda().to(ops.act_torch_dtype(act)).contiguous()
    plan = cp.plan_wgrad(g, B, src_hw, dst_hw, act)
    dw = torch.zeros(plan.dw_elems, dtype=torch.float32, device="cuda")
    ops.conv_wgrad(plan, ops.view(dy), ops.view(x), dw)
    assert ops.device_error() == 0
    grad = torch.zeros(npar, dtype=torch.float32, device="cuda")
    pi = torch.from_numpy(plan.scatter[0]).cuda()
    di = torch.from_numpy(plan.scatter[1]).cuda()
    grad[pi] = dw[di]
    ref = cp.gconv_wgrad_reference(g, x.float(), dy.float(), npar)
    _close(grad, ref, _lib.RD_F32 if act == _lib.RD_F32 else _lib.RD_BF16, (name, "wgrad"))
