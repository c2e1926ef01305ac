"""Parity of EXACTLY the convolution programs bench.py times: every forward / data-gradient / weight-gradient plan of
resnet18_latefusion at B=16, 352x1216, bf16 -- i.e. every entry of radar_depth_b200/tuned_tiles.json as the engine
instantiates it (tile shape, tap-row folding, TMA mode) -- is run at full size through the C ABI and compared with the
slow torch evaluation of the same GConv description (pinned to F.conv2d / autograd by tests/test_convplan.py).
tools/autotune.py applies the same check before it accepts a candidate."""
import zlib

import pytest
import torch

from radar_depth_b200 import _lib, convplan as cp, ops
from radar_depth_b200.engine import LatefusionEngine
from radar_depth_b200.model.models import ResNet_latefusion

pytestmark = pytest.mark.gpu
B, H, W = 16, 352, 1216


def _programs():
    """(name, kind, GConv, plan, src_hw, dst_hw) of every distinct tuned key, planned exactly as the engine does."""
    act = _lib.RD_BF16
    m = ResNet_latefusion(18, "upproj", (H, W), 4, pretrained=False)
    eng = LatefusionEngine(m, 4, (H, W), act)
    eng.det = False
    eng.adopt("cpu")
    eng.configure(B, H, W)          # planning only: nothing is launched
    seen, out = set(), []
    for rec in eng.convs:
        g = rec["g"]
        fp = rec["fplan"].params
        src_hw, dst_hw = (fp.srcH, fp.srcW), (fp.dstH, fp.dstW)
        jobs = [("f", g, rec["fplan"], src_hw, dst_hw)]
        if rec["dplan"] is not None:
            jobs.append(("f", g.transposed(), rec["dplan"], dst_hw, src_hw))
        if rec["wplan"] is not None:
            jobs.append(("w", g, rec["wplan"], src_hw, dst_hw))
        sms = rec["wargs"][3]                          # SM budget of the layer's chain: part of the table key (engine lanes)
        for kind, gg, plan, s_hw, d_hw in jobs:
            key = cp.tune_key(kind, gg, B, s_hw, d_hw, act) + (f"|sm{sms}" if sms != cp.NUM_SMS else "")
            if key in seen:
                continue
            seen.add(key)
            # the blocking of the launch that transforms its source tile in shared memory, when the engine uses one
            out.append((rec["name"], kind, gg, plan, s_hw, d_hw, key, rec.get("wplan_bn") if kind == "w" else None))
    return out


def _rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-20))


def test_every_timed_conv_program_matches_the_torch_evaluation():
    table = cp.tuned_table()
    progs = _programs()
    assert sum(1 for p in progs if p[-2] in table) >= 60, "the tuned table no longer matches the engine's programs"
    worst = []
    for name, kind, g, plan, s_hw, d_hw, key, plan_bn in progs:
        gen = torch.Generator(device="cuda").manual_seed(zlib.crc32(key.encode()) % (1 << 30))
        npar = int(max(int(t.widx.max()) for t in g.taps)) + 1
        x = torch.randn(B, s_hw[0], s_hw[1], g.Cx, device="cuda", generator=gen).bfloat16()
        w = torch.randn(npar, device="cuda", generator=gen) * 0.05
        if kind == "f":
            wpk = ops.pack_weights(w, torch.from_numpy(plan.pack_idx).cuda())
            out = torch.full((B, d_hw[0], d_hw[1], g.N), float("nan"), device="cuda", dtype=torch.bfloat16)
            stats = torch.zeros(2, g.N, dtype=torch.float64, device="cuda")
            ops.conv_fprop(plan, ops.view(x), wpk, ops.view(out), stats=(stats, g.N))
            assert ops.device_error() == 0, key
            ref = cp.gconv_reference(g, x.float(), w.bfloat16().float(), d_hw)
            covered = {t.ph for t in g.taps}
            for a in range(g.OS):
                for b in range(g.OS):
                    if (a, b) not in covered:          # phases without taps are not written by the kernel
                        out[:, a::g.OS, b::g.OS] = 0
            assert not torch.isnan(out.float()).any(), key
            r = _rel(out, ref)
            rs = _rel(stats[0], ref.double().sum(dim=(0, 1, 2)))
            rq = _rel(stats[1], (ref.double() ** 2).sum(dim=(0, 1, 2)))
            worst.append((max(r, rq), key))
            assert r < 8e-3 and rq < 8e-3, (name, key, r, rs, rq)
            del out, ref, wpk
        else:
            dy = torch.randn(B, d_hw[0], d_hw[1], g.N, device="cuda", generator=gen).bfloat16()
            dw = torch.zeros(plan.dw_elems, dtype=torch.float32, device="cuda")
            ops.conv_wgrad(plan, ops.view(dy), ops.view(x), dw)
            assert ops.device_error() == 0, key
            grad = torch.zeros(npar, dtype=torch.float32, device="cuda")
            grad[torch.from_numpy(plan.scatter[0]).cuda()] = dw[torch.from_numpy(plan.scatter[1]).cuda()]
            ref = cp.gconv_wgrad_reference(g, x.float(), dy.float(), npar)
            r = _rel(grad, ref)
            worst.append((r, key))
            assert r < 8e-3, (name, key, r)
            if plan_bn is not None:
                sc = torch.rand(g.Cx, device="cuda", generator=gen) + 0.5
                sh = torch.randn(g.Cx, device="cuda", generator=gen) * 0.3
                dw.zero_()
                ops.conv_wgrad(plan_bn, ops.view(dy), ops.view(x), dw, ld=(sc, sh, 0.0))
                assert ops.device_error() == 0, key
                grad.zero_()
                grad[torch.from_numpy(plan_bn.scatter[0]).cuda()] = dw[torch.from_numpy(plan_bn.scatter[1]).cuda()]
                ref = cp.gconv_wgrad_reference(g, torch.relu(x.float() * sc + sh).bfloat16().float(), dy.float(), npar)
                r = _rel(grad, ref)
                worst.append((r, key + "|bn"))
                assert r < 8e-3, (name, key + "|bn", r)
            del dy, dw, ref
        del x, w
    worst.sort(reverse=True)
    print("[tuned tiles] programs:", len(progs), "worst rel-L2:", worst[:3])
