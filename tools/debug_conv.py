"""First-contact diagnostics for the tcgen05 conv programs (run on the GPU box): tiny problems with structured
weights so that a wrong descriptor / layout shows up as a recognisable pattern.  Prints, never asserts."""
import sys
import numpy as np
import torch

from radar_depth_b200 import _lib, convplan as cp, ops


def run(g, B, s_hw, d_hw, act, x, w, label):
    plan = cp.plan_fprop(g, B, s_hw, d_hw, act)
    p = plan.params
    print(f"[{label}] geo={plan.info['geo']} IS={p.IS} WS={p.WS} N={p.N} nblk={p.nblk} P={p.P} ntaps={p.ntaps} groups={plan.info['grp_n']}")
    wpk = ops.pack_weights(w, torch.from_numpy(plan.pack_idx).cuda())
    out = torch.full((B, d_hw[0], d_hw[1], g.N), float("nan"), device="cuda", dtype=ops.act_torch_dtype(act))
    try:
        ops.conv_fprop(plan, ops.view(x), wpk, ops.view(out))
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        print(f"[{label}] LAUNCH/EXEC ERROR: {e}")
        try:
            print("device error code:", hex(ops.device_error()))
        except Exception as e2:  # noqa
            print("device_error failed:", e2)
        return None
    wref = w.bfloat16().float() if act == _lib.RD_BF16 else w
    ref = cp.gconv_reference(g, x.float(), wref, d_hw)
    o = out.float()
    nan = torch.isnan(o).sum().item()
    err = (torch.nan_to_num(o) - ref).abs().max().item()
    rel = ((torch.nan_to_num(o) - ref).norm() / (ref.norm() + 1e-12)).item()
    print(f"[{label}] nan={nan} maxerr={err:.4g} rel={rel:.4g} refmax={ref.abs().max().item():.4g}")
    if rel > 1e-2:
        print("out[0,0,:4,:8]\n", o[0, 0, :4, :8].cpu().numpy())
        print("ref[0,0,:4,:8]\n", ref[0, 0, :4, :8].cpu().numpy())
        print("out[0,1,:4,:8]\n", o[0, 1, :4, :8].cpu().numpy())
        print("ref[0,1,:4,:8]\n", ref[0, 1, :4, :8].cpu().numpy())
    return rel


def main():
    torch.manual_seed(0)
    act = _lib.RD_BF16
    # 1) 1x1 identity, 16 channels
    g = cp.gconv_standard(0, 16, 16, 1, 1, 0)
    x = torch.randn(1, 8, 12, 16).cuda().bfloat16()
    w = torch.eye(16).reshape(-1).cuda()
    run(g, 1, (8, 12), (8, 12), act, x, w, "1x1 identity")
    # 2) 1x1 random 32->48
    g = cp.gconv_standard(0, 48, 32, 1, 1, 0)
    x = torch.randn(1, 8, 12, 32).cuda().bfloat16()
    w = torch.randn(48 * 32).cuda() * 0.2
    run(g, 1, (8, 12), (8, 12), act, x, w, "1x1 random")
    # 3) 3x3 with a single non-zero tap (shift test)
    g = cp.gconv_standard(0, 16, 16, 3, 1, 1)
    x = torch.randn(1, 8, 12, 16).cuda().bfloat16()
    for ky, kx in ((1, 1), (0, 0), (2, 1), (1, 2)):
        wt = torch.zeros(16, 16, 3, 3)
        wt[:, :, ky, kx] = torch.eye(16)
        run(g, 1, (8, 12), (8, 12), act, x, wt.reshape(-1).cuda(), f"3x3 one-hot tap ({ky},{kx})")
    # 4) 3x3 random, multi tile, multi cblk
    g = cp.gconv_standard(0, 64, 64, 3, 1, 1)
    x = torch.randn(2, 20, 37, 64).cuda().bfloat16()
    w = torch.randn(64 * 64 * 9).cuda() * 0.05
    run(g, 2, (20, 37), (20, 37), act, x, w, "3x3 64->64")
    run(g, 2, (20, 37), (20, 37), _lib.RD_F32, x.float(), w, "3x3 64->64 f32 split")
    # 5) stride 2
    g = cp.gconv_standard(0, 32, 16, 3, 2, 1)
    x = torch.randn(1, 12, 16, 16).cuda().bfloat16()
    w = torch.randn(32 * 16 * 9).cuda() * 0.1
    run(g, 1, (12, 16), (6, 8), act, x, w, "3x3 s2")
    run(g.transposed(), 1, (6, 8), (12, 16), act, torch.randn(1, 6, 8, 32).cuda().bfloat16(), w, "3x3 s2 dgrad")
    # wgrad first contact
    g = cp.gconv_standard(0, 32, 16, 3, 1, 1)
    x = torch.randn(1, 8, 12, 16).cuda().bfloat16()
    dy = torch.randn(1, 8, 12, 32).cuda().bfloat16()
    plan = cp.plan_wgrad(g, 1, (8, 12), (8, 12), act)
    print("[wgrad] info", plan.info)
    dw = torch.zeros(plan.dw_elems, device="cuda")
    try:
        ops.conv_wgrad(plan, ops.view(dy), ops.view(x), dw)
        torch.cuda.synchronize()
        grad = torch.zeros(32 * 16 * 9, device="cuda")
        grad[torch.from_numpy(plan.scatter[0]).cuda()] = dw[torch.from_numpy(plan.scatter[1]).cuda()]
        ref = cp.gconv_wgrad_reference(g, x.float(), dy.float(), grad.numel())
        rel = ((grad - ref).norm() / ref.norm()).item()
        print(f"[wgrad] rel={rel:.4g}")
        if rel > 1e-2:
            print("got", grad[:16].cpu().numpy())
            print("ref", ref[:16].cpu().numpy())
    except Exception as e:  # noqa
        print("[wgrad] ERROR", e)
        try:
            print("device error code:", hex(ops.device_error()))
        except Exception as e2:  # noqa
            print("device_error failed:", e2)


if __name__ == "__main__":
    main()
