"""Times the tcgen05 weight-gradient program on one layer shape under different blockings (GPU box)."""
import itertools
import statistics
import sys
import torch
from radar_depth_b200 import _lib, convplan as cp, ops

shapes = {"up4": None, "stem": None, "l1": (64, 64, 3, 1, 1, (88, 304), (88, 304)), "l2": (128, 128, 3, 1, 1, (44, 152), (44, 152)),
          "l4": (512, 512, 3, 1, 1, (11, 38), (11, 38)), "d16": (16, 16, 3, 1, 1, (176, 608), (176, 608))}
name = sys.argv[1] if len(sys.argv) > 1 else "l1"
B = 16
if name == "stem":
    g = cp.gconv_stem(0, 64 * 3 * 49, 1)
    Cout, Cin, k, shw, dhw = 80, 16, 4, (176, 608), (176, 608)
elif name.startswith("up"):
    cin = {"up1": 256, "up2": 128, "up3": 64, "up4": 32}[name]
    hw = {"up1": (11, 38), "up2": (22, 76), "up3": (44, 152), "up4": (88, 304)}[name]
    g = cp.gconv_upproj(0, 10 ** 7, cin, cin // 2)
    Cout, Cin, k, shw, dhw = cin, cin, 5, hw, (2 * hw[0], 2 * hw[1])
else:
    Cout, Cin, k, s, pad, shw, dhw = shapes[name]
    g = cp.gconv_standard(0, Cout, Cin, k, s, pad)
x = torch.randn(B, shw[0], shw[1], g.Cx, device="cuda").bfloat16()
dy = torch.randn(B, dhw[0], dhw[1], g.N, device="cuda").bfloat16()
flops = 2.0 * B * (dhw[0] // g.OS) * (dhw[1] // g.OS) * len(g.taps) * g.Cx * g.N
for nc, ks, maxc in itertools.product((None,), (None,), (0,)):
    if nc is not None and (nc > g.Cx or g.Cx % nc):
        continue
    try:
        plan = cp.plan_wgrad(g, B, shw, dhw, _lib.RD_BF16) if nc is None else cp.plan_wgrad(g, B, shw, dhw, _lib.RD_BF16, ks_target=ks, nc=nc)
    except Exception as e:  # noqa
        print(nc, ks, "infeasible", e)
        continue
    dw = torch.zeros(plan.dw_elems, device="cuda")
    mc = plan.params.max_ctas if maxc == 0 else max(1, maxc // (plan.params.ncob * plan.params.ncib * plan.params.ntg))
    for _ in range(3):
        ops.conv_wgrad(plan, ops.view(dy), ops.view(x), dw, max_ctas=mc)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.conv_wgrad(plan, ops.view(dy), ops.view(x), dw, max_ctas=mc); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = statistics.median(ts)
    ncta = mc * plan.params.ncob * plan.params.ncib * plan.params.ntg
    dbg = torch.zeros(4, ncta, dtype=torch.int64, device="cuda")
    ops.conv_wgrad(plan, ops.view(dy), ops.view(x), dw, max_ctas=mc, dbg=dbg)
    torch.cuda.synchronize()
    d = dbg.double().mean(dim=1).cpu().numpy() / 1e3
    tiles_per_cta = plan.info["ntiles"] / mc
    for fl, nm in ((1, "no-MMA"), (2, "no-load"), (3, "neither")):
        for _ in range(2):
            ops.conv_wgrad(plan, ops.view(dy), ops.view(x), dw, max_ctas=mc, dbg_flags=fl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.conv_wgrad(plan, ops.view(dy), ops.view(x), dw, max_ctas=mc, dbg_flags=fl); e1.record()
        torch.cuda.synchronize()
        dbg2 = torch.zeros(4, ncta, dtype=torch.int64, device="cuda")
        ops.conv_wgrad(plan, ops.view(dy), ops.view(x), dw, max_ctas=mc, dbg=dbg2, dbg_flags=fl)
        torch.cuda.synchronize()
        d2 = dbg2.double().mean(dim=1).cpu().numpy() / 1e3
        print(f"      {nm}: {e0.elapsed_time(e1) * 1e3:7.1f} us  [loader wait {d2[0]:6.1f} fill {d2[1]:6.1f} | issuer wait {d2[2]:6.1f} issue {d2[3]:6.1f}]")
    dbg3 = torch.zeros(4, ncta, dtype=torch.int64, device="cuda")
    ops.conv_wgrad(plan, ops.view(dy), ops.view(x), dw, max_ctas=mc, dbg=dbg3, dbg_flags=8)
    torch.cuda.synchronize()
    d3, d3m = dbg3.double().mean(dim=1).cpu().numpy() / 1e3, dbg3.double().max(dim=1).values.cpu().numpy() / 1e3
    print(f"      timeline kcyc since entry (mean/max): first stage ready {d3[0]:.1f}/{d3m[0]:.1f} last UMMA issued {d3[1]:.1f}/{d3m[1]:.1f} "
          f"accumulators complete {d3[2]:.1f}/{d3m[2]:.1f} epilogue done {d3[3]:.1f}/{d3m[3]:.1f}")
    print(f"      kcycles/CTA: loader wait {d[0]:7.1f} fill {d[1]:7.1f} | issuer wait {d[2]:7.1f} issue {d[3]:7.1f} | tiles/CTA {tiles_per_cta:.1f}")
    i = plan.info
    print(f"{name} nc={nc} ks_t={ks} -> KS={i['geo']['KS']:3d} Wl={i['geo']['Wl']:3d} Ht={i['geo']['Ht']:2d} tg={i['tg_size']:2d}x{i['ntg']} NS={i['NS']} "
          f"gx={mc:3d} ctas={mc * plan.params.ncob * plan.params.ncib * plan.params.ntg:4d}  {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TFLOP/s")
