"""Times the tcgen05 forward/data-gradient program on one layer shape with per-role cycle counters (GPU box)."""
import statistics
import sys
import torch
from radar_depth_b200 import _lib, convplan as cp, ops

shapes = {"l1": (64, 64, 3, 1, 1, (88, 304), (88, 304)), "l2": (128, 128, 3, 1, 1, (44, 152), (44, 152)),
          "l3": (256, 256, 3, 1, 1, (22, 76), (22, 76)), "l4": (512, 512, 3, 1, 1, (11, 38), (11, 38)),
          "l2s2": (128, 64, 3, 2, 1, (88, 304), (44, 152)), "d16": (16, 16, 3, 1, 1, (176, 608), (176, 608)),
          "dep1": (16, 16, 3, 1, 1, (88, 304), (88, 304)), "dep3": (64, 64, 3, 1, 1, (22, 76), (22, 76)),
          "stem": None, "up4": None, "up3": None}
name = sys.argv[1] if len(sys.argv) > 1 else "l1"
bn = len(sys.argv) > 2 and sys.argv[2] == "bn"
B = 16
if name == "stem":
    g = cp.gconv_stem(0, 64 * 3 * 49, 1)
    shw = dhw = (176, 608)
elif name in ("up4", "up3"):
    cin = {"up4": 32, "up3": 64}[name]
    shw = {"up4": (88, 304), "up3": (44, 152)}[name]
    dhw = (2 * shw[0], 2 * shw[1])
    g = cp.gconv_upproj(0, (cin // 2) * cin * 25, cin, cin // 2)
else:
    Cout, Cin, k, s, pad, shw, dhw = shapes[name]
    g = cp.gconv_standard(0, Cout, Cin, k, s, pad)
Cin, Cout = g.Cx, g.N
x = torch.randn(B, shw[0], shw[1], Cin, device="cuda").bfloat16()
w = torch.randn(int(max(int(t.widx.max()) for t in g.taps)) + 1, device="cuda") * 0.05
sc, sh = torch.rand(Cin, device="cuda") + 0.5, torch.randn(Cin, device="cuda") * 0.1
flops = 2.0 * B * (dhw[0] // g.OS) * (dhw[1] // g.OS) * sum(int((t.widx >= 0).sum()) for t in g.taps)
abytes = (x.numel() + B * dhw[0] * dhw[1] * Cout) * 2
overrides = [None]
if len(sys.argv) > 3:
    for spec in sys.argv[3:]:
        ht, wt = (int(v) for v in spec.split("x"))
        overrides.append(dict(Ht=ht, Wt=wt))
for ov in overrides:
    plan = cp.plan_fprop(g, B, shw, dhw, _lib.RD_BF16, tile_override=ov)
    wpk = ops.pack_weights(w, torch.from_numpy(plan.pack_idx).cuda())
    out = torch.empty(B, dhw[0], dhw[1], Cout, device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(2, Cout, dtype=torch.float64, device="cuda")
    ld = (sc, sh, 0.0) if bn else None
    p = plan.params
    ncta = min(p.max_ctas, plan.ntiles) * p.nblk
    i = plan.info
    print(f"{name}{' +bn' if bn else ''}: geo={ {k_: i['geo'][k_] for k_ in ('MB', 'Wl', 'Wt', 'Ht', 'tiles_y', 'tiles_x')} } N={i['N']}x{i['nblk']} IS={i['IS']} WS={i['WS']} tiles={plan.ntiles} ctas={ncta}")
    for fl, nm in ((0, "full"), (1, "no-MMA"), (2, "no-load"), (4, "no-store"), (7, "neither")):
        for _ in range(3):
            ops.conv_fprop(plan, ops.view(x), wpk, ops.view(out), ld=ld, stats=(stats, Cout), dbg_flags=fl)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _r in range(4):      # back to back: the queue hides the host launch latency
                ops.conv_fprop(plan, ops.view(x), wpk, ops.view(out), ld=ld, stats=(stats, Cout), dbg_flags=fl)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 4)
        dbg = torch.zeros(6, ncta, dtype=torch.int64, device="cuda")
        ops.conv_fprop(plan, ops.view(x), wpk, ops.view(out), ld=ld, stats=(stats, Cout), dbg=dbg, dbg_flags=fl)
        torch.cuda.synchronize()
        d = dbg.double().mean(dim=1).cpu().numpy() / 1e3
        ms = statistics.median(ts)
        print(f"   {nm:8s} {ms * 1e3:7.1f} us {flops / ms / 1e9:7.1f} TF | kcyc/CTA loader wait {d[0]:6.1f} fill {d[1]:6.1f} | issuer wait {d[2]:6.1f} issue {d[3]:6.1f} | epi wait {d[4]:6.1f} work {d[5]:6.1f}")
    dbg = torch.zeros(6, ncta, dtype=torch.int64, device="cuda")
    ops.conv_fprop(plan, ops.view(x), wpk, ops.view(out), ld=ld, stats=(stats, Cout), dbg=dbg, dbg_flags=8)
    torch.cuda.synchronize()
    d = dbg.double().mean(dim=1).cpu().numpy() / 1e3
    dm = dbg.double().max(dim=1).values.cpu().numpy() / 1e3
    print(f"   timeline kcyc since entry (mean/max over CTAs): prologue {d[0]:.1f}/{dm[0]:.1f} loaders done {d[1]:.1f}/{dm[1]:.1f} first stage consumed {d[2]:.1f}/{dm[2]:.1f} "
          f"last commit {d[3]:.1f}/{dm[3]:.1f} first acc ready {d[4]:.1f}/{dm[4]:.1f} epilogue done {d[5]:.1f}/{dm[5]:.1f}")
    # how the launch scales with the number of CTAs (SMs) it may use: separates fixed latency from per-SM throughput
    for mc in (148, 111, 74, 37, 24, 16, 12, 8):
        mcx = max(1, mc // p.nblk)
        for _ in range(2):
            ops.conv_fprop(plan, ops.view(x), wpk, ops.view(out), ld=ld, stats=(stats, Cout), max_ctas=mcx)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _r in range(4):
            ops.conv_fprop(plan, ops.view(x), wpk, ops.view(out), ld=ld, stats=(stats, Cout), max_ctas=mcx)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 4
        print(f"   ctas {mcx * p.nblk:4d}: {ms * 1e3:7.1f} us  {abytes / ms / 1e6:7.0f} GB/s  {flops / ms / 1e9:7.1f} TF")
