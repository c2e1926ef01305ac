"""Where the time between two consecutive conv launches goes (GPU box): per-CTA global-timer stamps at entry / exit of
rd::conv_fprop_kernel (dbg_flags 8|16) for launches issued back to back.
usage: python tools/launch_gap.py <shape> [bn]"""
import sys
import torch
from radar_depth_b200 import _lib, convplan as cp, ops

shapes = {"l1": (64, 64, 3, 1, 1, (88, 304), (88, 304)), "l2": (128, 128, 3, 1, 1, (44, 152), (44, 152)),
          "l4": (512, 512, 3, 1, 1, (11, 38), (11, 38)), "dep1": (16, 16, 3, 1, 1, (88, 304), (88, 304)),
          "dep3": (64, 64, 3, 1, 1, (22, 76), (22, 76))}
name = sys.argv[1] if len(sys.argv) > 1 else "l1"
bn = len(sys.argv) > 2 and sys.argv[2] == "bn"
B = 16
Cout, Cin, k, s, pad, shw, dhw = shapes[name]
g = cp.gconv_standard(0, Cout, Cin, k, s, pad)
x = torch.randn(B, shw[0], shw[1], Cin, device="cuda").bfloat16()
w = torch.randn(Cout * Cin * k * k, device="cuda") * 0.05
sc, sh = torch.rand(Cin, device="cuda") + 0.5, torch.randn(Cin, device="cuda") * 0.1
plan = cp.plan_fprop(g, B, shw, dhw, _lib.RD_BF16)
wpk = ops.pack_weights(w, torch.from_numpy(plan.pack_idx).cuda())
out = torch.empty(B, dhw[0], dhw[1], Cout, device="cuda", dtype=torch.bfloat16)
stats = torch.zeros(2, Cout, dtype=torch.float64, device="cuda")
ld = (sc, sh, 0.0) if bn else None
p = plan.params
ncta = min(p.max_ctas, plan.ntiles) * p.nblk
R = 6
dbg = [torch.zeros(6, ncta, dtype=torch.int64, device="cuda") for _ in range(R)]
for _ in range(3):
    ops.conv_fprop(plan, ops.view(x), wpk, ops.view(out), ld=ld, stats=(stats, Cout))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for r in range(R):
    ops.conv_fprop(plan, ops.view(x), wpk, ops.view(out), ld=ld, stats=(stats, Cout), dbg=dbg[r], dbg_flags=8 | 16)
e1.record()
torch.cuda.synchronize()
print(f"{name}{' +bn' if bn else ''}: {R} launches back to back, {e0.elapsed_time(e1) * 1e3 / R:.1f} us per launch by events, ctas={ncta}")
prev_end = None
for r in range(R):
    d = dbg[r].cpu()
    ent, ext = d[0], d[1]
    t0 = int(ent.min())
    line = (f"  launch {r}: entry skew {int(ent.max()) - t0:6d} ns | kernel (first entry -> last exit) {int(ext.max()) - t0:7d} ns | "
            f"epilogue done {float(d[5].double().mean()) / 1e3:6.1f}/{int(d[5].max()) / 1e3:6.1f} kcyc, exit {float(d[2].double().mean()) / 1e3:6.1f}/{int(d[2].max()) / 1e3:6.1f} kcyc (mean/max)"
            f" | first exit {int(ext.min()) - t0:7d} ns")
    if prev_end is not None:
        line += f" | gap to previous launch {t0 - prev_end:6d} ns"
    prev_end = int(ext.max())
    print(line)
