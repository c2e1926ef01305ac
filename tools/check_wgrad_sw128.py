"""Weight-gradient kernel: output check against the torch evaluation of the same GConv + timing of one layer shape per
line, for the staging mode selected by RD_WG_SW128 (0 = 16-byte chunk planes, 1 = 128-byte swizzled blocks, 2 = the same with
the descriptors' base_offset field set).  Run once per mode (the switch is read once per process):
    for m in 0 1 2; do RD_WG_SW128=$m python tools/check_wgrad_sw128.py; done"""
import os
import statistics
import sys
import torch
from radar_depth_b200 import _lib, convplan as cp, ops

B = int(os.environ.get("B", "16"))
SHAPES = [  # name, Cout, Cin, k, hw, forced nc (None = plan as the engine would), fused BatchNorm on the source
    ("l1", 64, 64, 3, (88, 304), None, True),
    ("l1", 64, 64, 3, (88, 304), None, False),
    ("l2", 128, 128, 3, (44, 152), None, True),
    ("l2", 128, 128, 3, (44, 152), 128, True),
    ("l3", 256, 256, 3, (22, 76), None, True),
    ("l3", 256, 256, 3, (22, 76), 128, True),
    ("l4", 512, 512, 3, (11, 38), 64, True),
    ("l4", 512, 512, 3, (11, 38), 128, True),
    ("l4", 512, 512, 3, (11, 38), 256, True),
    ("d1", 128, 128, 3, (22, 76), None, True),
    ("f1", 512, 640, 1, (11, 38), 128, False),
    # stride-2 sources (parity planes) and the sub-pixel programs (parity planes of the gradient)
    ("s2", 128, 64, 3, (88, 304), 64, False),
    ("s3", 256, 128, 3, (44, 152), 64, False),
    ("s3", 256, 128, 3, (44, 152), 128, False),
    ("ds3", 256, 128, 1, (44, 152), 128, False),
    ("up1", 256, 256, 5, (11, 38), 64, True),
    ("up1", 256, 256, 5, (11, 38), 128, False),
    ("up2", 128, 128, 5, (22, 76), 64, True),
    ("up3", 64, 64, 5, (44, 152), 64, True),
]
only = sys.argv[1:] or None
mode = os.environ.get("RD_WG_SW128", "1")
for name, Cout, Cin, k, hw, nc, bn in SHAPES:
    if only and name not in only:
        continue
    ghw = hw
    if name.startswith("up"):
        g = cp.gconv_upproj(0, 10 ** 7, Cin, Cin // 2)
        Cout = g.N
        ghw = (2 * hw[0], 2 * hw[1])
    elif name.startswith("s") or name.startswith("ds"):
        g = cp.gconv_standard(0, Cout, Cin, k, 2, k // 2)
        ghw = ((hw[0] + 1) // 2, (hw[1] + 1) // 2)
    else:
        g = cp.gconv_standard(0, Cout, Cin, k, 1, k // 2)
    npar = int(max(int(t.widx.max()) for t in g.taps)) + 1
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(B, hw[0], hw[1], Cin, generator=gen).cuda().bfloat16()
    dy = torch.randn(B, ghw[0], ghw[1], Cout, generator=gen).cuda().bfloat16()
    sc = (torch.rand(Cin, generator=gen) + 0.5).cuda()
    sh = (torch.randn(Cin, generator=gen) * 0.3).cuda()
    try:
        plan = cp.plan_wgrad(g, B, hw, ghw, _lib.RD_BF16) if nc is None else cp.plan_wgrad(g, B, hw, ghw, _lib.RD_BF16, nc=nc, ks_target=256)
    except Exception as e:  # noqa
        print(name, nc, "infeasible", e)
        continue
    ld = (sc, sh, 0.0) if bn else None
    dw = torch.zeros(plan.dw_elems, device="cuda")
    ops.conv_wgrad(plan, ops.view(dy), ops.view(x), dw, ld=ld)
    err_dev = ops.device_error()
    grad = torch.zeros(npar, device="cuda")
    grad[torch.from_numpy(plan.scatter[0]).cuda()] = dw[torch.from_numpy(plan.scatter[1]).cuda()]
    xr = x.float()
    if bn:
        xr = torch.relu(xr * sc + sh).bfloat16().float()
    ref = cp.gconv_wgrad_reference(g, xr, dy.float(), npar)
    rel = ((grad - ref).norm() / ref.norm()).item()
    ts = []
    for _ in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.conv_wgrad(plan, ops.view(dy), ops.view(x), dw, ld=ld); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    us = statistics.median(ts[2:]) * 1e3
    i = plan.info
    flops = 2.0 * B * (ghw[0] // g.OS) * (ghw[1] // g.OS) * sum(int((t.widx >= 0).sum()) for t in g.taps)
    print(f"mode {mode} {name} bn={int(bn)} nc={i['Nc']:3d} KS={i['geo']['KS']:3d} Wl={i['geo']['Wl']:3d} Ht={i['geo']['Ht']:2d} tg={i['tg_size']}x{i['ntg']} "
          f"NS={i['NS']} gc={i['gcopies']}  rel_err {rel:.2e} dev_err {err_dev}  {us:7.1f} us  {flops / us / 1e6:6.1f} TFLOP/s  "
          f"{'OK' if rel < 8e-3 and err_dev == 0 else 'WRONG'}", flush=True)
