"""Per-launch CUDA-event timing of one training step (eager), with algorithmic TFLOP/s per convolution program."""
import os
import statistics
import sys
import torch
from radar_depth_b200.model.models import ResNet_latefusion
from radar_depth_b200.evaluation.criteria_new import MaskedL1Loss
sys.path.insert(0, ".")
from bench import synth_host_batch, H, W

b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
torch.manual_seed(0)
m = ResNet_latefusion(18, "upproj", (H, W), 4, pretrained=False).cuda().train()
m.precision = prec
x, t = synth_host_batch(b, 1234)
x, t = x.cuda(), t.cuda()
crit = MaskedL1Loss()
m._get_engine().use_graphs = False
for _ in range(3):
    crit(m(x), t).backward()
torch.cuda.synchronize()
eng = m._engine
flops = {}
for rec in eng.convs:
    g = rec["g"]
    p = rec["fplan"].params
    base = b * p.Hb * p.Wb
    # exact MACs: every tap touches every base pixel once (border zeros included, as cuDNN counts them)
    f = 2.0 * base * len(g.taps) * g.Cx * g.N
    if rec["name"] == "stem":
        f = 2.0 * b * p.Hb * p.Wb * (49 * 3 * 64 + 49 * (g.Cx // 4 - 3 if g.Cx == 16 else 2) * 16)
    flops[rec["name"]] = f
def ew_bytes(L):
    """Algorithmic bytes of the HBM-bound launches (tensors read + written, bf16 = 2 B)."""
    k = L.name.split(":")[0]
    es = 2 if prec == "bf16" else 4
    a = L.args
    if k == "bn_bwd_apply":
        return 3 * a[6] * a[7] * es
    if k == "join_bwd":
        return (4 + (1 if a[3].ptr else 0)) * a[5] * a[6] * es
    if k == "join":
        return (2 + (1 if a[3].ptr else 0)) * a[7] * a[8] * es
    if k == "maxpool":
        return (a[3] * a[4] * a[5] * a[6] + a[3] * a[13] * a[14] * a[6]) * es + a[3] * a[13] * a[14] * a[6]
    if k == "maxpool_bwd":
        return (2 * a[6] * a[7] * a[8] * a[9] + a[6] * a[13] * a[14] * a[9]) * es + a[6] * a[13] * a[14] * a[9]
    return 0
ebytes = {}
for L in eng.fwd + eng.bwd:
    ebytes[L.name] = ebytes.get(L.name, 0) + ew_bytes(L)
st = torch.cuda.current_stream().cuda_stream
rows = []
for prog in (eng.fwd, eng.bwd):
    for L in prog:
        # 4 launches back to back inside one event pair: the queue hides the host launch latency, which would
        # otherwise be charged to short kernels
        reps = 4
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            rc = L.fn(*L.args, st)
            assert rc == 0
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        rows.append((L.name, ms))
tot = sum(r[1] for r in rows)
print(f"total {tot:.3f} ms over {len(rows)} launches, b={b} {prec}")
for name, ms in sorted(rows, key=lambda r: -r[1])[:int(os.environ.get('TOPN', '70'))]:
    kind, _, lname = name.partition(":")
    lname = lname.replace("(eval)", "")
    tf = ""
    if kind in ("conv_f", "conv_d", "wgrad") and lname in flops:
        tf = f"{flops[lname] / (ms * 1e-3) / 1e12:8.1f} TFLOP/s"
    elif ebytes.get(name, 0) > 0:
        tf = f"{ebytes[name] / (ms * 1e-3) / 1e9:8.0f} GB/s"
    print(f"{ms:8.4f} ms  {100 * ms / tot:5.1f}%  {name:60s} {tf}")
import json, os
os.makedirs("gpurun_out", exist_ok=True)
json.dump(dict(rows=rows, flops=flops), open(f"gpurun_out/layers_all_{prec}_b{b}.json", "w"))
agg = {}
for name, ms in rows:
    k = name.split(":")[0]
    agg[k] = agg.get(k, 0) + ms
print({k: round(v, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])})
