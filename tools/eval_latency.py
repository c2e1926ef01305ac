"""Inference latency of the eval path (model.eval(), torch.no_grad(), CUDA graphs): what validate() times as t_GPU
(reference main.py:584-595).  usage: python tools/eval_latency.py [batch]"""
import sys
import torch
from radar_depth_b200.model.models import ResNet_latefusion
sys.path.insert(0, ".")
from bench import synth_host_batch, H, W

b = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.manual_seed(0)
m = ResNet_latefusion(18, "upproj", (H, W), 4, pretrained=False).cuda().eval()
x, _ = synth_host_batch(b, 1234)
x = x.cuda()
with torch.no_grad():
    for _ in range(5):
        y = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n):
        y = m(x)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
eng = m._engine
print(f"eval b={b} {H}x{W} bf16: {ms:.3f} ms/forward = {b / ms * 1e3:.1f} images/s, {len(eng.fwd_eval)} launches per forward in the eval program, {len(eng.fwd_infer)} in the inference program (RD_INFER_FOLD)")
