"""Replays selected launches of the training step (for ncu): one full eager warm-up step fills every buffer with real
data, then each launch whose name contains one of the given substrings is issued `reps` times.
usage: python tools/run_launch.py <reps> <batch> <name-substring> [<name-substring> ...]"""
import sys
import torch
from radar_depth_b200.model.models import ResNet_latefusion
from radar_depth_b200.evaluation.criteria_new import MaskedL1Loss
sys.path.insert(0, ".")
from bench import synth_host_batch, H, W

reps, b, pats = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3:]
torch.manual_seed(0)
m = ResNet_latefusion(18, "upproj", (H, W), 4, pretrained=False).cuda().train()
m.precision = "bf16"
x, t = synth_host_batch(b, 1234)
x, t = x.cuda(), t.cuda()
m._get_engine().use_graphs = False
crit = MaskedL1Loss()
for _ in range(2):
    crit(m(x), t).backward()
torch.cuda.synchronize()
eng = m._engine
st = torch.cuda.current_stream().cuda_stream
torch.cuda.profiler.start()      # ncu --profile-from-start off: only the replayed launches are captured
for L in eng.fwd + eng.bwd:
    if any(p == L.name or (p.endswith("*") and L.name.startswith(p[:-1])) for p in pats):
        for _ in range(reps):
            assert L.fn(*L.args, st) == 0
        torch.cuda.synchronize()
        print("replayed", L.name)
torch.cuda.profiler.stop()
