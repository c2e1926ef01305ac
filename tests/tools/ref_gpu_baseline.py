"""The bar: the reference's PyTorch/cuDNN path on the B200 (run on the GPU box; NOT part of the product or bench.py).
The reference modules cannot travel to the box, so this times the oracle restatement (same torch ops: F.conv2d ->
cuDNN, batch-norm formula, max_pool2d, interpolate) on CUDA tensors, as written (fp32 NCHW, TF32 as torch defaults)
and under autocast(bf16) + channels_last, and reports the autocast-vs-fp32 forward gap on the synthetic weights."""
import json
import statistics
import sys
import torch
from oracle import torch_oracle as O

b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H, W = 352, 1216
torch.backends.cudnn.benchmark = True
sd = {k: v.cuda() for k, v in O.synth_state_dict(O.latefusion_entries(4)).items()}
inputs, target = O.synth_batch(b, H, W)
inputs, target = inputs.cuda(), target.cuda()


def step(autocast, channels_last):
    params, work = O._leaf_params(sd, torch.float32)
    x = inputs.contiguous(memory_format=torch.channels_last) if channels_last else inputs
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        pred = O.latefusion_forward(work, x, (H, W), True, {})
        loss = O.masked_l1(pred.float(), target)
    loss.backward()
    return pred.detach().float(), loss.detach()


out = {}
for name, ac, cl in (("fp32_nchw_tf32default", False, False), ("bf16_autocast_channels_last", True, True)):
    for _ in range(5):
        step(ac, cl)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(ac, cl); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out[name] = dict(ms_per_step=statistics.median(ts), images_per_s=b / (statistics.median(ts) * 1e-3))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
p32, l32 = step(False, False)
p16, l16 = step(True, True)
out["autocast_vs_fp32_pred_rel_l2"] = float((p16 - p32).norm() / p32.norm())
out["autocast_vs_fp32_loss_rel"] = float((l16 - l32).abs() / l32.abs())
out["batch"] = b
out["allow_tf32_cudnn_default"] = True
print(json.dumps(out))
