"""BASELINE.json configs[3]: resnet18_multistage_uncertainty_fixs --decoder upproj, b=8, 352x1216, one GPU --
forward (stage 1, SID filter, stage 2) + the fixs loss of main.py:416-429 + backward + fused SGD, CUDA-event timed.
Both stages start from the same latefusion weights, as ResNet_multistage does from a latefusion checkpoint
(multistage_model.py:40-49).  Prints one JSON line (a parity-config measurement, not the bench.py headline).

  python tools/bench_multistage.py [--batch 8] [--steps 20] [--warmup 5] [--precision bf16]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
H, W = 352, 1216
# algorithmic conv FLOPs per image, forward + backward, structural zeros of Unpool excluded
FLOP_PER_IMAGE = 240.8e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    args = ap.parse_args()

    import torch
    from bench import synth_host_batch
    from radar_depth_b200.evaluation.criteria_new import MaskedL1Loss, SmoothnessLoss
    from radar_depth_b200.model.models import ResNet_latefusion
    from radar_depth_b200.model.multistage_model import ResNet_multistage
    from radar_depth_b200.optim import FusedSGD

    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device (sm_100a); there is no CPU fallback for the product path")
    torch.cuda.set_device(0)
    torch.manual_seed(0)
    late = ResNet_latefusion(18, "upproj", (H, W), 4, pretrained=False)
    model = ResNet_multistage(18, "upproj", (H, W), pretrained=False)
    sd = late.state_dict()
    model.stage1.load_state_dict(sd, strict=True)
    model.stage2.load_state_dict(model.filter_state_dict(dict(sd), model.stage2.state_dict()), strict=False)
    model.register_parameter("w_stage1", torch.nn.Parameter(torch.tensor(1.0)))      # main.py:166-172
    model.register_parameter("w_stage2", torch.nn.Parameter(torch.tensor(1.0)))
    model = model.cuda().train()
    model.stage1.precision = model.stage2.precision = args.precision
    l1, sm = MaskedL1Loss(), SmoothnessLoss()
    opt = FusedSGD(model, lr=0.01, momentum=0.9, weight_decay=1e-4)
    h_in, h_tg = synth_host_batch(args.batch, 1234, p_lidar=0.05)
    x, t = h_in.cuda(), h_tg.cuda()

    def step():
        out = model(x)
        d1, d2 = l1(out["stage1"], t), l1(out["stage2"], t)
        s = sm(out["stage1"], x)
        loss = torch.exp(-model.w_stage1) * (d1 + 0.1 * s) + torch.exp(-model.w_stage2) * d2 + model.w_stage1 + model.w_stage2
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    for _ in range(max(args.warmup, 3)):
        loss0 = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({
        "metric": "images/sec fwd+bwd resnet18_multistage_uncertainty_fixs b=%d 352x1216" % args.batch,
        "value": args.batch / (ms * 1e-3), "unit": "images/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "resnet18_multistage_uncertainty_fixs --decoder upproj: stage 1 + SID filter + stage 2 + "
                               "fixs loss (2x MaskedL1, smoothness, uncertainty weights) + bwd + SGD, both stages from the "
                               "same latefusion weights (BASELINE.json configs[3])", "batch": args.batch},
        "conv_tflops": FLOP_PER_IMAGE * args.batch / (ms * 1e-3) / 1e12,
        "loss_after_warmup": float(loss0), "loss_after_timed_steps": float(loss)}))


if __name__ == "__main__":
    main()
