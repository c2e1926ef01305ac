"""The GPU speed bar of north_star ("the reference's cuDNN forward+backward step"), SURVEY.md 8(d) "Reference GPU timing".

Runs the reference's layers AS MODULES on stock PyTorch (nn.Conv2d / nn.BatchNorm2d / F.conv_transpose2d / nn.MaxPool2d
semantics on cuDNN + ATen): the UNMODIFIED reference classes when /root/reference (or $RADAR_DEPTH_REFERENCE) exists,
otherwise oracle/torch_modules.py's walk over the same nn layers (pinned to the real reference by
tests/test_oracle_golden.py::test_module_walk_*).  Two arms, both with cudnn.benchmark = True like main.py:11,47:
  fp32  -- as the reference runs it (NCHW fp32; TF32 for cuDNN convolutions is torch's default, stated in the output);
  bf16  -- the same modules under torch.autocast(bfloat16) + channels_last.
Step = forward + loss + zero_grad + backward + SGD(lr .01, momentum .9, wd 1e-4), CUDA-event timed, 10 warm-up + N timed.

    python tests/tools/cudnn_bar.py [--arch latefusion|multistage] [--batch B] [--steps N] [--out profiles/x.json]

This script is measurement infrastructure (it imports oracle/); nothing in radar_depth_b200/ depends on it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import torch_modules as TM  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402

H, W = 352, 1216


def build(arch):
    ref_dir = os.environ.get("RADAR_DEPTH_REFERENCE", "/root/reference")
    if os.path.isdir(ref_dir):
        from oracle.gen_golden import import_reference
        keep = (torch.Tensor.cuda, torch.nn.Module.cuda)
        ref = import_reference()          # installs no-op .cuda shims for GPU-less containers: undo them here
        torch.Tensor.cuda, torch.nn.Module.cuda = keep
        if arch == "latefusion":
            m = ref.models.ResNet_latefusion(18, "upproj", (H, W), 4, pretrained=False)
            fwd = lambda x: m(x)
        else:
            m = ref.multistage.ResNet_multistage(18, "upproj", (H, W), pretrained=False)
            fwd = lambda x: m(x)
        return m, fwd, "unmodified reference classes"
    from radar_depth_b200.model.models import ResNet_latefusion
    from radar_depth_b200.model.multistage_model import ResNet_multistage
    if arch == "latefusion":
        m = ResNet_latefusion(18, "upproj", (H, W), 4, pretrained=False)
        return m, (lambda x: TM.latefusion_forward(m, x)), "oracle/torch_modules.py walk over nn.Conv2d/nn.BatchNorm2d/F.conv_transpose2d"
    m = ResNet_multistage(18, "upproj", (H, W), pretrained=False)
    return m, (lambda x: TM.multistage_forward(m, x)), "oracle/torch_modules.py walk over nn.Conv2d/nn.BatchNorm2d/F.conv_transpose2d"


def clocks():
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active",
                              "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True, timeout=10).stdout.strip()
        return out
    except Exception as e:  # noqa
        return str(e)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="latefusion", choices=["latefusion", "multistage"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    b = args.batch or (16 if args.arch == "latefusion" else 8)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    res = {"arch": args.arch, "batch": b, "hw": [H, W], "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
           "gpu": torch.cuda.get_device_name(0), "cudnn_benchmark": True,
           "allow_tf32_conv": bool(torch.backends.cudnn.allow_tf32), "allow_tf32_matmul": bool(torch.backends.cuda.matmul.allow_tf32)}
    inputs, target = O.synth_batch(b, H, W)
    for mode in ("fp32", "bf16_autocast_channels_last"):
        m, fwd, how = build(args.arch)
        res["modules"] = how
        if args.arch == "multistage":
            m.register_parameter("w_stage1", torch.nn.Parameter(torch.tensor(1.0)))
            m.register_parameter("w_stage2", torch.nn.Parameter(torch.tensor(1.0)))
        m = m.cuda().train()
        x, t = inputs.cuda(), target.cuda()
        if mode != "fp32":
            m = m.to(memory_format=torch.channels_last)
            x = x.contiguous(memory_format=torch.channels_last)
        opt = torch.optim.SGD(m.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode != "fp32")):
                out = fwd(x)
                if args.arch == "latefusion":
                    loss = TM.masked_l1(out.float(), t)
                else:
                    d1, d2 = TM.masked_l1(out["stage1"].float(), t), TM.masked_l1(out["stage2"].float(), t)
                    s = TM.smoothness(out["stage1"].float(), x.float())
                    loss = torch.exp(-m.w_stage1) * (d1 + 0.1 * s) + torch.exp(-m.w_stage2) * d2 + m.w_stage1 + m.w_stage2
            opt.zero_grad()
            loss.backward()
            opt.step()
            return loss

        for _ in range(10):
            loss = step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        res[mode] = {"ms_per_step": ms, "images_per_s": b / ms * 1e3, "loss_after": float(loss), "clocks_after": clocks(),
                     "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        print(f"[cudnn bar] {args.arch} b={b} {mode}: {ms:.2f} ms/step = {b / ms * 1e3:.0f} img/s", flush=True)
        del m, opt
        torch.cuda.empty_cache()
    if args.arch == "latefusion":
        # how far the reference's OWN bf16-autocast forward is from its fp32 forward on the tests' synthetic weights (the
        # level the bf16 throughput mode of this repo is held to in tests/test_model_gpu.py), strict fp32 as the anchor
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        m, fwd, _ = build("latefusion")
        m.load_state_dict(O.synth_state_dict(O.latefusion_entries(4)), strict=True)
        m = m.cuda().train()
        xi, ti = O.synth_batch(2, H, W)
        xi, ti = xi.cuda(), ti.cuda()
        with torch.no_grad():
            p32 = fwd(xi).float()
            m.load_state_dict(O.synth_state_dict(O.latefusion_entries(4)), strict=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                p16 = fwd(xi).float()
        res["autocast_vs_fp32_pred_rel_l2"] = float((p16 - p32).norm() / p32.norm())
        res["autocast_vs_fp32_loss_rel"] = float((TM.masked_l1(p16, ti) - TM.masked_l1(p32, ti)).abs() / TM.masked_l1(p32, ti).abs())
        del m
    best = min(res["fp32"]["ms_per_step"], res["bf16_autocast_channels_last"]["ms_per_step"])
    res["bar_ms_per_step"] = best
    res["target_ms_per_step_1p5x"] = best / 1.5
    line = json.dumps(res)
    print(line)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
