#!/usr/bin/env python
"""bench.py -- headline benchmark of the radar_depth hot path on B200 (BASELINE.json metric):
images/s for one forward + MaskedL1 + backward (+ gradient all-reduce + SGD) training step of
resnet18_latefusion --decoder upproj, per-GPU batch 16, 352x1216, synthetic RGB + sparse radar.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--arch latefusion|multistage] [--batch B]
                    [--precision bf16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

--arch latefusion (default) is BASELINE.json configs[1]/[2] (per-GPU b=16); --arch multistage is configs[3]/[4]:
resnet18_multistage_uncertainty_fixs, per-GPU b=8, the fixs uncertainty loss of main.py:416-429 (two MaskedL1 + 0.1 *
Smoothness), stage 2 initialised from the stage-1 (latefusion) weights like multistage_model.py:40-49.

Prints ONE JSON line (rank 0).  `value`: device-timed throughput with inputs resident in HBM.  `e2e`: the same step
through the public API (model(inputs) -> MaskedL1Loss -> backward -> optimizer.step) with the H2D copy of the batch
from pinned host memory and the D2H read of the loss inside the timed region.  `roofline`: tcgen05 convolution
kernels (fprop/dgrad/wgrad programs), algorithmic conv FLOPs / summed CUDA-event launch durations, against the
measured dense bf16 peak.  `cpu_baseline`: the CPU oracle (port of the reference's PyTorch path) on the host cores,
bounded sample.  `--impl reference` times that CPU path alone (the reference owns no GPU kernels of its own).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 352, 1216
# algorithmic conv FLOPs per image (2*MAC, Unpool's structural zeros excluded, no dgrad for the stems)
FLOP_FWD_BWD_PER_IMAGE = {"latefusion": 120.167e9, "multistage": 240.8e9}
METRICS = {"latefusion": "images/sec fwd+bwd resnet18_latefusion b=16 352x1216",
           "multistage": "images/sec fwd+bwd resnet18_multistage_uncertainty_fixs b=8 352x1216"}
DEFAULT_BATCH = {"latefusion": 16, "multistage": 8}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_sustained=d.get("bf16_tflops_sustained", 1374.9), bf16_burst=d.get("bf16_tflops", 1608.4),
                    hbm=d.get("hbm_gbs", 6553.6), source="measured")
    return dict(bf16_sustained=1400.0, bf16_burst=1590.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        busy = [v for v in sm if v > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_host_batch(b, seed, p_lidar=0.05):
    """SURVEY.md 8(d) synthetic batch (same recipe as oracle.torch_oracle.synth_batch), built on the host."""
    import torch
    g = torch.Generator().manual_seed(seed)
    rgb = torch.rand(b, 3, H, W, generator=g)
    mk = torch.rand(b, 1, H, W, generator=g) < 100.0 / (H * W)
    radar = torch.zeros(b, 1, H, W)
    radar[mk] = torch.rand(int(mk.sum()), generator=g) * 79 + 1
    inputs = torch.cat((rgb, radar), dim=1)
    tm = torch.rand(b, 1, H, W, generator=g) < p_lidar
    target = torch.zeros(b, 1, H, W)
    target[tm] = torch.rand(int(tm.sum()), generator=g) * 79 + 1
    return inputs, target


def cpu_reference_throughput(steps: int, warmup: int, batch: int = 2, arch: str = "latefusion"):
    """Times the CPU oracle (the reference's PyTorch path restated in oracle/) on the host cores: bounded sample."""
    import torch
    from oracle import torch_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind = "latefusion" if arch == "latefusion" else "multistage_fixs"
    sd = O.synth_state_dict(O.latefusion_entries(4) if arch == "latefusion" else O.multistage_entries())
    inputs, target = O.synth_batch(batch, H, W)
    for _ in range(max(warmup, 1)):
        O.train_step(sd, inputs, target, kind)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        O.train_step(sd, inputs, target, kind)
        ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    return dict(value=batch / med, unit="images/s", cores=cores, kind="port",
                sample=f"{steps} timed fwd+loss+bwd iterations of {arch} b={batch} 352x1216 fp32 (oracle/torch_oracle.train_step, "
                       f"torch {torch.__version__} CPU kernels, {cores} threads), median {med:.3f} s/iter"), med


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(args.steps, 6)
    base, med = cpu_reference_throughput(steps, min(args.warmup, 2), arch=args.arch)
    line = {"impl": "reference", "metric": METRICS[args.arch], "value": base["value"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 2), "ms_per_step": med * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"resnet18_{args.arch} upproj fwd+loss+bwd, 352x1216, CPU sample b=2 per step",
                       "global_batch": 2, "note": "the reference has no GPU kernels of its own; this arm is its PyTorch CPU path "
                                                  "(oracle port) on the box's host cores"},
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--arch", default="latefusion", choices=["latefusion", "multistage"])
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default 16 latefusion / 8 multistage)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-timing", action="store_true")
    ap.add_argument("--dump-launches", default=None, help="write the per-launch CUDA-event profile of one step to this file")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = DEFAULT_BATCH[args.arch]
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from radar_depth_b200.model.models import ResNet_latefusion
    from radar_depth_b200.model.multistage_model import ResNet_multistage
    from radar_depth_b200.evaluation.criteria_new import MaskedL1Loss, SmoothnessLoss
    from radar_depth_b200.optim import FusedSGD
    from radar_depth_b200 import ddp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (sm_100a); there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W_ = max(args.warmup, 3)
    K = args.steps
    b = args.batch

    torch.manual_seed(0)
    crit = MaskedL1Loss()
    if args.arch == "latefusion":
        model = ResNet_latefusion(18, "upproj", (H, W), 4, pretrained=False).cuda()
        model.precision = args.precision
        nets = [model]

        def loss_fn(out, x, t):
            return crit(out, t)
    else:
        model = ResNet_multistage(18, "upproj", (H, W), pretrained=False)
        # configs[3]: "init from latefusion weights" = multistage_model.py:40-49 with a random latefusion state
        sd1 = model.stage1.state_dict()
        model.stage2.load_state_dict(model.filter_state_dict(dict(sd1), model.stage2.state_dict()), strict=False)
        model.register_parameter("w_stage1", torch.nn.Parameter(torch.tensor(1.0)))      # main.py:166-172
        model.register_parameter("w_stage2", torch.nn.Parameter(torch.tensor(1.0)))
        model = model.cuda()
        model.stage1.precision = model.stage2.precision = args.precision
        nets = [model.stage1, model.stage2]
        smooth = SmoothnessLoss()

        def loss_fn(out, x, t):                                                          # main.py:416-429
            d1, d2 = crit(out["stage1"], t), crit(out["stage2"], t)
            return torch.exp(-model.w_stage1) * (d1 + 0.1 * smooth(out["stage1"], x)) + torch.exp(-model.w_stage2) * d2 + \
                model.w_stage1 + model.w_stage2
    model.train()
    ddp.broadcast_parameters(model)
    opt = FusedSGD(model, lr=0.01, momentum=0.9, weight_decay=1e-4)
    # Bucketed all-reduce inside the backward pass: measured SLOWER than one exposed collective at N=2 (latefusion 9.52 vs
    # 9.36 ms/step, multistage 13.47 vs 13.21: NCCL's CTAs displace persistent 148-CTA conv kernels, whose stragglers then
    # wait for an SM, and the three-segment backward adds two graph launches), so it is opt-in (RD_DDP_OVERLAP=1).
    overlap = world > 1 and os.environ.get("RD_DDP_OVERLAP", "0") == "1"
    if overlap:
        ddp.enable_overlap(model)          # bucketed all-reduce started from inside the backward pass

    h_in, h_tg = synth_host_batch(b, 1234 + rank, p_lidar=0.05)
    h_in, h_tg = h_in.pin_memory(), h_tg.pin_memory()
    d_in, d_tg = h_in.cuda(non_blocking=True), h_tg.cuda(non_blocking=True)
    h_loss = torch.zeros((), dtype=torch.float32).pin_memory()
    collectives = [0]

    def step(x, t):
        pred = model(x)
        loss = loss_fn(pred, x, t)
        opt.zero_grad()
        loss.backward()
        collectives[0] = ddp.allreduce_gradients(model, opt)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms

    # ---- warm-up (first call eager, second captures the CUDA graphs)
    for _ in range(W_):
        loss = step(d_in, d_tg)
    torch.cuda.synchronize()
    loss0 = float(loss)

    # ---- device-resident timing
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    last = {}

    def timed_step():
        last["loss"] = step(d_in, d_tg)

    ms = timed(timed_step, K)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / K
    value = world * b / (ms_per_step * 1e-3)

    # ---- end to end: every step's batch comes from pinned host memory and its loss goes back to the host.  Like a
    # DataLoader with pin_memory + non_blocking copies, the H2D copy of step i+1 is issued on a copy stream while step i
    # computes (two device buffers); every copy and every loss read-back is inside the timed region.
    copy_stream = torch.cuda.Stream()
    dev = [(torch.empty_like(d_in), torch.empty_like(d_tg)) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0}

    def prefetch(slot):
        copy_stream.wait_event(freed[slot])
        with torch.cuda.stream(copy_stream):
            dev[slot][0].copy_(h_in, non_blocking=True)
            dev[slot][1].copy_(h_tg, non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_step():
        slot = state["i"] & 1
        torch.cuda.current_stream().wait_event(ready[slot])
        prefetch(slot ^ 1)
        l = step(dev[slot][0], dev[slot][1])
        freed[slot].record()
        h_loss.copy_(l, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        state["i"] += 1

    # what the host link delivers for exactly these two copies, alone (an e2e step cannot be shorter than this)
    ce0, ce1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h2d_alone = []
    for _ in range(3):
        torch.cuda.synchronize()
        with torch.cuda.stream(copy_stream):
            ce0.record(copy_stream)
            dev[0][0].copy_(h_in, non_blocking=True)
            dev[0][1].copy_(h_tg, non_blocking=True)
            ce1.record(copy_stream)
        torch.cuda.synchronize()
        h2d_alone.append(ce0.elapsed_time(ce1))
    h2d_alone_ms = min(h2d_alone)
    for ev in freed:
        ev.record()
    prefetch(0)
    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, K)
    e2e_value = world * b / (ms_e2e / K * 1e3 * 1e-6)
    h2d = h_in.numel() * 4 + h_tg.numel() * 4
    d2h = 4

    engs = [n._engine for n in nets]
    # engine programs + per engine one SGD kernel; losses: MaskedL1 = l1_fwd + ordered_sum + finalize + l1_bwd (4 kernels),
    # Smoothness = image_sum + smoothness x3 + ordered_sum x3 + finalize (8), SID filter (1)
    launches = sum(e.launches_per_step() for e in engs) + len(engs) + (4 if args.arch == "latefusion" else 4 + 4 + 8 + 1)
    eng = engs[0]
    line = {"metric": METRICS[args.arch], "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W_,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32(bf16x3 split)", "data": "synthetic",
            "config": {"workload": ("resnet18_latefusion --decoder upproj, fwd + MaskedL1 + bwd + grad all-reduce + SGD, "
                                    f"per-GPU b={b}, 352x1216 RGB + sparse radar (BASELINE.json configs[1])") if args.arch == "latefusion"
                       else ("resnet18_multistage_uncertainty_fixs --decoder upproj, fwd + fixs loss (2x MaskedL1 + 0.1 Smoothness) + bwd + "
                             f"grad all-reduce + SGD, per-GPU b={b}, 352x1216, stage 2 initialised from the latefusion weights "
                             "(BASELINE.json configs[3])"),
                       "arch": args.arch, "deterministic": bool(eng.det),
                       "global_batch": world * b, "parallelism": f"dp{world}", "cuda_graphs": bool(eng.use_graphs),
                       "l2": "working set (>1 GB of activations per step) exceeds the 126 MB L2; no explicit flush",
                       "collectives_per_step": collectives[0], "allreduce_overlapped_with_backward": bool(overlap), "loss_after_warmup": loss0,
                       "loss_after_timed_steps": float(last["loss"])},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / K, "h2d_copy_alone_ms": h2d_alone_ms,
                    "h2d_gbs_alone": (h_in.numel() + h_tg.numel()) * 4 / (h2d_alone_ms * 1e-3) / 1e9,
                    "note": "the copy of step i+1 runs on a copy stream beside step i; when h2d_copy_alone_ms exceeds the device step the host link, not the GPU, sets e2e"},
            "gpu_launches": launches * K, "clocks": clocks}

    # ---- roofline of the tcgen05 convolution programs: per-launch CUDA-event timing on the launch stream
    if rank == 0 and not args.no_kernel_timing:
        line["roofline"] = kernel_roofline(model, engs, d_in, d_tg, loss_fn, b, args.arch, args.dump_launches)
        line["roofline"]["frac_step"] = FLOP_FWD_BWD_PER_IMAGE[args.arch] * b / (ms_per_step * 1e-3) / 1e12 / line["roofline"]["peak"]
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base, _ = cpu_reference_throughput(4, 1, arch=args.arch)
        line["cpu_baseline"] = base
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def kernel_roofline(model, engs, d_in, d_tg, loss_fn, b, arch, dump_path=None):
    """Times every launch of one training step with CUDA events (graph-replayed copies).  The headline entry aggregates the
    tcgen05 convolution programs (achieved = algorithmic conv FLOPs of the step / summed duration of those launches).
    `classes` splits the launches by the roof that bounds them: convolution programs whose arithmetic
    intensity (Launch.meta: algorithmic FLOPs / bytes) is above the ridge of the measured peaks count against the tensor
    roof, the others and every elementwise kernel against the HBM roof."""
    import torch
    from radar_depth_b200.convplan import NUM_SMS as cp_num_sms
    peaks = _peaks()
    ridge = peaks["bf16_sustained"] * 1e12 / (peaks["hbm"] * 1e9)
    saved = [e.use_graphs for e in engs]
    for e in engs:
        e.use_graphs = False
    try:
        per = {}
        for rep in range(3):
            out = model(d_in)            # warm (eager)
            loss = loss_fn(out, d_in, d_tg)
            loss.backward()
        torch.cuda.synchronize()
        reps = 6
        # depth_lane_convs: the depth encoder's conv programs, timed alone on their lane's SMs (hidden beside lane 0 in the step)
        cls = {"tensor_bound_convs": [0.0, 0, 0.0, 0.0], "hbm_bound_convs": [0.0, 0, 0.0, 0.0], "depth_lane_convs": [0.0, 0, 0.0, 0.0],
               "elementwise": [0.0, 0, 0.0, 0.0]}
        worst = []
        lanes = {"lane1_ms": 0.0, "lane0_ms_beside_lane1": 0.0, "exposed_lane1_ms": 0.0, "lane1_conv_ms": 0.0}
        from radar_depth_b200 import determinism
        for eng in engs:
            for prog_name, prog in (("fwd", eng.fwd), ("bwd", eng.bwd)):
                in_par, reg0, reg1 = False, 0.0, 0.0
                for L in prog:
                    if L.sync == "join" and in_par:
                        lanes["lane1_ms"] += reg1
                        lanes["lane0_ms_beside_lane1"] += reg0
                        lanes["exposed_lane1_ms"] += max(0.0, reg1 - reg0)
                        in_par, reg0, reg1 = False, 0.0, 0.0
                    if L.sync == "fork":
                        in_par = True
                    # `reps` copies of the launch captured into one CUDA graph and replayed: what the launch costs on the
                    # GPU inside the step's graph.  (Eager back-to-back launches measure the HOST for the short kernels: one
                    # ctypes call + tensor-map encoding is ~10 us, a 16-channel convolution runs 12 us.)
                    g = torch.cuda.CUDAGraph()
                    with determinism.mode(eng.det_scratch if eng.det else None):
                        with torch.cuda.graph(g):
                            cst = torch.cuda.current_stream().cuda_stream
                            for _ in range(reps):
                                rc = L.fn(*L.args, cst)
                                assert rc == 0, L.name
                    g.replay()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    g.replay()
                    e1.record()
                    torch.cuda.synchronize()
                    t = e0.elapsed_time(e1) / reps
                    del g
                    kind = L.name.split(":")[0]
                    per.setdefault(kind, [0.0, 0])
                    per[kind][0] += t
                    per[kind][1] += 1
                    meta = L.meta or {}
                    is_conv = kind in ("conv_f", "conv_d", "wgrad")
                    if in_par:
                        if L.lane == 1:
                            reg1 += t
                            lanes["lane1_conv_ms"] += t if is_conv else 0.0
                        else:
                            reg0 += t
                    if is_conv and in_par and L.lane == 1:
                        key = "depth_lane_convs"
                    elif is_conv:
                        key = "tensor_bound_convs" if meta["flops"] / meta["bytes"] >= ridge else "hbm_bound_convs"
                    else:
                        key = "elementwise"
                    c = cls[key]
                    c[0] += t
                    c[1] += 1
                    c[2] += meta.get("flops", 0.0)
                    c[3] += meta.get("bytes", 0.0)
                    if meta.get("bytes"):
                        frac = (meta["flops"] / (t * 1e-3) / 1e12 / peaks["bf16_sustained"]) if key == "tensor_bound_convs" \
                            else (meta["bytes"] / (t * 1e-3) / 1e9 / peaks["hbm"])
                        worst.append((t, L.name + (f" [lane 1: {eng._depth_sms} SMs]" if (in_par and L.lane == 1) else
                                                   (f" [lane 0: {cp_num_sms - eng._depth_sms} SMs]" if in_par else "")), key, round(frac, 3)))
        conv_ms = sum(v[0] for k, v in per.items() if k in ("conv_f", "conv_d", "wgrad"))
        conv_n = sum(v[1] for k, v in per.items() if k in ("conv_f", "conv_d", "wgrad"))
        total_ms = sum(v[0] for v in per.values())
        flops = FLOP_FWD_BWD_PER_IMAGE[arch] * b
        achieved = flops / (conv_ms * 1e-3) / 1e12
        classes = {}
        for k, (ms, n, fl, by) in cls.items():
            ent = {"launches": n, "ms_per_step": round(ms, 4), "algorithmic_gflop": round(fl / 1e9, 2), "algorithmic_gbytes": round(by / 1e9, 3)}
            if k == "depth_lane_convs":
                ent["note"] = "each launch timed alone on the lane's SMs (depth_sms of 148); in the step this time runs beside lane 0 (roofline.lanes)"
            if ms > 0:
                if k == "tensor_bound_convs":
                    ent.update(bound="tensor", achieved=fl / (ms * 1e-3) / 1e12, unit="TFLOP/s", peak=peaks["bf16_sustained"])
                else:
                    ent.update(bound="hbm", achieved=by / (ms * 1e-3) / 1e9, unit="GB/s", peak=peaks["hbm"])
                ent["frac"] = ent["achieved"] / ent["peak"]
            classes[k] = ent
        worst.sort(reverse=True)
        if dump_path:
            os.makedirs(os.path.dirname(dump_path) or ".", exist_ok=True)
            with open(dump_path, "w") as fh:
                fh.write(f"# {arch} b={b}: per-launch CUDA-event times of one step ({reps} copies of the launch replayed from a CUDA graph); "
                         f"frac = fraction of the roof that bounds the launch (tensor {peaks['bf16_sustained']} TFLOP/s sustained / HBM {peaks['hbm']} GB/s)\n")
                fh.write(f"# total {total_ms:.3f} ms over {sum(v[1] for v in per.values())} launches; by kind: "
                         + json.dumps({k: round(v[0], 4) for k, v in sorted(per.items())}) + "\n")
                for t, n, k, f in worst:
                    fh.write(f"{t:9.4f} ms  {k:20s} {f:6.3f}  {n}\n")
        # DRAM bytes per conv launch from the committed ncu launch list of this same command (profiles/): only valid for
        # the configuration that capture was taken on (latefusion b=16 bf16)
        traffic, traffic_src = None, None
        for name in ("r02_conv_traffic.json", "r01_conv_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", name)) as fh:
                    if b == 16 and arch == "latefusion":
                        traffic, traffic_src = json.load(fh)["traffic_bytes_per_launch"], name
                        break
            except Exception:
                continue
        # In the step the depth encoder's chain (lane 1) runs on its own SMs BESIDE the RGB encoder's chain (lane 0, the other
        # SMs): the time the convolution programs occupy the GPU is the UNION of the two lanes' intervals, i.e. the serial sum
        # minus the lane-1 convolution time that runs hidden beside lane 0.  `achieved` / `frac` = algorithmic conv FLOPs of
        # the step / that union time, against the whole GPU's tensor peak; `frac_serial_sum` divides by the plain sum of every
        # launch timed alone (double-counts the overlapped interval; what an ncu launch list adds up to), `frac_step` by the
        # whole measured step.
        hidden = max(0.0, lanes["lane1_ms"] - lanes["exposed_lane1_ms"])
        hidden_conv = hidden * (lanes["lane1_conv_ms"] / lanes["lane1_ms"]) if lanes["lane1_ms"] > 0 else 0.0
        conv_crit = conv_ms - hidden_conv
        achieved_serial = achieved
        achieved = flops / (conv_crit * 1e-3) / 1e12
        return {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_sustained"], "traffic": traffic,
                "frac_serial_sum": achieved_serial / peaks["bf16_sustained"], "conv_ms_serial_sum": conv_ms,
                "frac_note": "conv FLOPs / union of the conv launches' intervals (two encoder lanes run concurrently on disjoint SM sets)",
                "conv_ms_critical_path": conv_crit, "lanes": {k: round(v, 4) for k, v in lanes.items()},
                "depth_sms": int(getattr(engs[0], "_depth_sms", 0)) if getattr(engs[0], "_par", False) else 0,
                "traffic_note": f"dram__bytes_read+write per conv launch, ncu launch list (cold cache per launch), bytes; profiles/{traffic_src}",
                "peak_source": peaks["source"] + " (sustained bf16)",
                "kernel": "conv_fprop_kernel + conv_wgrad_kernel (tcgen05 implicit-GEMM programs)",
                "launches": conv_n, "avg_launch_ms": conv_ms / max(conv_n, 1), "conv_ms_per_step": conv_crit,
                "all_kernels_ms_per_step": total_ms, "conv_share_of_step": conv_ms / total_ms,
                "by_kind_ms": {k: round(v[0], 4) for k, v in sorted(per.items())},
                "algorithmic_flop_per_step": flops, "ridge_flop_per_byte": round(ridge, 1), "classes": classes,
                "slowest_launches": [dict(ms=round(t, 4), name=n, cls=k, frac_of_roof=f) for t, n, k, f in worst[:12]]}
    finally:
        for e, u in zip(engs, saved):
            e.use_graphs = u


if __name__ == "__main__":
    main()
