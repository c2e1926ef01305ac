"""Generate tests/golden/*.npz from the REAL reference modules (runs only where /root/reference exists).

TEST INFRASTRUCTURE ONLY.  The reference is imported unmodified with the three shims of
SURVEY.md section 8c (no source edits):
  1. ``Tensor.cuda`` / ``Module.cuda`` -> identity, because ``Unpool.__init__`` calls ``.cuda()``
     (models.py:23) and this container has no GPU;
  2. a stub ``attrdict`` module (config/config_nuscenes.py:8 imports it; the real package is
     broken on python >= 3.10);
  3. ``torchvision.models.resnet18`` forced to ``weights=None`` because ResNet_multistage
     hard-codes ``pretrained=True`` (multistage_model.py:29-30) and there is no network.

Weights come from ``torch_oracle.synth_state_dict`` (key-seeded, independent of construction
order); inputs from ``torch_oracle.synth_batch``.  Only small tensors / summaries are stored.

Usage:  python -m oracle.gen_golden            (from the repo root)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("RADAR_DEPTH_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def import_reference():
    """Import the reference's model/criteria modules with the shims; returns a namespace."""
    if not os.path.isdir(REF):
        raise FileNotFoundError(f"reference not present at {REF}")
    import torchvision.models as tvm

    torch.Tensor.cuda = lambda self, *a, **k: self            # shim 1
    torch.nn.Module.cuda = lambda self, *a, **k: self
    if "attrdict" not in sys.modules:                          # shim 2
        m = types.ModuleType("attrdict")

        class AttrDict(dict):
            __getattr__ = dict.__getitem__

        m.AttrDict = AttrDict
        sys.modules["attrdict"] = m
    orig = tvm.resnet18
    if not getattr(orig, "_rd_nodl", False):                   # shim 3
        def resnet18(pretrained=False, **k):
            return orig(weights=None)
        resnet18._rd_nodl = True
        tvm.__dict__["resnet18"] = resnet18
        tvm.resnet18 = resnet18
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import importlib
    models = importlib.import_module("model.models")
    multistage = importlib.import_module("model.multistage_model")
    criteria = importlib.import_module("evaluation.criteria_new")
    metrics = importlib.import_module("evaluation.metrics")
    return types.SimpleNamespace(models=models, multistage=multistage, criteria=criteria, metrics=metrics)


def _subsample(t: torch.Tensor, step: int = 8) -> np.ndarray:
    return t.detach()[..., ::step, ::step].contiguous().numpy()


def _grad_summary(named_params):
    names, norms, heads = [], [], []
    for k, p in named_params:
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        names.append(k)
        norms.append(float(g.double().norm()))
        flat = g.reshape(-1)
        h = torch.zeros(8, dtype=torch.float32)
        h[: min(8, flat.numel())] = flat[:8]
        heads.append(h.numpy())
    return np.array(names), np.array(norms, dtype=np.float64), np.stack(heads)


def run_latefusion(ref, b, h, w, seed_sd=7, in_channels=4, training=True, full_pred=True):
    from oracle import torch_oracle as O
    sd = O.synth_state_dict(O.latefusion_entries(in_channels), seed=seed_sd)
    cls = ref.models.ResNet_latefusion if in_channels == 4 else ref.multistage.ResNet_latefusion2
    model = cls(18, "upproj", (h, w), in_channels, pretrained=False)
    assert list(model.state_dict().keys()) == list(sd.keys()), "state_dict key order mismatch"
    model.load_state_dict(sd, strict=True)
    model.train(training)
    inputs, target = O.synth_batch(b, h, w, seed=1234)
    if in_channels == 5:
        g = torch.Generator().manual_seed(99)
        inputs = torch.cat((inputs, torch.rand(b, 1, h, w, generator=g) * 40), dim=1)
    pred = model(inputs)
    crit = ref.criteria.MaskedL1Loss()
    loss = crit(pred, target)
    out = {"loss": np.float64(loss.item()), "b": b, "h": h, "w": w}
    out["pred"] = pred.detach().numpy() if full_pred else _subsample(pred)
    if training:
        loss.backward()
        names, norms, heads = _grad_summary(model.named_parameters())
        out.update(grad_names=names, grad_norms=norms, grad_heads=heads)
        bufs = {k: v for k, v in model.state_dict().items()
                if k.endswith("running_mean") or k.endswith("running_var")}
        out["buf_names"] = np.array(list(bufs.keys()))
        out["buf_values"] = np.concatenate([v.reshape(-1).numpy() for v in bufs.values()])
        out["nbt"] = np.int64(model.state_dict()["bn1.num_batches_tracked"].item())
    return out


def run_variant(ref, b, h, w, arch, in_channels, decoder, seed_sd=7):
    """The other constructors on the same kernels: ResNet (models.py:233-303) and the UpConv / DeConv
    decoders (models.py:135-176) under either encoder."""
    from oracle import torch_oracle as O
    ent = O.resnet_entries(in_channels, decoder) if arch == "resnet" else O.latefusion_entries(in_channels, decoder)
    sd = O.synth_state_dict(ent, seed=seed_sd)
    cls = ref.models.ResNet if arch == "resnet" else ref.models.ResNet_latefusion
    model = cls(18, decoder, (h, w), in_channels, pretrained=False)
    assert list(model.state_dict().keys()) == list(sd.keys()), "state_dict key order mismatch"
    model.load_state_dict(sd, strict=True)
    model.train()
    inputs, target = O.synth_batch(b, h, w, seed=1234)
    pred = model(inputs[:, :in_channels])
    loss = ref.criteria.MaskedL1Loss()(pred, target)
    loss.backward()
    names, norms, heads = _grad_summary(model.named_parameters())
    bufs = {k: v for k, v in model.state_dict().items() if k.endswith("running_mean") or k.endswith("running_var")}
    return {"loss": np.float64(loss.item()), "b": b, "h": h, "w": w, "pred": _subsample(pred, 2), "grad_names": names,
            "grad_norms": norms, "grad_heads": heads, "buf_names": np.array(list(bufs.keys())),
            "buf_values": np.concatenate([v.reshape(-1).numpy() for v in bufs.values()])}


def run_pnp(ref, b, h, w, training, seed_sd=7):
    """pnp_forward_front / pnp_forward_rear of the real reference (models.py:669-707) and the gradient of the masked L1
    loss with respect to the bottleneck feature (what a PnP-Depth refinement loop differentiates)."""
    from oracle import torch_oracle as O
    sd = O.synth_state_dict(O.latefusion_entries(4), seed=seed_sd)
    model = ref.models.ResNet_latefusion(18, "upproj", (h, w), 4, pretrained=False)
    model.load_state_dict(sd, strict=True)
    model.train(training)
    inputs, target = O.synth_batch(b, h, w, seed=1234)
    with torch.no_grad():
        feat = model.pnp_forward_front(inputs)
    feat = feat.clone().requires_grad_(True)
    pred = model.pnp_forward_rear(feat)
    loss = ref.criteria.MaskedL1Loss()(pred, target)
    loss.backward()
    return {"loss": np.float64(loss.item()), "b": b, "h": h, "w": w, "training": training, "feature": feat.detach().numpy(),
            "pred": pred.detach().numpy(), "dfeature": feat.grad.numpy()}


def run_multistage(ref, b, h, w, seed_sd=7, full_pred=True):
    from oracle import torch_oracle as O
    ent = O.multistage_entries()
    sd = O.synth_state_dict(ent, seed=seed_sd)
    model = ref.multistage.ResNet_multistage(18, "upproj", (h, w), pretrained=False)
    w1 = torch.nn.Parameter(torch.tensor(1.0))                 # main.py:166-172
    w2 = torch.nn.Parameter(torch.tensor(1.0))
    model.register_parameter("w_stage1", w1)
    model.register_parameter("w_stage2", w2)
    assert list(model.state_dict().keys()) == list(sd.keys()), "multistage key order mismatch"
    model.load_state_dict(sd, strict=True)
    model.train()
    inputs, target = O.synth_batch(b, h, w, seed=1234)
    # make the stage-1 prediction plausible so the SID filter keeps a mix of points
    o = model(inputs)
    l1, sm = ref.criteria.MaskedL1Loss(), ref.criteria.SmoothnessLoss()
    d1, d2 = l1(o["stage1"], target), l1(o["stage2"], target)
    s = sm(o["stage1"], inputs)
    loss = torch.exp(-w1) * (d1 + 0.1 * s) + torch.exp(-w2) * d2 + w1 + w2   # main.py:420-429
    loss.backward()
    names, norms, heads = _grad_summary(model.named_parameters())
    return {
        "loss": np.float64(loss.item()), "l1_stage1": np.float64(d1.item()), "l1_stage2": np.float64(d2.item()),
        "smooth": np.float64(s.item()), "b": b, "h": h, "w": w,
        "stage1": o["stage1"].detach().numpy() if full_pred else _subsample(o["stage1"]),
        "stage2": o["stage2"].detach().numpy() if full_pred else _subsample(o["stage2"]),
        "mask_sum": np.float64(o["mask"].sum().item()),
        "radar_filtered_sum": np.float64(o["radar_filtered"].sum().item()),
        "grad_names": names, "grad_norms": norms, "grad_heads": heads,
    }


def run_losses(ref):
    g = torch.Generator().manual_seed(5)
    pred = (torch.rand(2, 1, 24, 40, generator=g) * 30 + 1).requires_grad_(True)
    img = torch.rand(2, 4, 24, 40, generator=g)
    tgt = torch.rand(2, 1, 24, 40, generator=g) * 50
    tgt[torch.rand(2, 1, 24, 40, generator=g) < 0.7] = 0
    l1 = ref.criteria.MaskedL1Loss()(pred, tgt)
    sm = ref.criteria.SmoothnessLoss()(pred, img)
    (l1 + sm).backward()
    sparse = torch.zeros(2, 1, 24, 40)
    mk = torch.rand(2, 1, 24, 40, generator=g) < 0.2
    sparse[mk] = torch.rand(int(mk.sum()), generator=g) * 60 + 1
    rf, mask = ref.multistage.Filter_layer()(sparse, pred.detach())
    return {"pred": pred.detach().numpy(), "img": img.numpy(), "tgt": tgt.numpy(), "l1": np.float64(l1.item()),
            "smooth": np.float64(sm.item()), "grad": pred.grad.numpy(), "sparse": sparse.numpy(),
            "radar_filtered": rf.numpy(), "mask": mask.numpy()}


_METRIC_FIELDS = ("mse", "rmse", "mae", "lg10", "absrel", "delta1", "delta2", "delta3", "irmse", "imae")


def run_metrics(ref):
    """Result.evaluate and Result_multidist.evaluate of the real reference (evaluation/metrics.py) on a seeded pair."""
    g = torch.Generator().manual_seed(11)
    tgt = torch.rand(2, 1, 48, 80, generator=g) * 110
    tgt[torch.rand(2, 1, 48, 80, generator=g) < 0.6] = 0
    out = (tgt * (0.7 + 0.6 * torch.rand(2, 1, 48, 80, generator=g)) + torch.rand(2, 1, 48, 80, generator=g) * 3 + 0.5)
    r = ref.metrics.Result()
    r.evaluate(out, tgt)
    md = ref.metrics.Result_multidist()
    md.evaluate(out, tgt)
    res = {"output": out.numpy(), "target": tgt.numpy(),
           "result": np.array([getattr(r, k) for k in _METRIC_FIELDS], dtype=np.float64),
           "multidist": np.array([[getattr(x, k) for k in _METRIC_FIELDS] for x in md.result_lst], dtype=np.float64),
           "valid_label": np.array(md.valid_label, dtype=np.int64), "l1": np.float64(r.mae)}
    return res


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    ref = import_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    jobs = {
        "latefusion_train_b2_64x96": lambda: run_latefusion(ref, 2, 64, 96),
        "latefusion_eval_b1_64x96": lambda: run_latefusion(ref, 1, 64, 96, training=False),
        "latefusion_train_b2_90x160": lambda: run_latefusion(ref, 2, 90, 160),       # odd sizes (450x800 / 5)
        "latefusion2_c5_train_b2_64x96": lambda: run_latefusion(ref, 2, 64, 96, in_channels=5),
        "latefusion_train_b2_352x1216": lambda: run_latefusion(ref, 2, 352, 1216, full_pred=False),
        "multistage_fixs_train_b2_64x96": lambda: run_multistage(ref, 2, 64, 96),
        "multistage_fixs_train_b2_352x1216": lambda: run_multistage(ref, 2, 352, 1216, full_pred=False),   # configs[3] shape
        "resnet_rgbd_upproj_b2_64x96": lambda: run_variant(ref, 2, 64, 96, "resnet", 4, "upproj"),
        "resnet_rgb_deconv3_b2_64x96": lambda: run_variant(ref, 2, 64, 96, "resnet", 3, "deconv3"),
        "resnet_rgbd_upconv_b2_64x96": lambda: run_variant(ref, 2, 64, 96, "resnet", 4, "upconv"),
        "latefusion_deconv2_b2_64x96": lambda: run_variant(ref, 2, 64, 96, "latefusion", 4, "deconv2"),
        "latefusion_upconv_b2_64x96": lambda: run_variant(ref, 2, 64, 96, "latefusion", 4, "upconv"),
        "pnp_train_b2_64x96": lambda: run_pnp(ref, 2, 64, 96, True),
        "pnp_eval_b2_64x96": lambda: run_pnp(ref, 2, 64, 96, False),
        "losses_filter": lambda: run_losses(ref),
        "metrics": lambda: run_metrics(ref),
    }
    only = sys.argv[1:]
    for name, fn in jobs.items():
        if only and name not in only:
            continue
        out = fn()
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
        print(f"[golden] {name}: loss={out.get('loss', out.get('l1'))}")


if __name__ == "__main__":
    main()
