"""Pins oracle/torch_oracle.py against outputs of the REAL reference (tests/golden/*.npz,
made by oracle/gen_golden.py) and, when /root/reference is present, against the live reference."""
import os

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O

REF_PRESENT = os.path.isdir(os.environ.get("RADAR_DEPTH_REFERENCE", "/root/reference"))


def _close(a, b, tol=2e-4):
    """rel-L2 agreement; fp32 CPU noise between F.batch_norm and the explicit formula is ~1e-6."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    rel = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12)
    assert rel < tol, rel


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)


def _check_grads(g, res, rtol=1e-2):  # fp32 noise: sign() in the L1 grad + 12-sample BNs (measured up to 4e-3)
    names = [str(n) for n in g["grad_names"]]
    assert names == list(res["grads"].keys())
    for i, k in enumerate(names):
        gr = res["grads"][k]
        ref_norm = float(g["grad_norms"][i])
        got = float(gr.double().norm())
        assert abs(got - ref_norm) <= rtol * max(ref_norm, 1e-6) + 1e-7, (k, got, ref_norm)
        head = gr.reshape(-1)[:8].numpy()
        np.testing.assert_allclose(head, g["grad_heads"][i][: head.size], rtol=1e-2, atol=1e-2 * max(ref_norm, 1e-3))  # fp32-vs-fp64 noise measured at ~1e-4*norm


@pytest.mark.parametrize("name,cin", [("latefusion_train_b2_64x96", 4), ("latefusion_train_b2_90x160", 4),
                                      ("latefusion2_c5_train_b2_64x96", 5)])
def test_latefusion_train_matches_reference_golden(golden_dir, name, cin):
    g = _load(golden_dir, name)
    b, h, w = int(g["b"]), int(g["h"]), int(g["w"])
    sd = O.synth_state_dict(O.latefusion_entries(cin))
    inputs, target = O.synth_batch(b, h, w)
    if cin == 5:
        gen = torch.Generator().manual_seed(99)
        inputs = torch.cat((inputs, torch.rand(b, 1, h, w, generator=gen) * 40), dim=1)
    res = O.train_step(sd, inputs, target, "latefusion")
    _close(res["pred"].numpy(), g["pred"])
    assert abs(float(res["loss"]) - float(g["loss"])) < 1e-4       # north_star loss tolerance
    _check_grads(g, res)
    flat = np.concatenate([res["new_buffers"][str(k)].reshape(-1).numpy() for k in g["buf_names"]])
    np.testing.assert_allclose(flat, g["buf_values"], rtol=1e-5, atol=1e-6)
    assert int(res["new_buffers"]["bn1.num_batches_tracked"]) == int(g["nbt"])


def test_latefusion_eval_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "latefusion_eval_b1_64x96")
    sd = O.synth_state_dict(O.latefusion_entries(4))
    inputs, target = O.synth_batch(1, 64, 96)
    with torch.no_grad():
        pred = O.latefusion_forward(sd, inputs, (64, 96), training=False)
    _close(pred.numpy(), g["pred"])
    assert abs(float(O.masked_l1(pred, target)) - float(g["loss"])) < 1e-4


@pytest.mark.parametrize("name,training", [("pnp_train_b2_64x96", True), ("pnp_eval_b2_64x96", False)])
def test_pnp_front_rear_match_reference_golden(golden_dir, name, training):
    """latefusion_front / latefusion_rear against pnp_forward_front / pnp_forward_rear of the real reference
    (models.py:669-707), incl. the gradient of the masked L1 loss w.r.t. the bottleneck feature."""
    g = _load(golden_dir, name)
    sd = O.synth_state_dict(O.latefusion_entries(4))
    inputs, target = O.synth_batch(2, 64, 96)
    with torch.no_grad():
        feat = O.latefusion_front(sd, inputs, training=training)
    _close(feat.numpy(), g["feature"])
    f = torch.from_numpy(g["feature"]).clone().requires_grad_(True)
    pred = O.latefusion_rear(sd, f, (64, 96), training=training)
    loss = O.masked_l1(pred, target)
    loss.backward()
    _close(pred.detach().numpy(), g["pred"])
    assert abs(float(loss) - float(g["loss"])) < 1e-4
    _close(f.grad.numpy(), g["dfeature"], tol=2e-2)          # sign() of the L1 gradient flips on fp32 noise
    with torch.no_grad():                                      # rear(front(x)) is the un-cut forward
        full = O.latefusion_forward(sd, inputs, (64, 96), training=training)
    _close(full.numpy(), g["pred"])


def test_latefusion_full_size_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "latefusion_train_b2_352x1216")
    sd = O.synth_state_dict(O.latefusion_entries(4))
    inputs, target = O.synth_batch(2, 352, 1216)
    res = O.train_step(sd, inputs, target, "latefusion")
    _close(res["pred"][..., ::8, ::8].numpy(), g["pred"])
    assert abs(float(res["loss"]) - float(g["loss"])) < 1e-4
    _check_grads(g, res)


def test_multistage_fixs_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "multistage_fixs_train_b2_64x96")
    sd = O.synth_state_dict(O.multistage_entries())
    inputs, target = O.synth_batch(2, 64, 96)
    res = O.train_step(sd, inputs, target, "multistage_fixs")
    _close(res["stage1"].numpy(), g["stage1"])
    _close(res["stage2"].numpy(), g["stage2"])
    assert abs(float(res["loss"]) - float(g["loss"])) < 1e-4
    assert float(res["mask"].sum()) == float(g["mask_sum"])
    assert abs(float(res["radar_filtered"].sum()) - float(g["radar_filtered_sum"])) < 1e-3
    _check_grads(g, res)


def test_losses_and_filter_match_reference_golden(golden_dir):
    g = _load(golden_dir, "losses_filter")
    pred = torch.from_numpy(g["pred"]).requires_grad_(True)
    l1 = O.masked_l1(pred, torch.from_numpy(g["tgt"]))
    sm = O.smoothness(pred, torch.from_numpy(g["img"]))
    (l1 + sm).backward()
    assert abs(float(l1) - float(g["l1"])) < 1e-5
    assert abs(float(sm) - float(g["smooth"])) < 1e-6
    np.testing.assert_allclose(pred.grad.numpy(), g["grad"], rtol=1e-5, atol=1e-8)
    rf, mask = O.filter_layer(torch.from_numpy(g["sparse"]), pred.detach())
    np.testing.assert_array_equal(mask.numpy(), g["mask"])
    np.testing.assert_array_equal(rf.numpy(), g["radar_filtered"])


def test_unpool_is_zero_stuffing():
    x = torch.arange(12.0).reshape(1, 2, 2, 3)
    u = O.unpool(x)
    assert u.shape == (1, 2, 4, 6)
    assert torch.equal(u[:, :, ::2, ::2], x) and float(u.sum()) == float(x.sum())


@pytest.mark.skipif(not REF_PRESENT, reason="live reference only exists in the build container")
def test_key_order_and_shapes_match_live_reference():
    from oracle.gen_golden import import_reference
    ref = import_reference()
    m = ref.models.ResNet_latefusion(18, "upproj", (64, 96), 4, pretrained=False)
    ent = O.latefusion_entries(4)
    sd = m.state_dict()
    assert list(sd.keys()) == list(ent.keys()) and len(ent) == 325
    for k, (shape, _) in ent.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    assert len(O.multistage_entries()) == 652


_METRIC_FIELDS = ("mse", "rmse", "mae", "lg10", "absrel", "delta1", "delta2", "delta3", "irmse", "imae")


def test_depth_metrics_match_reference_golden(golden_dir):
    """oracle.depth_metrics against Result.evaluate / Result_multidist.evaluate of the real reference
    (evaluation/metrics.py, golden written by oracle/gen_golden.py)."""
    g = _load(golden_dir, "metrics")
    out, tgt = torch.from_numpy(g["output"]), torch.from_numpy(g["target"])
    r = O.depth_metrics(out, tgt)
    np.testing.assert_allclose([r[k] for k in _METRIC_FIELDS], g["result"], rtol=1e-6, atol=1e-9)
    edges = [10., 20., 30., 40., 50., 60., 70., 80., 90., 100.]
    for i in range(10):
        lo = 0.0 if i == 0 else edges[i - 1]
        hi = float("inf") if i == 9 else edges[i]
        ri = O.depth_metrics(out, tgt, lo, hi)
        np.testing.assert_allclose([ri[k] for k in _METRIC_FIELDS], g["multidist"][i], rtol=1e-6, atol=1e-9, equal_nan=True)
        assert (ri["count"] > 0) == bool(g["valid_label"][i])


# ---- oracle/torch_modules.py: the module-level walk that the cuDNN speed bar (tests/tools/cudnn_bar.py) times
def test_module_walk_matches_reference_golden(golden_dir):
    from oracle import torch_modules as TM
    from radar_depth_b200.model.models import ResNet_latefusion
    g = _load(golden_dir, "latefusion_train_b2_64x96")
    m = ResNet_latefusion(18, "upproj", (64, 96), 4, pretrained=False)
    m.load_state_dict(O.synth_state_dict(O.latefusion_entries(4)), strict=True)
    m.train()
    inputs, target = O.synth_batch(2, 64, 96)
    pred = TM.latefusion_forward(m, inputs)
    loss = TM.masked_l1(pred, target)
    loss.backward()
    _close(pred.detach().numpy(), g["pred"], 1e-5)
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    names = [str(n) for n in g["grad_names"]]
    got = dict(m.named_parameters())
    for i, k in enumerate(names):
        rn = float(g["grad_norms"][i])
        assert abs(float(got[k].grad.double().norm()) - rn) <= 1e-2 * max(rn, 1e-6) + 1e-7, k
    assert int(m.bn1.num_batches_tracked) == int(g["nbt"])


def test_module_walk_multistage_matches_reference_golden(golden_dir):
    from oracle import torch_modules as TM
    from radar_depth_b200.model.multistage_model import ResNet_multistage
    g = _load(golden_dir, "multistage_fixs_train_b2_64x96")
    m = ResNet_multistage(18, "upproj", (64, 96), pretrained=False)
    m.register_parameter("w_stage1", torch.nn.Parameter(torch.tensor(1.0)))
    m.register_parameter("w_stage2", torch.nn.Parameter(torch.tensor(1.0)))
    m.load_state_dict(O.synth_state_dict(O.multistage_entries()), strict=True)
    m.train()
    inputs, target = O.synth_batch(2, 64, 96)
    o = TM.multistage_forward(m, inputs)
    d1, d2 = TM.masked_l1(o["stage1"], target), TM.masked_l1(o["stage2"], target)
    s = TM.smoothness(o["stage1"], inputs)
    loss = torch.exp(-m.w_stage1) * (d1 + 0.1 * s) + torch.exp(-m.w_stage2) * d2 + m.w_stage1 + m.w_stage2
    _close(o["stage1"].detach().numpy(), g["stage1"], 1e-5)
    _close(o["stage2"].detach().numpy(), g["stage2"], 1e-4)
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert abs(float(s) - float(g["smooth"])) <= 1e-5 * abs(float(g["smooth"]))
    assert float(o["mask"].sum()) == float(g["mask_sum"])


# ---- the other constructors on the same kernels: ResNet and the UpConv / DeConv decoders
VARIANTS = [("resnet_rgbd_upproj_b2_64x96", "resnet", 4, "upproj"), ("resnet_rgb_deconv3_b2_64x96", "resnet", 3, "deconv3"),
            ("resnet_rgbd_upconv_b2_64x96", "resnet", 4, "upconv"), ("latefusion_deconv2_b2_64x96", "latefusion", 4, "deconv2"),
            ("latefusion_upconv_b2_64x96", "latefusion", 4, "upconv")]


@pytest.mark.parametrize("name,arch,cin,decoder", VARIANTS)
def test_variant_constructors_match_reference_golden(golden_dir, name, arch, cin, decoder):
    g = _load(golden_dir, name)
    ent = O.resnet_entries(cin, decoder) if arch == "resnet" else O.latefusion_entries(cin, decoder)
    sd = O.synth_state_dict(ent)
    inputs, target = O.synth_batch(2, 64, 96)
    res = O.train_step(sd, inputs[:, :cin], target, arch, decoder=decoder)
    _close(res["pred"][..., ::2, ::2].numpy(), g["pred"], 2e-5)
    assert abs(float(res["loss"]) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    _check_grads(g, res)
