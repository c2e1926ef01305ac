"""GPU input pipeline against the CPU restatement of the reference's transform_train / transform_val
(oracle/dataset_oracle.py = scipy.ndimage.rotate + PIL resize + PIL ImageEnhance): BIT-EXACT, byte and index work."""
import numpy as np
import pytest
import torch

from oracle import dataset_oracle as D
from radar_depth_b200 import _lib
from radar_depth_b200.dataset.gpu_pipeline import GpuInputPipeline, draw_train_params

pytestmark = pytest.mark.gpu


def _batch(seeds, h=450, w=800):
    s = [D.synth_sample(k, h, w) for k in seeds]
    img = torch.from_numpy(np.stack([a for a, _, _ in s])).cuda()
    lid = torch.from_numpy(np.stack([b for _, b, _ in s])).cuda()
    rad = torch.from_numpy(np.stack([c for _, _, c in s])).cuda()
    return s, img, lid, rad


def _check(out, refs, keys=("inputs", "labels", "radar_depth", "lidar_depth", "rgb")):
    for b, ref in enumerate(refs):
        for k in keys:
            got = out[k][b].cpu().numpy()
            assert got.shape == ref[k].shape, (k, got.shape, ref[k].shape)
            bad = int((got != ref[k]).sum())
            assert bad == 0, (b, k, bad, float(np.abs(got - ref[k]).max()))


def test_train_batch_is_bit_exact_with_scipy_and_pil():
    seeds = [11, 12, 13, 14]
    raw, img, lid, rad = _batch(seeds)
    pipe = GpuInputPipeline(mode="train", modality="rgbd", sparsifier="radar", max_depth=80.0, seed=3)
    out = pipe(img, lid, rad)
    torch.cuda.synchronize()
    assert out["inputs"].shape == (4, 4, 450, 800)
    refs = [D.transform_train(raw[b][0], raw[b][1], raw[b][2], out["params"][b], max_depth=80.0) for b in range(len(seeds))]
    _check(out, refs)
    # the pipeline's own generator makes the reference's draw sequence
    rs = np.random.RandomState(3)
    assert [p["scale"] for p in out["params"]] == [D.draw_train_params(rs)["scale"] for _ in seeds]


@pytest.mark.parametrize("case", ["no_rotation_flip", "unit_factors", "max_scale_dark", "zero_factor"])
def test_train_edge_cases(case):
    raw, img, lid, rad = _batch([21, 22])
    p0 = draw_train_params(np.random.RandomState(9))
    p1 = draw_train_params(np.random.RandomState(10))
    if case == "no_rotation_flip":
        p0.update(angle=0.0, flip=True)
        p1.update(angle=-5.0, flip=False, scale=1.0, i=0, j=0)
    elif case == "unit_factors":                       # Image.blend returns copies at factor 1 (and the image itself at scale 1)
        p0.update(factors=[1.0, 1.0, 1.0])
        p1.update(factors=[1.2, 1.0, 0.8], order=[2, 0, 1])
    elif case == "max_scale_dark":                     # dark image: scipy's bytescale stretches it before the resize
        p0.update(scale=1.5, i=225, j=400)
        img[0] = img[0] // 3 + 7
        raw[0] = ((raw[0][0] // 3 + 7).astype(np.uint8), raw[0][1], raw[0][2])
        p1.update(scale=1.4999, i=0, j=399)
    else:
        p0.update(factors=[0.0, 1.1, 0.9])
        p1.update(factors=[1.1, 0.0, 0.9], order=[1, 2, 0])
    pipe = GpuInputPipeline(mode="train", max_depth=100.0)
    out = pipe(img, lid, rad, params=[p0, p1])
    torch.cuda.synchronize()
    refs = [D.transform_train(raw[b][0], raw[b][1], raw[b][2], [p0, p1][b], max_depth=100.0) for b in range(2)]
    _check(out, refs)


@pytest.mark.parametrize("modality", ["rgbd", "rgb"])
def test_val_batch_is_bit_exact(modality):
    raw, img, lid, rad = _batch([31, 32, 33], h=460, w=816)            # larger than the crop: CenterCrop does something
    pipe = GpuInputPipeline(mode="val", modality=modality, max_depth=60.0)
    out = pipe(img, lid, rad)
    torch.cuda.synchronize()
    refs = [D.transform_val(r[0], r[1], r[2], max_depth=60.0) for r in raw]
    if modality == "rgb":
        for r in refs:
            r["inputs"] = r["rgb"]
    _check(out, refs)


def test_rejects_host_tensors_and_wrong_types():
    pipe = GpuInputPipeline(mode="val")
    img = torch.zeros(1, 450, 800, 3, dtype=torch.uint8)
    d = torch.zeros(1, 450, 800, dtype=torch.int16)
    with pytest.raises(_lib.RdError):
        pipe(img, d, d)
    with pytest.raises(_lib.RdError):
        pipe(img.cuda().float(), d.cuda(), d.cuda())
    with pytest.raises(_lib.RdError):
        pipe(img.cuda()[:, :100], d.cuda()[:, :100], d.cuda()[:, :100])


def test_throughput_note(capsys):
    """Not a benchmark gate: prints the batch-16 latency of the train pipeline next to the oracle's per-sample CPU time."""
    import time
    raw, img, lid, rad = _batch(list(range(40, 56)))
    pipe = GpuInputPipeline(mode="train", seed=1)
    for _ in range(2):
        pipe(img, lid, rad)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = pipe(img, lid, rad)
    torch.cuda.synchronize()
    gpu_ms = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    D.transform_train(raw[0][0], raw[0][1], raw[0][2], out["params"][0])
    cpu_ms = (time.perf_counter() - t0) * 1e3
    with capsys.disabled():
        print(f"\n[input pipeline] b=16 train batch on the GPU (incl. host tables): {gpu_ms:.2f} ms; one sample through scipy+PIL on one core: {cpu_ms:.1f} ms")
