"""CPU tests of the input pipeline: the oracle's restated scipy.misc.imresize and the HOST tables of
radar_depth_b200/dataset/gpu_pipeline.py against the reference's real dependencies called directly (PIL Image.resize,
ImageEnhance, scipy.ndimage.rotate of this image: Pillow 12.2, scipy 1.1x)."""
import math

import numpy as np
import pytest

from oracle import dataset_oracle as D
from radar_depth_b200.dataset import gpu_pipeline as G

PIL = pytest.importorskip("PIL")
from PIL import Image, ImageEnhance  # noqa: E402


def _apply_bilinear(a_u8, rows, cols):
    """numpy evaluation of what aug_rgb_kernel does with the two tables (horizontal pass rounded to uint8, then vertical)."""
    H, W, C = a_u8.shape
    a = a_u8.astype(np.int64)
    tmp = np.zeros((H, len(cols), C), np.int64)
    for x, (x0, n, k0, k1, k2) in enumerate(cols):
        acc = np.full((H, C), 1 << 21, np.int64)
        for t, k in enumerate((k0, k1, k2)[:n]):
            acc += a[:, x0 + t, :] * k
        tmp[:, x, :] = np.clip(acc >> 22, 0, 255)
    out = np.zeros((len(rows), len(cols), C), np.int64)
    for y, (y0, n, k0, k1, k2) in enumerate(rows):
        acc = np.full((len(cols), C), 1 << 21, np.int64)
        for t, k in enumerate((k0, k1, k2)[:n]):
            acc += tmp[y0 + t] * k
        out[y] = np.clip(acc >> 22, 0, 255)
    return out.astype(np.uint8)


def test_bilinear_tables_reproduce_pil_resize_bit_for_bit():
    rs = np.random.RandomState(0)
    for _ in range(25):
        H, W = rs.randint(6, 70), rs.randint(6, 90)
        s = rs.uniform(1.0, 1.5)
        oh, ow = int(H * s), int(W * s)
        a = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
        ref = np.array(Image.fromarray(a).resize((ow, oh), resample=2))
        got = _apply_bilinear(a, G.pil_bilinear_table(H, oh, 0, oh), G.pil_bilinear_table(W, ow, 0, ow))
        assert np.array_equal(ref, got)
    # a window of the table equals the same rows of the full table
    full = G.pil_bilinear_table(450, 611, 0, 611)
    assert np.array_equal(full[37:37 + 450], G.pil_bilinear_table(450, 611, 37, 450))
    with pytest.raises(ValueError):
        G.pil_bilinear_table(100, 80, 0, 80)


def _bilinear_loop(in_size, out_size):
    """PIL's precompute_coeffs / normalize_coeffs_8bpc as the C code loops (scalar doubles) -- the vectorised host table must
    equal it entry for entry."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support, ss = 1.0 * filterscale, 1.0 / filterscale
    rows = []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k, ww = [], 0.0
        for x in range(xmax):
            v = abs((x + xmin - center + 0.5) * ss)
            w = 1.0 - v if v < 1.0 else 0.0
            k.append(w)
            ww += w
        k = [w / ww if ww != 0.0 else w for w in k] + [0.0] * (3 - xmax)
        rows.append([xmin, xmax] + [int(0.5 + w * (1 << 22)) if i < xmax else 0 for i, w in enumerate(k)])
    return np.array(rows, dtype=np.int32)


def test_vectorised_host_tables_equal_the_scalar_loops():
    rs = np.random.RandomState(11)
    for _ in range(40):
        n = int(rs.randint(5, 900))
        m = int(n * rs.uniform(1.0, 1.5))
        assert np.array_equal(G.pil_bilinear_table(n, m, 0, m), _bilinear_loop(n, m))
        a, xo, ref = n / m, 0.0 + (n / m) * 0.5, []
        for _x in range(m):
            ref.append(int(xo))
            xo += a
        assert np.array_equal(G.pil_nearest_table(n, m, 0, m), np.array(ref, dtype=np.int32))


def test_nearest_tables_reproduce_pil_resize_in_mode_F():
    rs = np.random.RandomState(1)
    for _ in range(40):
        H, W = rs.randint(5, 470), rs.randint(5, 820)
        s = rs.uniform(1.0, 1.5)
        oh, ow = int(H * s), int(W * s)
        a = rs.rand(H, W).astype(np.float32)
        im = Image.frombytes("F", (W, H), a.tobytes())
        ref = np.array(im.resize((ow, oh), resample=0))
        yt, xt = G.pil_nearest_table(H, oh, 0, oh), G.pil_nearest_table(W, ow, 0, ow)
        assert np.array_equal(ref, a[yt][:, xt])


def _rotate_model(a, angle):
    """numpy evaluation of aug_rot_src: c = (y*m0 + x*m1) + off, valid iff 0 <= c <= len-1, index floor(c + 0.5)."""
    H, W = a.shape[:2]
    rot, off = G.rotation_affine(angle, H, W)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    cy = (yy * rot[0, 0] + xx * rot[0, 1]) + off[0]
    cx = (yy * rot[1, 0] + xx * rot[1, 1]) + off[1]
    ok = (cy >= 0) & (cy <= H - 1) & (cx >= 0) & (cx <= W - 1)
    iy = np.clip(np.floor(cy + 0.5).astype(int), 0, H - 1)
    ix = np.clip(np.floor(cx + 0.5).astype(int), 0, W - 1)
    out = a[iy, ix]
    out[~ok] = 0
    return out


@pytest.mark.parametrize("angle", [0.0, 1e-3, 3.7, -4.9, 5.0, -0.25])
def test_rotation_rule_reproduces_scipy_ndimage_rotate(angle):
    rs = np.random.RandomState(2)
    img = rs.randint(0, 256, (90, 160, 3)).astype(np.float32)
    dep = (rs.rand(90, 160) * 80).astype(np.float32)
    assert np.array_equal(D.rotate(img, angle), _rotate_model(img, angle))
    assert np.array_equal(D.rotate(dep, angle), _rotate_model(dep, angle))


def test_restated_imresize_is_bytescale_plus_pil():
    rs = np.random.RandomState(3)
    a = rs.randint(0, 256, (40, 60, 3)).astype(np.float32)
    a[0, 0, 0], a[0, 0, 1] = 0.0, 255.0                       # full range: bytescale is the identity
    out = D.imresize(a, np.float64(1.25), "bilinear")
    assert out.dtype == np.uint8 and out.shape == (50, 75, 3)
    assert np.array_equal(out, np.array(Image.fromarray(a.astype(np.uint8)).resize((75, 50), resample=2)))
    b = a * 0.5 + 20.0                                        # reduced range: the min-max stretch of scipy's toimage
    lo, hi = b.min(), b.max()
    stretched = ((b - lo) * np.float32(255.0 / float(hi - lo))).clip(0, 255) + np.float32(0.5)
    assert np.array_equal(D.bytescale(b), stretched.astype(np.uint8))
    d = rs.rand(40, 60).astype(np.float32) * 90
    dn = D.imresize(d, np.float64(1.5), "nearest", "F")
    assert dn.dtype == np.float32 and dn.shape == (60, 90) and set(np.unique(dn)).issubset(set(np.unique(d)))


def _L(a):
    a = a.astype(np.int64)
    return ((a[..., 0] * 19595 + a[..., 1] * 38470 + a[..., 2] * 7471 + 0x8000) >> 16)


def _blend(d, v, f):
    """numpy evaluation of aug_blend / aug_jitter_kernel."""
    if f == 1.0:
        return v.copy()
    if f == 0.0:
        return d.astype(np.uint8)
    t = (d.astype(np.float32) + np.float32(f) * (v.astype(np.int32) - d.astype(np.int32)).astype(np.float32)).astype(np.float32)
    if 0.0 <= f <= 1.0:
        return t.astype(np.int32).astype(np.uint8)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, t.astype(np.int32))).astype(np.uint8)


def test_jitter_arithmetic_reproduces_pil_imageenhance():
    rs = np.random.RandomState(4)
    for t in range(60):
        a = rs.randint(0, 256, (rs.randint(3, 40), rs.randint(3, 50), 3)).astype(np.uint8)
        if t % 5 == 0:
            a //= 4
        f = [0.8, 1.0, 1.2, 0.0][t % 4] if t < 8 else rs.uniform(0.8, 1.2)
        im = Image.fromarray(a)
        L = _L(a)
        mean = int(L.sum() / L.size + 0.5)
        assert np.array_equal(np.array(ImageEnhance.Brightness(im).enhance(f)), _blend(np.zeros_like(a), a, f))
        assert np.array_equal(np.array(ImageEnhance.Contrast(im).enhance(f)), _blend(np.full_like(a, mean), a, f))
        assert np.array_equal(np.array(ImageEnhance.Color(im).enhance(f)), _blend(np.stack([L] * 3, -1).astype(np.uint8), a, f))


def test_random_draws_follow_the_reference_call_order():
    """transform_train draws scale, angle, flip, crop row, crop column (nuscenes_dataset_torch_new.py:247-249,284-285), then
    ColorJitter.get_params draws three factors and shuffles the three operations (transforms.py:457-474) -- replayed here
    literally on the global np.random the reference uses."""
    for seed in (0, 7, 123):
        np.random.seed(seed)
        scale = np.random.uniform(1., 1.5)
        angle = np.random.uniform(-5., 5.)
        flip = np.random.uniform(0.0, 1.0) < 0.5
        h_bound, w_bound = math.floor(450 * scale) - 450, math.floor(800 * scale) - 800
        i = round(np.random.uniform(0, h_bound))
        j = round(np.random.uniform(0, w_bound))
        fac = [np.random.uniform(0.8, 1.2) for _ in range(3)]
        ops = ["b", "c", "s"]
        np.random.shuffle(ops)
        for mod in (G, D):
            p = mod.draw_train_params(np.random.RandomState(seed))
            assert (p["scale"], p["angle"], p["flip"], p["i"], p["j"]) == (scale, angle, bool(flip), i, j)
            assert list(p["factors"]) == fac and ["bcs"[o] for o in p["order"]] == ops


def test_constructor_mirrors_the_reference_errors():
    with pytest.raises(ValueError):
        G.GpuInputPipeline(mode="test")
    with pytest.raises(ValueError):
        G.GpuInputPipeline(modality="depth")
    with pytest.raises(ValueError):
        G.GpuInputPipeline(sparsifier="nonsense")
    with pytest.raises(NotImplementedError):
        G.GpuInputPipeline(sparsifier="uniform")
    p = G.GpuInputPipeline(mode="val", modality="rgb")
    assert p.output_size == [450, 800]


def test_oracle_train_and_val_shapes_and_semantics():
    img, lid, rad = D.synth_sample(5)
    p = D.draw_train_params(np.random.RandomState(5))
    out = D.transform_train(img, lid, rad, p, max_depth=80.0)
    assert out["inputs"].shape == (4, 450, 800) and out["labels"].shape == (1, 450, 800)
    assert out["inputs"].dtype == np.float32 and 0.0 <= out["rgb"].min() and out["rgb"].max() <= 1.0
    assert out["radar_depth"].max() <= 80.0 and np.array_equal(out["inputs"][3:], out["radar_depth"])
    # depth values are source values divided by the scale: every non-zero label is one of them
    src = np.unique((lid / 256.).astype(np.float32) / np.float32(p["scale"]))
    assert set(np.unique(out["labels"])).issubset(set(src) | {0.0})
    v = D.transform_val(img, lid, rad, max_depth=80.0)
    assert np.array_equal(v["rgb"], (img / 255.).astype(np.float32).transpose(2, 0, 1))
    assert np.array_equal(v["labels"][0], (lid / 256.).astype(np.float32))
