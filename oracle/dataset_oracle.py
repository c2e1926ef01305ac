"""CPU restatement of the reference's per-sample input pipeline -- TEST INFRASTRUCTURE ONLY (tests/ and the checker
legs of bench.py may import it; nothing under radar_depth_b200/ does).

Follows dataset/nuscenes_dataset_torch_new.py:180-197 (h5 decode), :237-412 (``transform_train``), :415-560
(``transform_val``) and dataset/transforms.py (Rotate :278-296, Resize :299-327, Crop :536-573, HorizontalFlip :406-428,
ColorJitter :431-489, CenterCrop :332-389, ToTensor :190-216) for the configuration main.py uses
(``transform_mode="sparse-to-dense"``, ``modality="rgbd"``, ``sparsifier="radar"``).

The reference's ``Resize`` calls ``scipy.misc.imresize``, which scipy removed in 1.3; the reference pins no scipy version
(no requirements file), so ``imresize`` / ``toimage`` / ``bytescale`` / ``fromimage`` are restated here from scipy 1.2.x
``scipy/misc/pilutil.py`` (the last release that shipped them) on top of Pillow -- including the famous quirk that a
float image is min-max "bytescaled" to uint8 before it is resized.  Everything else is the reference's own dependency
called directly: ``scipy.ndimage.rotate`` and PIL ``Image.resize`` / ``ImageEnhance`` (versions of this image: see
tests/test_dataset_oracle_cpu.py, which also pins the restated pieces against direct PIL calls).

Random numbers: the reference draws from the global ``np.random`` in a fixed order (scale, angle, flip, crop row, crop
column, then inside ColorJitter brightness, contrast, saturation factors and the shuffle of the three operations);
``draw_train_params`` makes the same calls on a ``np.random.RandomState`` so that seeding reproduces the sequence.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

CROP_TRAIN = (450, 800)          # config/config_nuscenes.py:43-50 (sparse_transform_config)
CROP_VAL = (450, 800)
ROTATION = 5.0
SCALE_RANGE = (1.0, 1.5)
JITTER = (0.2, 0.2, 0.2)         # transforms.ColorJitter(0.2, 0.2, 0.2), nuscenes_dataset_torch_new.py:252


# ------------------------------------------------------------------------------------------ h5 decode
def decode_depth(raw_i16: np.ndarray) -> np.ndarray:
    """nuscenes_dataset_torch_new.py:190-195: depth is stored as (metres * 256).astype(int16) (nuscenes_export.py:24-27)."""
    return raw_i16 / 256.


# ------------------------------------------------------------------------------------------ scipy 1.2 pilutil, restated
def bytescale(data, cmin=None, cmax=None, high=255, low=0):
    if data.dtype == np.uint8:
        return data
    if cmin is None:
        cmin = data.min()
    if cmax is None:
        cmax = data.max()
    cscale = cmax - cmin
    if cscale < 0:
        raise ValueError("`cmax` should be larger than `cmin`.")
    elif cscale == 0:
        cscale = 1
    scale = np.float32(float(high - low) / float(cscale))
    bytedata = (data - cmin) * scale + np.float32(low)
    return (bytedata.clip(low, high) + np.float32(0.5)).astype(np.uint8)


def toimage(arr, mode=None):
    from PIL import Image
    data = np.asarray(arr)
    shape = list(data.shape)
    if len(shape) == 2:
        shape = (shape[1], shape[0])
        assert mode == "F", "the reference only resizes 2-D arrays in mode 'F' (transforms.py:327)"
        data32 = data.astype(np.float32)
        return Image.frombytes(mode, shape, data32.tobytes())
    ca = int(np.flatnonzero(np.asarray(shape) == 3)[0])
    assert ca == 2, "H x W x 3 images only"
    bytedata = bytescale(data)
    return Image.frombytes("RGB", (shape[1], shape[0]), bytedata.tobytes())


def imresize(arr, size, interp="bilinear", mode=None):
    im = toimage(arr, mode=mode)
    ts = type(size)
    if np.issubdtype(ts, np.signedinteger):
        percent = size / 100.0
        size = tuple((np.array(im.size) * percent).astype(int))
    elif np.issubdtype(type(size), np.floating):
        size = tuple((np.array(im.size) * size).astype(int))
    else:
        size = (size[1], size[0])
    func = {"nearest": 0, "lanczos": 1, "bilinear": 2, "bicubic": 3, "cubic": 3}
    imnew = im.resize(size, resample=func[interp])
    return np.array(imnew)


# ------------------------------------------------------------------------------------------ transforms.py
def rotate(img, angle):
    import scipy.ndimage as ndi
    return ndi.rotate(img, angle, reshape=False, prefilter=False, order=0)            # transforms.py:296


def resize(img, size, interpolation):
    if img.ndim == 3:
        return imresize(img, size, interpolation)                                     # transforms.py:325
    return imresize(img, size, interpolation, "F")                                    # transforms.py:327


def color_jitter(img_u8: np.ndarray, factors, order) -> np.ndarray:
    """ColorJitter.__call__ (transforms.py:479-489) with the three factors and the shuffled order already drawn."""
    from PIL import Image, ImageEnhance
    pil = Image.fromarray(img_u8)
    enh = (ImageEnhance.Brightness, ImageEnhance.Contrast, ImageEnhance.Color)        # transforms.py:46-97
    for op in order:
        pil = enh[op](pil).enhance(factors[op])
    return np.array(pil)


# ------------------------------------------------------------------------------------------ random draws
def draw_train_params(rs: np.random.RandomState, crop=CROP_TRAIN, rotation=ROTATION, scale_range=SCALE_RANGE,
                      jitter=JITTER) -> Dict:
    """The np.random calls of transform_train (nuscenes_dataset_torch_new.py:247-249,284-285) followed by those of
    ColorJitter.get_params (transforms.py:457-474), in the reference's order."""
    scale = rs.uniform(scale_range[0], scale_range[1])
    angle = rs.uniform(-rotation, rotation)
    flip = bool(rs.uniform(0.0, 1.0) < 0.5)
    h_new, w_new = crop
    h_scaled, w_scaled = math.floor(h_new * scale), math.floor(w_new * scale)
    h_bound, w_bound = h_scaled - crop[0], w_scaled - crop[1]
    i = round(rs.uniform(0, h_bound))
    j = round(rs.uniform(0, w_bound))
    factors = [rs.uniform(max(0, 1 - jitter[0]), 1 + jitter[0]), rs.uniform(max(0, 1 - jitter[1]), 1 + jitter[1]),
               rs.uniform(max(0, 1 - jitter[2]), 1 + jitter[2])]
    order = [0, 1, 2]
    rs.shuffle(order)
    return dict(scale=scale, angle=angle, flip=flip, i=int(i), j=int(j), factors=factors, order=order)


# ------------------------------------------------------------------------------------------ the two transforms
def transform_train(image_u8: np.ndarray, lidar_i16: np.ndarray, radar_i16: np.ndarray, p: Dict, max_depth: float = 100.0,
                    crop=CROP_TRAIN) -> Dict[str, np.ndarray]:
    """image_u8 [H,W,3], lidar_i16 / radar_i16 [H,W] as stored in the exported h5 files.  Returns CHW float32 arrays named
    like the reference's output_dict (nuscenes_dataset_torch_new.py:397-412)."""
    rgb = np.array(image_u8).astype(np.float32)
    lidar = np.array(decode_depth(lidar_i16)).astype(np.float32)
    radar = np.array(decode_depth(radar_i16)).astype(np.float32)
    s, ang, flip, i, j = p["scale"], p["angle"], p["flip"], p["i"], p["j"]

    def geo(img, interp):
        img = rotate(img, ang)
        img = resize(img, np.float64(s), interp)
        img = img[i:i + crop[0], j:j + crop[1]]
        return np.fliplr(img) if flip else img

    rgb = geo(rgb, "bilinear")
    rgb = color_jitter(np.ascontiguousarray(rgb), p["factors"], p["order"])
    rgb = rgb / 255.
    lidar /= float(s)
    lidar = geo(lidar, "nearest")
    rgb = np.array(rgb).astype(np.float32).transpose(2, 0, 1).copy()
    lidar = np.array(lidar).astype(np.float32)[None]
    radar /= float(s)
    radar = np.array(geo(radar, "nearest")).astype(np.float32)[None]
    radar = radar.copy()
    radar[radar > max_depth] = 0                      # :373-374 (in place: the returned radar_depth is the masked one)
    inputs = np.concatenate((rgb, radar), axis=0)
    return dict(rgb=rgb, lidar_depth=lidar, radar_depth=radar, inputs=inputs, labels=lidar)


def transform_val(image_u8: np.ndarray, lidar_i16: np.ndarray, radar_i16: np.ndarray, max_depth: float = 100.0,
                  crop=CROP_VAL) -> Dict[str, np.ndarray]:
    rgb = np.array(image_u8).astype(np.float32)
    lidar = np.array(decode_depth(lidar_i16)).astype(np.float32)
    radar = np.array(decode_depth(radar_i16)).astype(np.float32)
    h, w = rgb.shape[:2]
    th, tw = crop
    i, j = int(round((h - th) / 2.)), int(round((w - tw) / 2.))                       # transforms.py:352-361

    def cc(img):
        return img[i:i + th, j:j + tw]

    rgb = cc(rgb) / 255.
    rgb = np.array(rgb).astype(np.float32).transpose(2, 0, 1).copy()
    lidar = np.array(cc(lidar)).astype(np.float32)[None]
    radar = np.array(cc(radar)).astype(np.float32)[None].copy()
    radar[radar > max_depth] = 0
    inputs = np.concatenate((rgb, radar), axis=0)
    return dict(rgb=rgb, lidar_depth=lidar, radar_depth=radar, inputs=inputs, labels=lidar)


def synth_sample(seed: int, h: int = 450, w: int = 800):
    """A synthetic exported sample: uint8 image, int16 x256 lidar (~5 % of the pixels) and radar (~100 returns)."""
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = (np.sin(yy / 37.0 + seed) * 60 + np.cos(xx / 53.0) * 60 + 128)[..., None] + rs.randint(-40, 40, (h, w, 3))
    image = np.clip(base, 0, 255).astype(np.uint8)
    lidar = np.where(rs.rand(h, w) < 0.05, rs.rand(h, w) * 100 + 1, 0.0)
    radar = np.where(rs.rand(h, w) < 100.0 / (h * w), rs.rand(h, w) * 120 + 1, 0.0)
    return image, (lidar * 256).astype(np.int16), (radar * 256).astype(np.int16)
