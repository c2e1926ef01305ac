"""The reference's input pipeline on the GPU -- SURVEY.md 8(f)-4.

dataset/nuscenes_dataset_torch_new.py runs ``transform_train`` (:237-412) / ``transform_val`` (:415-560) per sample on CPU
DataLoader workers: h5 decode, scipy.ndimage rotation, scipy.misc.imresize (PIL) scaling, crop, flip, PIL ColorJitter,
/255, radar max-depth filter, concatenation.  At ~1700 images/s per GPU that is no longer affordable on host cores, so
this module does the same work for a whole batch in a handful of kernels (csrc/rd_dataset.cuh) on RAW exported samples
that were copied to the device as they are stored (uint8 image + two int16 depth maps: 3.1 bytes per pixel instead of the
20 bytes per pixel of the fp32 tensors the reference moves).

What stays on the host, per sample and per step (microseconds of numpy): the random draws, in the reference's order on a
``np.random.RandomState`` (seed it like the reference seeds ``np.random`` and the parameter sequence is the same), the
rotation matrix / offset exactly as scipy.ndimage.rotate computes them, and the small index / coefficient tables of PIL's
resampler for the 450 + 800 rows and columns of the crop window.  Everything per pixel happens on the device, bit for bit
what scipy + PIL produce (tests/test_dataset_gpu.py).

Scope: ``transform_mode="sparse-to-dense"`` with ``modality`` "rgb" | "rgbd" and ``sparsifier="radar"`` -- what main.py
trains with.  The lidar-sampling sparsifiers and the ``radar_filtered*`` variants (which need the exported point-index map)
raise NotImplementedError; the DORN mode's ImageNet normalisation is not part of this path.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .. import _lib
from ..ops import ptr, stream_ptr

CROP_TRAIN = (450, 800)          # config/config_nuscenes.py:43-50
CROP_VAL = (450, 800)
PRECISION_BITS = 32 - 8 - 2      # PIL Resample.c


def draw_train_params(rs: np.random.RandomState, crop=CROP_TRAIN, rotation: float = 5.0, scale_range=(1.0, 1.5),
                      jitter=(0.2, 0.2, 0.2)) -> Dict:
    """The random draws of transform_train (nuscenes_dataset_torch_new.py:247-249,284-285) and of ColorJitter.get_params
    (transforms.py:457-474) in the reference's order."""
    scale = rs.uniform(scale_range[0], scale_range[1])
    angle = rs.uniform(-rotation, rotation)
    flip = bool(rs.uniform(0.0, 1.0) < 0.5)
    h_scaled, w_scaled = math.floor(crop[0] * scale), math.floor(crop[1] * scale)
    i = round(rs.uniform(0, h_scaled - crop[0]))
    j = round(rs.uniform(0, w_scaled - crop[1]))
    factors = [rs.uniform(max(0, 1 - jitter[k]), 1 + jitter[k]) for k in range(3)]
    order = [0, 1, 2]
    rs.shuffle(order)
    return dict(scale=float(scale), angle=float(angle), flip=flip, i=int(i), j=int(j), factors=[float(f) for f in factors],
                order=[int(o) for o in order])


def rotation_affine(angle: float, H: int, W: int):
    """Matrix and offset of scipy.ndimage.rotate(img, angle, reshape=False) (scipy/ndimage/_interpolation.py: cosdg / sindg,
    centre-to-centre offset), in double precision like scipy."""
    from scipy import special
    c, s = special.cosdg(angle), special.sindg(angle)
    rot = np.array([[c, s], [-s, c]])
    shape = np.asarray([H, W])
    out_center = rot @ ((shape - 1) / 2)
    in_center = (shape - 1) / 2
    off = in_center - out_center
    return rot, off


def pil_bilinear_table(in_size: int, out_size: int, first: int, count: int) -> np.ndarray:
    """Rows ``first .. first+count`` of PIL's 8-bit bilinear resampling table for in_size -> out_size (Resample.c:
    precompute_coeffs + normalize_coeffs_8bpc): [count][5] = {first source index, taps, k0, k1, k2} (22-bit fixed point).
    Vectorised over the rows with the same double-precision operations, in the same order, as the C loop."""
    if out_size < in_size:
        raise ValueError("the GPU pipeline supports scale factors >= 1 (the reference draws from [1, 1.5])")
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ss = 1.0 / filterscale
    xx = np.arange(first, first + count, dtype=np.float64)
    center = (xx + 0.5) * scale
    xmin = np.maximum(np.trunc(center - support + 0.5), 0.0)
    xmax = np.minimum(np.trunc(center + support + 0.5), float(in_size))
    n = (xmax - xmin).astype(np.int64)
    assert n.min() > 0 and n.max() <= 3
    w = np.zeros((3, count), dtype=np.float64)
    ww = np.zeros(count, dtype=np.float64)
    for x in range(3):
        v = np.abs((x + xmin - center + 0.5) * ss)
        w[x] = np.where((v < 1.0) & (x < n), 1.0 - v, 0.0)
        ww = ww + w[x]
    out = np.zeros((count, 5), dtype=np.int32)
    out[:, 0], out[:, 1] = xmin.astype(np.int32), n.astype(np.int32)
    nz = ww != 0.0
    for x in range(3):
        k = np.where(nz, w[x] / np.where(nz, ww, 1.0), w[x])
        out[:, 2 + x] = np.where(x < n, np.trunc(0.5 + k * (1 << PRECISION_BITS)), 0.0).astype(np.int32)
    return out


def pil_nearest_table(in_size: int, out_size: int, first: int, count: int) -> np.ndarray:
    """Source index of PIL's NEAREST resize for output positions first .. first+count (Geometry.c ImagingScaleAffine: the
    source coordinate is a RUNNING SUM xo += in/out started at in/out * 0.5, truncated; np.cumsum adds sequentially)."""
    a = in_size / out_size
    steps = np.full(out_size, a, dtype=np.float64)
    steps[0] = 0.0 + a * 0.5
    xo = np.cumsum(steps)
    return np.trunc(xo).astype(np.int32)[first:first + count].copy()


class GpuInputPipeline:
    """Batch version of the reference's ``nuscenes_dataset_torch.__getitem__`` post-processing.

    ``pipe(images_u8, lidar_i16, radar_i16, params=None)`` with device tensors [B,H,W,3] uint8 / [B,H,W] int16 returns the
    reference's ``output_dict`` entries as device tensors: inputs [B,3|4,h,w], labels = lidar_depth [B,1,h,w], radar_depth
    [B,1,h,w] (masked by max_depth like the reference's in-place edit), rgb (a view of inputs[:, :3])."""

    def __init__(self, mode: str = "train", transform_mode: str = "sparse-to-dense", modality: str = "rgbd",
                 sparsifier: Optional[str] = None, max_depth: float = 100.0, seed: Optional[int] = None):
        if mode not in ("train", "val"):
            raise ValueError("[Error] Unknown dataset mode")
        if modality not in ("rgb", "rgbd"):
            raise ValueError("[Error] Unsupported modality. Consider ", ["rgb", "rgbd"])
        if transform_mode != "sparse-to-dense":
            raise NotImplementedError("only the sparse-to-dense transform mode (what main.py trains with) runs on the GPU")
        if sparsifier is None:
            sparsifier = "radar"
        if sparsifier not in ("uniform", "lidar_radar", "radar", "radar_filtered", "radar_filtered2"):
            raise ValueError("[Error] Invalid sparsifier.")
        if sparsifier != "radar":
            raise NotImplementedError(f"sparsifier {sparsifier!r} needs the lidar samplers / the exported point-index map; "
                                      "only 'radar' runs on the GPU")
        self.mode, self.modality, self.max_depth = mode, modality, float(max_depth)
        self.crop = CROP_TRAIN if mode == "train" else CROP_VAL
        self.output_size = list(self.crop)
        self.rs = np.random.RandomState(seed)

    # ------------------------------------------------------------------ host-side tables
    def _sample_struct(self, p: Optional[Dict], H: int, W: int):
        s = _lib.AugSample()
        ch, cw = self.crop
        if p is None:                                            # validation: centre crop (transforms.py:352-361)
            s.identity_rot, s.flip = 1, 0
            s.crop_i, s.crop_j = int(round((H - ch) / 2.)), int(round((W - cw) / 2.))
            s.depth_div = 1.0
            s.m00 = s.m11 = 1.0
            for k in range(3):
                s.op[k], s.factor[k] = k, 1.0
            return s, None, None
        rot, off = rotation_affine(p["angle"], H, W)
        s.m00, s.m01, s.m10, s.m11 = float(rot[0, 0]), float(rot[0, 1]), float(rot[1, 0]), float(rot[1, 1])
        s.off0, s.off1 = float(off[0]), float(off[1])
        s.identity_rot, s.flip = 0, 1 if p["flip"] else 0
        s.crop_i, s.crop_j = int(p["i"]), int(p["j"])
        s.depth_div = float(np.float32(p["scale"]))
        for k, op in enumerate(p["order"]):
            s.op[k], s.factor[k] = int(op), float(p["factors"][op])
        # scipy.misc.imresize: size = (array(im.size) * scale).astype(int)
        rw, rh = (np.array([W, H]) * np.float64(p["scale"])).astype(int)
        if s.crop_i + ch > rh or s.crop_j + cw > rw:
            raise _lib.RdError(f"crop window {ch}x{cw} at ({s.crop_i},{s.crop_j}) leaves the {rh}x{rw} resized image")
        bil = np.concatenate([pil_bilinear_table(H, int(rh), s.crop_i, ch), pil_bilinear_table(W, int(rw), s.crop_j, cw)])
        near = np.concatenate([pil_nearest_table(H, int(rh), s.crop_i, ch), pil_nearest_table(W, int(rw), s.crop_j, cw)])
        return s, bil, near

    def draw(self) -> Dict:
        return draw_train_params(self.rs, self.crop)

    # ------------------------------------------------------------------ the batch call
    def __call__(self, images: torch.Tensor, lidar: torch.Tensor, radar: torch.Tensor,
                 params: Optional[Sequence[Dict]] = None) -> Dict[str, torch.Tensor]:
        if not (images.is_cuda and lidar.is_cuda and radar.is_cuda):
            raise _lib.RdError("GpuInputPipeline needs the raw samples on the CUDA device (there is no CPU path)")
        if images.dtype != torch.uint8 or lidar.dtype != torch.int16 or radar.dtype != torch.int16:
            raise _lib.RdError("raw samples are uint8 images and int16 (x256) depth maps, as exported")
        B, H, W, C3 = images.shape
        if C3 != 3 or tuple(lidar.shape) != (B, H, W) or tuple(radar.shape) != (B, H, W):
            raise _lib.RdError("expected images [B,H,W,3], lidar / radar [B,H,W]")
        ch, cw = self.crop
        if H < ch or W < cw:
            raise _lib.RdError(f"raw samples ({H}x{W}) are smaller than the crop ({ch}x{cw})")
        images, lidar, radar = images.contiguous(), lidar.contiguous(), radar.contiguous()
        dev = images.device
        train = self.mode == "train"
        if train and params is None:
            params = [self.draw() for _ in range(B)]
        structs = (_lib.AugSample * B)()
        bils, nears = [], []
        for b in range(B):
            s, bil, near = self._sample_struct(params[b] if train else None, H, W)
            structs[b] = s
            if train:
                bils.append(bil)
                nears.append(near)
        samples = torch.frombuffer(bytearray(bytes(structs)), dtype=torch.uint8).to(dev)
        cin = 4 if self.modality == "rgbd" else 3
        inputs = torch.empty(B, cin, ch, cw, dtype=torch.float32, device=dev)
        labels = torch.empty(B, 1, ch, cw, dtype=torch.float32, device=dev)
        radar_out = torch.empty(B, 1, ch, cw, dtype=torch.float32, device=dev)
        st = stream_ptr()
        img8 = near_t = None
        if train:
            bil_t = torch.from_numpy(np.stack(bils)).to(dev)
            near_t = torch.from_numpy(np.stack(nears)).to(dev)
            img8 = torch.empty(B, ch, cw, 3, dtype=torch.uint8, device=dev)
            scratch = torch.empty(B * 32 + 64, dtype=torch.uint8, device=dev)
            off = (-scratch.data_ptr()) % 8
            _lib.call("rd_aug_rgb", ptr(images), ptr(samples), ptr(bil_t), scratch.data_ptr() + off, B, H, W, ch, cw, ptr(img8), st)
        _lib.call("rd_aug_pack", ptr(img8) if train else None, ptr(images), ptr(lidar), ptr(radar), ptr(samples),
                  ptr(near_t) if train else None, B, H, W, ch, cw, 0 if train else 1, 1 if cin == 4 else 0,
                  self.max_depth, ptr(inputs), ptr(labels), ptr(radar_out), st)
        out = dict(rgb=inputs[:, :3], lidar_depth=labels, radar_depth=radar_out, inputs=inputs, labels=labels)
        if train:
            out["params"] = list(params)
        return out
