"""Thin torch-tensor wrappers over the C ABI (raw pointers + the current CUDA stream).  No compute happens in
Python or in torch here: torch only owns the device memory and the stream."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import RD_BF16, RD_F32, View

NULL_VIEW = View(None, 0, 0)


def act_torch_dtype(act_dtype: int):
    return torch.bfloat16 if act_dtype == RD_BF16 else torch.float32


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def view(t: torch.Tensor, coff: int = 0) -> View:
    """NHWC view of a contiguous [..., pitch] tensor starting at channel ``coff``."""
    assert t.is_cuda and t.is_contiguous()
    return View(t.data_ptr(), t.shape[-1], coff)


def ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous()
    return t.data_ptr()


def pack_weights(wflat: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    assert wflat.dtype == torch.float32 and idx.dtype == torch.int32
    if out is None:
        out = torch.empty(idx.numel(), dtype=torch.bfloat16, device=wflat.device)
    _lib.call("rd_pack_weights", ptr(wflat), ptr(idx), ptr(out), idx.numel(), stream_ptr())
    return out


def conv_fprop(plan, src: View, wpk, dst: View, ld=None, epi: int = 0, addend: Optional[View] = None,
               zsrc: Optional[View] = None, ep=None, stats: Optional[Tuple[torch.Tensor, int]] = None,
               max_ctas: Optional[int] = None, dbg=None, dbg_flags: int = 0, tail=None):
    """Launch one tcgen05 convolution program.  ``wpk``: packed weights (tensor or raw pointer);
    ``ld`` = (scale, shift, slope) fuses the producer's BN+activation on load; ``ep`` likewise for the
    activation-gradient epilogue (epi=1); ``stats`` = (fp64 tensor [2, stride], stride)."""
    p = type(plan.params).from_buffer_copy(plan.params)
    p.dbg = ptr(dbg) if dbg is not None else None
    p.dbg_flags = dbg_flags
    p.src = src
    p.dst = dst
    p.wpk = wpk if isinstance(wpk, int) else ptr(wpk)
    if ld is not None:
        p.ld_scale, p.ld_shift, p.ld_slope = ptr(ld[0]), ptr(ld[1]), float(ld[2])
    else:
        p.ld_scale, p.ld_shift, p.ld_slope = None, None, 1.0
    p.epi = epi
    p.addend = addend if addend is not None else NULL_VIEW
    p.zsrc = zsrc if zsrc is not None else NULL_VIEW
    if ep is not None:
        p.ep_scale, p.ep_shift, p.ep_slope = ptr(ep[0]), ptr(ep[1]), float(ep[2])
    else:
        p.ep_scale, p.ep_shift, p.ep_slope = None, None, 0.0
    if stats is not None:
        p.stats, p.stats_stride = ptr(stats[0]) if not isinstance(stats[0], int) else stats[0], int(stats[1])
    else:
        p.stats, p.stats_stride = None, 0
    if max_ctas is not None:
        p.max_ctas = max_ctas
    if tail is not None:
        p.tail = tail            # _lib.BnTail: BatchNorm finalisation fused into the kernel tail
    _lib.call("rd_conv_fprop", C.byref(p), stream_ptr())


def conv_wgrad(plan, gy: View, x: View, dw, ld=None, max_ctas: Optional[int] = None, dbg=None, dbg_flags: int = 0):
    p = type(plan.params).from_buffer_copy(plan.params)
    p.dbg = ptr(dbg) if dbg is not None else None
    p.dbg_flags = dbg_flags
    p.gy = gy
    p.x = x
    p.dw = dw if isinstance(dw, int) else ptr(dw)
    if ld is not None:
        p.ld_scale, p.ld_shift, p.ld_slope = ptr(ld[0]), ptr(ld[1]), float(ld[2])
    else:
        p.ld_scale, p.ld_shift, p.ld_slope = None, None, 1.0
    if max_ctas is not None:
        p.max_ctas = max_ctas
    _lib.call("rd_conv_wgrad", C.byref(p), stream_ptr())


def device_error() -> int:
    return _lib.load().rd_device_error(stream_ptr())
