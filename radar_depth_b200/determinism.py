"""Deterministic (fixed-order) reductions -- rd_set_deterministic in include/radar_depth_b200.h.

The reference's CPU path is reproducible run to run; with floating-point atomics in BatchNorm statistics, weight
gradients and loss sums this path would only be reproducible to ~1e-4 on small-population channels.  In deterministic
mode every such sum runs in a fixed order (per-warp / per-CTA partials added by index), so two executions of the same
step on the same inputs are bit-identical.

Defaults: the losses are always deterministic (two tiny extra launches); the network engines are deterministic in the
fp32 parity mode and use atomics in the bf16 throughput mode.  ``RD_DETERMINISTIC=1`` / ``=0`` forces the engines either way.
"""
from __future__ import annotations

import contextlib
import os
from typing import Dict, Optional

import torch

from . import _lib

LOSS_SCRATCH_BYTES = 1 << 20
_loss_scratch: Dict[torch.device, torch.Tensor] = {}


def engine_default(act_dtype: int) -> bool:
    env = os.environ.get("RD_DETERMINISTIC")
    if env is not None and env != "":
        return env != "0"
    return act_dtype == _lib.RD_F32


def loss_scratch(device) -> torch.Tensor:
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    t = _loss_scratch.get(device)
    if t is None:
        t = torch.zeros(LOSS_SCRATCH_BYTES, dtype=torch.uint8, device=device)
        _loss_scratch[device] = t
    return t


@contextlib.contextmanager
def mode(scratch: Optional[torch.Tensor]):
    """Launches issued inside the block run deterministically (scratch = uint8 device tensor) or, with None, with atomics.
    The flag is thread-local inside the library and restored on exit."""
    lib = _lib.load()
    if scratch is None:
        yield
        return
    _lib.check(lib.rd_set_deterministic(1, scratch.data_ptr(), scratch.numel() * scratch.element_size()), "rd_set_deterministic")
    try:
        yield
    finally:
        lib.rd_set_deterministic(0, None, 0)
