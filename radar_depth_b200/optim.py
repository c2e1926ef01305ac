"""Fused SGD over the flat parameter arena -- SURVEY.md 8(f)-1.  One kernel replaces the ~163 per-tensor updates of
torch.optim.SGD as configured by the reference (main.py:285-290: lr 0.01, momentum 0.9, weight_decay 1e-4,
dampening 0, nesterov False); ``adjust_learning_rate`` (utils.py:85-89) keeps working through ``param_groups``.
Parameters that do not live in an engine arena (the uncertainty scalars w_stage1/w_stage2 that main.py:166-172
registers on the model) are updated by the same rule with torch ops."""
from __future__ import annotations

from typing import List

import torch

from . import _lib
from .ops import ptr, stream_ptr


def _engines(model) -> List:
    found = []
    for m in model.modules():
        eng = getattr(m, "_engine", None)
        if eng is not None and eng.flat is not None:
            found.append(eng)
    return found


class FusedSGD:
    def __init__(self, model: torch.nn.Module, lr: float = 0.01, momentum: float = 0.9, weight_decay: float = 1e-4):
        self.model = model
        self.param_groups = [dict(lr=lr, momentum=momentum, weight_decay=weight_decay)]
        self._mom = {}
        self._loose = {}

    def zero_grad(self, set_to_none: bool = True):
        for p in self.model.parameters():
            p.grad = None

    def step(self):
        g = self.param_groups[0]
        lr, mo, wd = float(g["lr"]), float(g["momentum"]), float(g["weight_decay"])
        engines = _engines(self.model)
        if not engines:
            raise _lib.RdError("FusedSGD.step() before the first forward: no parameter arena exists yet")
        owned = set()
        for eng in engines:
            key = id(eng.flat)
            first = key not in self._mom
            if first:
                self._mom[key] = torch.zeros_like(eng.flat)
            _lib.call("rd_sgd", ptr(eng.flat), ptr(eng.gflat), ptr(self._mom[key]), eng.flat.numel(), lr, mo, wd,
                      1 if first else 0, stream_ptr())
            for _, p in eng.module.named_parameters():
                owned.add(id(p))
        with torch.no_grad():
            for p in self.model.parameters():
                if id(p) in owned or p.grad is None:
                    continue
                d = p.grad + wd * p
                buf = self._loose.get(id(p))
                buf = d.clone() if buf is None else buf.mul_(mo).add_(d)
                self._loose[id(p)] = buf
                p.add_(buf, alpha=-lr)
