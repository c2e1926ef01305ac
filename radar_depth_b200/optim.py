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
    """The parameter set is fixed at construction, like torch.optim.SGD(model.parameters(), ...) (main.py:285-290 builds
    the optimizer after create_model has registered w_stage1 / w_stage2).  Which parameters live in an engine arena is
    worked out at the first step and re-derived only when an engine object changes (precision switch, re-adoption)."""

    def __init__(self, model: torch.nn.Module, lr: float = 0.01, momentum: float = 0.9, weight_decay: float = 1e-4):
        self.model = model
        self.param_groups = [dict(lr=lr, momentum=momentum, weight_decay=weight_decay)]
        self._params = list(model.parameters())
        self._mom = {}
        self._loose = {}
        self._eng_modules = None           # modules that own an engine, in model.modules() order
        self._engines: List = []
        self._loose_params: List[torch.nn.Parameter] = []
        self._arena_keys = None
        # set by ddp.allreduce_gradients(model, optimizer): gradients hold the SUM over ranks, 1/world is applied in the update
        self.grad_scale = 1.0

    def zero_grad(self, set_to_none: bool = True):
        for p in self._params:
            p.grad = None

    def _refresh(self):
        """(engines, loose parameters); cached while every engine object and its arena stay the same."""
        if self._eng_modules is not None:
            engs = [getattr(m, "_engine", None) for m in self._eng_modules]
            if all(e is not None and e.flat is not None for e in engs) and \
                    [(id(e), id(e.flat)) for e in engs] == self._arena_keys:
                return
        self._eng_modules = [m for m in self.model.modules()
                             if getattr(m, "_engine", None) is not None and m._engine.flat is not None]
        self._engines = [m._engine for m in self._eng_modules]
        self._arena_keys = [(id(e), id(e.flat)) for e in self._engines]
        owned = set()
        for eng in self._engines:
            for _, p in eng.module.named_parameters():
                owned.add(id(p))
        self._loose_params = [p for p in self._params if id(p) not in owned]

    def step(self):
        g = self.param_groups[0]
        lr, mo, wd = float(g["lr"]), float(g["momentum"]), float(g["weight_decay"])
        self._refresh()
        if not self._engines:
            raise _lib.RdError("FusedSGD.step() before the first forward: no parameter arena exists yet")
        for eng in self._engines:
            key = id(eng.flat)
            first = key not in self._mom
            if first:
                self._mom[key] = torch.zeros_like(eng.flat)
            if self.grad_scale == 1.0:
                _lib.call("rd_sgd", ptr(eng.flat), ptr(eng.gflat), ptr(self._mom[key]), eng.flat.numel(), lr, mo, wd,
                          1 if first else 0, stream_ptr())
            else:
                _lib.call("rd_sgd_scaled", ptr(eng.flat), ptr(eng.gflat), ptr(self._mom[key]), eng.flat.numel(), lr, mo, wd,
                          1 if first else 0, float(self.grad_scale), stream_ptr())
        with torch.no_grad():
            for p in self._loose_params:
                if p.grad is None:
                    continue
                d = p.grad * self.grad_scale + wd * p
                buf = self._loose.get(id(p))
                buf = d.clone() if buf is None else buf.mul_(mo).add_(d)
                self._loose[id(p)] = buf
                p.add_(buf, alpha=-lr)
