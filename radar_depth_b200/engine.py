"""Forward / backward schedule of ResNet_latefusion (reference model/models.py:519-664; ResNet_latefusion2 in
model/multistage_model.py:123-276 is the same graph with a 2-channel depth stem) on the sm_100a kernels.

The engine owns
  * a flat fp32 parameter arena (the nn.Parameters of the module become views of it) and a flat gradient arena
    (``param.grad`` become views of it): one bucket for the fused SGD step and the NCCL all-reduce;
  * packed bf16 weight tiles for every convolution program (re-packed from the arena at the start of every
    forward, one gather kernel for the whole network);
  * NHWC activation buffers, BatchNorm vectors and fp64 statistic slots, all allocated once per input shape so that
    the launch sequence is static (CUDA-graph capturable);
  * two launch programs (lists of pre-bound C-ABI calls): forward and backward.

What is fused where (vs. the reference's one-library-call-per-module execution):
  * both 7x7 stems are one 4x4-tap convolution over the space-to-depth input;
  * relu(bn(conv(x))) is never materialised: convs store the raw output z and emit per-channel sum / sum-of-squares
    from their epilogue; the consumer applies the BN affine + activation while staging its operand;
  * Unpool + the two 5x5 convs of an UpProj block are one 4-phase sub-pixel program with N = 2*Cout;
  * activation-gradient masks and BatchNorm-backward reductions live in the data-gradient epilogues;
  * torch.cat (models.py:652) is a channel offset into one 640-channel buffer.
"""
from __future__ import annotations

import ctypes as C
import os
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib, convplan as cp, determinism
from ._lib import RD_BF16, RD_F32, View

# copies of every statistics array (rd_bn_tail.slots spreads same-address fp64 atomics).  Measured on B200 (same box,
# bench.py): 1 -> 11.02, 4 -> 11.01, 16 -> 11.16 ms/step: the conv CTAs do not finish close enough together for the atomics
# to queue up, and the finalising CTA pays for summing the copies.  Kept at 1.
STAT_SLOTS = int(os.environ.get("RD_STAT_SLOTS", "1"))
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def _p(t: torch.Tensor, off: int = 0) -> int:
    return t.data_ptr() + off * t.element_size()


def _v(t: torch.Tensor, coff: int = 0) -> View:
    return View(t.data_ptr(), t.shape[-1], coff)


NULLV = View(None, 0, 0)


class BNGroup:
    """Per-channel vectors of one or more BatchNorm2d layers laid side by side over a fused channel range."""

    def __init__(self, eng: "LatefusionEngine", members: List[Tuple[str, int, int]]):
        self.members = members                                   # (bn name, first channel, channels)
        self.C = sum(m[2] for m in members)
        d = eng.device
        self.vec = eng.hold(torch.zeros(7, self.C, dtype=torch.float32, device=d))    # scale shift mean invstd A B C
        self.scale, self.shift, self.mean, self.invstd, self.cA, self.cB, self.cC = self.vec.unbind(0)
        # [slot][row][C]; launches are handed slot 0, block b accumulates into slot b % STAT_SLOTS
        self.fstats = eng.alloc_stats(STAT_SLOTS * 2 * self.C).view(STAT_SLOTS, 2, self.C)[0]
        self.bstats = eng.alloc_stats(STAT_SLOTS * 3 * self.C).view(STAT_SLOTS, 3, self.C)[0]


class Launch:
    """One pre-bound C-ABI call.  ``meta`` = algorithmic work of the launch for the roofline report of bench.py:
    {"flops": 2*MAC without structural zeros, "bytes": tensors read + written once}."""
    __slots__ = ("fn", "args", "name", "meta", "lane", "sync")

    def __init__(self, name, fn, args, meta=None, lane=0, sync=None):
        self.name, self.fn, self.args, self.meta = name, fn, args, meta
        # lane 1 = the depth encoder's chain, which runs on a side stream next to the RGB encoder's chain (lane 0);
        # sync = "fork" on the first launch of the two chains, "join" on the first launch after them (see _run)
        self.lane, self.sync = lane, sync


class LatefusionEngine:
    def __init__(self, module: torch.nn.Module, in_channels: int, output_size, act_dtype: int = RD_BF16,
                 arch: str = "latefusion", decoder: str = "upproj"):
        """arch: "latefusion" (RGB + depth encoders, models.py:519-664) or "resnet" (one encoder over all input channels,
        models.py:233-303); decoder: "upproj" | "upconv" | "deconv2" | "deconv3" (models.py:135-230)."""
        assert arch in ("latefusion", "resnet") and decoder in ("upproj", "upconv", "deconv2", "deconv3"), (arch, decoder)
        self.arch, self.decoder = arch, decoder
        self.module = module
        self.in_channels = in_channels
        self.output_size = tuple(int(v) for v in output_size)
        self.act_dtype = act_dtype
        self.tdtype = torch.bfloat16 if act_dtype == RD_BF16 else torch.float32
        self.lib = _lib.load()
        self.device = None
        self.flat = None
        self.gflat = None
        self._plist = None                 # [(owner module, attribute, parameter, arena offset, numel, shape)]
        self._blist: List[torch.Tensor] = []
        self._gviews: List[torch.Tensor] = []
        self._adopt_checks = 0
        self.offs: Dict[str, Tuple[int, tuple]] = {}
        self.cfg = None
        self._stats_chunks: List[torch.Tensor] = []
        self._keep = []                    # keeps ctypes structs referenced by launches alive
        self._buffers = []                 # keeps every device buffer referenced by raw pointer alive
        self.use_graphs = os.environ.get("RADAR_DEPTH_B200_GRAPHS", "1") != "0"
        # fixed-order reductions (determinism.py): default on in the fp32 parity mode, off in the bf16 throughput mode
        self.det = determinism.engine_default(act_dtype)
        self.det_scratch = None
        # tile shapes: measured table (tuned_tiles.json) or, with use_tuned = False, the analytic cost model only
        self.use_tuned = os.environ.get("RD_USE_TUNED", "1") != "0"
        # SMs given to the depth encoder while it runs next to the RGB encoder (0 = one stream, every launch owns the GPU).
        # Measured on B200 (bench.py, same box, after the round-2 kernels): 20 SMs for the depth chain leave the RGB chain
        # 128 -- every N-block split of its launches (2 / 4 blocks -> 64 / 32 CTAs) then divides its tile counts evenly, which
        # 127 or 126 do not (a 32-tile layer4 launch needs two rounds on 31 CTAs per block): latefusion b=16 8.20 ms/step on
        # one stream, 8.42 (14), 7.96 (16), 7.84 (18), 7.65 (20), 8.05 (22), 8.12 (24); multistage b=8 12.13 on one stream,
        # 10.86 (16), 10.64 (18), 10.48 (20), 11.36 (21 ... 24), 11.47 (32).  (Round 1/early round 2, slower small-layer kernels:
        # the split only paid at b=8.)  Default: 20 at every batch size.
        env = os.environ.get("RD_DEPTH_SMS")
        self.depth_sms = int(env) if env not in (None, "") else None       # None = choose in configure()
        self._side = None
        self._graphs = {}

    # ------------------------------------------------------------------ parameter arena
    # Host-side bookkeeping runs on every forward / backward: it walks cached (owner module, attribute, parameter) triples
    # instead of module.named_parameters() (0.5 ms per traversal of the 150-module tree) and re-binds cached gradient
    # views instead of slicing the arena anew (163 x 2 torch ops); measured 4.0 -> 0.3 ms of Python per engine and step.
    def params_adopted(self) -> bool:
        if self.flat is None or self._plist is None or self.flat.device != self.device:
            return False
        self._adopt_checks += 1
        if self._adopt_checks % 128 == 0 and not self._param_set_unchanged():
            return False                   # a parameter was ADDED to the module since adoption (full walk, rarely)
        base = self.flat.data_ptr()
        for owner, attr, p, off, _, _ in self._plist:
            if owner._parameters.get(attr) is not p or p.data_ptr() != base + 4 * off:
                return False               # parameter object replaced, or its storage moved (.to(), .data = ...)
        for b in self._blist:
            if b.device != self.device:
                return False
        return True

    def _param_set_unchanged(self) -> bool:
        cur = [id(p) for _, p in self.module.named_parameters()]
        return cur == [id(e[2]) for e in self._plist] and \
            [id(b) for b in self.module.buffers()] == [id(b) for b in self._blist]

    def adopt(self, device) -> None:
        """Flatten the module's parameters into one fp32 arena; parameters become views of it."""
        self.device = torch.device(device)
        named = [(n, p) for n, p in self.module.named_parameters()]
        offs, total = OrderedDict(), 0
        for n, p in named:
            offs[n] = (total, tuple(p.shape))
            total += (p.numel() + 3) // 4 * 4                   # keep every parameter 16-byte aligned
        flat = torch.zeros(total, dtype=torch.float32, device=self.device)
        gflat = torch.zeros(total, dtype=torch.float32, device=self.device)
        with torch.no_grad():
            for n, p in named:
                off, shape = offs[n]
                flat[off:off + p.numel()].copy_(p.detach().reshape(-1).to(self.device, torch.float32))
                p.data = flat[off:off + p.numel()].view(shape)
                p.grad = None
        for b in list(self.module.buffers()):
            if b.device != self.device:
                raise RuntimeError("move the module to the CUDA device before the first forward (model.cuda())")
        self.flat, self.gflat, self.offs, self.nparams = flat, gflat, offs, total
        self._plist = []
        for n, p in named:
            owner_name, _, attr = n.rpartition(".")
            owner = self.module.get_submodule(owner_name) if owner_name else self.module
            assert owner._parameters[attr] is p
            self._plist.append((owner, attr, p, offs[n][0], p.numel(), offs[n][1]))
        self._blist = list(self.module.buffers())
        self._gviews = [gflat[off:off + n].view(shape) for _, _, _, off, n, shape in self._plist]
        self._adopt_checks = 0
        self.cfg = None

    def bind_grads(self) -> None:
        """param.grad = its view of the gradient arena (the SAME view object every step; a .grad that already aliases
        the right arena slot -- e.g. after optimizer.zero_grad(set_to_none=False) -- is left alone)."""
        for e, v in zip(self._plist, self._gviews):
            g = e[2].grad
            if g is not v and (g is None or g.data_ptr() != v.data_ptr()):
                e[2].grad = v

    def grads_bound(self) -> bool:
        for e, v in zip(self._plist, self._gviews):
            g = e[2].grad
            if g is not v and (g is None or g.data_ptr() != v.data_ptr()):
                return False
        return True

    # ------------------------------------------------------------------ allocation helpers
    def alloc_stats(self, n: int) -> torch.Tensor:
        off = self._stats_used
        self._stats_used += n
        assert self._stats_used <= self.stats.numel()
        return self.stats[off:off + n]

    def act(self, B, H, W, Cc) -> torch.Tensor:
        return self.hold(torch.zeros(B, H, W, Cc, dtype=self.tdtype, device=self.device))

    def hold(self, t):
        """Every device buffer referenced by raw pointer from a launch must stay alive as long as the program."""
        self._buffers.append(t)
        return t

    # ------------------------------------------------------------------ configuration for one input shape
    def configure(self, B: int, H: int, W: int) -> None:
        key = (B, H, W)
        if self.cfg is not None and self.cfg["key"] == key:
            return
        o = {k: v[0] for k, v in self.offs.items()}
        self._keep = []
        self._buffers = []
        self._graphs = {}
        self.stats = torch.zeros(1 << 20, dtype=torch.float64, device=self.device)
        self._stats_used = 0
        self.convs = []                    # (name, GConv, fplan, dplan, wplan)
        self._wpk_tables: List[np.ndarray] = []
        self._wpk_total = 0
        self._dw_total = 0
        self._scatter_p: List[np.ndarray] = []
        self._scatter_d: List[np.ndarray] = []
        self.fwd: List[Launch] = []
        self.bwd: List[Launch] = []
        self.fwd_eval: List[Launch] = []
        lib = self.lib
        act = self.act_dtype
        cin_d = self.in_channels - 3
        Cs = 4 if self.in_channels <= 4 else 8
        H2, W2 = (H + 1) // 2, (W + 1) // 2
        H4, W4 = (H2 + 1) // 2, (W2 + 1) // 2

        # deterministic mode: the conv kernels keep per-warp statistic arrays behind their rings (8 warps x [2][N] floats)
        det_reserve = 8 * 2 * 256 * 4 if self.det else 0
        det_bytes = [8 << 20]

        # -------- helpers that register a conv and emit launches
        depth_sms = self.depth_sms if self.depth_sms is not None else 20
        single = self.arch == "resnet"                    # one encoder over all input channels (models.py:233-303)
        par = depth_sms > 0 and not self.det and not single    # deterministic mode shares one scratch buffer: one stream
        self._par, self._depth_sms = par, depth_sms
        sm_of = {None: cp.NUM_SMS, 0: cp.NUM_SMS - depth_sms if par else cp.NUM_SMS, 1: depth_sms if par else cp.NUM_SMS}

        def reg(name, g: cp.GConv, src_hw, dst_hw, need_dgrad=True, need_wgrad=True, lane=None):
            sms = sm_of[lane]
            # measured tile table: entries are keyed by the SM budget of the launch (convplan.tuned_lookup)
            tuned = self.use_tuned
            fplan = cp.plan_fprop(g, B, src_hw, dst_hw, act, smem_reserve=det_reserve, use_tuned=tuned, sm_budget=sms)
            f_off = self._wpk_total
            self._wpk_tables.append(fplan.pack_idx)
            self._wpk_total += fplan.wpk_elems
            dplan, d_off = None, None
            if need_dgrad:
                dplan = cp.plan_fprop(g.transposed(), B, dst_hw, src_hw, act, smem_reserve=det_reserve, use_tuned=tuned,
                                      sm_budget=sms)
                d_off = self._wpk_total
                self._wpk_tables.append(dplan.pack_idx)
                self._wpk_total += dplan.wpk_elems
            wplan, w_off = None, None
            if need_wgrad:
                wplan = cp.plan_wgrad(g, B, src_hw, dst_hw, act, use_tuned=tuned, sm_budget=sms)
                w_off = self._dw_total
                self._scatter_p.append(wplan.scatter[0])
                self._scatter_d.append(wplan.scatter[1] + w_off)
                self._dw_total += wplan.dw_elems
                det_bytes.append(int(wplan.params.max_ctas) * wplan.dw_elems * 4)
            nnz = int(sum(int((t.widx >= 0).sum()) for t in g.taps))
            es = 2 if act == RD_BF16 else 4
            sH, sW = src_hw
            dH, dW = dst_hw
            in_b, out_b = B * sH * sW * g.Cx * es, B * dH * dW * g.N * es
            fpp, dpp = fplan.params, (dplan.params if dplan is not None else None)
            work = dict(f=dict(flops=2.0 * B * fpp.Hb * fpp.Wb * nnz, bytes=in_b + out_b + 2 * nnz),
                        d=dict(flops=2.0 * B * dpp.Hb * dpp.Wb * nnz, bytes=in_b + out_b + 2 * nnz) if dpp is not None else None,
                        w=dict(flops=2.0 * B * fpp.Hb * fpp.Wb * nnz, bytes=in_b + out_b + 4 * nnz))
            rec = dict(name=name, g=g, fplan=fplan, dplan=dplan, wplan=wplan, f_off=f_off, d_off=d_off, w_off=w_off, work=work,
                       lane=(lane or 0) if par else 0, wargs=(src_hw, dst_hw, tuned, sms))
            self.convs.append(rec)
            return rec

        self._pending = []                 # launches whose weight / dw pointers are patched after arenas exist

        def emit_conv(prog, rec, which, src: View, dst: View, ld=None, epi=0, addend=None, zsrc=None, ep=None,
                      stats=None, tag="", tail=None, ep_split=None, ep_slope_b=0.0):
            plan = rec["fplan"] if which == "f" else rec["dplan"]
            p = type(plan.params).from_buffer_copy(plan.params)
            p.src, p.dst = src, dst
            if ld is not None:
                p.ld_scale, p.ld_shift, p.ld_slope = _p(ld[0]), _p(ld[1]), float(ld[2])
            else:
                p.ld_scale, p.ld_shift, p.ld_slope = None, None, 1.0
            p.epi = epi
            p.addend = addend if addend is not None else NULLV
            p.zsrc = zsrc if zsrc is not None else NULLV
            if ep is not None:
                p.ep_scale, p.ep_shift, p.ep_slope = _p(ep[0]), _p(ep[1]), float(ep[2])
            p.ep_split = int(ep_split) if ep_split is not None else (1 << 30)
            p.ep_slope_b = float(ep_slope_b)
            if stats is not None:
                p.stats, p.stats_stride = _p(stats), stats.shape[1]
            if tail is not None:
                assert stats is not None
                p.tail = tail
            self._pending.append((p, "wpk", rec["f_off"] if which == "f" else rec["d_off"]))
            self._keep.append(p)
            meta = dict(rec["work"][which])
            extra = (1 if addend is not None else 0) + (1 if zsrc is not None else 0)       # tensors of the output's size
            if extra:
                es_ = 2 if act == RD_BF16 else 4
                meta["bytes"] += extra * B * p.dstH * p.dstW * p.N * p.nblk * es_
            prog.append(Launch(f"conv_{which}:{rec['name']}{tag}", lib.rd_conv_fprop, (C.byref(p),), meta, lane=rec["lane"]))
            return p

        def emit_wgrad(prog, rec, gy: View, x: View, ld=None):
            plan = rec["wplan"]
            if ld is not None and not self.det:
                # launches that transform the source tile in shared memory may have their own measured blocking (the dw
                # layout does not depend on it)
                if "wplan_bn" not in rec:
                    src_hw, dst_hw, tuned, sms = rec["wargs"]
                    has = tuned and cp.tuned_lookup(cp.tune_key("w", rec["g"], B, src_hw, dst_hw, act), sms, "|bn") is not None
                    rec["wplan_bn"] = cp.plan_wgrad(rec["g"], B, src_hw, dst_hw, act, use_tuned=True, sm_budget=sms, bn=True) if has else None
                if rec["wplan_bn"] is not None:
                    plan = rec["wplan_bn"]
                    assert plan.dw_elems == rec["wplan"].dw_elems
            p = type(plan.params).from_buffer_copy(plan.params)
            p.gy, p.x = gy, x
            if ld is not None:
                p.ld_scale, p.ld_shift, p.ld_slope = _p(ld[0]), _p(ld[1]), float(ld[2])
            else:
                p.ld_scale, p.ld_shift, p.ld_slope = None, None, 1.0
            self._pending.append((p, "dw", rec["w_off"]))
            self._keep.append(p)
            prog.append(Launch(f"wgrad:{rec['name']}", lib.rd_conv_wgrad, (C.byref(p),), dict(rec["work"]["w"]), lane=rec["lane"]))

        def bn_buffers(name):
            m = self.module.get_submodule(name)
            return m.running_mean, m.running_var, m.num_batches_tracked

        # BatchNorm finalisation is fused into the tail of the kernel that produces its statistics (rd_bn_tail: the
        # last CTA turns the fp64 sums into scale/shift or into the backward coefficients).  Only the eval-mode
        # forward, which has no statistics pass, still uses the stand-alone rd_bn_finalize launch.
        def new_tail(jobs, rows: int, Ctot: int) -> "_lib.BnTail":
            """rows x Ctot = shape of one copy of the statistics array the launch accumulates into."""
            t = _lib.BnTail()
            assert 1 <= len(jobs) <= _lib.RD_MAX_BN_JOBS
            t.counter = _p(self.alloc_stats(1))            # 8-byte slot of the fp64 arena, zeroed with the statistics
            t.njobs = len(jobs)
            t.slots, t.slot_stride = STAT_SLOTS, rows * Ctot
            for i, j in enumerate(jobs):
                t.job[i] = j
            self._keep.append(t)
            return t

        def fwd_jobs(grp: BNGroup, count: float):
            jobs = []
            for (name, c0, Cn) in grp.members:
                rm, rv, nbt = bn_buffers(name)
                j = _lib.BnJob()
                j.kind, j.C, j.count = 1, Cn, float(count)
                j.sum_a, j.sum_b = _p(grp.fstats[0], c0), _p(grp.fstats[1], c0)
                j.gamma, j.beta = _p(self.flat, o[name + ".weight"]), _p(self.flat, o[name + ".bias"])
                j.running_mean, j.running_var, j.nbt = _p(rm), _p(rv), _p(nbt)
                j.v0, j.v1, j.v2, j.v3 = _p(grp.scale, c0), _p(grp.shift, c0), _p(grp.mean, c0), _p(grp.invstd, c0)
                j.momentum, j.eps = BN_MOMENTUM, BN_EPS
                jobs.append(j)
            return jobs

        self._bn_eval_rows: List[List[int]] = []

        def emit_bn_eval(grp: BNGroup, count: float):
            # eval mode: scale/shift depend only on parameters and running statistics -> every BatchNorm of the network is
            # finalised by ONE launch at the head of the eval program (rd_bn_finalize_eval_multi, table built below)
            for (name, c0, Cn) in grp.members:
                rm, rv, nbt = bn_buffers(name)
                go, bo = o[name + ".weight"], o[name + ".bias"]
                self._bn_eval_rows.append([_p(self.flat, go), _p(self.flat, bo), _p(rm), _p(rv), _p(grp.scale, c0),
                                           _p(grp.shift, c0), _p(grp.mean, c0), _p(grp.invstd, c0), Cn])

        def bwd_job(grp: BNGroup, midx: int, sum_g_ptr: int, sum_gz_ptr: int, count: float):
            name, c0, Cn = grp.members[midx]
            go, bo = o[name + ".weight"], o[name + ".bias"]
            j = _lib.BnJob()
            j.kind, j.C, j.count = 2, Cn, float(count)
            j.sum_a, j.sum_b = sum_g_ptr, sum_gz_ptr
            j.gamma = _p(self.flat, go)
            j.v0, j.v1, j.v2, j.v3 = _p(grp.mean, c0), _p(grp.invstd, c0), _p(self.gflat, go), _p(self.gflat, bo)
            j.cA, j.cB, j.cC = _p(grp.cA, c0), _p(grp.cB, c0), _p(grp.cC, c0)
            return j

        def both(launch: Launch):          # identical in train and eval forward
            self.fwd.append(launch)
            self.fwd_eval.append(launch)

        def emit_conv_fwd(rec, src, dst, ld, grp: BNGroup, count: float):
            """conv + (training) statistics epilogue + fused BatchNorm finalisation; eval: conv, then running-stat affine."""
            emit_conv(self.fwd, rec, "f", src, dst, ld=ld, stats=grp.fstats, tail=new_tail(fwd_jobs(grp, count), 2, grp.C))
            emit_conv(self.fwd_eval, rec, "f", src, dst, ld=ld, stats=None, tag="(eval)")
            emit_bn_eval(grp, count)

        # ============================== buffers + forward program ==============================
        self.x_in = torch.zeros(B, self.in_channels, H, W, dtype=torch.float32, device=self.device)
        xs = self.act(B, H2, W2, 4 * Cs)
        es = 2 if act == RD_BF16 else 4                 # bytes per stored activation element (Launch.meta["bytes"])
        # The programs keep this launch (tools and the per-launch profile replay it from x_in), but forward() issues it
        # EAGERLY from the caller's tensor in front of the CUDA-graph replay and _run() skips it: the network input is read
        # once, where it lies, instead of being copied into a static buffer first (137 MB of reads + writes at b=16).
        both(Launch("input_pack", lib.rd_input_pack, (_p(self.x_in), _p(xs), B, self.in_channels, H, W, Cs, act),
                    dict(bytes=self.x_in.numel() * 4 + xs.numel() * es)))
        self._ipack_tail = (_p(xs), B, self.in_channels, H, W, Cs, act)

        # ---- stem
        if single:
            stem = reg("stem", cp.gconv_stem_single(o["conv1.weight"], self.in_channels), (H2, W2), (H2, W2), need_dgrad=False)
            Cst = 64
            g_stem = BNGroup(self, [("bn1", 0, 64)])
        else:
            stem = reg("stem", cp.gconv_stem(o["conv1.weight"], o["conv1_depth.weight"], cin_d), (H2, W2), (H2, W2),
                       need_dgrad=(self.in_channels > 4))
            Cst = 80
            g_stem = BNGroup(self, [("bn1", 0, 64), ("bn1_depth", 64, 16)])
        z_stem = self.act(B, H2, W2, Cst)
        n_stem = float(B * H2 * W2)
        emit_conv_fwd(stem, _v(xs), _v(z_stem), None, g_stem, n_stem)
        p_rgb = self.act(B, H4, W4, 64)
        p_d = None if single else self.act(B, H4, W4, 16)
        amax = self.hold(torch.zeros(B, H4, W4, Cst, dtype=torch.uint8, device=self.device))
        pool_out = (p_rgb.numel() + (0 if single else p_d.numel())) * es
        # bf16 training: the forward also keeps every window winner's PRE-activation value, so that the backward forms the
        # stem BatchNorm's sums at pool resolution and writes dz in one pass over the stem tensor (rd_maxpool_bwd_stats/apply)
        self._stem_bwd2 = act == RD_BF16 and os.environ.get("RD_STEM_BWD2", "1") != "0"
        zarg = self.hold(torch.zeros(B, H4, W4, Cst, dtype=torch.bfloat16, device=self.device)) if self._stem_bwd2 else None
        mp_args = lambda za: (_v(z_stem), _p(g_stem.scale), _p(g_stem.shift), B, H2, W2, Cst, 64, 0.0, 0.2, _v(p_rgb),
                              NULLV if single else _v(p_d), _p(amax), H4, W4, (_p(za) if za is not None else None), act)
        mp_bytes = z_stem.numel() * es + pool_out + amax.numel()
        self.fwd.append(Launch("maxpool", lib.rd_maxpool_fwd, mp_args(zarg),
                               dict(bytes=mp_bytes + (zarg.numel() * 2 if zarg is not None else 0))))
        self.fwd_eval.append(Launch("maxpool", lib.rd_maxpool_fwd, mp_args(None), dict(bytes=mp_bytes)))

        # ---- encoders
        enc_specs = [("", (64, 128, 256, 512), 64, p_rgb, 0)]
        if not single:
            enc_specs.append(("_depth", (16, 32, 64, 128), 16, p_d, 512))
        Ccat = 512 if single else 640
        h32, w32 = H4, W4
        for _ in range(3):
            h32, w32 = (h32 + 1) // 2, (w32 + 1) // 2
        concat = self.act(B, h32, w32, Ccat)
        d_concat = self.act(B, h32, w32, Ccat)
        blocks_all = []
        enc_seg = []                        # (lane, first index, end index) of each encoder chain in fwd / fwd_eval
        for lane, (suffix, widths, cin0, x0, cat_off) in enumerate(enc_specs):
            x_cur, cin = x0, cin0
            h, w = H4, W4
            blks = []
            seg0 = (len(self.fwd), len(self.fwd_eval))
            for li, cw in enumerate(widths, start=1):
                for bi in range(2):
                    pfx = f"layer{li}{suffix}.{bi}"
                    stride = 2 if (li > 1 and bi == 0) else 1
                    ho, wo = ((h + 1) // 2, (w + 1) // 2) if stride == 2 else (h, w)
                    ci = cin if bi == 0 else cw
                    c1 = reg(pfx + ".conv1", cp.gconv_standard(o[pfx + ".conv1.weight"], cw, ci, 3, stride, 1), (h, w), (ho, wo), lane=lane)
                    c2 = reg(pfx + ".conv2", cp.gconv_standard(o[pfx + ".conv2.weight"], cw, cw, 3, 1, 1), (ho, wo), (ho, wo), lane=lane)
                    ds = None
                    if stride == 2:
                        ds = reg(pfx + ".downsample.0", cp.gconv_standard(o[pfx + ".downsample.0.weight"], cw, ci, 1, 2, 0), (h, w), (ho, wo),
                                 lane=lane)
                    n = float(B * ho * wo)
                    z1, z2 = self.act(B, ho, wo, cw), self.act(B, ho, wo, cw)
                    b1 = BNGroup(self, [(pfx + ".bn1", 0, cw)])
                    b2 = BNGroup(self, [(pfx + ".bn2", 0, cw)])
                    emit_conv_fwd(c1, _v(x_cur), _v(z1), None, b1, n)
                    emit_conv_fwd(c2, _v(z1), _v(z2), (b1.scale, b1.shift, 0.0), b2, n)
                    zd, bd = None, None
                    if ds is not None:
                        zd = self.act(B, ho, wo, cw)
                        bd = BNGroup(self, [(pfx + ".downsample.1", 0, cw)])
                        emit_conv_fwd(ds, _v(x_cur), _v(zd), None, bd, n)
                    last = (li == 4 and bi == 1)
                    if last:
                        out_t, out_v = concat, _v(concat, cat_off)
                    else:
                        out_t = self.act(B, ho, wo, cw)
                        out_v = _v(out_t)
                    idv = _v(zd) if zd is not None else (_v(x_cur) if isinstance(x_cur, torch.Tensor) else x_cur)
                    both(Launch("join:" + pfx, lib.rd_bn_add_act,
                                (_v(z2), _p(b2.scale), _p(b2.shift), idv, _p(bd.scale) if bd else None,
                                 _p(bd.shift) if bd else None, out_v, int(B * ho * wo), cw, 0.0, act),
                                dict(bytes=3 * B * ho * wo * cw * es)))
                    blks.append(dict(pfx=pfx, c1=c1, c2=c2, ds=ds, b1=b1, b2=b2, bd=bd, z1=z1, z2=z2, zd=zd, x_in=x_cur,
                                     out_v=out_v, hw_in=(h, w), hw=(ho, wo), cw=cw, ci=ci, n=n, last=last, cat_off=cat_off))
                    x_cur, h, w = out_t, ho, wo
                cin = cw
            blocks_all.append(blks)
            enc_seg.append((lane, seg0, (len(self.fwd), len(self.fwd_eval))))
        # The two encoders are independent chains between the shared stem and the fusion convolution: with `par` the depth
        # chain (3.5 % of the FLOPs, ~45 % of the encoder's launches, every one of them latency-bound) runs on a side stream
        # on its own `depth_sms` SMs while the RGB chain keeps the rest -- the launches were planned with those SM budgets.
        if par:
            for lane, (f0, e0), (f1, e1) in enc_seg:
                for prog, a, b in ((self.fwd, f0, f1), (self.fwd_eval, e0, e1)):
                    for L in prog[a:b]:
                        L.lane = lane
            self.fwd[enc_seg[0][1][0]].sync = "fork"
            self.fwd_eval[enc_seg[0][1][1]].sync = "fork"

        # ---- fusion 1x1s
        nf = float(B * h32 * w32)
        cc2 = reg("conv2", cp.gconv_standard(o["conv2.weight"], 256, 512, 1, 1, 0), (h32, w32), (h32, w32))
        zc2 = self.act(B, h32, w32, 256)
        bc2 = BNGroup(self, [("bn2", 0, 256)])
        if single:
            cf = zf = bf = None
            emit_conv_fwd(cc2, _v(concat), _v(zc2), None, bc2, nf)           # layer4's output feeds conv2 directly
        else:
            cf = reg("conv_fusion", cp.gconv_standard(o["conv_fusion.weight"], 512, 640, 1, 1, 0), (h32, w32), (h32, w32))
            zf = self.act(B, h32, w32, 512)
            bf = BNGroup(self, [("bn_fusion", 0, 512)])
            emit_conv_fwd(cf, _v(concat), _v(zf), None, bf, nf)
            if par:
                self.fwd[-1].sync = "join"
                self.fwd_eval[-1].sync = "join"
            emit_conv_fwd(cc2, _v(zf), _v(zc2), (bf.scale, bf.shift, 1.0), bc2, nf)
        # graph cut of pnp_forward_front / pnp_forward_rear (models.py:669-707): everything up to here is the "front"
        split_f, split_fe = len(self.fwd), len(self.fwd_eval)
        self.bneck = self.hold(torch.zeros(B, 256, h32, w32, dtype=torch.float32, device=self.device))
        self.d_bneck = self.hold(torch.zeros(B, 256, h32, w32, dtype=torch.float32, device=self.device))

        # ---- decoder (UpProj x4)
        dec = []
        d_in_t, d_in_ld = zc2, (bc2.scale, bc2.shift, 1.0)
        h, w, cin = h32, w32, 256
        upproj = self.decoder == "upproj"
        for li in range(1, 5):
            pfx = f"decoder.layer{li}"
            co = cin // 2
            ho, wo = 2 * h, 2 * w
            n = float(B * ho * wo)
            if not upproj:
                # UpConv (unpool -> 5x5 conv -> BN -> ReLU, models.py:160-169) / DeConv (ConvTranspose2d -> BN -> ReLU,
                # models.py:140-151): ONE 4-phase program per stage; its BN + ReLU are applied by the next stage on load
                if self.decoder == "upconv":
                    up = reg(pfx + ".conv", cp.gconv_upconv(o[pfx + ".conv.weight"], cin, co), (h, w), (ho, wo))
                else:
                    k = int(self.decoder[6])
                    up = reg(pfx + f".deconv{k}", cp.gconv_deconv(o[pfx + f".deconv{k}.weight"], cin, co, k), (h, w), (ho, wo))
                z = self.act(B, ho, wo, co)
                grp = BNGroup(self, [(pfx + ".batchnorm", 0, co)])
                emit_conv_fwd(up, _v(d_in_t), _v(z), d_in_ld, grp, n)
                dec.append(dict(pfx=pfx, up=up, z=z, grp=grp, co=co, cin=cin, x_in=d_in_t, x_ld=d_in_ld, hw_in=(h, w),
                                hw=(ho, wo), n=n))
                d_in_t, d_in_ld = z, (grp.scale, grp.shift, 0.0)
                h, w, cin = ho, wo, co
                continue
            up = reg(pfx + ".up5x5", cp.gconv_upproj(o[pfx + ".upper_branch.conv1.weight"], o[pfx + ".bottom_branch.conv.weight"], cin, co),
                     (h, w), (ho, wo))
            c3 = reg(pfx + ".upper_branch.conv2", cp.gconv_standard(o[pfx + ".upper_branch.conv2.weight"], co, co, 3, 1, 1), (ho, wo), (ho, wo))
            zcat, zu2, out_t = self.act(B, ho, wo, 2 * co), self.act(B, ho, wo, co), self.act(B, ho, wo, co)
            gcat = BNGroup(self, [(pfx + ".upper_branch.batchnorm1", 0, co), (pfx + ".bottom_branch.batchnorm", co, co)])
            bu2 = BNGroup(self, [(pfx + ".upper_branch.batchnorm2", 0, co)])
            emit_conv_fwd(up, _v(d_in_t), _v(zcat), d_in_ld, gcat, n)
            emit_conv_fwd(c3, _v(zcat, 0), _v(zu2), (gcat.scale, gcat.shift, 0.0), bu2, n)
            both(Launch("join:" + pfx, lib.rd_bn_add_act,
                        (_v(zu2), _p(bu2.scale), _p(bu2.shift), _v(zcat, co), _p(gcat.scale, co), _p(gcat.shift, co),
                         _v(out_t), int(B * ho * wo), co, 0.0, act), dict(bytes=3 * B * ho * wo * co * es)))
            dec.append(dict(pfx=pfx, up=up, c3=c3, zcat=zcat, zu2=zu2, out=out_t, gcat=gcat, bu2=bu2, co=co, cin=cin,
                            x_in=d_in_t, x_ld=d_in_ld, hw_in=(h, w), hw=(ho, wo), n=n))
            d_in_t, d_in_ld = out_t, None
            h, w, cin = ho, wo, co
        Hd, Wd = h, w
        self.Hd, self.Wd = Hd, Wd
        if not upproj:
            # the 3x3 head reads a materialised activation: relu(bn(z)) of the last stage
            L = dec[-1]
            L["out"] = self.act(B, Hd, Wd, L["co"])
            both(Launch("bn_act:" + L["pfx"], lib.rd_bn_add_act,
                        (_v(L["z"]), _p(L["grp"].scale), _p(L["grp"].shift), NULLV, None, None, _v(L["out"]), int(B * Hd * Wd),
                         L["co"], 0.0, act), dict(bytes=2 * B * Hd * Wd * L["co"] * es)))
            d_in_t = L["out"]

        # ---- head
        OH, OW = self.output_size
        c3map = self.hold(torch.zeros(B, Hd, Wd, dtype=torch.float32, device=self.device))
        self.pred = torch.zeros(B, 1, OH, OW, dtype=torch.float32, device=self.device)
        w3 = _p(self.flat, o["conv3.weight"])
        both(Launch("head_conv", lib.rd_head_conv_fwd, (_v(d_in_t), w3, B, Hd, Wd, _p(c3map), act),
                    dict(bytes=d_in_t.numel() * es + c3map.numel() * 4, flops=2.0 * B * Hd * Wd * 144)))
        both(Launch("bilinear", lib.rd_bilinear_fwd, (_p(c3map), B, Hd, Wd, _p(self.pred), OH, OW),
                    dict(bytes=(c3map.numel() + self.pred.numel()) * 4)))

        # ============================== inference program (model.eval() + torch.no_grad(), main.py:584-595) ==============
        # with running statistics every BatchNorm is a fixed per-channel affine map, so it moves into the
        # epilogue of the convolution that feeds it (rd_conv_params.epi = 2) together with the residual add and the
        # activation: every conv then reads a MATERIALISED activation through the raw-TMA path (two epilogue groups, no
        # in-place transform hop), and the 12 residual-join launches disappear.  Same buffers as the training program.
        # (The collapse of conv_fusion o bn_fusion o conv2 o bn2 into one 640->256 1x1 needs a 256x512 @ 512x640 product of
        # the WEIGHTS before every evaluation pass; both 1x1s together take ~50 us at 11x38, so it is not done.)
        self.fwd_infer_body: Optional[List[Launch]] = None
        if upproj:
            fi: List[Launch] = [self.fwd[0]]                                  # input_pack
            ones = self.hold(torch.ones(Cst, dtype=torch.float32, device=self.device))
            zeros = self.hold(torch.zeros(Cst, dtype=torch.float32, device=self.device))
            emit_conv(fi, stem, "f", _v(xs), _v(z_stem), epi=2, ep=(g_stem.scale, g_stem.shift, 0.0), ep_split=64, ep_slope_b=0.2,
                      tag="(infer)")
            fi.append(Launch("maxpool", lib.rd_maxpool_fwd,
                             (_v(z_stem), _p(ones), _p(zeros), B, H2, W2, Cst, 64, 1.0, 1.0, _v(p_rgb),
                              NULLV if single else _v(p_d), _p(amax), H4, W4, None, act), dict(bytes=z_stem.numel() * es + pool_out + amax.numel())))
            enc0 = len(fi)
            for blks in blocks_all:
                for Bk in blks:
                    b1, b2, bd = Bk["b1"], Bk["b2"], Bk["bd"]
                    x_v = _v(Bk["x_in"])
                    emit_conv(fi, Bk["c1"], "f", x_v, _v(Bk["z1"]), epi=2, ep=(b1.scale, b1.shift, 0.0), tag="(infer)")
                    if bd is not None:
                        emit_conv(fi, Bk["ds"], "f", x_v, _v(Bk["zd"]), epi=2, ep=(bd.scale, bd.shift, 1.0), tag="(infer)")
                        idv = _v(Bk["zd"])
                    else:
                        idv = x_v
                    emit_conv(fi, Bk["c2"], "f", _v(Bk["z1"]), Bk["out_v"], epi=2, ep=(b2.scale, b2.shift, 0.0), addend=idv,
                              tag="(infer)")
            if single:
                emit_conv(fi, cc2, "f", _v(concat), _v(zc2), epi=2, ep=(bc2.scale, bc2.shift, 1.0), tag="(infer)")
            else:
                emit_conv(fi, cf, "f", _v(concat), _v(zf), epi=2, ep=(bf.scale, bf.shift, 1.0), tag="(infer)")
                if par:                    # two-lane encoder: the depth chain (lane 1, set by emit_conv) forks after the
                    fi[enc0].sync = "fork"  # max-pool and joins before the fusion convolution, as in the training program
                    fi[-1].sync = "join"
                emit_conv(fi, cc2, "f", _v(zf), _v(zc2), epi=2, ep=(bc2.scale, bc2.shift, 1.0), tag="(infer)")
            self._infer_split = len(fi)                                     # graph cut of pnp_forward_front / rear
            x_t = zc2
            for L in dec:
                co_, gcat_, bu2_ = L["co"], L["gcat"], L["bu2"]
                emit_conv(fi, L["up"], "f", _v(x_t), _v(L["zcat"]), epi=2, ep=(gcat_.scale, gcat_.shift, 0.0), ep_split=co_,
                          ep_slope_b=1.0, tag="(infer)")
                emit_conv(fi, L["c3"], "f", _v(L["zcat"], 0), _v(L["out"]), epi=2, ep=(bu2_.scale, bu2_.shift, 0.0),
                          addend=_v(L["zcat"], co_), tag="(infer)")
                x_t = L["out"]
            fi += self.fwd[-2:]                                               # head_conv, bilinear
            self.fwd_infer_body = fi

        # ============================== backward program ==============================
        bw = self.bwd
        self.dpred = torch.zeros(B, 1, OH, OW, dtype=torch.float32, device=self.device)
        dc3 = self.hold(torch.zeros(B, Hd, Wd, dtype=torch.float32, device=self.device))
        bw.append(Launch("bilinear_bwd", lib.rd_bilinear_bwd, (_p(self.dpred), B, Hd, Wd, _p(dc3), OH, OW),
                         dict(bytes=(self.dpred.numel() + dc3.numel()) * 4)))
        d_out = self.act(B, Hd, Wd, dec[-1]["co"])
        bw.append(Launch("head_conv_bwd", lib.rd_head_conv_bwd,
                         (_p(dc3), _v(dec[-1]["out"]), w3, B, Hd, Wd, _v(d_out), _p(self.gflat, o["conv3.weight"]), act),
                         dict(bytes=dc3.numel() * 4 + 2 * d_out.numel() * es, flops=4.0 * B * Hd * Wd * 144)))
        g_cur = None                       # non-UpProj decoders: g = d(loss)/d(bn output) masked by the ReLU, per stage
        for li in range(3, -1, -1):
            L = dec[li]
            co, (ho, wo), n = L["co"], L["hw"], L["n"]
            npix = int(B * ho * wo)
            if not upproj:
                grp = L["grp"]
                if li == 3:
                    g_cur = self.act(B, ho, wo, co)
                    tj = new_tail([bwd_job(grp, 0, _p(grp.bstats[0]), _p(grp.bstats[1]), n)], 3, grp.C)
                    bw.append(Launch("join_bwd:" + L["pfx"], lib.rd_join_bwd,
                                     (_v(d_out), _v(L["out"]), _v(L["z"]), NULLV, _v(g_cur), npix, co, 0.0,
                                      _p(grp.bstats[0]), _p(grp.bstats[1]), _p(grp.bstats[2]), C.byref(tj), act),
                                     dict(bytes=4 * npix * co * es)))
                bw.append(Launch("bn_bwd_apply:dec", lib.rd_bn_bwd_apply,
                                 (_v(g_cur), _v(L["z"]), _v(g_cur), _p(grp.cA), _p(grp.cB), _p(grp.cC), npix, co, act),
                                 dict(bytes=3 * npix * co * es)))
                emit_wgrad(bw, L["up"], _v(g_cur), _v(L["x_in"]), ld=L["x_ld"])
                hin, win = L["hw_in"]
                if li > 0:
                    P = dec[li - 1]
                    g_prev = self.act(B, hin, win, L["cin"])
                    emit_conv(bw, L["up"], "d", _v(g_cur), _v(g_prev), epi=1, zsrc=_v(P["z"]), ep=(P["grp"].scale, P["grp"].shift, 0.0),
                              stats=P["grp"].bstats[:2],
                              tail=new_tail([bwd_job(P["grp"], 0, _p(P["grp"].bstats[0]), _p(P["grp"].bstats[1]), P["n"])], 3, P["grp"].C))
                    g_cur = g_prev
                else:
                    g_c2 = self.act(B, h32, w32, 256)
                    emit_conv(bw, L["up"], "d", _v(g_cur), _v(g_c2), epi=1, zsrc=_v(zc2), ep=(bc2.scale, bc2.shift, 1.0),
                              stats=bc2.bstats[:2], tail=new_tail([bwd_job(bc2, 0, _p(bc2.bstats[0]), _p(bc2.bstats[1]), nf)], 3, bc2.C))
                    split_b = len(bw)
                continue
            gcat, bu2 = L["gcat"], L["bu2"]
            g_t = self.act(B, ho, wo, co)
            dzcat = self.act(B, ho, wo, 2 * co)
            dzu2 = self.act(B, ho, wo, co)
            tj = new_tail([bwd_job(bu2, 0, _p(bu2.bstats[0]), _p(bu2.bstats[1]), n),
                           bwd_job(gcat, 1, _p(bu2.bstats[0]), _p(bu2.bstats[2]), n)], 3, bu2.C)
            bw.append(Launch("join_bwd:" + L["pfx"], lib.rd_join_bwd,
                             (_v(d_out), _v(L["out"]), _v(L["zu2"]), _v(L["zcat"], co), _v(g_t), npix, co, 0.0,
                              _p(bu2.bstats[0]), _p(bu2.bstats[1]), _p(bu2.bstats[2]), C.byref(tj), act),
                             dict(bytes=5 * npix * co * es)))
            bw.append(Launch("bn_bwd_apply:bottom", lib.rd_bn_bwd_apply,
                             (_v(g_t), _v(L["zcat"], co), _v(dzcat, co), _p(gcat.cA, co), _p(gcat.cB, co), _p(gcat.cC, co), npix, co, act),
                             dict(bytes=3 * npix * co * es)))
            bw.append(Launch("bn_bwd_apply:u2", lib.rd_bn_bwd_apply,
                             (_v(g_t), _v(L["zu2"]), _v(dzu2), _p(bu2.cA), _p(bu2.cB), _p(bu2.cC), npix, co, act),
                             dict(bytes=3 * npix * co * es)))
            emit_wgrad(bw, L["c3"], _v(dzu2), _v(L["zcat"], 0), ld=(gcat.scale, gcat.shift, 0.0))
            emit_conv(bw, L["c3"], "d", _v(dzu2), _v(dzcat, 0), epi=1, zsrc=_v(L["zcat"], 0), ep=(gcat.scale, gcat.shift, 0.0),
                      stats=gcat.bstats[:2], tail=new_tail([bwd_job(gcat, 0, _p(gcat.bstats[0]), _p(gcat.bstats[1]), n)], 3, gcat.C))
            bw.append(Launch("bn_bwd_apply:u1", lib.rd_bn_bwd_apply,
                             (_v(dzcat, 0), _v(L["zcat"], 0), _v(dzcat, 0), _p(gcat.cA), _p(gcat.cB), _p(gcat.cC), npix, co, act),
                             dict(bytes=3 * npix * co * es)))
            emit_wgrad(bw, L["up"], _v(dzcat), _v(L["x_in"]), ld=L["x_ld"])
            if li > 0:
                hin, win = L["hw_in"]
                d_prev = self.act(B, hin, win, L["cin"])
                emit_conv(bw, L["up"], "d", _v(dzcat), _v(d_prev))
                d_out = d_prev
            else:
                g_c2 = self.act(B, h32, w32, 256)
                emit_conv(bw, L["up"], "d", _v(dzcat), _v(g_c2), epi=1, zsrc=_v(zc2), ep=(bc2.scale, bc2.shift, 1.0),
                          stats=bc2.bstats[:2], tail=new_tail([bwd_job(bc2, 0, _p(bc2.bstats[0]), _p(bc2.bstats[1]), nf)], 3, bc2.C))
                split_b = len(bw)          # g_c2 = d(loss)/d(bn2 output) here (slope 1: the activation mask is the identity)
        npf = int(B * h32 * w32)
        bw.append(Launch("bn_bwd_apply:bn2", lib.rd_bn_bwd_apply,
                         (_v(g_c2), _v(zc2), _v(g_c2), _p(bc2.cA), _p(bc2.cB), _p(bc2.cC), npf, 256, act), dict(bytes=3 * npf * 256 * es)))
        if single:
            emit_wgrad(bw, cc2, _v(g_c2), _v(concat))
            emit_conv(bw, cc2, "d", _v(g_c2), _v(d_concat))
        else:
            emit_wgrad(bw, cc2, _v(g_c2), _v(zf), ld=(bf.scale, bf.shift, 1.0))
            g_f = self.act(B, h32, w32, 512)
            emit_conv(bw, cc2, "d", _v(g_c2), _v(g_f), epi=1, zsrc=_v(zf), ep=(bf.scale, bf.shift, 1.0), stats=bf.bstats[:2],
                      tail=new_tail([bwd_job(bf, 0, _p(bf.bstats[0]), _p(bf.bstats[1]), nf)], 3, bf.C))
            bw.append(Launch("bn_bwd_apply:bn_fusion", lib.rd_bn_bwd_apply,
                             (_v(g_f), _v(zf), _v(g_f), _p(bf.cA), _p(bf.cB), _p(bf.cC), npf, 512, act), dict(bytes=3 * npf * 512 * es)))
            emit_wgrad(bw, cf, _v(g_f), _v(concat))
            emit_conv(bw, cf, "d", _v(g_f), _v(d_concat))

        dpool = []
        cut_enc, cut_l4 = len(bw), None     # gradient-bucket boundaries of the backward program (see _grad_buckets)
        for lane, blks in enumerate(blocks_all):
            bseg0 = len(bw)
            d_out_v = _v(d_concat, blks[-1]["cat_off"])
            for Bk in reversed(blks):
                cw, (ho, wo), (hi, wi), n = Bk["cw"], Bk["hw"], Bk["hw_in"], Bk["n"]
                npix = int(B * ho * wo)
                b1, b2, bd = Bk["b1"], Bk["b2"], Bk["bd"]
                g_t, dz2, g1 = self.act(B, ho, wo, cw), self.act(B, ho, wo, cw), self.act(B, ho, wo, cw)
                jobs = [bwd_job(b2, 0, _p(b2.bstats[0]), _p(b2.bstats[1]), n)]
                if bd:
                    jobs.append(bwd_job(bd, 0, _p(b2.bstats[0]), _p(b2.bstats[2]), n))
                tj = new_tail(jobs, 3, b2.C)
                bw.append(Launch("join_bwd:" + Bk["pfx"], lib.rd_join_bwd,
                                 (d_out_v, Bk["out_v"], _v(Bk["z2"]), _v(Bk["zd"]) if bd else NULLV, _v(g_t), npix, cw, 0.0,
                                  _p(b2.bstats[0]), _p(b2.bstats[1]), _p(b2.bstats[2]), C.byref(tj), act),
                                 dict(bytes=(5 if bd else 4) * npix * cw * es)))
                bw.append(Launch("bn_bwd_apply:bn2", lib.rd_bn_bwd_apply,
                                 (_v(g_t), _v(Bk["z2"]), _v(dz2), _p(b2.cA), _p(b2.cB), _p(b2.cC), npix, cw, act),
                                 dict(bytes=3 * npix * cw * es)))
                dzd = None
                if bd:
                    dzd = self.act(B, ho, wo, cw)
                    bw.append(Launch("bn_bwd_apply:ds", lib.rd_bn_bwd_apply,
                                     (_v(g_t), _v(Bk["zd"]), _v(dzd), _p(bd.cA), _p(bd.cB), _p(bd.cC), npix, cw, act),
                                     dict(bytes=3 * npix * cw * es)))
                emit_wgrad(bw, Bk["c2"], _v(dz2), _v(Bk["z1"]), ld=(b1.scale, b1.shift, 0.0))
                emit_conv(bw, Bk["c2"], "d", _v(dz2), _v(g1), epi=1, zsrc=_v(Bk["z1"]), ep=(b1.scale, b1.shift, 0.0),
                          stats=b1.bstats[:2], tail=new_tail([bwd_job(b1, 0, _p(b1.bstats[0]), _p(b1.bstats[1]), n)], 3, b1.C))
                bw.append(Launch("bn_bwd_apply:bn1", lib.rd_bn_bwd_apply,
                                 (_v(g1), _v(Bk["z1"]), _v(g1), _p(b1.cA), _p(b1.cB), _p(b1.cC), npix, cw, act),
                                 dict(bytes=3 * npix * cw * es)))
                x_in_v = _v(Bk["x_in"])
                emit_wgrad(bw, Bk["c1"], _v(g1), x_in_v)
                dx = self.act(B, hi, wi, Bk["ci"])
                Bk["dx_t"], Bk["g_t"], Bk["dz2_t"], Bk["g1_t"] = dx, g_t, dz2, g1
                if bd:
                    emit_conv(bw, Bk["c1"], "d", _v(g1), _v(dx))
                    emit_wgrad(bw, Bk["ds"], _v(dzd), x_in_v)
                    emit_conv(bw, Bk["ds"], "d", _v(dzd), _v(dx), addend=_v(dx))
                else:
                    emit_conv(bw, Bk["c1"], "d", _v(g1), _v(dx), addend=_v(g_t))
                d_out_v = _v(dx)
                if lane == 0 and Bk["pfx"] == "layer4.0":
                    cut_l4 = len(bw)
            dpool.append(d_out_v)
            if par:
                for L in bw[bseg0:]:
                    L.lane = lane
                if lane == 0:
                    bw[bseg0].sync = "fork"

        gz_stem = self.act(B, H2, W2, Cst)
        stem_jobs = [bwd_job(g_stem, 0, _p(g_stem.bstats[0]), _p(g_stem.bstats[1]), n_stem)]
        if not single:
            stem_jobs.append(bwd_job(g_stem, 1, _p(g_stem.bstats[0], 64), _p(g_stem.bstats[1], 64), n_stem))
        tj = new_tail(stem_jobs, 3, g_stem.C)
        dp_b = NULLV if single else dpool[1]
        if self._stem_bwd2:
            bw.append(Launch("maxpool_bwd_stats", lib.rd_maxpool_bwd_stats,
                             (dpool[0], dp_b, _p(zarg), _p(g_stem.scale), _p(g_stem.shift), B, H4, W4, Cst, 64, 0.0, 0.2,
                              _p(g_stem.bstats[0]), _p(g_stem.bstats[1]), C.byref(tj)),
                             dict(bytes=pool_out + zarg.numel() * 2), sync="join" if par else None))
            bw.append(Launch("maxpool_bwd_apply", lib.rd_maxpool_bwd_apply,
                             (dpool[0], dp_b, _p(amax), _v(z_stem), _p(g_stem.scale), _p(g_stem.shift), _p(g_stem.cA), _p(g_stem.cB),
                              _p(g_stem.cC), B, H2, W2, Cst, 64, 0.0, 0.2, H4, W4, _v(gz_stem)),
                             dict(bytes=pool_out + amax.numel() + 2 * z_stem.numel() * es)))
        else:
            bw.append(Launch("maxpool_bwd", lib.rd_maxpool_bwd,
                             (dpool[0], dp_b, _p(amax), _v(z_stem), _p(g_stem.scale), _p(g_stem.shift), B, H2, W2,
                              Cst, 64, 0.0, 0.2, H4, W4, _v(gz_stem), _p(g_stem.bstats[0]), _p(g_stem.bstats[1]), C.byref(tj), act),
                             dict(bytes=pool_out + amax.numel() + 2 * z_stem.numel() * es),
                             sync="join" if par else None))
            bw.append(Launch("bn_bwd_apply:stem", lib.rd_bn_bwd_apply,
                             (_v(gz_stem), _v(z_stem), _v(gz_stem), _p(g_stem.cA), _p(g_stem.cB), _p(g_stem.cC), int(B * H2 * W2), Cst, act),
                             dict(bytes=3 * z_stem.numel() * es)))
        emit_wgrad(bw, stem, _v(gz_stem), _v(xs))
        self.dxs = None
        if self.in_channels > 4:
            self.dxs = self.act(B, H2, W2, 4 * Cs)
            emit_conv(bw, stem, "d", _v(gz_stem), _v(self.dxs))

        # ============================== arenas that depend on the totals ==============================
        self.det_scratch = torch.zeros(max(det_bytes), dtype=torch.uint8, device=self.device) if self.det else None
        self.wpk = torch.zeros(self._wpk_total, dtype=torch.bfloat16, device=self.device)
        self.dw = torch.zeros(max(self._dw_total, 1), dtype=torch.float32, device=self.device)
        pack_np = np.concatenate(self._wpk_tables)
        # compact (base, stride) table of rd_pack_weights_g8: 8 bytes per 8 outputs instead of 32
        grp, fbk = cp.compact_pack_table(pack_np)
        self.pack_groups = self.hold(torch.from_numpy(grp).to(self.device))
        self.pack_fallback = self.hold(torch.from_numpy(fbk).to(self.device))
        self.pack_elems = int(pack_np.size)
        unpack = np.full(self.nparams, -1, dtype=np.int32)
        unpack[np.concatenate(self._scatter_p)] = np.concatenate(self._scatter_d).astype(np.int32)
        self.unpack_idx = torch.from_numpy(unpack).to(self.device)
        for p, kind, off in self._pending:
            if kind == "wpk":
                p.wpk = _p(self.wpk, off)
            else:
                p.dw = _p(self.dw, off)
        self._pending = []
        self._wpk_tables = []
        pack = Launch("pack_weights", lib.rd_pack_weights_g8,
                      (_p(self.flat), _p(self.pack_groups), _p(self.pack_fallback), _p(self.wpk), self.pack_elems // 8, None),
                      dict(bytes=self.pack_elems * (1 + 4 + 2)))
        self.fwd.insert(0, pack)
        self.fwd_eval.insert(0, pack)
        self.bn_eval_table = self.hold(torch.tensor(self._bn_eval_rows, dtype=torch.int64, device=self.device))
        self.fwd_eval.insert(1, Launch("bn_fin_eval_all", lib.rd_bn_finalize_eval_multi,
                                       (_p(self.bn_eval_table), len(self._bn_eval_rows), BN_EPS)))
        # inference program: same two head launches; decoders without the folded program fall back to the eval program
        self.fwd_infer = (self.fwd_eval[:2] + self.fwd_infer_body) if self.fwd_infer_body is not None else self.fwd_eval
        if self.fwd_infer_body is not None and os.environ.get("RD_INFER_PACK_HASH", "1") != "0":
            # the weights rarely change between inference forwards: re-pack only when the arena's content hash changed
            nchunks = 1184                                    # 8 blocks per SM
            self._whash = self.hold(torch.zeros(nchunks, dtype=torch.int64, device=self.device))
            self._wdirty = self.hold(torch.zeros(1, dtype=torch.int32, device=self.device))
            hashl = Launch("weights_hash", lib.rd_weights_hash, (_p(self.flat), self.nparams, _p(self._whash), nchunks, _p(self._wdirty)),
                           dict(bytes=self.nparams * 4))
            packif = Launch("pack_weights_if", lib.rd_pack_weights_g8,
                            (_p(self.flat), _p(self.pack_groups), _p(self.pack_fallback), _p(self.wpk), self.pack_elems // 8, _p(self._wdirty)),
                            dict(bytes=self.pack_elems * (1 + 4 + 2)))
            self.fwd_infer = [hashl, packif, self.fwd_eval[1]] + self.fwd_infer_body
        bw.append(Launch("unpack_grads", lib.rd_unpack_grads, (_p(self.dw), _p(self.unpack_idx), _p(self.gflat), self.nparams),
                         dict(bytes=self.nparams * (4 + 4 + 8))))
        self._grad_buckets(cut_enc, cut_l4, single)
        self.stats_used = self.stats[:self._stats_used]
        # programs of the graph cut (pack_weights / bn_fin_eval_all were inserted at the heads above)
        exp = Launch("feature_export", lib.rd_feature_export,
                     (_v(zc2), _p(bc2.scale), _p(bc2.shift), _p(self.bneck), B, h32, w32, 256, act))
        imp = Launch("feature_import", lib.rd_feature_import, (_p(self.bneck), _v(zc2), B, h32, w32, 256, act))
        self.front = self.fwd[:split_f + 1] + [exp]
        self.front_eval = self.fwd_eval[:split_fe + 2] + [exp]
        self.rear = [pack, imp] + self.fwd[split_f + 1:]
        self.rear_eval = self.fwd_eval[:2] + [imp] + self.fwd_eval[split_fe + 2:]
        # the same cut in the inference program (bn2 is already applied by conv2's epilogue: export without the affine map)
        self.front_infer = self.rear_infer = None
        if self.fwd_infer_body is not None:
            exp_raw = Launch("feature_export(infer)", lib.rd_feature_export, (_v(zc2), None, None, _p(self.bneck), B, h32, w32, 256, act))
            self.front_infer = self.fwd_eval[:2] + self.fwd_infer_body[:self._infer_split] + [exp_raw]
            self.rear_infer = self.fwd_eval[:2] + [imp] + self.fwd_infer_body[self._infer_split:]
        self.rear_bwd = bw[:split_b] + [Launch("feature_export(grad)", lib.rd_feature_export,
                                               (_v(g_c2), None, None, _p(self.d_bneck), B, h32, w32, 256, act))]
        self.bc2 = bc2
        # BatchNorm-backward jobs of the backward program: (job view, batch count).  See _bn_backward_mode.
        self._bwd_jobs = []
        for L in bw:
            for a in L.args:
                obj = getattr(a, "_obj", None)
                tail = obj if isinstance(obj, _lib.BnTail) else getattr(obj, "tail", None)
                if isinstance(tail, _lib.BnTail):
                    for i in range(tail.njobs):
                        if tail.job[i].kind == 2:
                            self._bwd_jobs.append((tail.job[i], float(tail.job[i].count)))
        self.cfg = dict(key=key, B=B, H=H, W=W, Hd=Hd, Wd=Wd, h32=h32, w32=w32)
        self.dec, self.blocks_all = dec, blocks_all
        self.dbg = dict(z_stem=z_stem, gz_stem=gz_stem, p_rgb=p_rgb, p_d=p_d, amax=amax, xs=xs, concat=concat, d_concat=d_concat,
                        zf=zf, zc2=zc2, g_stem=g_stem)

    # ------------------------------------------------------------------ gradient buckets (overlapped all-reduce, ddp.py)
    def _grad_buckets(self, cut_enc: int, cut_l4: int, single: bool) -> None:
        """The backward program in three segments whose parameter gradients are FINAL when the segment ends, in the order
        the backward produces them ("decoder grads are ready first, RGB stem last"):
          0: head, decoder, conv2, fusion (arena tail)        -- ends where the encoders' backward starts;
          1: RGB layer4 (8.4 M of the 14.7 M parameters)      -- ends after layer4.0's weight gradients;
          2: everything else (RGB layers 1-3, depth encoder, stems).
        Each segment ends with the rd_unpack_grads launches of its own arena ranges, so ddp.py can start the all-reduce of
        a bucket while the next segment computes.  self.bwd (one segment, one unpack launch) stays the default program."""
        names = list(self.offs)
        first_tail = "conv2.weight" if single else "conv_fusion.weight"
        t0 = self.offs[first_tail][0]
        l4 = [i for i, n in enumerate(names) if n.startswith("layer4.")]
        a4 = self.offs[names[l4[0]]][0]
        b4 = self.offs[names[l4[-1] + 1]][0]
        rest = [(0, a4)] + ([(b4, t0)] if b4 < t0 else [])
        self.grad_ranges = [[(t0, self.nparams)], [(a4, b4)], rest]
        body = self.bwd[:-1]                                   # without the whole-arena unpack launch
        cuts = [0, cut_enc, cut_l4, len(body)]

        def unpack(a, b):
            return Launch(f"unpack_grads[{a}:{b}]", self.lib.rd_unpack_grads,
                          (_p(self.dw), _p(self.unpack_idx, a), _p(self.gflat, a), b - a), dict(bytes=(b - a) * 16))
        self.bwd_segments = [body[cuts[k]:cuts[k + 1]] + [unpack(a, b) for a, b in self.grad_ranges[k]] for k in range(3)]

    # ------------------------------------------------------------------ execution
    def _run(self, prog: List[Launch]):
        # Any program that re-packs the weights unconditionally makes the inference program's stored content hash (which
        # describes the weights of ITS last re-packing) meaningless: it is cleared before the next hashed forward.
        names = prog[0].name if prog else ""
        if names == "weights_hash":
            if getattr(self, "_whash_stale", False):
                self._whash.zero_()
                self._whash_stale = False
        elif any(L.name == "pack_weights" for L in prog[:2]):
            self._whash_stale = True
        # (Weight-gradient launches on a second, event-forked stream were tried: 10.91 vs 10.92 ms/step on B200 --
        # full-grid kernels with ~200 KB of shared memory per CTA do not overlap; the program stays single-stream.)
        main = torch.cuda.current_stream()
        st = main.cuda_stream
        lib = self.lib
        side = None
        # a program slice that starts inside the two-lane region (segmented backward) forks at its first launch
        fork_first = any(L.lane == 1 for L in prog) and not any(L.sync == "fork" for L in prog)
        with determinism.mode(self.det_scratch if self.det else None):
            for i, L in enumerate(prog):
                if L.name == "input_pack" and self.use_graphs:
                    continue                   # issued eagerly from the caller's tensor (see _pack_input)
                if L.sync == "fork" or (fork_first and i == 0):
                    # the side stream joins here (under CUDA-graph capture this event edge forks the graph)
                    if self._side is None or self._side.device != main.device:
                        self._side = torch.cuda.Stream(device=main.device)
                    side = self._side
                    ev = torch.cuda.Event()
                    ev.record(main)
                    side.wait_event(ev)
                elif L.sync == "join" and side is not None:
                    ev = torch.cuda.Event()
                    ev.record(side)
                    main.wait_event(ev)
                    side = None
                rc = L.fn(*L.args, side.cuda_stream if (L.lane == 1 and side is not None) else st)
                if rc != 0:
                    raise _lib.RdError(f"{L.name} failed ({rc}): {lib.rd_last_error().decode()}")
            if side is not None:               # a program slice that ends inside the parallel region (never the case today)
                ev = torch.cuda.Event()
                ev.record(side)
                main.wait_event(ev)

    def _pack_input(self, x: torch.Tensor) -> None:
        """rd_input_pack straight from the caller's NCHW fp32 tensor (outside the captured graphs: its address varies).
        Without CUDA graphs (use_graphs = False: tools, debugging) the input is copied to x_in and the programs' own
        input_pack launch runs, so that every launch of a program can be replayed on its own afterwards."""
        parts = x if isinstance(x, (tuple, list)) else None
        if not self.use_graphs:
            self.x_in.copy_(torch.cat([t.float() for t in parts], dim=1) if parts is not None else x)
            return
        st = torch.cuda.current_stream().cuda_stream
        if parts is not None:
            # channel planes of several source tensors (rd_input_pack_parts): the concatenated input is never materialised
            planes, strides = [], []
            for t in parts:
                assert t.dtype == torch.float32 and t.dim() == 4 and t.stride(3) == 1 and t.stride(2) == t.shape[3] and \
                    t.stride(1) == t.shape[2] * t.shape[3], "input parts must be fp32 NCHW with dense channel planes"
                for c in range(t.shape[1]):
                    planes.append(t.data_ptr() + 4 * c * t.stride(1))
                    strides.append(t.stride(0))
            assert len(planes) == self.in_channels
            pl = (C.c_void_p * len(planes))(*planes)
            bs = (C.c_longlong * len(strides))(*strides)
            rc = self.lib.rd_input_pack_parts(pl, bs, *self._ipack_tail, st)
        else:
            assert x.dtype == torch.float32 and x.is_contiguous()
            rc = self.lib.rd_input_pack(x.data_ptr(), *self._ipack_tail, st)
        if rc != 0:
            raise _lib.RdError(f"input_pack failed ({rc}): {self.lib.rd_last_error().decode()}")

    def forward(self, x: torch.Tensor, training: bool, inference: bool = False) -> torch.Tensor:
        """inference = the pass will never be differentiated (torch.no_grad()): in eval mode it then runs the program with
        BatchNorm, residual add and activation folded into the conv epilogues, which overwrites the raw conv outputs the
        backward program would need."""
        if isinstance(x, (tuple, list)):                # channel groups of the input as separate tensors (see _pack_input)
            B, _, H, W = x[0].shape
            Cc = sum(int(t.shape[1]) for t in x)
            dev = x[0].device
        else:
            B, Cc, H, W = x.shape
            dev = x.device
        assert Cc == self.in_channels, (Cc, self.in_channels)
        if not self.params_adopted():
            self.adopt(dev)
        self.configure(B, H, W)
        self._pack_input(x)
        if training:
            self._replay("fwd", lambda: self._fwd_body(True))
        elif self._fold(training, inference):
            self._replay("fwd_infer", lambda: self._run(self.fwd_infer))
        else:
            self._replay("fwd_eval", lambda: self._fwd_body(False))
        return self.pred

    def _fwd_body(self, training: bool):
        if training:
            self.stats_used.zero_()
        self._run(self.fwd if training else self.fwd_eval)

    def _bn_backward_mode(self, training: bool) -> None:
        """Eval-mode BatchNorm (running statistics) is a per-channel affine map, so its backward is dz = gamma*invstd*g
        without the two batch-mean terms.  Those terms carry 1/count in rd_bn_tail's coefficients (cB, cC), so an
        infinite count turns every job into the eval-mode backward (dgamma/dbeta and cA are the same expressions; mean
        and invstd already hold the running statistics after an eval forward).  Eager launches only: a captured graph
        keeps the counts it was captured with."""
        for job, count in self._bwd_jobs:
            job.count = count if training else float("inf")

    def _bwd_body(self):
        self.dw.zero_()
        # backward statistics and tail tickets start from zero in EVERY backward (the forward's own sums were consumed
        # by its tails), so a second backward over the same forward (retain_graph) accumulates correctly
        self.stats_used.zero_()
        self._run(self.bwd)

    def _bwd_segment(self, k: int):
        if k == 0:
            self.dw.zero_()
            self.stats_used.zero_()
        self._run(self.bwd_segments[k])

    # ------------------------------------------------------------------ graph cut at the bottleneck (models.py:669-707)
    # PnP-Depth refinement runs the front once, then iterates the rear (forward + gradient w.r.t. the bottleneck feature).
    # These run eagerly (no CUDA graph): they are API surface, main.py never calls them.
    def _fold(self, training: bool, inference: bool) -> bool:
        return (not training) and inference and self.fwd_infer_body is not None and os.environ.get("RD_INFER_FOLD", "1") != "0"

    def forward_front(self, x: torch.Tensor, training: bool, inference: bool = False) -> torch.Tensor:
        B, Cc, H, W = x.shape
        assert Cc == self.in_channels, (Cc, self.in_channels)
        if not self.params_adopted():
            self.adopt(x.device)
        self.configure(B, H, W)
        self._pack_input(x)
        if training:
            self.stats_used.zero_()
        if self._fold(training, inference):
            self._run(self.front_infer)
        else:
            self._run(self.front if training else self.front_eval)
        return self.bneck

    def forward_rear(self, feat: torch.Tensor, training: bool, image_hw=None, inference: bool = False) -> torch.Tensor:
        B, Cc, h, w = feat.shape
        if Cc != 256:
            raise _lib.RdError(f"pnp_forward_rear expects the 256-channel bn2 output, got {Cc} channels")
        if not self.params_adopted():
            self.adopt(feat.device)
        if self.cfg is None or (self.cfg["B"], self.cfg["h32"], self.cfg["w32"]) != (B, h, w):
            H, W = image_hw if image_hw is not None else self.output_size     # the feature alone does not determine H, W
            self.configure(B, H, W)
            if (self.cfg["h32"], self.cfg["w32"]) != (h, w):
                raise _lib.RdError(f"pnp_forward_rear: a {h}x{w} feature does not belong to a {H}x{W} image "
                                   f"(expected {self.cfg['h32']}x{self.cfg['w32']})")
        self.bneck.copy_(feat)
        if self._fold(training, inference):
            self._run(self.rear_infer)         # the decoder's first conv reads the imported feature as it is
            return self.pred
        if training:
            self.stats_used.zero_()
        prog = self.rear if training else self.rear_eval
        head = 2 if training else 3
        self._run(prog[:head])             # pack_weights (+ eval: every BatchNorm's running-stat vectors), feature_import
        # the decoder's first conv applies bn2 on load in the full graph; here its input already IS bn2's output
        self.bc2.scale.fill_(1.0)
        self.bc2.shift.zero_()
        self._run(prog[head:])
        return self.pred

    def backward_rear(self, dpred: torch.Tensor, training: bool) -> torch.Tensor:
        """d(loss)/d(feature) of the last forward_rear, in the BatchNorm mode that forward ran in.  The decoder's
        parameter gradients are not produced on this path (the gradient arena is left as it was)."""
        self.dpred.copy_(dpred.reshape(self.dpred.shape))
        keep = self.gflat.clone()          # the BatchNorm tails and the head add dgamma/dbeta/dW into the arena
        self.dw.zero_()
        self.stats_used.zero_()
        self._bn_backward_mode(training)
        try:
            self._run(self.rear_bwd)
        finally:
            self._bn_backward_mode(True)
        self.gflat.copy_(keep)
        return self.d_bneck

    def backward(self, dpred: torch.Tensor, accumulate: bool, training: bool = True) -> None:
        """Fills the gradient arena from d(loss)/d(pred).  ``accumulate`` keeps what is already there.  ``training`` is
        the mode the forward ran in (eval-mode forwards are differentiated with the running statistics, eagerly)."""
        self.dpred.copy_(dpred.reshape(self.dpred.shape))
        if not accumulate:
            self.gflat.zero_()
        hook = getattr(self.module, "_rd_grad_hook", None)
        if training and hook is not None and not accumulate:
            # data-parallel training (ddp.enable_overlap): three segments, each followed by the hook that starts the
            # all-reduce of the arena ranges that segment completed
            for k in range(3):
                self._replay(f"bwd_seg{k}", lambda k=k: self._bwd_segment(k))
                hook(self, k)
            return
        if training:
            self._replay("bwd", self._bwd_body)
            return
        self._bn_backward_mode(False)
        try:
            self._bwd_body()
        finally:
            self._bn_backward_mode(True)

    # ------------------------------------------------------------------ CUDA graphs
    # The launch programs are static per input shape, so after one eager (warm-up) execution each program is
    # captured into a CUDA graph and replayed: ~360 kernel launches per step cost one cudaGraphLaunch each way.
    def _replay(self, which: str, body) -> None:
        # (see _run: graph replays do not pass through it, so the hash bookkeeping is repeated here, outside any capture)
        if which == "fwd_infer":
            if getattr(self, "_whash_stale", False) and getattr(self, "_whash", None) is not None:
                self._whash.zero_()
                self._whash_stale = False
        elif which in ("fwd", "fwd_eval"):
            self._whash_stale = True
        if not self.use_graphs:
            body()
            return
        state = self._graphs.get(which)
        if state is None:                      # first use: run eagerly (also sets kernel attributes, warms caches)
            body()
            self._graphs[which] = "warm"
            return
        if state == "warm":
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                body()
            self._graphs[which] = g
            state = g
        state.replay()

    def launches_per_step(self) -> int:
        return len(self.fwd) + len(self.bwd)

    def input_grad(self) -> torch.Tensor:
        """d(loss)/d(x) in NCHW fp32.  Only the stage-2 network of ResNet_multistage needs it (its 5th input channel
        is the stage-1 prediction, multistage_model.py:75); the 4-channel network never propagates into its input."""
        if self.dxs is None:
            raise RuntimeError("input gradients are only produced for in_channels > 4")
        B, H, W = self.cfg["B"], self.cfg["H"], self.cfg["W"]
        Cs = self.dxs.shape[-1] // 4
        H2, W2 = self.dxs.shape[1], self.dxs.shape[2]
        d = self.dxs.float().view(B, H2, W2, 2, 2, Cs)[..., : self.in_channels]      # [B,H2,W2,py,px,c]
        d = d.permute(0, 5, 1, 3, 2, 4).reshape(B, self.in_channels, 2 * H2, 2 * W2)
        return d[:, :, :H, :W].contiguous()

    def input_grad_channel(self, c: int) -> torch.Tensor:
        """Channel c of d(loss)/d(x), fp32 [B,1,H,W] (one kernel instead of the permute / slice chain above)."""
        if self.dxs is None:
            raise RuntimeError("input gradients are only produced for in_channels > 4")
        B, H, W = self.cfg["B"], self.cfg["H"], self.cfg["W"]
        out = torch.empty(B, 1, H, W, dtype=torch.float32, device=self.dxs.device)
        rc = self.lib.rd_input_grad_channel(self.dxs.data_ptr(), out.data_ptr(), B, H, W, self.dxs.shape[-1] // 4, c, self.act_dtype,
                                            torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise _lib.RdError(f"input_grad_channel failed ({rc}): {self.lib.rd_last_error().decode()}")
        return out
