"""Host-side planning of the tcgen05 convolution programs.

Every convolution on the hot path -- 7x7/3x3/1x1 stride 1|2 (reference models.py:539,559,88-91,605-608,573,582),
the 5x5 convs on the zero-stuffed Unpool output (models.py:13-27,191,198) -- is described ONCE as a ``GConv``:

    out[b, OS*y + a] = sum over taps t with phase a:   X[b, (y + s_t)*S + pl_t, :] @ W_t          (2-D indices)

with ``W_t`` given as an int32 index matrix ``widx[Cx, N]`` into the flat fp32 parameter arena (-1 = structural
zero).  From that single description this module derives

  * the forward program      (``plan_fprop(g)``),
  * the data-gradient program (``plan_fprop(g.transposed())``: swap S<->OS, phase<->plane, negate shifts, transpose W),
  * the weight-gradient program (``plan_wgrad(g)``),
  * the gather table that packs bf16 weight tiles from the parameter arena, and the scatter table that routes
    the weight-gradient accumulators back to OIHW ``.grad`` layout,

plus a slow torch reference (``gconv_reference`` / ``gconv_wgrad_reference``) used by the tests to validate both
the descriptions (against F.conv2d on CPU) and the kernels (on the GPU).
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib

SMEM_BUDGET = 232448
_TUNED_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tuned_tiles.json")
_TUNED = None


def tuned_table() -> dict:
    """Tile choices measured on a B200 by tools/autotune.py (committed with the repo); the analytic cost models below
    are the fallback for shapes that are not in the table."""
    global _TUNED
    if _TUNED is None:
        try:
            with open(_TUNED_PATH) as f:
                _TUNED = json.load(f)
        except Exception:
            _TUNED = {}
    return _TUNED


def tune_key(kind: str, g: "GConv", B: int, src_hw, dst_hw, act_dtype: int) -> str:
    return f"{kind}|Cx{g.Cx}|N{g.N}|S{g.S}|OS{g.OS}|T{len(g.taps)}|{src_hw[0]}x{src_hw[1]}>{dst_hw[0]}x{dst_hw[1]}|B{B}|dt{act_dtype}"


def tuned_lookup(key: str, sm_budget: int, suffix: str = ""):
    """Measured entry for a launch that may use `sm_budget` SMs: the entry measured with exactly that budget (key + "|smNN",
    the encoder chains that run side by side on disjoint SM sets) or, for budgets of at least half the GPU, the whole-GPU
    entry.  A chain that owns a few SMs only is planned by the cost model when it has no entry of its own (measured,
    multistage b=8: 12.4 vs 13.2 ms/step with the whole-GPU tiles on the small lane)."""
    tab = tuned_table()
    if sm_budget != NUM_SMS:
        t = tab.get(key + f"|sm{sm_budget}" + suffix)
        if t is not None or 2 * sm_budget < NUM_SMS:
            return t
    return tab.get(key + suffix)
FPROP_HEADER = 16384
WGRAD_HEADER = 10240
NUM_SMS = 148
L2_BYTES_PER_CYCLE = 24.0      # sustained L2 -> SM bytes per cycle per SM with all SMs pulling (planning constant)


@dataclass
class GTap:
    ph: Tuple[int, int]      # output phase (y, x) in [0, OS)
    pl: Tuple[int, int]      # source parity plane (y, x) in [0, S)
    s: Tuple[int, int]       # shift in plane coordinates
    widx: np.ndarray         # int32 [Cx, N] indices into the flat parameter arena, -1 = zero


@dataclass
class GConv:
    Cx: int
    N: int
    S: int
    OS: int
    taps: List[GTap]
    name: str = ""

    def transposed(self) -> "GConv":
        """Data-gradient form: dX = sum_t dOut[...] @ W_t^T."""
        taps = [GTap(ph=t.pl, pl=t.ph, s=(-t.s[0], -t.s[1]), widx=np.ascontiguousarray(t.widx.T)) for t in self.taps]
        return GConv(Cx=self.N, N=self.Cx, S=self.OS, OS=self.S, taps=taps, name=self.name + ".T")

    def sorted_taps(self) -> Tuple[List[GTap], List[Tuple[int, int]]]:
        phases = sorted({t.ph for t in self.taps})
        taps = sorted(self.taps, key=lambda t: (phases.index(t.ph), t.pl, t.s))
        return taps, phases


# ------------------------------------------------------------------------------------------ builders
def _oihw_index(off: int, Cout: int, Cin: int, kh: int, kw: int, ky: int, kx: int) -> np.ndarray:
    """[Cin, Cout] matrix of flat indices of W[co][ci][ky][kx] (OIHW at arena offset ``off``)."""
    co = np.arange(Cout, dtype=np.int64)[None, :]
    ci = np.arange(Cin, dtype=np.int64)[:, None]
    return (off + ((co * Cin + ci) * kh + ky) * kw + kx).astype(np.int32)


def gconv_standard(off: int, Cout: int, Cin: int, k: int, stride: int, pad: int, name: str = "") -> GConv:
    """nn.Conv2d(Cin, Cout, k, stride, pad, bias=False)."""
    assert stride in (1, 2)
    taps = []
    for ky in range(k):
        for kx in range(k):
            dy, dx = ky - pad, kx - pad
            if stride == 1:
                pl, s = (0, 0), (dy, dx)
            else:
                pl, s = (dy & 1, dx & 1), (dy >> 1, dx >> 1)
            taps.append(GTap(ph=(0, 0), pl=pl, s=s, widx=_oihw_index(off, Cout, Cin, k, k, ky, kx)))
    return GConv(Cx=Cin, N=Cout, S=stride, OS=1, taps=taps, name=name)


def gconv_stem(off_rgb: int, off_depth: int, cin_depth: int, name: str = "stem") -> GConv:
    """conv1 (3->64, 7x7 s2 p3, models.py:539) and conv1_depth (cin_depth->16, models.py:559 /
    multistage_model.py:163-164) fused into ONE stride-1 4x4-tap convolution over the space-to-depth input
    produced by rd_input_pack: channel = parity*Cs + c, Cs = 4 (4 input channels) or 8 (5 input channels)."""
    C = 3 + cin_depth
    Cs = 4 if C <= 4 else 8
    Cx, N = 4 * Cs, 80
    taps = []
    for sy in range(-2, 2):
        for sx in range(-2, 2):
            widx = np.full((Cx, N), -1, dtype=np.int32)
            for py in range(2):
                for px in range(2):
                    ky, kx = 2 * sy + py + 3, 2 * sx + px + 3
                    if not (0 <= ky < 7 and 0 <= kx < 7):
                        continue
                    q = py * 2 + px
                    for c in range(C):
                        if c < 3:
                            n = np.arange(64)
                            widx[q * Cs + c, :64] = off_rgb + ((n * 3 + c) * 7 + ky) * 7 + kx
                        else:
                            n = np.arange(16)
                            widx[q * Cs + c, 64:] = off_depth + ((n * cin_depth + (c - 3)) * 7 + ky) * 7 + kx
            taps.append(GTap(ph=(0, 0), pl=(0, 0), s=(sy, sx), widx=widx))
    return GConv(Cx=Cx, N=N, S=1, OS=1, taps=taps, name=name)


def gconv_stem_single(off: int, C: int, name: str = "stem") -> GConv:
    """ResNet.conv1 (C->64, 7x7 s2 p3, models.py:238-245) as a stride-1 4x4-tap convolution over the space-to-depth
    input of rd_input_pack (channel = parity*4 + c, C <= 4)."""
    assert 1 <= C <= 4
    Cs, N = 4, 64
    taps = []
    for sy in range(-2, 2):
        for sx in range(-2, 2):
            widx = np.full((4 * Cs, N), -1, dtype=np.int32)
            for py in range(2):
                for px in range(2):
                    ky, kx = 2 * sy + py + 3, 2 * sx + px + 3
                    if not (0 <= ky < 7 and 0 <= kx < 7):
                        continue
                    n = np.arange(N)
                    for c in range(C):
                        widx[(py * 2 + px) * Cs + c, :] = off + ((n * C + c) * 7 + ky) * 7 + kx
            taps.append(GTap(ph=(0, 0), pl=(0, 0), s=(sy, sx), widx=widx))
    return GConv(Cx=4 * Cs, N=N, S=1, OS=1, taps=taps, name=name)


def gconv_upconv(off: int, Cin: int, Cout: int, name: str = "") -> GConv:
    """UpConv's 5x5 p2 conv on Unpool(x) (models.py:160-169) as 4 sub-pixel convolutions on the un-stuffed x."""
    taps = []
    for a in range(2):
        for b in range(2):
            for ky in range(5):
                if (a + ky - 2) % 2:
                    continue
                for kx in range(5):
                    if (b + kx - 2) % 2:
                        continue
                    taps.append(GTap(ph=(a, b), pl=(0, 0), s=((a + ky - 2) // 2, (b + kx - 2) // 2),
                                     widx=_oihw_index(off, Cout, Cin, 5, 5, ky, kx)))
    return GConv(Cx=Cin, N=Cout, S=1, OS=2, taps=taps, name=name)


def gconv_deconv(off: int, Cin: int, Cout: int, k: int, name: str = "") -> GConv:
    """nn.ConvTranspose2d(Cin, Cout, k, stride 2, padding (k-1)//2, output_padding k%2, bias=False) (DeConv, models.py:140-151):
    the transpose of the stride-2 convolution whose OIHW weight is the ConvTranspose2d weight [Cin, Cout, k, k]."""
    assert k in (2, 3)
    return gconv_standard(off, Cin, Cout, k, 2, (k - 1) // 2, name=name).transposed()


def gconv_upproj(off_upper: int, off_bottom: int, Cin: int, Cout: int, name: str = "") -> GConv:
    """Both 5x5 p2 convs of an UpProjModule (upper_branch.conv1 | bottom_branch.conv, models.py:191,198) applied to
    Unpool(x) (models.py:13-27), as 4 sub-pixel convolutions on the un-stuffed x with N = 2*Cout."""
    taps = []
    for a in range(2):
        for b in range(2):
            for ky in range(5):
                if (a + ky - 2) % 2:
                    continue
                for kx in range(5):
                    if (b + kx - 2) % 2:
                        continue
                    w = np.concatenate([_oihw_index(off_upper, Cout, Cin, 5, 5, ky, kx),
                                        _oihw_index(off_bottom, Cout, Cin, 5, 5, ky, kx)], axis=1)
                    taps.append(GTap(ph=(a, b), pl=(0, 0), s=((a + ky - 2) // 2, (b + kx - 2) // 2), widx=w))
    return GConv(Cx=Cin, N=2 * Cout, S=1, OS=2, taps=taps, name=name)


# ------------------------------------------------------------------------------------------ references
def _gather_source(x, t: GTap, S: int, Hb: int, Wb: int):
    """x: [B,H,W,C] torch tensor -> [B,Hb,Wb,C] of X[(y+sy)*S+py, (x+sx)*S+px] with zeros out of range."""
    import torch
    B, H, W, C = x.shape
    ys = (torch.arange(Hb, device=x.device) + t.s[0]) * S + t.pl[0]
    xs = (torch.arange(Wb, device=x.device) + t.s[1]) * S + t.pl[1]
    vy = (ys >= 0) & (ys < H)
    vx = (xs >= 0) & (xs < W)
    g = x[:, ys.clamp(0, H - 1)][:, :, xs.clamp(0, W - 1)]
    return g * (vy[:, None] & vx[None, :]).to(x.dtype)[None, :, :, None]


def _tap_weight(wflat, t: GTap):
    import torch
    idx = torch.from_numpy(t.widx.astype(np.int64)).to(wflat.device)
    w = wflat[idx.clamp(min=0)]
    return w * (idx >= 0).to(wflat.dtype)


def gconv_reference(g: GConv, x, wflat, dst_hw):
    """Slow exact evaluation of a GConv with torch ops.  x: [B,H,W,Cx]; returns [B,dstH,dstW,N].  Output positions
    belonging to phases without taps are left at zero."""
    import torch
    B = x.shape[0]
    dH, dW = dst_hw
    Hb, Wb = -(-dH // g.OS), -(-dW // g.OS)
    out = torch.zeros(B, Hb * g.OS, Wb * g.OS, g.N, dtype=x.dtype, device=x.device)
    for t in g.taps:
        out[:, t.ph[0]::g.OS, t.ph[1]::g.OS] += _gather_source(x, t, g.S, Hb, Wb) @ _tap_weight(wflat, t)
    return out[:, :dH, :dW]


def gconv_wgrad_reference(g: GConv, x, dout, nparams: int):
    """Flat gradient (length nparams) of sum(out * dout) w.r.t. the parameter arena."""
    import torch
    B, dH, dW, _ = dout.shape
    Hb, Wb = -(-dH // g.OS), -(-dW // g.OS)
    pad = torch.zeros(B, Hb * g.OS, Wb * g.OS, g.N, dtype=dout.dtype, device=dout.device)
    pad[:, :dH, :dW] = dout
    grad = torch.zeros(nparams, dtype=x.dtype, device=x.device)
    for t in g.taps:
        xs = _gather_source(x, t, g.S, Hb, Wb).reshape(-1, g.Cx)
        go = pad[:, t.ph[0]::g.OS, t.ph[1]::g.OS].reshape(-1, g.N)
        dw = xs.t() @ go                                    # [Cx, N]
        idx = torch.from_numpy(t.widx.astype(np.int64)).to(x.device)
        m = idx >= 0
        grad.index_add_(0, idx[m], dw[m])
    return grad


# ------------------------------------------------------------------------------------------ fprop planning
def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def _chunk_stride(ps: int, nchunks: int) -> int:
    """Smallest stride (in 16-byte slots) >= ps between chunk planes such that the 8 lanes of a quarter-warp, which
    store (column, chunk) items with the chunk index fastest, hit 8 distinct 16-byte bank groups."""
    if nchunks >= 8:
        want = lambda v: v % 2 == 1                    # odd: chunks 0..7 land 16 B apart (mod 128)
    elif nchunks == 4:
        want = lambda v: v % 8 in (2, 6)
    elif nchunks == 2:
        want = lambda v: v % 8 == 4
    else:
        want = lambda v: True
    v = ps
    while not want(v):
        v += 1
    return v


@dataclass
class FpropPlan:
    g: GConv
    params: "_lib.ConvParams"          # geometry / program filled in; pointers bound at call time
    pack_idx: np.ndarray               # int32 gather table for the packed weights of this conv
    wpk_elems: int
    ntiles: int
    info: dict = field(default_factory=dict)


def _choose_fprop_tile(Hb, Wb, halo_y, halo_x, S, P, N, nblk, ntaps, ncblk, parts, wstage_bytes, B_=16, budget=SMEM_BUDGET,
                       sms=NUM_SMS):
    """Tile search.  The cost model was fitted to per-role cycle counters measured on B200 (tools/bench_fprop.py):
    UMMA ~max(48, N/2) cycles each, epilogue ~40 cycles per 16 columns per 32 rows (overlapped with the next tile
    when two accumulator sets fit in TMEM), ~2.5k cycles of pipeline hand-off per tile, and a strong preference for
    wide tiles (long contiguous runs per row for the loaders) with at most 4 accumulator blocks."""
    best = None
    MBmax = max(1, 512 // (P * N))
    if N >= 64:
        MBmax = min(MBmax, 4)
    planes = S * S
    wl_min = min(24, Wb + halo_x)
    for MB in range(1, MBmax + 1):
        M = MB * 128
        for Wl in range(max(halo_x + 1, wl_min), min(Wb + halo_x, M, 256 // S) + 1):
            Wt = Wl - halo_x
            Ht = min(M // Wl, Hb)
            if Ht < 1:
                continue
            if MB > 1 and Ht * Wl <= (MB - 1) * 128:
                continue                                   # a smaller MB covers this tile
            plane_rows = Ht + halo_y
            plane_slots = _round_up(M + halo_y * Wl + halo_x, 8)
            istage = _round_up(parts * 2 * _chunk_stride(planes * plane_slots, 2) * 16, 128)
            if FPROP_HEADER + 2 * istage + 2 * wstage_bytes > budget:
                continue
            ty, tx = -(-Hb // Ht), -(-Wb // Wt)
            mma = MB * ntaps * max(N, 96) / 2.0 * (3 if parts == 2 else 1)
            load = planes * plane_rows * Wl * 2 * 0.9 * parts
            epi = MB * P * (N / 16.0) * 160.0
            stage = max(mma, load)
            if 2 * P * MB * N <= 512:          # two accumulator sets: the epilogue overlaps the next tile's MMAs
                per_tile = max(stage * ncblk, epi) + 2500.0
            else:
                per_tile = stage * ncblk + epi + 2500.0
            ctas = max(1, sms // nblk)
            rounds = -(-(ty * tx * B_) // ctas)
            cost = rounds * per_tile
            if best is None or cost < best[0]:
                best = (cost, dict(MB=MB, Wl=Wl, Wt=Wt, Ht=Ht, plane_rows=plane_rows, plane_slots=plane_slots,
                                   istage=istage, tiles_y=ty, tiles_x=tx))
    if best is None:
        raise ValueError("no feasible tile for this convolution")
    return best[1]


def plan_fprop(g: GConv, B: int, src_hw, dst_hw, act_dtype: int = _lib.RD_BF16, n_per_cta: Optional[int] = None,
               tile_override: Optional[dict] = None, use_tuned: bool = True, smem_reserve: int = 0,
               sm_budget: int = NUM_SMS) -> FpropPlan:
    """smem_reserve: bytes of dynamic shared memory the launcher needs behind the rings (deterministic statistics).
    sm_budget: SMs (= persistent CTAs) this launch may occupy; less than the whole GPU when the engine runs the RGB and the
    depth encoder side by side on disjoint sets of SMs (engine.LatefusionEngine, `lane`)."""
    assert g.Cx % 16 == 0 and g.N % 16 == 0, (g.Cx, g.N)
    budget = SMEM_BUDGET - smem_reserve
    parts = 2 if act_dtype == _lib.RD_F32 else 1
    taps, phases = g.sorted_taps()
    P = len(phases)
    ntaps = len(taps)
    assert ntaps <= _lib.RD_MAX_TAPS
    if tile_override is None and use_tuned:
        tile_override = tuned_lookup(tune_key("f", g, B, src_hw, dst_hw, act_dtype), sm_budget)
    if n_per_cta is None and tile_override and "N" in tile_override:
        n_per_cta = int(tile_override["N"])            # measured: output channels per CTA (tools/autotune.py)
    if n_per_cta is None:
        n_per_cta = min(g.N, 128)
        while g.N % n_per_cta:
            n_per_cta -= 16
    N = n_per_cta
    assert g.N % N == 0 and P * N <= 512
    nblk = g.N // N
    ncblk = g.Cx // 16
    dH, dW = dst_hw
    Hb, Wb = -(-dH // g.OS), -(-dW // g.OS)
    sy = [t.s[0] for t in taps]
    sx = [t.s[1] for t in taps]
    sy_min, sx_min = min(sy), min(sx)
    halo_y, halo_x = max(sy) - sy_min, max(sx) - sx_min
    # weight groups
    tap_bytes = parts * N * 32
    wcap = 36864
    max_g = max(1, wcap // tap_bytes)
    ngroups = -(-ntaps // max_g)
    assert ngroups <= _lib.RD_MAX_GROUPS
    base, rem = divmod(ntaps, ngroups)
    grp_n = [base + (1 if i < rem else 0) for i in range(ngroups)]
    wstage = _round_up(max(grp_n) * tap_bytes, 128)
    geo = None
    if tile_override:
        geo = {k: v for k, v in tile_override.items() if k != "N"}
        geo.setdefault("Wl", geo["Wt"] + halo_x)
        geo.setdefault("MB", -(-geo["Ht"] * geo["Wl"] // 128))
        M = geo["MB"] * 128
        geo["plane_rows"] = geo["Ht"] + halo_y
        geo["plane_slots"] = _round_up(M + halo_y * geo["Wl"] + halo_x, 8)
        geo["istage"] = _round_up(parts * 2 * _chunk_stride(g.S * g.S * geo["plane_slots"], 2) * 16, 128)
        geo["tiles_y"], geo["tiles_x"] = -(-Hb // geo["Ht"]), -(-Wb // geo["Wt"])
        if FPROP_HEADER + 2 * geo["istage"] + 2 * wstage > budget:
            geo = None                                     # a measured tile that no longer fits (smem_reserve): re-plan
    if geo is None:
        geo = _choose_fprop_tile(Hb, Wb, halo_y, halo_x, g.S, P, N, nblk, ntaps, ncblk, parts, wstage, B, budget, sm_budget)
    istage = geo["istage"]
    # ring depths within the shared-memory budget
    avail = budget - FPROP_HEADER
    IS = 2
    WS = 2
    while True:
        grown = False
        if WS < 6 and FPROP_HEADER + IS * istage + (WS + 1) * wstage <= budget:
            WS += 1
            grown = True
        if IS < 6 and (IS + 1) * istage <= 98304 and FPROP_HEADER + (IS + 1) * istage + WS * wstage <= budget:
            IS += 1
            grown = True
        if not grown:
            break
    assert IS * istage + WS * wstage <= avail

    p = _lib.ConvParams()
    p.Cin, p.S, p.B = g.Cx, g.S, B
    p.srcH, p.srcW = src_hw
    p.Hb, p.Wb = Hb, Wb
    p.Ht, p.Wt, p.Wl = geo["Ht"], geo["Wt"], geo["Wl"]
    p.plane_rows, p.plane_slots = geo["plane_rows"], geo["plane_slots"]
    p.chunk_stride = _chunk_stride(g.S * g.S * geo["plane_slots"], 2)
    p.sy_min, p.sx_min = sy_min, sx_min
    p.MB = geo["MB"]
    p.tiles_y, p.tiles_x = geo["tiles_y"], geo["tiles_x"]
    p.P, p.OS, p.ntaps, p.ngroups = P, g.OS, ntaps, ngroups
    for i, ph in enumerate(phases):
        p.phase_y[i], p.phase_x[i] = ph
    seen = set()
    for i, t in enumerate(taps):
        pi = phases.index(t.ph)
        q = t.pl[0] * g.S + t.pl[1]
        p.taps[i].a_shift = q * geo["plane_slots"] + (t.s[0] - sy_min) * geo["Wl"] + (t.s[1] - sx_min)
        p.taps[i].phase = pi
        p.taps[i].first = 0 if pi in seen else 1
        seen.add(pi)
    first = 0
    for i, n in enumerate(grp_n):
        p.grp_first[i], p.grp_n[i] = first, n
        first += n
    p.N, p.nblk = N, nblk
    p.dstH, p.dstW = dH, dW
    p.IS, p.WS, p.istage_bytes, p.wstage_bytes = IS, WS, istage, wstage
    p.act_dtype = act_dtype
    # stride-2 sources are staged as four parity planes; taps that only ever read plane (0,0) (1x1 stride-2 downsample
    # convolutions) need just that one
    p.src_planes = 1 if (g.S == 2 and all(t.pl == (0, 0) for t in taps)) else 0
    ntiles = geo["tiles_y"] * geo["tiles_x"] * B
    p.max_ctas = max(1, sm_budget // nblk)
    # gather table [nblk][ncblk][tap][part][j][n][k]
    Wst = np.stack([t.widx for t in taps], axis=0)                       # [T, Cx, Ntot]
    Wst = Wst.reshape(ntaps, ncblk, 2, 8, nblk, N).transpose(4, 1, 0, 2, 5, 3)   # [nblk, ncblk, T, j, n, k]
    if parts == 2:
        lo = np.where(Wst >= 0, Wst | (1 << 30), -1)
        Wst = np.stack([Wst, lo], axis=3)                                # [nblk, ncblk, T, part, j, n, k]
    pack_idx = np.ascontiguousarray(Wst).reshape(-1).astype(np.int32)
    assert pack_idx.size == nblk * ncblk * ntaps * parts * 2 * N * 8
    return FpropPlan(g=g, params=p, pack_idx=pack_idx, wpk_elems=int(pack_idx.size), ntiles=ntiles,
                     info=dict(geo=geo, IS=IS, WS=WS, grp_n=grp_n, P=P, N=N, nblk=nblk))


def compact_pack_table(idx: np.ndarray):
    """(groups [G,2] int32, fallback [F,8] int32) for rd_pack_weights_g8 from a rd_pack_weights index table (length 8 G).
    A group of 8 consecutive outputs whose indices form an arithmetic progression with one hi/lo flag becomes (base, stride);
    eight invalid entries become (-1, 0); anything else (holes: zero-padded stem taps, channel counts that are no
    multiple of 8) keeps its eight indices in the fallback table and becomes (row, -1)."""
    assert idx.size % 8 == 0
    t = idx.reshape(-1, 8).astype(np.int64)
    valid = t >= 0
    allneg = ~valid.any(axis=1)
    allpos = valid.all(axis=1)
    raw = t & 0x3FFFFFFF
    flag = t & 0x40000000
    d = np.diff(raw, axis=1)
    ap = allpos & (d == d[:, :1]).all(axis=1) & (d[:, 0] >= 0) & (flag == flag[:, :1]).all(axis=1)
    groups = np.zeros((t.shape[0], 2), dtype=np.int32)
    groups[ap, 0] = t[ap, 0].astype(np.int32)
    groups[ap, 1] = d[ap, 0].astype(np.int32)
    groups[allneg, 0] = -1
    other = ~(ap | allneg)
    fb = idx.reshape(-1, 8)[other].astype(np.int32)
    groups[other, 0] = np.arange(int(other.sum()), dtype=np.int32)
    groups[other, 1] = -1
    if fb.shape[0] == 0:
        fb = np.full((1, 8), -1, dtype=np.int32)
    return np.ascontiguousarray(groups), np.ascontiguousarray(fb)


# ------------------------------------------------------------------------------------------ wgrad planning
@dataclass
class WgradPlan:
    g: GConv
    params: "_lib.WgradParams"
    dw_elems: int                      # ntaps * N * Cx fp32 accumulators
    scatter: Tuple[np.ndarray, np.ndarray]   # (param flat index, index into this conv's dw block)
    info: dict = field(default_factory=dict)


def _search_wgrad_tile(g, Nc, Mc, ncob, ntaps, parts, Hb, Wb, halo_y, halo_x, ks_target, R=1, tile_oy=0):
    """KS slots per plane / tile shape minimising the modelled cycles subject to >= 2 ring stages in shared memory.
    ntaps = number of accumulators (jobs when the gradient tile is staged R times, see plan_wgrad)."""
    tg_cap = max(1, min(16, 512 // Nc))
    ntg = -(-ntaps // tg_cap)
    tg_size = -(-ntaps // ntg)
    best = None
    for KS0 in (32, 64, 96, 128, 192, 256, 384, 512):
        if KS0 > max(ks_target, 32):
            continue
        for Wl in range(halo_x + 1, min(Wb + halo_x, KS0) + 1):
            Wt = Wl - halo_x
            Ht = min(KS0 // Wl, Hb)
            if Ht < 1:
                continue
            # A tile of exactly KS = Ht*Wl slots lets the gradient tile be one dense TMA box (rd_conv_wgrad: TMA writes the
            # chunk planes of a box back to back, so a plane must not have a tail); such tiles are preferred.
            tma_g = parts == 1 and (Ht * Wl) % 16 == 0 and Wl * g.OS <= 256
            if R > 1 and not tma_g:
                continue                                   # gradient copies are TMA boxes
            KS = Ht * Wl if tma_g else KS0
            xrows = Ht + halo_y
            xslots = _round_up(KS + halo_y * Wl + halo_x + (4 if g.Cx == 16 else 0), 8)     # + the junk 4th tap of a folded row
            GPS = _chunk_stride(g.OS * g.OS * KS, Mc // 8)
            XPS = _chunk_stride(g.S * g.S * xslots, Nc // 8)
            g_bytes = _round_up(parts * R * (Mc // 8) * GPS * 16, 128)
            stage = _round_up(g_bytes + parts * (Nc // 8) * XPS * 16, 128)
            # the M=128 operand always spans 16 chunk planes; rows past Mc are junk but must stay inside smem
            pad = max(0, ((parts - 1) * (Mc // 8) + 16) * GPS * 16 - stage)
            if WGRAD_HEADER + 2 * stage + pad > SMEM_BUDGET:
                continue
            ty, tx = -(-(Hb + tile_oy) // Ht), -(-Wb // Wt)
            load = R * (Mc // 8) * GPS + (Nc // 8) * XPS
            mma = (KS // 16) * tg_size * max(Nc, 32) / 2.0 * (3 if parts == 2 else 1)
            cost = ty * tx * (max(mma, load * 0.35 * parts * (0.6 if tma_g else 1.0)) + 300.0)
            if best is None or cost < best[0]:
                best = (cost, dict(KS=KS, Wl=Wl, Wt=Wt, Ht=Ht, xrows=xrows, xslots=xslots, g_bytes=g_bytes,
                                   stage=stage, tiles_y=ty, tiles_x=tx, pad=pad, GPS=GPS, XPS=XPS, tma_g=tma_g))
    return best


def _gcopy_layout(g: GConv, taps, Mc: int, parts: int):
    """Gradient copies stacked in M (rd_wgrad_params.gcopies): (R, tap rows, tap columns) for stride-1 convolutions with
    Cout <= 64 whose taps form a full rows x columns grid, else None."""
    if parts != 1 or g.S != 1 or g.OS not in (1, 2) or Mc > 64 or os.environ.get("RD_TMA", "1") == "0" or \
            os.environ.get("RD_WGRAD_GCOPY", "1") == "0":
        return None
    if g.OS == 2:
        # sub-pixel (UpProj / UpConv / ConvTranspose2d) programs: the four parity planes of the gradient tile are staged side
        # by side anyway, so R adjacent planes ARE the stacked operand -- taps of those phases that share a source shift
        # become one accumulator, with no extra load at all
        R = min(128 // Mc, 4)
        return (R, None, None) if R >= 2 else None
    rows = sorted({t.s[0] for t in taps})
    cols = sorted({t.s[1] for t in taps})
    if len(rows) < 2 or len(rows) * len(cols) != len(taps) or [t.s for t in taps] != [(r, c) for r in rows for c in cols]:
        return None
    if rows != list(range(rows[0], rows[0] + len(rows))) or cols != list(range(cols[0], cols[0] + len(cols))):
        return None
    R = min(128 // Mc, len(rows), 8)
    return (R, rows, cols) if R >= 2 else None


def plan_wgrad(g: GConv, B: int, x_hw, g_hw, act_dtype: int = _lib.RD_BF16, ks_target: int = 256,
               nc: Optional[int] = None, use_tuned: bool = True, sm_budget: int = NUM_SMS, gcopy: Optional[bool] = None,
               bn: bool = False) -> WgradPlan:
    """x_hw: spatial size of the source activation; g_hw: spatial size of the output gradient.
    gcopy: stack shifted copies of the gradient tile in M where the layer allows it (None: measured table, else on).
    bn: the launch applies the producer's BatchNorm + activation to the source tile in shared memory (every tap group
    repeats that pass, so the measured table may hold a different blocking under key + "|bn")."""
    assert g.Cx % 16 == 0 and g.N % 8 == 0
    if use_tuned and nc is None:
        key = tune_key("w", g, B, x_hw, g_hw, act_dtype)
        t = (tuned_lookup(key, sm_budget, "|bn") if bn else None) or tuned_lookup(key, sm_budget)
        if t:
            nc, ks_target = t["nc"], t["ks"]
            if gcopy is None and "gc" in t:
                gcopy = bool(t["gc"])
    if gcopy is None:
        gcopy = True
    parts = 2 if act_dtype == _lib.RD_F32 else 1
    taps, phases = g.sorted_taps()
    ntaps = len(taps)
    gH, gW = g_hw
    Hb, Wb = -(-gH // g.OS), -(-gW // g.OS)
    sy = [t.s[0] for t in taps]
    sx = [t.s[1] for t in taps]
    sy_min, sx_min = min(sy), min(sx)
    halo_y, halo_x = max(sy) - sy_min, max(sx) - sx_min
    Mc = min(_round_up(g.N, 8), 128)
    ncob = -(-g.N // Mc)
    # ---- gradient copies: jobs instead of taps
    gl = _gcopy_layout(g, taps, Mc, parts) if gcopy else None
    R, tile_oy = 1, 0
    jobs = None                                       # [(sy of copy 0, sx or None when the row is N-folded)]
    fold_n = False
    plane_jobs = None                                 # OS = 2: [(first plane, (sy, sx), [tap index or -1 per stacked plane])]
    if gl is not None and gl[1] is None:
        R = gl[0]
        phases_ = sorted({t.ph for t in taps})
        assert all(ph in [(a, b) for a in range(2) for b in range(2)] for ph in phases_)
        by = {}
        for i, t in enumerate(taps):
            by[(t.ph[0] * 2 + t.ph[1], t.s)] = i
        plane_jobs = []
        for q0 in range(0, 4, R):
            shifts = sorted({s_ for (q_, s_) in by if q0 <= q_ < q0 + R})
            for s_ in shifts:
                plane_jobs.append((q0, s_, [by.get((q0 + r, s_), -1) for r in range(R)]))
        jobs = plane_jobs
    elif gl is not None:
        R, rows, cols = gl
        tile_oy = R - 1
        job_rows = [rows[-1] - m * R for m in range(-(-len(rows) // R))]
        fold_n = g.Cx == 16 and (nc in (None, 16)) and 2 <= len(cols) <= 4 and len(job_rows) * 64 <= 512 and \
            os.environ.get("RD_WGRAD_FOLD", "1") != "0"
        jobs = [(sy_, None) for sy_ in job_rows] if fold_n else [(sy_, sx_) for sy_ in job_rows for sx_ in cols]
        if fold_n:
            nc = 16
    nacc = len(jobs) if jobs is not None else ntaps
    if nc is None:
        # prefer ONE tap group per CTA (the gradient and source tiles are then staged once per pixel tile instead
        # of once per tap group): the widest Nc (multiple of 16 dividing Cx) with ntaps*Nc <= 512 TMEM columns
        tpc = min(nacc, 16)
        nc = 16
        for cand in range(16, min(g.Cx, 256) + 1, 16):
            if g.Cx % cand == 0 and tpc * cand <= 512:
                nc = cand
    nc_candidates = [nc] + [c for c in (128, 64, 32, 16) if c < nc and g.Cx % c == 0]
    best = None
    for Nc in nc_candidates:
        best = _search_wgrad_tile(g, Nc, Mc, ncob, nacc, parts, Hb, Wb, halo_y, halo_x, ks_target, R, tile_oy)
        if best is not None:
            break
    if best is None and R > 1:
        return plan_wgrad(g, B, x_hw, g_hw, act_dtype, ks_target, nc, use_tuned, sm_budget, gcopy=False, bn=bn)
    if best is None:
        raise ValueError("no feasible wgrad tile")
    ncib = g.Cx // Nc
    tg_cap = max(1, min(16, 512 // Nc))
    if fold_n:
        tg_cap = 16                                   # folded jobs own 64 columns each (checked above)
    ntg = -(-nacc // tg_cap)
    tg_size = -(-nacc // ntg)
    geo = best[1]
    NS = 2
    while NS < 4 and WGRAD_HEADER + (NS + 1) * geo["stage"] + geo["pad"] <= SMEM_BUDGET:
        NS += 1
    p = _lib.WgradParams()
    p.gH, p.gW, p.Cout, p.Sg = gH, gW, g.N, g.OS
    p.xH, p.xW = x_hw
    p.Cin, p.Sx = g.Cx, g.S
    p.B = B
    p.Hb, p.Wb, p.Ht, p.Wt, p.Wl = Hb, Wb, geo["Ht"], geo["Wt"], geo["Wl"]
    p.KS = geo["KS"]
    p.g_chunk_stride, p.x_chunk_stride = geo["GPS"], geo["XPS"]
    p.x_plane_rows, p.x_plane_slots = geo["xrows"], geo["xslots"]
    p.sy_min, p.sx_min = sy_min, sx_min
    p.tiles_y, p.tiles_x = geo["tiles_y"], geo["tiles_x"]
    p.ntaps, p.tg_size, p.ntg = ntaps, tg_size, ntg
    for i, t in enumerate(taps):
        p.taps[i].g_off = (t.ph[0] * g.OS + t.ph[1]) * geo["KS"]
        q = t.pl[0] * g.S + t.pl[1]
        p.taps[i].x_shift = q * geo["xslots"] + (t.s[0] - sy_min) * geo["Wl"] + (t.s[1] - sx_min)
    p.gcopies, p.njobs, p.tile_oy, p.tile_ox = (R if jobs is not None else 0), (nacc if jobs is not None else 0), tile_oy, 0
    if plane_jobs is not None:
        for j, (q0, s_, tl) in enumerate(plane_jobs):
            p.taps[j].g_off = q0 * geo["KS"]                     # plane offset (rewritten to the TMA layout by the launcher)
            p.taps[j].x_shift = (s_[0] - sy_min) * geo["Wl"] + (s_[1] - sx_min)
            for r in range(8):
                p.job_tap[j][r] = tl[r] if r < R else -1
    elif jobs is not None:
        ncols = len(cols)
        for r in range(R):
            p.gcopy_dy[r], p.gcopy_dx[r] = r, 0
        for j, (sy_, sx_) in enumerate(jobs):
            # copy r of the gradient tile holds g[p + (r, 0)]: against the source at shift (sy_, sx_) it produces tap (sy_ - r, sx_)
            p.taps[j].g_off = 0
            p.taps[j].x_shift = (sy_ - sy_min) * geo["Wl"] + ((sx_ if sx_ is not None else cols[0]) - sx_min)
            for r in range(8):
                ky = sy_ - r
                ok = r < R and ky >= rows[0]
                p.job_tap[j][r] = ((ky - rows[0]) * ncols + ((sx_ - cols[0]) if sx_ is not None else 0)) if ok else -1
    p.Mc, p.ncob, p.Nc, p.ncib = Mc, ncob, Nc, ncib
    p.NS, p.stage_bytes, p.g_bytes = NS, geo["stage"], geo["g_bytes"]
    p.act_dtype = act_dtype
    p.x_planes = 1 if (g.S == 2 and all(t.pl == (0, 0) for t in taps)) else 0
    # tap-row folding (rd_wgrad_params.fold_rows): 16-channel stride-1 sources whose taps form full rows of <= 4 adjacent taps
    p.fold_rows, p.fold_len = 0, 0
    if jobs is not None:
        if fold_n:
            assert Nc == 16 and ntg == 1
            p.fold_rows, p.fold_len = len(jobs), len(cols)
    elif parts == 1 and g.Cx == 16 and Nc == 16 and g.S == 1 and g.OS == 1 and ntg == 1 and os.environ.get("RD_WGRAD_FOLD", "1") != "0":
        rows = sorted({t.s[0] for t in taps})
        cols = sorted({t.s[1] for t in taps})
        full = len(rows) * len(cols) == ntaps and cols == list(range(cols[0], cols[0] + len(cols))) and \
            [t.s for t in taps] == [(r, c) for r in rows for c in cols]
        if full and 2 <= len(cols) <= 4 and len(rows) * 64 <= 512:
            p.fold_rows, p.fold_len = len(rows), len(cols)
    ntiles = geo["tiles_y"] * geo["tiles_x"] * B
    ctas_other = ncob * ncib * ntg
    # one resident wave: every extra wave pays the CTA prologue (TMEM alloc, ring zero-fill) and the final fp32
    # reduction of its accumulators again (measured: 1 wave is ~5 % faster than 2 on the layer1 shapes)
    p.max_ctas = max(1, min(ntiles, sm_budget // ctas_other))
    # scatter table: dw[(t*N + n)*Cx + c]  ->  parameter widx_t[c, n]
    pi, di = [], []
    for ti, t in enumerate(taps):
        c, n = np.nonzero(t.widx >= 0)
        pi.append(t.widx[c, n].astype(np.int64))
        di.append(((ti * g.N + n) * g.Cx + c).astype(np.int64))
    scatter = (np.concatenate(pi), np.concatenate(di))
    return WgradPlan(g=g, params=p, dw_elems=ntaps * g.N * g.Cx, scatter=scatter,
                     info=dict(geo=geo, NS=NS, Mc=Mc, Nc=Nc, tg_size=tg_size, ntg=ntg, ntiles=ntiles, gcopies=int(p.gcopies),
                               njobs=int(p.njobs)))
