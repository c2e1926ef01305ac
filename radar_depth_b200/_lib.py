"""ctypes binding of libradar_depth_b200.so (C ABI in include/radar_depth_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (or ``python -m radar_depth_b200.build``).
There is NO fallback: if the shared object is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libradar_depth_b200.so")

RD_BF16, RD_F32 = 0, 1
RD_MAX_TAPS, RD_MAX_GROUPS, RD_MAX_PHASES = 32, 16, 4


class RdError(RuntimeError):
    pass


class View(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("pitch", C.c_int32), ("coff", C.c_int32)]


class Tap(C.Structure):
    _fields_ = [("a_shift", C.c_int32), ("phase", C.c_int32), ("first", C.c_int32), ("pad_", C.c_int32)]


class BnJob(C.Structure):
    _fields_ = [("kind", C.c_int32), ("C", C.c_int32), ("sum_a", C.c_void_p), ("sum_b", C.c_void_p), ("count", C.c_double),
                ("gamma", C.c_void_p), ("beta", C.c_void_p), ("running_mean", C.c_void_p), ("running_var", C.c_void_p),
                ("nbt", C.c_void_p), ("v0", C.c_void_p), ("v1", C.c_void_p), ("v2", C.c_void_p), ("v3", C.c_void_p),
                ("cA", C.c_void_p), ("cB", C.c_void_p), ("cC", C.c_void_p), ("momentum", C.c_float), ("eps", C.c_float)]


RD_MAX_BN_JOBS = 2


class BnTail(C.Structure):
    """rd_bn_tail: BatchNorm finalisation(s) executed by the last CTA of the statistics-producing kernel."""
    _fields_ = [("counter", C.c_void_p), ("njobs", C.c_int32), ("slots", C.c_int32), ("slot_stride", C.c_int64),
                ("job", BnJob * RD_MAX_BN_JOBS)]


class ConvParams(C.Structure):
    _fields_ = [
        ("src", View), ("srcH", C.c_int32), ("srcW", C.c_int32), ("Cin", C.c_int32), ("S", C.c_int32),
        ("ld_scale", C.c_void_p), ("ld_shift", C.c_void_p), ("ld_slope", C.c_float), ("B", C.c_int32),
        ("Hb", C.c_int32), ("Wb", C.c_int32), ("Ht", C.c_int32), ("Wt", C.c_int32), ("Wl", C.c_int32),
        ("plane_rows", C.c_int32), ("plane_slots", C.c_int32), ("chunk_stride", C.c_int32),
        ("sy_min", C.c_int32), ("sx_min", C.c_int32),
        ("MB", C.c_int32), ("tiles_y", C.c_int32), ("tiles_x", C.c_int32),
        ("P", C.c_int32), ("OS", C.c_int32), ("ntaps", C.c_int32), ("ngroups", C.c_int32),
        ("phase_y", C.c_int32 * RD_MAX_PHASES), ("phase_x", C.c_int32 * RD_MAX_PHASES),
        ("taps", Tap * RD_MAX_TAPS),
        ("grp_first", C.c_int32 * RD_MAX_GROUPS), ("grp_n", C.c_int32 * RD_MAX_GROUPS),
        ("wpk", C.c_void_p), ("N", C.c_int32), ("nblk", C.c_int32),
        ("dst", View), ("dstH", C.c_int32), ("dstW", C.c_int32),
        ("epi", C.c_int32), ("addend", View), ("zsrc", View),
        ("ep_scale", C.c_void_p), ("ep_shift", C.c_void_p), ("ep_slope", C.c_float),
        ("stats", C.c_void_p), ("stats_stride", C.c_int32), ("tail", BnTail),
        ("IS", C.c_int32), ("WS", C.c_int32), ("istage_bytes", C.c_int32), ("wstage_bytes", C.c_int32),
        ("act_dtype", C.c_int32), ("max_ctas", C.c_int32), ("dbg", C.c_void_p), ("dbg_flags", C.c_int32), ("src_planes", C.c_int32),
        ("ep_split", C.c_int32), ("ep_slope_b", C.c_float),
    ]


class WTap(C.Structure):
    _fields_ = [("g_off", C.c_int32), ("x_shift", C.c_int32)]


class WgradParams(C.Structure):
    _fields_ = [
        ("gy", View), ("gH", C.c_int32), ("gW", C.c_int32), ("Cout", C.c_int32), ("Sg", C.c_int32),
        ("x", View), ("xH", C.c_int32), ("xW", C.c_int32), ("Cin", C.c_int32), ("Sx", C.c_int32),
        ("ld_scale", C.c_void_p), ("ld_shift", C.c_void_p), ("ld_slope", C.c_float), ("B", C.c_int32),
        ("Hb", C.c_int32), ("Wb", C.c_int32), ("Ht", C.c_int32), ("Wt", C.c_int32), ("Wl", C.c_int32),
        ("KS", C.c_int32), ("g_chunk_stride", C.c_int32), ("x_chunk_stride", C.c_int32), ("x_plane_rows", C.c_int32), ("x_plane_slots", C.c_int32),
        ("sy_min", C.c_int32), ("sx_min", C.c_int32), ("tiles_y", C.c_int32), ("tiles_x", C.c_int32),
        ("ntaps", C.c_int32), ("tg_size", C.c_int32), ("ntg", C.c_int32),
        ("taps", WTap * RD_MAX_TAPS),
        ("Mc", C.c_int32), ("ncob", C.c_int32), ("Nc", C.c_int32), ("ncib", C.c_int32),
        ("dw", C.c_void_p), ("NS", C.c_int32), ("stage_bytes", C.c_int32), ("g_bytes", C.c_int32),
        ("act_dtype", C.c_int32), ("max_ctas", C.c_int32), ("dbg", C.c_void_p), ("dbg_flags", C.c_int32), ("x_planes", C.c_int32),
        ("fold_rows", C.c_int32), ("fold_len", C.c_int32), ("pad2_", C.c_int32),
        ("gcopies", C.c_int32), ("njobs", C.c_int32), ("gcopy_dy", C.c_int32 * 8), ("gcopy_dx", C.c_int32 * 8),
        ("tile_oy", C.c_int32), ("tile_ox", C.c_int32), ("job_tap", (C.c_int8 * 8) * RD_MAX_TAPS),
    ]


class AugSample(C.Structure):
    """rd_aug_sample: per-sample parameters of the GPU input pipeline (radar_depth_b200/dataset/gpu_pipeline.py)."""
    _fields_ = [("m00", C.c_double), ("m01", C.c_double), ("m10", C.c_double), ("m11", C.c_double), ("off0", C.c_double),
                ("off1", C.c_double), ("factor", C.c_double * 3), ("depth_div", C.c_float), ("identity_rot", C.c_int32),
                ("flip", C.c_int32), ("crop_i", C.c_int32), ("crop_j", C.c_int32), ("op", C.c_int32 * 3), ("pad_", C.c_int32)]


_lib = None

_I, _F, _D, _P, _LL = C.c_int, C.c_float, C.c_double, C.c_void_p, C.c_longlong

_PROTOS = {
    "rd_version": ([], _I),
    "rd_sizeof": ([_I], _I),
    "rd_device_error": ([_P], _I),
    "rd_set_deterministic": ([_I, _P, _LL], _I),
    "rd_get_deterministic": ([], _I),
    "rd_conv_fprop": ([C.POINTER(ConvParams), _P], _I),
    "rd_conv_wgrad": ([C.POINTER(WgradParams), _P], _I),
    "rd_input_pack": ([_P, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "rd_input_pack_parts": ([_P, _P, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "rd_input_grad_channel": ([_P, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "rd_bn_finalize": ([_P, _P, _D, _P, _P, _P, _P, _P, _I, _I, _F, _F, _P, _P, _P, _P, _P], _I),
    "rd_bn_finalize_eval_multi": ([_P, _I, _F, _P], _I),
    "rd_bn_bwd_finalize": ([_P, _P, _D, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P], _I),
    "rd_bn_add_act": ([View, _P, _P, View, _P, _P, View, _LL, _I, _F, _I, _P], _I),
    "rd_join_bwd": ([View, View, View, View, View, _LL, _I, _F, _P, _P, _P, C.POINTER(BnTail), _I, _P], _I),
    "rd_bn_bwd_apply": ([View, View, View, _P, _P, _P, _LL, _I, _I, _P], _I),
    "rd_grad_stats": ([View, View, _LL, _I, _P, _P, _I, _P], _I),
    "rd_maxpool_fwd": ([View, _P, _P, _I, _I, _I, _I, _I, _F, _F, View, View, _P, _I, _I, _P, _I, _P], _I),
    "rd_maxpool_bwd_stats": ([View, View, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _P, _P, C.POINTER(BnTail), _P], _I),
    "rd_maxpool_bwd_apply": ([View, View, _P, View, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _I, _I, View, _P], _I),
    "rd_maxpool_bwd": ([View, View, _P, View, _P, _P, _I, _I, _I, _I, _I, _F, _F, _I, _I, View, _P, _P, C.POINTER(BnTail), _I, _P], _I),
    "rd_head_conv_fwd": ([View, _P, _I, _I, _I, _P, _I, _P], _I),
    "rd_head_conv_bwd": ([_P, View, _P, _I, _I, _I, View, _P, _I, _P], _I),
    "rd_bilinear_fwd": ([_P, _I, _I, _I, _P, _I, _I, _P], _I),
    "rd_bilinear_bwd": ([_P, _I, _I, _I, _P, _I, _I, _P], _I),
    "rd_l1_fwd": ([_P, _P, _LL, _P, _P, _P], _I),
    "rd_l1_bwd": ([_P, _P, _LL, _P, _P, _P, _I, _P], _I),
    "rd_smoothness_fwd": ([_P, _P, _I, _I, _I, _I, _P, _P, _P], _I),
    "rd_smoothness_bwd": ([_P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _P], _I),
    "rd_depth_metrics": ([_P, _P, _LL, _F, _F, _P, _P], _I),
    "rd_sid_filter": ([_P, _P, _LL, _P, _P, _P], _I),
    "rd_pack_weights": ([_P, _P, _P, _LL, _P], _I),
    "rd_weights_hash": ([_P, _LL, _P, _I, _P, _P], _I),
    "rd_pack_weights_g8": ([_P, _P, _P, _P, _LL, _P, _P], _I),
    "rd_pack_weights_if": ([_P, _P, _P, _LL, _P, _P], _I),
    "rd_unpack_grads": ([_P, _P, _P, _LL, _P], _I),
    "rd_sgd": ([_P, _P, _P, _LL, _F, _F, _F, _I, _P], _I),
    "rd_sgd_scaled": ([_P, _P, _P, _LL, _F, _F, _F, _I, _F, _P], _I),
    "rd_feature_export": ([View, _P, _P, _P, _I, _I, _I, _I, _I, _P], _I),
    "rd_feature_import": ([_P, View, _I, _I, _I, _I, _I, _P], _I),
    "rd_aug_rgb": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P], _I),
    "rd_aug_pack": ([_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P], _I),
}
EXPORTS = tuple(_PROTOS.keys()) + ("rd_last_error",)


def load():
    """Load the shared library (once).  Raises RdError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RdError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    lib.rd_last_error.argtypes = []
    lib.rd_last_error.restype = C.c_char_p
    for name, (args, res) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    if lib.rd_sizeof(2) != C.sizeof(BnTail):
        raise RdError(f"rd_bn_tail layout mismatch: {lib.rd_sizeof(2)} vs {C.sizeof(BnTail)}")
    if lib.rd_sizeof(0) != C.sizeof(ConvParams) or lib.rd_sizeof(1) != C.sizeof(WgradParams):
        raise RdError("parameter block layout mismatch between _lib.py and the built library: "
                      f"{lib.rd_sizeof(0)} vs {C.sizeof(ConvParams)}, {lib.rd_sizeof(1)} vs {C.sizeof(WgradParams)}")
    if lib.rd_sizeof(3) != C.sizeof(AugSample):
        raise RdError(f"rd_aug_sample layout mismatch: {lib.rd_sizeof(3)} vs {C.sizeof(AugSample)}")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().rd_last_error().decode("utf-8", "replace")
        raise RdError(f"{what or 'rd call'} failed ({rc}): {msg}")


def call(name: str, *args):
    lib = load()
    check(getattr(lib, name)(*args), name)
