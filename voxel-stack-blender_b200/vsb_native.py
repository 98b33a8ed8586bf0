"""
vsb_native -- ctypes binding of libvsb_b200.so (include/vsb.h).

This is the stub a maintainer of the reference would add: the reference has no FFI,
its hot path calls cv2 / numpy / scipy / numba; every such call site maps to one entry point here.
There is NO CPU fallback: if the shared library is missing or no CUDA device is present the import /
first call fails loudly.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int8, c_int32, c_size_t, c_uint8, c_ulonglong, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VSB_B200_LIB") or os.path.join(_HERE, "libvsb_b200.so")  # override: A/B builds of the same ABI

VSB_MAX_PRIORS = 16
VSB_MAX_WINDOW = 128     # receding_layers a stack engine can hold (csrc/vsb_stack.cu STACK_MAX_WINDOW)
VSB_MAX_XY_OPS = 16
VSB_FAST_REACH = 63

Z_FIXED, Z_WEIGHTED, Z_CONST, Z_DIST, Z_ENHANCED, Z_NONE, Z_MINMAX, Z_FIXED_FAR, Z_WEIGHTED_FAR, Z_ENHANCED_FAR = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9
METRIC_CHAMFER5, METRIC_EXACT = 0, 1
OP_NONE, OP_LUT, OP_GAUSSIAN, OP_BILATERAL, OP_MEDIAN, OP_UNSHARP, OP_RESIZE = 0, 1, 2, 3, 4, 5, 6
INTER_NEAREST, INTER_LINEAR, INTER_CUBIC, INTER_LANCZOS4, INTER_AREA, INTER_LINEAR_F32 = 0, 1, 2, 3, 4, 5
INTERP_BY_NAME = {"NEAREST": INTER_NEAREST, "BILINEAR": INTER_LINEAR, "BICUBIC": INTER_CUBIC, "LANCZOS4": INTER_LANCZOS4,
                  "AREA": INTER_AREA}   # xy_blend_processor.py:84-88; unknown names fall back to LANCZOS4 (:89)


class VsbError(RuntimeError):
    pass


class ZParams(ctypes.Structure):
    _fields_ = [("mode", c_int), ("metric", c_int), ("n_priors", c_int), ("reach", c_int), ("const_val", c_int),
                ("fade", c_float), ("den", c_float), ("total_weight", c_float),
                ("weights", c_float * VSB_MAX_PRIORS), ("fades", c_float * VSB_MAX_PRIORS),
                ("dens", c_float * VSB_MAX_PRIORS), ("enhanced_arith", c_int), ("fade_limit", c_double),
                ("aniso_w", c_int), ("aniso_h", c_int), ("aniso_reach", c_int)]


class XYOp(ctypes.Structure):
    _fields_ = [("type", c_int), ("lut", c_uint8 * 256), ("kx", c_int32 * 63), ("ky", c_int32 * 63),
                ("nkx", c_int), ("nky", c_int), ("radius", c_int), ("ntaps", c_int),
                ("tap_dy", c_int8 * 1024), ("tap_dx", c_int8 * 1024), ("tap_sw", c_float * 1024),
                ("cw", c_float * 256), ("median_ksize", c_int), ("unsharp_amount", c_float), ("unsharp_threshold", c_int),
                ("resize_w", c_int), ("resize_h", c_int), ("interp", c_int)]


class RoiStat(ctypes.Structure):
    _fields_ = [("root", c_int), ("area", c_int), ("x0", c_int), ("y0", c_int), ("x1", c_int), ("y1", c_int),
                ("first_block", c_int), ("pad", c_int), ("sum_x", c_ulonglong), ("sum_y", c_ulonglong)]


class XYFused(ctypes.Structure):
    _fields_ = [("has_bilateral", c_int), ("radius", c_int), ("ntaps", c_int), ("tap_dy", c_int8 * 96),
                ("tap_dx", c_int8 * 96), ("tap_sw", c_float * 96), ("cw", c_float * 256), ("nkx", c_int), ("nky", c_int),
                ("kx", c_int32 * 7), ("ky", c_int32 * 7), ("two_valued_min_same", c_int)]


_SIGS = {
    "vsb_last_error_string": (c_char_p, []),
    "vsb_version": (c_int, []),
    "vsb_launch_count": (c_ulonglong, []),
    "vsb_threshold_pack": (c_int, [c_void_p, c_size_t, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "vsb_bits_unpack": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "vsb_bits_or": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_void_p, c_void_p]),
    "vsb_maskdiff": (c_int, [c_void_p, POINTER(c_void_p), c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vsb_chamfer_table": (c_int, [c_void_p, c_int]),
    "vsb_dt_chamfer5": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_float, c_void_p, c_size_t, c_void_p]),
    "vsb_dt_exact_windowed": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_size_t, c_void_p]),
    "vsb_edt_exact_scratch_bytes": (c_size_t, [c_int, c_int]),
    "vsb_edt_exact": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vsb_zblend_fused": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p, POINTER(c_void_p), c_int,
                                 POINTER(ZParams), c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t,
                                 c_void_p]),
    "vsb_rec_dist": (c_int, [c_void_p, POINTER(c_void_p), c_int, c_int, c_int, POINTER(ZParams), c_void_p, c_void_p, c_void_p, c_size_t,
                             c_void_p]),
    "vsb_ccl8_max": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "vsb_enhanced_fade": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p, c_int, c_void_p, c_size_t, c_void_p,
                                  c_void_p, c_double, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vsb_enhanced_from_labels": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_int, c_void_p,
                                         c_void_p, c_void_p]),
    "vsb_lut_apply": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vsb_merge_max": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "vsb_gaussian_q88": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                 c_size_t, c_void_p]),
    "vsb_bilateral": (c_int, [c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_size_t, c_void_p]),
    "vsb_median": (c_int, [c_void_p, c_size_t, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "vsb_coldist_scratch_elems": (c_size_t, [c_int, c_int, c_int]),
    "vsb_coldist": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "vsb_dt_unbounded": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p,
                                 c_size_t, c_void_p]),
    "vsb_zblend_unbounded": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p, POINTER(c_void_p), c_int, c_int, c_int,
                                     c_int, c_float, c_float, c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_void_p,
                                     c_void_p, c_void_p, c_size_t, c_void_p]),
    "vsb_rec_dist_unbounded": (c_int, [c_void_p, POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p,
                                       c_void_p, c_size_t, c_void_p, c_void_p]),
    "vsb_weighted_unbounded": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p, POINTER(c_void_p), c_int, POINTER(ZParams),
                                       c_void_p, c_size_t, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_size_t,
                                       c_void_p]),
    "vsb_pack_flags": (c_int, [c_void_p, c_size_t, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "vsb_pack_flags_lut": (c_int, [c_void_p, c_size_t, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                   POINTER(c_void_p), POINTER(c_void_p), c_int, c_void_p, c_void_p]),
    "vsb_layer_patch": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p, POINTER(c_void_p), c_int, POINTER(ZParams), c_void_p,
                                c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vsb_layer_fused_enhanced": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p, POINTER(c_void_p), c_int, POINTER(ZParams),
                                         c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                         c_void_p, c_void_p, c_void_p]),
    "vsb_bilateral_two_valued_min_same": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int]),
    "vsb_layer_stats": (c_int, [c_int, c_void_p]),
    "vsb_layer_fast_path_limits": (c_int, [c_int, c_int]),
    "vsb_layer_fused": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p, POINTER(c_void_p), c_void_p, POINTER(c_void_p),
                                c_int, POINTER(ZParams), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p,
                                c_void_p]),
    "vsb_stack_create": (c_int, [c_int, c_int, c_int, POINTER(ZParams), POINTER(XYOp), c_int, POINTER(c_void_p)]),
    "vsb_roi_components": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "vsb_roi_fade": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p, POINTER(c_void_p), c_int, c_int, c_void_p, c_void_p,
                             c_int, c_int, c_int, c_void_p, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p, c_size_t,
                             c_void_p]),
    "vsb_stack_destroy": (c_int, [c_void_p]),
    "vsb_stack_out_size": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int)]),
    "vsb_unsharp_combine": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, c_int, c_int, c_float, c_int, c_void_p, c_size_t,
                                    c_void_p]),
    "vsb_resize_plan_create": (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    "vsb_resize_plan_destroy": (c_int, [c_void_p]),
    "vsb_resize_plan_dims": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "vsb_resize_u8": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]),
    "vsb_resize_f32_linear": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]),
    "vsb_resize_bits_nearest": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "vsb_stack_reset": (c_int, [c_void_p]),
    "vsb_stack_push_halo": (c_int, [c_void_p, c_void_p, c_size_t]),
    "vsb_stack_push_halo_device": (c_int, [c_void_p, c_void_p, c_size_t]),
    "vsb_stack_process_host": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t]),
    "vsb_stack_submit_host": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t]),
    "vsb_stack_wait": (c_int, [c_void_p, c_int]),
    "vsb_stack_slots": (c_int, [c_void_p]),
    "vsb_stack_process_device": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t]),
    "vsb_stack_process_device_batch": (c_int, [c_void_p, c_void_p, c_size_t, ctypes.c_longlong, c_void_p, c_size_t,
                                               ctypes.c_longlong, c_int, c_int]),
    "vsb_stack_sync": (c_int, [c_void_p]),
    "vsb_stack_profile_enable": (c_int, [c_void_p, c_int]),
    "vsb_stack_profile_read": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "vsb_stack_stream": (c_void_p, [c_void_p]),
    "vsb_synth_layer": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
}

_lib = None


def load():
    """dlopen libvsb_b200.so and type every entry point of include/vsb.h.  No device work."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VsbError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def exported_symbols():
    return sorted(_SIGS)


def check(rc: int, what: str = "") -> int:
    if rc < 0:
        msg = load().vsb_last_error_string().decode("utf-8", "replace")
        raise VsbError(f"{what or 'libvsb_b200'} failed ({rc}): {msg}")
    return rc


def launch_count() -> int:
    return int(load().vsb_launch_count())


def chamfer_table(R: int) -> np.ndarray:
    """T[(R+1),(R+1)] float32, host side (vsb_chamfer_table)."""
    T = np.empty((R + 1, R + 1), np.float32)
    check(load().vsb_chamfer_table(T.ctypes.data_as(c_void_p), int(R)), "vsb_chamfer_table")
    return T


def ptr_array(ptrs):
    arr = (c_void_p * max(1, len(ptrs)))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr
