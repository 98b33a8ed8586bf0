"""
vsb_params -- host-side parameter preparation for the device kernels: everything the reference computes
per call on the CPU that is *not* per-pixel work (fade parameters, Gaussian Q8.8 taps, bilateral weight
tables, LUT tables).  Pure Python / NumPy, no device access, no oracle import.

Numerics that must match third-party code bit for bit use the same C library routines that code uses
(math.exp / math.sqrt = libm, like OpenCV's std::exp), not NumPy's vectorised variants.
"""
from __future__ import annotations

import math
from typing import Optional, List, Sequence

import numpy as np

import vsb_native as N

f32 = np.float32


class UnsupportedOnDevice(NotImplementedError):
    """A parameter combination outside what the device kernels cover ("Not built")."""


def _mode_name(cfg) -> str:
    m = getattr(cfg, "blending_mode", "fixed_fade")
    return getattr(m, "value", m)


def chamfer_reach(F: float) -> int:
    """Largest Chebyshev offset whose distance can still be < F (both metrics are >= max|d|)."""
    return max(0, int(math.ceil(F)) - 1)


def anisotropic_dims(width: int, height: int, x_factor: float, y_factor: float):
    """Size of the resized mask of the anisotropic branch (processing_core.py:360-367); None = factors clamp to 1."""
    xf = max(0.1, min(10.0, x_factor))
    yf = max(0.1, min(10.0, y_factor))
    if xf == 1.0 and yf == 1.0:
        return None
    return int(width * xf), int(height * yf)


def zparams_from_config(cfg, metric: str = "chamfer5", width: Optional[int] = None, height: Optional[int] = None,
                        weighted_k: Optional[int] = None) -> N.ZParams:
    """Fade parameters as the reference derives them from Config (processing_core.py:112-113,
    :138-140, :227-257, :387).  `width`/`height` (the layer size) are needed by the anisotropic Enhanced EDT only;
    `weighted_k` = number of prior masks a per-call weighted blend was handed (default: config.receding_layers, what
    a stack holds).  Raises UnsupportedOnDevice for the combinations DESIGN.md lists as not built."""
    z = N.ZParams()
    z.metric = N.METRIC_EXACT if metric == "exact" else N.METRIC_CHAMFER5
    mode = _mode_name(cfg)
    F = float(cfg.fixed_fade_distance_receding)
    z.fade_limit = F
    z.enhanced_arith = 1 if getattr(cfg, "use_numba_jit", False) else 0
    if mode == "roi_fade":
        raise UnsupportedOnDevice("ROI_FADE carries tracker state from layer to layer: drive it through "
                                  "vsb_engine.RoiStackEngine (what ProcessingPipelineThread does), not StackEngine")
    if mode == "weighted_stack":
        # the reference only ever reads the first min(len(priors), len(weights)) entries (processing_core.py:231-233);
        # a stack never holds more than receding_layers priors
        held = weighted_k if weighted_k is not None else getattr(cfg, "receding_layers", len(cfg.manual_weights))
        k_max = min(len(cfg.manual_weights), max(0, int(held)))
        if k_max > N.VSB_MAX_PRIORS:
            raise UnsupportedOnDevice(f"weighted stack over more than {N.VSB_MAX_PRIORS} layers")
        weights = list(cfg.manual_weights)[:k_max]
        fades = list(cfg.fade_distances_receding)
        z.mode = N.Z_WEIGHTED
        z.n_priors = len(weights)
        fmax = 0.0
        for i, w in enumerate(weights):
            Fi = float(fades[i]) if i < len(fades) else F
            z.weights[i] = f32(w)
            z.fades[i] = f32(Fi)
            z.dens[i] = f32(Fi) if Fi > 0 else f32(1.0)
            if w > 0:
                fmax = max(fmax, Fi)
        z.total_weight = f32(float(sum(weights))) if weights else f32(0)
        z.reach = chamfer_reach(fmax)
        z.fade = f32(fmax)
        z.den = f32(1.0)
        if z.reach > N.VSB_FAST_REACH:      # a fade distance beyond the tile kernels' reach: unbounded distances
            z.mode = N.Z_WEIGHTED_FAR
            z.reach = 0
    elif mode == "enhanced_edt":
        if getattr(getattr(cfg, "anisotropic_params", None), "enabled", False):
            ap = cfg.anisotropic_params
            if anisotropic_dims(1000, 1000, ap.x_factor, ap.y_factor) is not None:
                if width is None or height is None:
                    raise ValueError("anisotropic Enhanced EDT: zparams_from_config needs the layer width and height")
                if metric != "chamfer5":
                    raise UnsupportedOnDevice("anisotropic correction with the exact metric")
                z.aniso_w, z.aniso_h = anisotropic_dims(width, height, ap.x_factor, ap.y_factor)
                if z.aniso_w < 1 or z.aniso_h < 1:
                    raise ValueError("anisotropic Enhanced EDT: the resized mask would be empty")   # cv2.resize raises too
                z.aniso_reach = chamfer_reach(max(F, 0.0) + 3.0)   # bilinear taps of a pixel with d < F lie within +3
                if z.aniso_reach > N.VSB_FAST_REACH:
                    raise UnsupportedOnDevice(f"anisotropic Enhanced EDT with fade distance {F} (reach {z.aniso_reach} > "
                                              f"{N.VSB_FAST_REACH})")
        z.mode = N.Z_ENHANCED
        z.n_priors = N.VSB_MAX_PRIORS
        z.fade = f32(F)
        z.den = f32(F) if F > 0 else f32(1.0)
        z.reach = chamfer_reach(F)
        if z.reach > N.VSB_FAST_REACH and z.aniso_w == 0:
            z.mode = N.Z_ENHANCED_FAR
            z.reach = 0
    else:  # fixed_fade (the dispatcher's default branch, processing_core.py:434-440)
        z.n_priors = N.VSB_MAX_PRIORS
        if not cfg.use_fixed_fade_receding:
            # global min-max normalisation over the receding pixels (:141-145): unbounded distances
            z.mode = N.Z_MINMAX
            z.fade = f32(F)
            z.den = f32(1.0)
            z.reach = 0
        elif F > 0 and chamfer_reach(F) > N.VSB_FAST_REACH:
            z.mode = N.Z_FIXED_FAR          # beyond the tile kernels' reach: column distances + row search
            z.fade = f32(F)
            z.den = f32(F)
            z.reach = 0
        elif F > 0:
            z.mode = N.Z_FIXED
            z.fade = f32(F)
            z.den = f32(F)
            z.reach = chamfer_reach(F)
        else:
            # clip(d, 0, F<=0) == F, denominator 1.0: one constant on every receding pixel (:138-149)
            z.mode = N.Z_CONST
            with np.errstate(invalid="ignore", over="ignore"):
                v = (f32(255) * (f32(1.0) - f32(F) / f32(1.0)))
                z.const_val = int(np.array([v], np.float32).astype(np.uint8)[0])
            z.reach = 0
            z.fade = f32(F)
            z.den = f32(1.0)
    if z.reach > N.VSB_FAST_REACH:
        raise UnsupportedOnDevice(f"{mode}: fade distance needs reach {z.reach} > {N.VSB_FAST_REACH}")
    return z


# ---- Gaussian taps ---------------------------------------------------------------------------------------
_SMALL_GAUSS = {1: [1.0], 3: [0.25, 0.5, 0.25], 5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
                7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125],
                9: [4 / 256, 13 / 256, 30 / 256, 51 / 256, 60 / 256, 51 / 256, 30 / 256, 13 / 256, 4 / 256]}


def gaussian_taps_q88(ksize: int, sigma: float) -> np.ndarray:
    """The Q8.8 taps cv2.GaussianBlur uses on uint8 input: float kernel of getGaussianKernel, rounded
    from the edges inwards with the rounding error carried along, centre tap = 256 - sum(others)."""
    if sigma <= 0 and ksize in _SMALL_GAUSS:
        k = list(_SMALL_GAUSS[ksize])
    else:
        s = sigma if sigma > 0 else 0.15 * ksize + 0.35
        scale = -0.125 / (s * s)
        half = (ksize - 1) // 2
        vals = [math.exp(float(x * x) * scale) for x in range(1 - ksize, 1 - ksize + 2 * half, 2)]
        mul = 1.0 / (2.0 * sum(vals) + 1.0)
        k = [v * mul for v in vals] + [mul] + [v * mul for v in reversed(vals)]
    n, half = ksize, ksize // 2
    q = [0] * n
    err = 0.0
    for i in range(half):
        v = k[i] * 256.0 + err
        r = int(round(v))          # round-half-even, like cvRound
        err = v - r
        q[i] = q[n - 1 - i] = r
    q[half] = 256 - 2 * sum(q[:half])
    return np.array(q, np.int32)


def bilateral_tables(d: int, sigma_color: float, sigma_space: float):
    """radius, tap offsets (row-major), float32 space weights and the 256 colour weights of
    cv2.bilateralFilter on uint8."""
    if sigma_color <= 0:
        sigma_color = 1.0
    if sigma_space <= 0:
        sigma_space = 1.0
    gc = -0.5 / (sigma_color * sigma_color)
    gs = -0.5 / (sigma_space * sigma_space)
    radius = int(round(sigma_space * 1.5)) if d <= 0 else d // 2
    radius = max(radius, 1)
    cw = np.array([math.exp(i * i * gc) for i in range(256)], np.float64).astype(np.float32)
    dy, dx, sw = [], [], []
    for i in range(-radius, radius + 1):
        for j in range(-radius, radius + 1):
            r = math.sqrt(float(i * i + j * j))
            if r > radius:
                continue
            dy.append(i)
            dx.append(j)
            sw.append(math.exp(r * r * gs))
    return (radius, np.array(dy, np.int8), np.array(dx, np.int8),
            np.array(sw, np.float64).astype(np.float32), cw)


def odd_ksize(k: int) -> int:
    if k <= 0:
        return 1
    return k if k % 2 else k + 1


def resize_dims(cur_w: int, cur_h: int, width, height):
    """Target size rule of apply_resize (xy_blend_processor.py:71-82); None = the op returns its input."""
    if width is None and height is None:
        return None
    if width is None:
        width = int(cur_w * (height / cur_h))
    elif height is None:
        height = int(cur_h * (width / cur_w))
    if width == cur_w and height == cur_h:
        return None
    return int(width), int(height)


def xy_ops_from_pipeline(pipeline_ops: Sequence, lut_for_op, width: Optional[int] = None,
                         height: Optional[int] = None) -> List[N.XYOp]:
    """XYBlendOperation list -> device op descriptors.  `lut_for_op(op)` returns the uint8[256] table of an
    apply_lut op (identity on any failure, xy_blend_processor.py:125-130).  `width`/`height` = size of the layers
    entering the chain; only a resize op needs them (its target size may depend on the current size)."""
    out: List[N.XYOp] = []
    cur_w, cur_h = width, height
    for op in pipeline_ops or []:
        t = str(getattr(op, "type", "none")).lower()
        x = N.XYOp()
        if t == "apply_lut":
            x.type = N.OP_LUT
            lut = np.ascontiguousarray(lut_for_op(op), np.uint8)
            ctypes_copy(x.lut, lut)
        elif t == "gaussian_blur":
            x.type = N.OP_GAUSSIAN
            kx, ky = odd_ksize(op.gaussian_ksize_x), odd_ksize(op.gaussian_ksize_y)
            sx, sy = float(op.gaussian_sigma_x), float(op.gaussian_sigma_y)
            if sy <= 0:
                sy = sx
            if kx > 63 or ky > 63:
                raise UnsupportedOnDevice("Gaussian kernels wider than 63 taps")
            qx, qy = gaussian_taps_q88(kx, sx), gaussian_taps_q88(ky, sy)
            x.nkx, x.nky = kx, ky
            ctypes_copy(x.kx, qx)
            ctypes_copy(x.ky, qy)
        elif t == "bilateral_filter":
            x.type = N.OP_BILATERAL
            radius, dy, dx, sw, cw = bilateral_tables(int(op.bilateral_d), float(op.bilateral_sigma_color),
                                                      float(op.bilateral_sigma_space))
            if len(sw) > 1024:
                raise UnsupportedOnDevice("bilateral windows with more than 1024 taps")
            x.radius, x.ntaps = radius, len(sw)
            ctypes_copy(x.tap_dy, dy)
            ctypes_copy(x.tap_dx, dx)
            ctypes_copy(x.tap_sw, sw)
            ctypes_copy(x.cw, cw)
        elif t == "median_blur":
            x.type = N.OP_MEDIAN
            x.median_ksize = odd_ksize(int(op.median_ksize))
        elif t == "unsharp_mask":
            x.type = N.OP_UNSHARP
            k = odd_ksize(int(op.unsharp_blur_ksize))
            if k > 63:
                raise UnsupportedOnDevice("Gaussian kernels wider than 63 taps")
            q = gaussian_taps_q88(k, float(op.unsharp_blur_sigma))
            x.nkx = x.nky = k
            ctypes_copy(x.kx, q)
            ctypes_copy(x.ky, q)
            x.unsharp_amount = float(op.unsharp_amount)
            x.unsharp_threshold = int(op.unsharp_threshold)
        elif t == "resize":
            if op.resize_width is None and op.resize_height is None:
                continue
            if cur_w is None or cur_h is None:
                raise ValueError("a resize op needs the layer size: pass width= and height= to xy_ops_from_pipeline")
            dims = resize_dims(cur_w, cur_h, op.resize_width, op.resize_height)
            if dims is None:
                continue
            if dims[0] < 1 or dims[1] < 1:
                raise ValueError(f"resize to {dims[0]}x{dims[1]}")   # cv2.resize raises on an empty target as well
            x.type = N.OP_RESIZE
            x.resize_w, x.resize_h = dims
            x.interp = N.INTERP_BY_NAME.get(str(op.resample_mode).upper(), N.INTER_LANCZOS4)
            cur_w, cur_h = dims
        elif t == "none":
            continue
        else:
            print(f"Warning: Unknown XY blend operation type '{t}'. Skipping.")
            continue
        out.append(x)
    return out


def ctypes_copy(dst, arr: np.ndarray) -> None:
    import ctypes
    arr = np.ascontiguousarray(arr)
    ctypes.memmove(dst, arr.ctypes.data, arr.nbytes)
