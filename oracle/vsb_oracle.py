"""
vsb_oracle -- TEST INFRASTRUCTURE ONLY (CPU oracle for the Z-blend -> LUT -> XY hot path).

A CPU restatement of the reference's per-layer algorithm, NumPy for the element-wise
float32/uint8 arithmetic (the reference itself is NumPy there, so the rounding sequence is
identical) and plain C (oracle/c/vsb_oracle.c, built by oracle/Makefile) for the pieces the
reference delegates to OpenCV / SciPy / Numba.  Every function cites the reference lines it
follows (paths relative to /root/reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module, and only as the checker.  The product (voxel-stack-blender_b200/) never
imports it and has no CPU fallback.

Parity pin: tests/golden/*.npz hold outputs of the *unmodified reference* (imported from
/root/reference by tests/golden/make_golden.py in the build container, cv2 4.13.0 / numpy 2.3.5 /
scipy 1.18.1 / numba 0.65.0); tests/test_oracle_golden.py checks this oracle against every one of
them, and against the numeric assertions of the reference's own tests/test_processing.py.
"""
from __future__ import annotations

import collections
import ctypes
import json
import math
import os
import subprocess
from typing import Iterable, List, Optional, Sequence

import numpy as np

import vsb_oracle_geom as geom

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libvsb_oracle.so")
_lib = None

f32 = np.float32


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc).  Building the checker is not using it."""
    src = os.path.join(_HERE, "c", "vsb_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, i, d = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        L.orc_chamfer5_table.argtypes = [vp, i]
        L.orc_chamfer5_brute.argtypes = [vp, i, i, i, vp, vp]
        L.orc_chamfer5.argtypes = [vp, i, i, i, vp, vp]
        L.orc_chamfer5.restype = i
        L.orc_chamfer5_twopass.argtypes = [vp, i, i, vp]
        L.orc_chamfer5_unbounded.argtypes = [vp, i, i, vp]
        L.orc_chamfer5_unbounded.restype = i
        L.orc_edt_exact_sq.argtypes = [vp, i, i, vp]
        L.orc_edt_exact_sq.restype = i
        L.orc_ccl8.argtypes = [vp, i, i, vp]
        L.orc_ccl8.restype = i
        L.orc_label_max.argtypes = [vp, vp, ctypes.c_size_t, i, vp]
        L.orc_gauss_q88.argtypes = [vp, i, i, vp, i, vp, i, vp]
        L.orc_bilateral.argtypes = [vp, i, i, i, d, d, i, vp]
        L.orc_median.argtypes = [vp, i, i, i, vp]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _u8(a) -> np.ndarray:
    a = np.ascontiguousarray(a)
    assert a.dtype == np.uint8 and a.ndim == 2
    return a


# --------------------------------------------------------------------------------------
# masks                                                        processing_core.py:44,47-67
# --------------------------------------------------------------------------------------
def threshold(gray: np.ndarray, thresh: int = 127) -> np.ndarray:
    """cv2.threshold(img,127,255,THRESH_BINARY): 255 where g>127 else 0 (processing_core.py:44)."""
    return np.where(gray > thresh, np.uint8(255), np.uint8(0)).astype(np.uint8)


def combine_priors(priors: Sequence[np.ndarray]) -> Optional[np.ndarray]:
    """find_prior_combined_white_mask (processing_core.py:47-67): OR of the list, None if empty."""
    if not priors:
        return None
    out = priors[0].copy()
    for p in priors[1:]:
        out |= p
    return out


# --------------------------------------------------------------------------------------
# distance metrics                                         processing_core.py:129,250,379
# --------------------------------------------------------------------------------------
def chamfer_reach(F: float) -> int:
    """Largest Chebyshev offset whose chamfer value can be < F (T >= max|d|)."""
    return max(0, int(math.ceil(F)) - 1)


def chamfer_table(R: int) -> np.ndarray:
    T = np.empty((R + 1, R + 1), np.float32)
    lib().orc_chamfer5_table(_p(T), R)
    return T


def chamfer5(cur_white: np.ndarray, R: int, brute: bool = False) -> np.ndarray:
    """distanceTransform(~cur, DIST_L2, 5): chamfer distance to the nearest cur_white!=0 pixel,
    exact for every value < R+1 (inf where no seed lies within Chebyshev reach R)."""
    cur_white = _u8(cur_white)
    H, W = cur_white.shape
    T = chamfer_table(R)
    out = np.empty((H, W), np.float32)
    if brute:
        lib().orc_chamfer5_brute(_p(cur_white), W, H, R, _p(T), _p(out))
    else:
        rc = lib().orc_chamfer5(_p(cur_white), W, H, R, _p(T), _p(out))
        assert rc == 0
    return out


def chamfer5_twopass(cur_white: np.ndarray) -> np.ndarray:
    """The two-pass raster algorithm of the library call itself (unbounded reach)."""
    cur_white = _u8(cur_white)
    H, W = cur_white.shape
    out = np.empty((H, W), np.float32)
    lib().orc_chamfer5_twopass(_p(cur_white), W, H, _p(out))
    return out


def chamfer5_unbounded(cur_white: np.ndarray) -> np.ndarray:
    """Unbounded-reach chamfer distance, exact integer costs rounded once to float32 (inf if no seed).  The
    definition of the unbounded device path; within a few ulps of the tabulated `chamfer5`."""
    cur_white = _u8(cur_white)
    H, W = cur_white.shape
    out = np.empty((H, W), np.float32)
    assert lib().orc_chamfer5_unbounded(_p(cur_white), W, H, _p(out)) == 0
    return out


FAST_REACH = 63   # the tabulated (sequential-sum) definition covers reach <= 63; beyond it the exact-integer one applies


def edt_exact_sq(cur_white: np.ndarray) -> np.ndarray:
    """Exact squared Euclidean distance (int32) to the nearest cur_white!=0 pixel; -1 if none."""
    cur_white = _u8(cur_white)
    H, W = cur_white.shape
    out = np.empty((H, W), np.int32)
    assert lib().orc_edt_exact_sq(_p(cur_white), W, H, _p(out)) == 0
    return out


def edt_exact(cur_white: np.ndarray) -> np.ndarray:
    """float32 sqrt of the exact squared distance (cv2.DIST_MASK_PRECISE equivalent)."""
    d2 = edt_exact_sq(cur_white)
    out = np.sqrt(d2.astype(np.float32))
    out[d2 < 0] = np.inf
    return out


def distance(cur_white: np.ndarray, F: float, metric: str = "chamfer5") -> np.ndarray:
    if metric == "chamfer5":
        if chamfer_reach(F) > FAST_REACH:
            return chamfer5_unbounded(cur_white)
        return chamfer5(cur_white, chamfer_reach(F))
    if metric == "exact":
        return edt_exact(cur_white)
    raise ValueError(metric)


# --------------------------------------------------------------------------------------
# gradient fades
# --------------------------------------------------------------------------------------
def _fade_u8(d_clipped: np.ndarray, den) -> np.ndarray:
    """(255 * (1 - d/den)).astype(uint8), every step rounded to float32 (processing_core.py:140-149)."""
    n = d_clipped / den
    inv = f32(1.0) - n
    with np.errstate(invalid="ignore"):
        return (f32(255.0) * inv).astype(np.uint8)


def fixed_fade(cur: np.ndarray, prior_combined: Optional[np.ndarray], F: float,
               use_fixed: bool = True, metric: str = "chamfer5") -> np.ndarray:
    """_calculate_receding_gradient_field_fixed_fade (processing_core.py:107-153)."""
    if prior_combined is None:
        return np.zeros_like(cur)
    rec = prior_combined & ~cur                                        # :119
    if not rec.any():                                                  # :122
        return np.zeros_like(cur)
    if use_fixed:                                                      # :137-140
        d = distance(cur, F, metric)
        rd = np.where(rec > 0, d, f32(0)).astype(np.float32)           # :132
        if rd.max() == 0:                                              # :133
            return np.zeros_like(cur)
        Ff = f32(F)
        den = Ff if F > 0 else f32(1.0)
        clipped = np.minimum(np.maximum(rd, f32(0)), Ff)               # np.clip(rd, 0, F)
        g = _fade_u8(clipped, den)
    else:                                                              # :141-145 global min-max
        d = chamfer5_unbounded(cur) if metric == "chamfer5" else edt_exact(cur)
        big = f32(np.finfo(np.float32).max)
        d = np.where(np.isinf(d), big, d)
        rd = np.where(rec > 0, d, f32(0)).astype(np.float32)
        if rd.max() == 0:
            return np.zeros_like(cur)
        vals = rd[rec > 0]
        mn, mx = float(vals.min()), float(vals.max())                  # minMaxLoc -> doubles
        if mx <= mn:
            return np.zeros_like(cur)
        n = (rd - f32(mn)) / f32(mx - mn)                              # float32 array op python float
        inv = f32(1.0) - n
        with np.errstate(invalid="ignore"):
            g = (f32(255.0) * inv).astype(np.uint8)
    return np.where(rec > 0, g, np.uint8(0)).astype(np.uint8)         # :152


def weighted_fade(cur: np.ndarray, priors: Sequence[np.ndarray], weights: Sequence[float],
                  fades: Sequence[float], default_fade: float, metric: str = "chamfer5") -> np.ndarray:
    """_calculate_weighted_receding_gradient_field (processing_core.py:222-284); priors newest first."""
    if not priors or not weights:
        return np.zeros_like(cur)
    k = min(len(priors), len(weights))
    if k == 0:
        return np.zeros_like(cur)
    priors, weights = list(priors[:k]), list(weights[:k])
    total = float(sum(weights))
    if total <= 0:
        return np.zeros_like(cur)
    acc = np.zeros(cur.shape, np.float32)
    fmax = max([fades[i] if i < len(fades) else default_fade for i in range(k)] + [1.0])
    d = distance(cur, fmax, metric)
    for i, (p, w) in enumerate(zip(priors, weights)):
        if w <= 0:
            continue
        Fi = fades[i] if i < len(fades) else default_fade
        den = f32(Fi) if Fi > 0 else f32(1.0)
        rec = p & ~cur
        if not rec.any():
            continue
        rd = np.where(rec > 0, d, f32(0)).astype(np.float32)
        clipped = np.minimum(np.maximum(rd, f32(0)), f32(Fi))
        n = f32(1.0) - clipped / den
        contrib = n * f32(w)
        acc += np.where(rec > 0, contrib, f32(0))
    fin = acc / f32(total)
    with np.errstate(invalid="ignore"):
        g = (f32(255.0) * fin).astype(np.uint8)
    union = np.zeros_like(cur)
    for p in priors:
        union |= p & ~cur
    return np.where(union > 0, g, np.uint8(0)).astype(np.uint8)


def ccl8(mask: np.ndarray):
    """cv2.connectedComponents(mask) (8-connectivity): (num_labels, int32 labels)."""
    mask = _u8(mask)
    H, W = mask.shape
    labels = np.empty((H, W), np.int32)
    n = lib().orc_ccl8(_p(mask), W, H, _p(labels))
    assert n >= 1
    return n, labels


def label_max(d: np.ndarray, labels: np.ndarray, num_labels: int) -> np.ndarray:
    d = np.ascontiguousarray(d, np.float32)
    labels = np.ascontiguousarray(labels, np.int32)
    mx = np.zeros(num_labels, np.float32)
    lib().orc_label_max(_p(d), _p(labels), d.size, num_labels, _p(mx))
    return mx


def enhanced_from_distance(rd: np.ndarray, labels: np.ndarray, num_labels: int, F: float,
                           flavour: str) -> np.ndarray:
    """_..._enhanced_edt_scipy (:287-303, float32) or _..._enhanced_edt_numba (:305-336, float64)."""
    mx = label_max(rd, labels, num_labels)
    if flavour == "scipy":
        max_map = mx.astype(np.float64)[labels]                         # concatenate([0], f32) -> f64
        den = np.minimum(max_map, F).astype(np.float32)
        den[den <= 0] = 1.0
        clipped = np.minimum(np.maximum(rd, f32(0)), den)
        return _fade_u8(clipped, den)
    if flavour == "numba":
        den = np.minimum(mx.astype(np.float64), float(F))[labels]
        ok = (labels > 0) & (den > 0)
        safe = np.where(ok, den, 1.0)
        inv = 1.0 - np.minimum(rd.astype(np.float64), safe) / safe
        g = (inv * 255.0).astype(np.int64)
        return np.where(ok, g, 0).astype(np.uint8)
    raise ValueError(flavour)


def enhanced_fade(cur: np.ndarray, priors: Sequence[np.ndarray], F: float, flavour: str = "scipy",
                  metric: str = "chamfer5", anisotropic: Optional[Sequence[float]] = None) -> np.ndarray:
    """_calculate_receding_gradient_field_enhanced_edt (processing_core.py:338-404).  `anisotropic` = (x_factor,
    y_factor) of an enabled AnisotropicParams: the distance map is then taken on the nearest-resized mask and
    resized back bilinearly (:357-377); distances stay in resized-pixel units, exactly like the reference."""
    if not priors:
        return np.zeros_like(cur)
    comb = combine_priors(priors)
    rec = comb & ~cur
    if not rec.any():
        return np.zeros_like(cur)
    # per-label denominators are min(max_label(d), F) == max_label(min(d, F)): clipped distances suffice
    d = None
    if anisotropic is not None:
        # reach F + 3: bilinear taps are adjacent pixels, so a tap beyond the reach only occurs where every tap is >= F
        far = f32(max(F, 0.0) + 4.0)   # stands in for "no seed within reach"; any value >= F fades to 0 all the same

        def clipped(m):
            dm = distance(m, max(F, 0.0) + 3.0, metric)
            return np.where(np.isfinite(dm), dm, far).astype(np.float32)
        d = geom.anisotropic_distance(cur, anisotropic[0], anisotropic[1], clipped)
    if d is None:
        d = distance(cur, F, metric)
    d = np.minimum(d, f32(max(F, 0.0)) if F > 0 else f32(np.finfo(np.float32).max))
    rd = np.where(rec > 0, d, f32(0)).astype(np.float32)
    n, labels = ccl8(rec)
    if n <= 1:
        return np.zeros_like(cur)
    g = enhanced_from_distance(rd, labels, n, F, flavour)
    return np.where(rec > 0, g, np.uint8(0)).astype(np.uint8)


def merge(original: np.ndarray, gradient: np.ndarray) -> np.ndarray:
    """merge_to_output (processing_core.py:442-460)."""
    return np.maximum(original, gradient)


def z_blend(cur: np.ndarray, priors_newest_first: Sequence[np.ndarray], cfg: dict) -> np.ndarray:
    """process_z_blending dispatch (processing_core.py:407-440) with the caller's prior handling
    (processing_pipeline.py:91-96).  cfg keys follow config.Config field names."""
    mode = cfg.get("blending_mode", "fixed_fade")
    metric = cfg.get("metric", "chamfer5")
    F = cfg.get("fixed_fade_distance_receding", 10.0)
    if mode == "weighted_stack":
        return weighted_fade(cur, list(priors_newest_first), cfg.get("manual_weights", [100, 75, 50, 25]),
                             cfg.get("fade_distances_receding", [10.0] * 4), F, metric)
    if mode == "enhanced_edt":
        ap = cfg.get("anisotropic_params") or {}
        return enhanced_fade(cur, list(priors_newest_first), F,
                             "numba" if cfg.get("use_numba_jit", False) else "scipy", metric,
                             (ap.get("x_factor", 1.0), ap.get("y_factor", 1.0)) if ap.get("enabled") else None)
    if mode == "fixed_fade":
        return fixed_fade(cur, combine_priors(list(priors_newest_first)), F,
                          cfg.get("use_fixed_fade_receding", False), metric)
    raise NotImplementedError(mode)


# --------------------------------------------------------------------------------------
# LUTs                                                              lut_manager.py:33-169
# --------------------------------------------------------------------------------------
def _curve_in_range(curve, in_min, in_max, out_min, out_max) -> np.ndarray:
    """_generate_curve_in_range (lut_manager.py:33-55): float64 ramp -> float32 table -> clip -> trunc."""
    table = np.arange(256, dtype=np.float32)
    if in_min >= in_max:
        return table.astype(np.uint8)
    ramp = np.linspace(0.0, 1.0, num=in_max - in_min + 1)
    table[in_min:in_max + 1] = curve(ramp) * (out_max - out_min) + out_min
    return np.clip(table, 0, 255).astype(np.uint8)


def generate_lut(kind: str, in_min=0, in_max=255, out_min=0, out_max=255, *, gamma_value=1.0,
                 s_curve_contrast=0.5, log_param=10.0, exp_param=2.0, sqrt_param=2.0, rodbard_param=1.0,
                 spline_points=((0, 0), (255, 255))) -> np.ndarray:
    """The eight generator families (lut_manager.py:92-169)."""
    rng = (in_min, in_max, out_min, out_max)
    if kind == "linear":
        return _curve_in_range(lambda x: x, *rng)
    if kind == "gamma":
        g = gamma_value if gamma_value > 0 else 0.01
        return _curve_in_range(lambda x: np.power(x, 1.0 / g), *rng)
    if kind == "s_curve":
        c = max(0.0, min(1.0, s_curve_contrast))

        def s_curve(x):
            if c == 0.0:
                return x
            if c == 1.0:
                return np.where(x < 0.5, 0.0, 1.0)
            gf = 1.0 / (1.0 + c * 4)
            return np.where(x < 0.5, np.power(x / 0.5, gf) * 0.5, 1.0 - np.power((1.0 - x) / 0.5, gf) * 0.5)
        return _curve_in_range(s_curve, *rng)
    if kind == "log":
        p = log_param if log_param > 0 else 0.01
        return _curve_in_range(lambda x: np.log1p(x * p) / np.log1p(p), *rng)
    if kind == "exp":
        p = exp_param if exp_param > 0 else 0.01
        return _curve_in_range(lambda x: np.power(x, p), *rng)
    if kind == "sqrt":
        r = sqrt_param if sqrt_param > 0 else 0.1
        return _curve_in_range(lambda x: np.power(x, 1.0 / r), *rng)
    if kind == "rodbard":
        k = rodbard_param

        def rod(x):
            num = x * (2.51 * x + 0.03)
            den = x * (2.43 * x + 0.59) + 0.14
            r = np.divide(num, den, out=np.zeros_like(x), where=den != 0)
            return (1 - k) * x + k * r
        return _curve_in_range(rod, *rng)
    if kind == "spline":
        from scipy.interpolate import CubicSpline
        pts = sorted([list(p) for p in spline_points], key=lambda p: p[0])
        if len(pts) < 2:
            return _curve_in_range(lambda x: x, *rng)
        xs = np.array([p[0] for p in pts]) / 255.0
        ys = np.array([p[1] for p in pts]) / 255.0
        if xs[0] > 0:
            xs, ys = np.insert(xs, 0, 0), np.insert(ys, 0, ys[0])
        if xs[-1] < 1.0:
            xs, ys = np.append(xs, 1.0), np.append(ys, ys[-1])
        return _curve_in_range(CubicSpline(xs, ys, bc_type="clamped"), *rng)
    raise ValueError(kind)


def load_lut(path: str) -> np.ndarray:
    """load_lut (lut_manager.py:75-88): JSON list of exactly 256 numbers."""
    with open(path, "r", encoding="utf-8") as fh:
        vals = json.load(fh)
    if not isinstance(vals, list) or len(vals) != 256:
        raise ValueError("Invalid LUT file format: Expected a list of 256 numbers.")
    return np.array(vals, dtype=np.uint8)


def apply_lut(img: np.ndarray, lut: np.ndarray) -> np.ndarray:
    """apply_z_lut (lut_manager.py:57-63)."""
    if img.dtype != np.uint8:
        raise TypeError("image must be uint8")
    if not isinstance(lut, np.ndarray) or lut.dtype != np.uint8 or lut.shape != (256,):
        raise ValueError("lut must be uint8[256]")
    return lut[img]


def lut_from_params(lp: dict) -> np.ndarray:
    """apply_lut_operation's table selection (xy_blend_processor.py:92-130): any failure -> identity."""
    lut = None
    try:
        if lp.get("lut_source", "generated") == "generated":
            lut = generate_lut(lp.get("lut_generation_type", "linear"), lp.get("input_min", 0),
                               lp.get("input_max", 255), lp.get("output_min", 0), lp.get("output_max", 255),
                               gamma_value=lp.get("gamma_value", 1.0),
                               s_curve_contrast=lp.get("s_curve_contrast", 0.5),
                               log_param=lp.get("log_param", 10.0), exp_param=lp.get("exp_param", 2.0),
                               sqrt_param=lp.get("sqrt_param", 2.0), rodbard_param=lp.get("rodbard_param", 1.0),
                               spline_points=lp.get("spline_points", [[0, 0], [255, 255]]))
        elif lp.get("lut_source") == "file":
            path = lp.get("fixed_lut_path", "")
            if path and os.path.exists(path):
                lut = load_lut(path)
    except Exception:
        lut = None
    return np.arange(256, dtype=np.uint8) if lut is None else lut


# --------------------------------------------------------------------------------------
# XY filters                                                   xy_blend_processor.py:27-46
# --------------------------------------------------------------------------------------
_SMALL_GAUSS = {1: [1.0], 3: [0.25, 0.5, 0.25], 5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
                7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125],
                9: [4 / 256, 13 / 256, 30 / 256, 51 / 256, 60 / 256, 51 / 256, 30 / 256, 13 / 256, 4 / 256]}


def gaussian_kernel_f64(ksize: int, sigma: float) -> np.ndarray:
    """cv2.getGaussianKernel(ksize, sigma) in float64 (fixed tables for sigma<=0, k<=9)."""
    if sigma <= 0 and ksize in _SMALL_GAUSS:
        return np.array(_SMALL_GAUSS[ksize], np.float64)
    s = sigma if sigma > 0 else 0.15 * ksize + 0.35
    scale2x = -0.125 / (s * s)
    half = (ksize - 1) // 2
    x = np.arange(1 - ksize, 1 - ksize + 2 * half, 2, dtype=np.float64)      # doubled offsets
    vals = np.exp(x * x * scale2x)
    total = 2.0 * vals.sum() + 1.0
    mul = 1.0 / total
    k = np.empty(ksize, np.float64)
    k[:half] = vals * mul
    k[ksize - half:] = (vals * mul)[::-1]
    k[half] = mul
    return k


def gaussian_kernel_q88(ksize: int, sigma: float) -> np.ndarray:
    """Q8.8 integer taps as cv2's fixed-point GaussianBlur uses for uint8: error-diffused rounding
    from the edge inwards, centre = 256 - sum(others)."""
    k = gaussian_kernel_f64(ksize, sigma)
    n, half = ksize, ksize // 2
    q = np.zeros(n, np.int64)
    err = 0.0
    for i in range(half):
        v = k[i] * 256.0 + err
        r = int(np.rint(v))
        err = v - r
        q[i] = q[n - 1 - i] = r
    q[half] = 256 - 2 * int(q[:half].sum())
    return q.astype(np.int32)


def gaussian_blur(img: np.ndarray, kx: int, ky: int, sx: float, sy: float) -> np.ndarray:
    """cv2.GaussianBlur(img,(kx,ky),sx,sy) on uint8 (xy_blend_processor.py:27-33)."""
    img = _u8(img)
    H, W = img.shape
    if sy <= 0:
        sy = sx
    if sx <= 0 and sy > 0:
        pass
    qx = gaussian_kernel_q88(kx, sx)
    qy = gaussian_kernel_q88(ky, sy)
    out = np.empty_like(img)
    lib().orc_gauss_q88(_p(img), W, H, _p(qx), len(qx), _p(qy), len(qy), _p(out))
    return out


def bilateral(img: np.ndarray, d: int, sigma_color: float, sigma_space: float, mode: str = "round") -> np.ndarray:
    """cv2.bilateralFilter(img,d,sc,ss) on uint8 (xy_blend_processor.py:35-40)."""
    img = _u8(img)
    H, W = img.shape
    out = np.empty_like(img)
    lib().orc_bilateral(_p(img), W, H, int(d), float(sigma_color), float(sigma_space),
                        1 if mode == "trunc" else 0, _p(out))
    return out


def median_blur(img: np.ndarray, k: int) -> np.ndarray:
    """apply_median_blur (xy_blend_processor.py:42-46): k<=1 returns the input object."""
    if k <= 1:
        return img
    img = _u8(img)
    H, W = img.shape
    out = np.empty_like(img)
    lib().orc_median(_p(img), W, H, int(k), _p(out))
    return out


def xy_pipeline(img: np.ndarray, ops: Iterable[dict]) -> np.ndarray:
    """process_xy_pipeline (xy_blend_processor.py:135-164).  ops are dicts with XYBlendOperation's
    field names; unknown types are skipped."""
    out = img.copy()
    for op in ops or []:
        t = str(op.get("type", "none")).lower()
        if t == "gaussian_blur":
            out = gaussian_blur(out, _odd(op.get("gaussian_ksize_x", 3)), _odd(op.get("gaussian_ksize_y", 3)),
                                op.get("gaussian_sigma_x", 0.0), op.get("gaussian_sigma_y", 0.0))
        elif t == "bilateral_filter":
            out = bilateral(out, op.get("bilateral_d", 9), op.get("bilateral_sigma_color", 75.0),
                            op.get("bilateral_sigma_space", 75.0))
        elif t == "median_blur":
            out = median_blur(out, _odd(op.get("median_ksize", 5)))
        elif t == "apply_lut":
            out = apply_lut(out, lut_from_params(op.get("lut_params", {})))
        elif t == "unsharp_mask":
            out = geom.unsharp_mask(out, op.get("unsharp_amount", 1.0), int(op.get("unsharp_threshold", 0)),
                                    _odd(op.get("unsharp_blur_ksize", 5)), op.get("unsharp_blur_sigma", 0.0), gaussian_blur)
        elif t == "resize":
            out = geom.apply_resize(out, op.get("resize_width"), op.get("resize_height"), op.get("resample_mode", "LANCZOS4"))
    return out


def _odd(k: int) -> int:
    """XYBlendOperation._ensure_odd_positive_ksize (config.py:127-129)."""
    if k <= 0:
        return 1
    return k if k % 2 else k + 1


# --------------------------------------------------------------------------------------
# the stack loop                                   processing_pipeline.py:74-113,163,188-223
# --------------------------------------------------------------------------------------
def process_layer(gray: np.ndarray, priors_newest_first: Sequence[np.ndarray], cfg: dict) -> np.ndarray:
    cur = threshold(gray)
    grad = z_blend(cur, priors_newest_first, cfg)
    return xy_pipeline(merge(gray, grad), cfg.get("xy_blend_pipeline", []))


def run_stack(layers: Sequence[np.ndarray], cfg: dict, halo: Sequence[np.ndarray] = ()) -> List[np.ndarray]:
    """Sliding deque(maxlen=receding_layers) of thresholded masks, newest-first snapshot per layer
    (processing_pipeline.py:163,216-223).  `halo` = gray layers preceding layers[0] (Z-shard halo)."""
    N = int(cfg.get("receding_layers", 4))
    window = collections.deque(maxlen=N)
    for h in halo:
        window.append(threshold(h))
    outs = []
    for g in layers:
        outs.append(process_layer(g, list(reversed(window)), cfg))
        window.append(threshold(g))
    return outs
