"""
lut_manager -- 256-entry uint8 look-up tables (reference lut_manager.py:29-169).

Table *generation* is 256 values of host arithmetic and stays in NumPy with the reference's exact
rounding sequence (float64 ramp -> float32 table -> clip -> truncate, lut_manager.py:42-55); table
*application* (`apply_z_lut`, lut_manager.py:57-63) is the device gather kernel vsb_lut_apply.
"""
from __future__ import annotations

import json
import os
from typing import Callable

import numpy as np

_IDENTITY = np.arange(256, dtype=np.uint8)


def get_default_z_lut() -> np.ndarray:
    return _IDENTITY.copy()


def _generate_curve_in_range(curve_func: Callable[[np.ndarray], np.ndarray], input_min: int, input_max: int,
                             output_min: int, output_max: int) -> np.ndarray:
    table = np.arange(256, dtype=np.float32)
    if input_min >= input_max:
        return table.astype(np.uint8)
    ramp = np.linspace(0.0, 1.0, num=input_max - input_min + 1)
    table[input_min:input_max + 1] = curve_func(ramp) * (output_max - output_min) + output_min
    return np.clip(table, 0, 255).astype(np.uint8)


def apply_z_lut(image_array: np.ndarray, lut_array: np.ndarray) -> np.ndarray:
    if image_array.dtype != np.uint8:
        raise TypeError("Input image_array for apply_z_lut must be of type np.uint8.")
    if not isinstance(lut_array, np.ndarray) or lut_array.dtype != np.uint8 or lut_array.shape != (256,):
        raise ValueError("Provided lut_array must be a 256-entry NumPy array of dtype uint8.")
    import vsb_engine as E
    shape = image_array.shape
    img2d = image_array.reshape(shape[0], -1) if image_array.ndim != 2 else image_array
    return E.lut_apply(E.DevImage.from_numpy(img2d), lut_array).numpy().reshape(shape)


def save_lut(filepath: str, lut_array: np.ndarray):
    if not isinstance(lut_array, np.ndarray) or lut_array.dtype != np.uint8 or lut_array.shape != (256,):
        raise ValueError("LUT must be a 256-entry NumPy array of dtype uint8 to save.")
    try:
        with open(filepath, "w", encoding="utf-8") as fh:
            json.dump(lut_array.tolist(), fh, indent=4)
    except Exception as exc:
        raise IOError(f"Failed to save LUT to '{filepath}': {exc}")


def load_lut(filepath: str) -> np.ndarray:
    if not os.path.exists(filepath):
        raise FileNotFoundError(f"LUT file not found: {filepath}")
    try:
        with open(filepath, "r", encoding="utf-8") as fh:
            values = json.load(fh)
        if not isinstance(values, list) or len(values) != 256:
            raise ValueError("Invalid LUT file format: Expected a list of 256 numbers.")
        return np.array(values, dtype=np.uint8)
    except json.JSONDecodeError as exc:
        raise ValueError(f"Invalid JSON format in LUT file '{filepath}': {exc}")
    except Exception as exc:
        raise IOError(f"Failed to load LUT from '{filepath}': {exc}")


# ---- curve families (reference lut_manager.py:92-169) ---------------------------------------------------------
# Every family is "shape(ramp in [0, 1]) stretched onto [output_min, output_max] over [input_min, input_max]"; a
# family is a function param -> shape.  The public generate_<family>_lut(param, in_min, in_max, out_min, out_max)
# entry points the GUI and xy_blend_processor call positionally are stamped out of this table.
def _power(exponent):
    return lambda x: np.power(x, exponent)


def _s_curve(contrast):
    contrast = max(0.0, min(1.0, contrast))
    if contrast == 0.0:
        return lambda x: x
    if contrast == 1.0:
        return lambda x: np.where(x < 0.5, 0.0, 1.0)
    g = 1.0 / (1.0 + contrast * 4)
    return lambda x: np.where(x < 0.5, np.power(x / 0.5, g) * 0.5, 1.0 - np.power((1.0 - x) / 0.5, g) * 0.5)


def _rodbard(contrast):
    def shape(x):                      # blend of the identity with the ACES-style rational curve
        num = x * (2.51 * x + 0.03)
        den = x * (2.43 * x + 0.59) + 0.14
        return (1 - contrast) * x + contrast * np.divide(num, den, out=np.zeros_like(x), where=den != 0)
    return shape


def _spline(control_points):
    if len(control_points) < 2:
        return lambda x: x
    from scipy.interpolate import CubicSpline
    pts = sorted(control_points, key=lambda p: p[0])
    xs = np.array([p[0] for p in pts]) / 255.0
    ys = np.array([p[1] for p in pts]) / 255.0
    if xs[0] > 0:                      # hold the first / last value out to the ends of the ramp
        xs, ys = np.insert(xs, 0, 0), np.insert(ys, 0, ys[0])
    if xs[-1] < 1.0:
        xs, ys = np.append(xs, 1.0), np.append(ys, ys[-1])
    return CubicSpline(xs, ys, bc_type="clamped")


def _floor(value, least):               # non-positive parameters fall back to a small positive one
    return least if value <= 0 else value


# family -> (LutParameters attribute carrying the parameter or None, param -> shape)
_FAMILIES = {
    "linear": (None, lambda _=None: (lambda x: x)),
    "gamma": ("gamma_value", lambda g: _power(1.0 / _floor(g, 0.01))),
    "s_curve": ("s_curve_contrast", _s_curve),
    "log": ("log_param", lambda k: (lambda x, k=_floor(k, 0.01): np.log1p(x * k) / np.log1p(k))),
    "exp": ("exp_param", lambda e: _power(_floor(e, 0.01))),
    "sqrt": ("sqrt_param", lambda r: _power(1.0 / _floor(r, 0.1))),
    "rodbard": ("rodbard_param", _rodbard),
    "spline": ("spline_points", _spline),
}


def _entry_point(family):
    _, shape_of = _FAMILIES[family]
    if family == "linear":
        def generate(input_min, input_max, output_min, output_max):
            return _generate_curve_in_range(shape_of(), input_min, input_max, output_min, output_max)
    else:
        def generate(param, input_min, input_max, output_min, output_max):
            return _generate_curve_in_range(shape_of(param), input_min, input_max, output_min, output_max)
    generate.__name__ = f"generate_{family}_lut"
    generate.__doc__ = f"uint8[256] table of the '{family}' family (reference lut_manager.py:92-169)."
    return generate


for _family in _FAMILIES:
    globals()[f"generate_{_family}_lut"] = _entry_point(_family)
del _family


def lut_from_params(lut_params) -> np.ndarray:
    """Table selection of apply_lut_operation (xy_blend_processor.py:92-130): generated or file table,
    identity (with the reference's printed warnings) on a missing file or any error."""
    table = None
    try:
        if lut_params.lut_source == "generated":
            family = _FAMILIES.get(lut_params.lut_generation_type)
            if family is not None:
                attr, shape_of = family
                shape = shape_of(getattr(lut_params, attr)) if attr else shape_of()
                table = _generate_curve_in_range(shape, lut_params.input_min, lut_params.input_max, lut_params.output_min,
                                                 lut_params.output_max)
        elif lut_params.lut_source == "file":
            if lut_params.fixed_lut_path and os.path.exists(lut_params.fixed_lut_path):
                table = load_lut(lut_params.fixed_lut_path)
            else:
                print(f"Warning: LUT file not found: '{lut_params.fixed_lut_path}'. Using default pass-through LUT.")
    except Exception as exc:
        print(f"Error processing LUT for operation 'apply_lut': {exc}. Using default pass-through LUT.")
    return get_default_z_lut() if table is None else table
