"""
lut_manager -- 256-entry uint8 look-up tables (reference lut_manager.py:29-169).

Table *generation* is 256 values of host arithmetic and stays in NumPy with the reference's exact
rounding sequence (float64 ramp -> float32 table -> clip -> truncate, lut_manager.py:42-55); table
*application* (`apply_z_lut`, lut_manager.py:57-63) is the device gather kernel vsb_lut_apply.
"""
from __future__ import annotations

import json
import os
from typing import Callable, List

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


def generate_linear_lut(input_min: int, input_max: int, output_min: int, output_max: int) -> np.ndarray:
    return _generate_curve_in_range(lambda x: x, input_min, input_max, output_min, output_max)


def _power_lut(exponent: float, rng) -> np.ndarray:
    return _generate_curve_in_range(lambda x: np.power(x, exponent), *rng)


def generate_gamma_lut(gamma_value: float, input_min: int, input_max: int, output_min: int, output_max: int) -> np.ndarray:
    if gamma_value <= 0:
        gamma_value = 0.01
    return _power_lut(1.0 / gamma_value, (input_min, input_max, output_min, output_max))


def generate_s_curve_lut(contrast: float, input_min: int, input_max: int, output_min: int, output_max: int) -> np.ndarray:
    contrast = max(0.0, min(1.0, contrast))

    def curve(x):
        if contrast == 0.0:
            return x
        if contrast == 1.0:
            return np.where(x < 0.5, 0.0, 1.0)
        g = 1.0 / (1.0 + contrast * 4)
        return np.where(x < 0.5, np.power(x / 0.5, g) * 0.5, 1.0 - np.power((1.0 - x) / 0.5, g) * 0.5)
    return _generate_curve_in_range(curve, input_min, input_max, output_min, output_max)


def generate_log_lut(param: float, input_min: int, input_max: int, output_min: int, output_max: int) -> np.ndarray:
    if param <= 0:
        param = 0.01
    return _generate_curve_in_range(lambda x: np.log1p(x * param) / np.log1p(param), input_min, input_max,
                                    output_min, output_max)


def generate_exp_lut(param: float, input_min: int, input_max: int, output_min: int, output_max: int) -> np.ndarray:
    if param <= 0:
        param = 0.01
    return _power_lut(param, (input_min, input_max, output_min, output_max))


def generate_sqrt_lut(root_value: float, input_min: int, input_max: int, output_min: int, output_max: int) -> np.ndarray:
    if root_value <= 0:
        root_value = 0.1
    return _power_lut(1.0 / root_value, (input_min, input_max, output_min, output_max))


def generate_rodbard_lut(contrast: float, input_min: int, input_max: int, output_min: int, output_max: int) -> np.ndarray:
    def curve(x):
        num = x * (2.51 * x + 0.03)
        den = x * (2.43 * x + 0.59) + 0.14
        aces = np.divide(num, den, out=np.zeros_like(x), where=den != 0)
        return (1 - contrast) * x + contrast * aces
    return _generate_curve_in_range(curve, input_min, input_max, output_min, output_max)


def generate_spline_lut(control_points: List[List[int]], input_min: int, input_max: int, output_min: int,
                        output_max: int) -> np.ndarray:
    if len(control_points) < 2:
        return generate_linear_lut(input_min, input_max, output_min, output_max)
    from scipy.interpolate import CubicSpline
    pts = sorted(control_points, key=lambda p: p[0])
    xs = np.array([p[0] for p in pts]) / 255.0
    ys = np.array([p[1] for p in pts]) / 255.0
    if xs[0] > 0:
        xs, ys = np.insert(xs, 0, 0), np.insert(ys, 0, ys[0])
    if xs[-1] < 1.0:
        xs, ys = np.append(xs, 1.0), np.append(ys, ys[-1])
    return _generate_curve_in_range(CubicSpline(xs, ys, bc_type="clamped"), input_min, input_max, output_min,
                                    output_max)


_GENERATORS = {
    "linear": lambda p, r: generate_linear_lut(*r),
    "gamma": lambda p, r: generate_gamma_lut(p.gamma_value, *r),
    "s_curve": lambda p, r: generate_s_curve_lut(p.s_curve_contrast, *r),
    "log": lambda p, r: generate_log_lut(p.log_param, *r),
    "exp": lambda p, r: generate_exp_lut(p.exp_param, *r),
    "sqrt": lambda p, r: generate_sqrt_lut(p.sqrt_param, *r),
    "rodbard": lambda p, r: generate_rodbard_lut(p.rodbard_param, *r),
    "spline": lambda p, r: generate_spline_lut(p.spline_points, *r),
}


def lut_from_params(lut_params) -> np.ndarray:
    """Table selection of apply_lut_operation (xy_blend_processor.py:92-130): generated or file table,
    identity (with the reference's printed warnings) on a missing file or any error."""
    table = None
    try:
        if lut_params.lut_source == "generated":
            gen = _GENERATORS.get(lut_params.lut_generation_type)
            if gen is not None:
                table = gen(lut_params, (lut_params.input_min, lut_params.input_max, lut_params.output_min,
                                         lut_params.output_max))
        elif lut_params.lut_source == "file":
            if lut_params.fixed_lut_path and os.path.exists(lut_params.fixed_lut_path):
                table = load_lut(lut_params.fixed_lut_path)
            else:
                print(f"Warning: LUT file not found: '{lut_params.fixed_lut_path}'. Using default pass-through LUT.")
    except Exception as exc:
        print(f"Error processing LUT for operation 'apply_lut': {exc}. Using default pass-through LUT.")
    return get_default_z_lut() if table is None else table
