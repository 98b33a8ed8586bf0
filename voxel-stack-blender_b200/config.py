"""
config -- the reference's configuration surface (config.py:30-305): same dataclasses, field names,
defaults, clamping rules and JSON layout, so presets/*.json, saved app_config.json files, main.py and
ui_components.py keep working.  Additive keys for the B200 engine (`distance_metric`, `devices`) default
to the reference's behaviour and are ignored by the reference itself (unknown keys are skipped there).
"""
from __future__ import annotations

import dataclasses
import enum
import json
import os
from dataclasses import field
from typing import List, Optional

DEFAULT_NUM_WORKERS = max(1, (os.cpu_count() or 2) - 1)


class ProcessingMode(enum.Enum):
    ENHANCED_EDT = "enhanced_edt"
    FIXED_FADE = "fixed_fade"
    ROI_FADE = "roi_fade"
    WEIGHTED_STACK = "weighted_stack"


class WeightingFalloff(enum.Enum):
    FLAT = "Flat"
    LINEAR = "Linear"
    EXPONENTIAL = "Exponential"
    LOGARITHMIC = "Logarithmic"
    GAUSSIAN = "Gaussian"


def _clamp(v, lo, hi):
    return max(lo, min(hi, v))


_LUT_KINDS = ("linear", "gamma", "s_curve", "log", "exp", "sqrt", "rodbard", "spline")
# field -> (low, high) applied in LutParameters.__post_init__ (config.py:73-84)
_LUT_LIMITS = {"gamma_value": (0.01, 10.0), "s_curve_contrast": (0.0, 1.0), "log_param": (0.01, 100.0),
               "exp_param": (0.01, 10.0), "sqrt_param": (0.1, 50.0), "rodbard_param": (0.0, 2.0)}


def _record(name, spec, methods=None):
    """A dataclass from a (field, type, default) table; callables are default factories."""
    cols = [(n, t, field(default_factory=d) if callable(d) else field(default=d)) for n, t, d in spec]
    return dataclasses.make_dataclass(name, cols, namespace=dict(methods or {}), module=__name__)


def _only_fields(cls, data: dict) -> dict:
    names = {f.name for f in dataclasses.fields(cls)}
    return {k: v for k, v in data.items() if k in names}


# ---- LutParameters (reference config.py:44-84) --------------------------------------------------------------
def _lut_post_init(self):
    self.lut_source = self.lut_source.lower()
    if self.lut_source not in ("generated", "file"):
        self.lut_source = "generated"
    if self.lut_generation_type not in _LUT_KINDS:
        self.lut_generation_type = "linear"
    for lo_name, hi_name in (("input_min", "input_max"), ("output_min", "output_max")):
        lo, hi = _clamp(getattr(self, lo_name), 0, 255), _clamp(getattr(self, hi_name), 0, 255)
        if lo > hi:
            lo, hi = hi, lo
        setattr(self, lo_name, lo)
        setattr(self, hi_name, hi)
    for name, (lo, hi) in _LUT_LIMITS.items():
        setattr(self, name, _clamp(getattr(self, name), lo, hi))


LutParameters = _record("LutParameters", [
    ("lut_source", str, "generated"), ("lut_generation_type", str, "linear"),
    ("input_min", int, 0), ("input_max", int, 255), ("output_min", int, 0), ("output_max", int, 255),
    ("gamma_value", float, 1.0), ("s_curve_contrast", float, 0.5), ("log_param", float, 10.0), ("exp_param", float, 2.0),
    ("sqrt_param", float, 2.0), ("rodbard_param", float, 1.0),
    ("spline_points", List[List[int]], lambda: [[0, 0], [255, 255]]), ("fixed_lut_path", str, ""),
], {"__post_init__": _lut_post_init})


# ---- XYBlendOperation (reference config.py:86-135) ----------------------------------------------------------
def _odd_positive(self, ksize: int) -> int:
    if ksize <= 0:
        return 1
    return ksize if ksize % 2 else ksize + 1


def _lut_params_from_dict(data: dict):
    return LutParameters(**_only_fields(LutParameters, data))


def _xy_post_init(self):
    self.type = self.type.lower()
    odd = self._ensure_odd_positive_ksize
    kernel_fields = {"gaussian_blur": ("gaussian_ksize_x", "gaussian_ksize_y"), "median_blur": ("median_ksize",),
                     "unsharp_mask": ("unsharp_blur_ksize",)}
    for name in kernel_fields.get(self.type, ()):
        setattr(self, name, odd(getattr(self, name)))
    if self.type == "resize":            # negative sizes clamp to 0, and 0 means "not given"
        for name in ("resize_width", "resize_height"):
            v = getattr(self, name)
            if v is not None:
                setattr(self, name, max(0, v) or None)
    if isinstance(self.lut_params, dict):
        self.lut_params = self.from_dict_to_lut_params(self.lut_params)


XYBlendOperation = _record("XYBlendOperation", [
    ("type", str, "none"),
    ("gaussian_ksize_x", int, 3), ("gaussian_ksize_y", int, 3), ("gaussian_sigma_x", float, 0.0), ("gaussian_sigma_y", float, 0.0),
    ("bilateral_d", int, 9), ("bilateral_sigma_color", float, 75.0), ("bilateral_sigma_space", float, 75.0),
    ("median_ksize", int, 5),
    ("unsharp_amount", float, 1.0), ("unsharp_threshold", int, 0), ("unsharp_blur_ksize", int, 5), ("unsharp_blur_sigma", float, 0.0),
    ("resize_width", Optional[int], None), ("resize_height", Optional[int], None), ("resample_mode", str, "LANCZOS4"),
    ("lut_params", LutParameters, LutParameters),
], {"__post_init__": _xy_post_init, "_ensure_odd_positive_ksize": _odd_positive,
    "from_dict_to_lut_params": staticmethod(_lut_params_from_dict)})

RoiParameters = _record("RoiParameters", [
    ("min_size", int, 100), ("enable_raft_support_handling", bool, False), ("raft_layer_count", int, 5),
    ("raft_min_size", int, 10000), ("support_max_size", int, 500), ("support_max_layer", int, 1000),
    ("support_max_growth", float, 2.5)])

AnisotropicParams = _record("AnisotropicParams", [("enabled", bool, False), ("x_factor", float, 1.0), ("y_factor", float, 1.0)])

_CONFIG_SCHEMA = [
    # I/O
    ("output_file_prefix", str, "Voxel_Blend_Processed_"), ("input_mode", str, "folder"), ("input_folder", str, ""),
    ("output_folder", str, ""), ("start_index", Optional[int], 0), ("stop_index", Optional[int], None),
    # UVTools mode
    ("uvtools_path", str, "C:\\Program Files\\UVTools\\UVToolsCmd.exe"), ("uvtools_temp_folder", str, ""),
    ("uvtools_input_file", str, ""), ("uvtools_output_location", str, "working_folder"),
    ("uvtools_delete_temp_on_completion", bool, True),
    # stack blending
    ("blending_mode", ProcessingMode, ProcessingMode.FIXED_FADE), ("receding_layers", int, 4),
    ("use_fixed_fade_receding", bool, False), ("fixed_fade_distance_receding", float, 10.0),
    ("anisotropic_params", AnisotropicParams, AnisotropicParams),
    # weighted stack
    ("weighted_falloff_type", WeightingFalloff, WeightingFalloff.LINEAR),
    ("manual_weights", List[int], lambda: [100, 75, 50, 25]),
    ("fade_distances_receding", List[float], lambda: [10.0, 10.0, 10.0, 10.0]),
    # overhang (unused by the reference as well)
    ("overhang_layers", int, 0), ("use_fixed_fade_overhang", bool, False), ("fixed_fade_distance_overhang", float, 10.0),
    # ROI mode
    ("roi_params", RoiParameters, RoiParameters),
    # general
    ("thread_count", int, DEFAULT_NUM_WORKERS), ("use_numba_jit", bool, False), ("debug_save", bool, False),
    ("xy_blend_pipeline", List[XYBlendOperation], lambda: [XYBlendOperation("none")]),
    # additive B200 engine keys (defaults = reference behaviour)
    ("distance_metric", str, "chamfer5"),      # "chamfer5" (cv2 DIST_L2 mask 5, bit parity) or "exact"
    ("devices", Optional[List[int]], None),    # CUDA devices for Z-sharding; None = current device
]


class _ConfigMethods:
    def to_dict(self) -> dict:
        d = dataclasses.asdict(self)
        for k, v in list(d.items()):
            if isinstance(v, enum.Enum):
                d[k] = v.value
        return d

    @classmethod
    def from_dict(cls, data: dict) -> "Config":
        cfg = cls()
        known = {f.name: f for f in dataclasses.fields(cls)}
        enums = {"blending_mode": (ProcessingMode, ProcessingMode.FIXED_FADE, "FIXED_FADE"),
                 "weighted_falloff_type": (WeightingFalloff, WeightingFalloff.LINEAR, "LINEAR")}
        nested = {"roi_params": RoiParameters, "anisotropic_params": AnisotropicParams}
        for key, value in data.items():
            if key not in known:
                print(f"Warning: Unrecognized config key '{key}' found. Skipping.")
                continue
            if key in enums:
                etype, default, dname = enums[key]
                try:
                    setattr(cfg, key, etype(value))
                except ValueError:
                    print(f"Warning: Invalid {key} '{value}'. Defaulting to {dname}.")
                    setattr(cfg, key, default)
            elif key == "xy_blend_pipeline":
                ops = []
                if isinstance(value, list):
                    for item in value:
                        if not isinstance(item, dict):
                            continue
                        kw = _only_fields(XYBlendOperation, item)
                        if isinstance(kw.get("lut_params"), dict):
                            kw["lut_params"] = XYBlendOperation.from_dict_to_lut_params(kw["lut_params"])
                        ops.append(XYBlendOperation(**kw))
                cfg.xy_blend_pipeline = ops
            elif key in nested:
                if isinstance(value, dict):
                    setattr(cfg, key, nested[key](**_only_fields(nested[key], value)))
            else:
                if known[key].type in (bool, "bool") and isinstance(value, str):
                    value = value.lower() in ("true", "1", "t", "y")
                try:
                    setattr(cfg, key, value)
                except (TypeError, ValueError):
                    print(f"Warning: Could not assign value '{value}' to '{key}'. Using default.")
        return cfg

    def save(self, filepath: str):
        with open(filepath, "w", encoding="utf-8") as fh:
            json.dump(self.to_dict(), fh, indent=4)

    @classmethod
    def load(cls, filepath: str) -> "Config":
        if not os.path.exists(filepath):
            cfg = cls()
            if not os.environ.get("VSB_NO_APP_CONFIG_WRITE"):
                try:
                    cfg.save(filepath)
                except Exception as exc:  # same tolerance as the reference
                    print(f"Error saving default config: {exc}")
            return cfg
        try:
            with open(filepath, "r", encoding="utf-8") as fh:
                return cls.from_dict(json.load(fh))
        except Exception as exc:
            print(f"Error loading config '{filepath}': {exc}. Using default.")
            return cls()


Config = _record("Config", _CONFIG_SCHEMA, {k: v for k, v in vars(_ConfigMethods).items() if not k.startswith("__")})


def upgrade_config(cfg):
    """legacy attribute names (config.py:292-301)"""
    for old, new in (("n_layers", "receding_layers"), ("use_fixed_norm", "use_fixed_fade_receding"),
                     ("fixed_fade_distance", "fixed_fade_distance_receding")):
        if hasattr(cfg, old):
            setattr(cfg, new, getattr(cfg, old))
            delattr(cfg, old)


_CONFIG_FILE = "app_config.json"
app_config = Config.load(_CONFIG_FILE)
upgrade_config(app_config)
