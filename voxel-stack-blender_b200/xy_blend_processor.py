"""
xy_blend_processor -- the reference's per-layer 2-D operation chain (xy_blend_processor.py:27-164) on the
device: same function names, `XYBlendOperation` semantics and printed warnings.  Each apply_* takes and
returns numpy uint8 [H, W]; process_xy_pipeline keeps the intermediate images on the device and composes
consecutive apply_lut operations into one table (final = cur[final], pyside_xy_blend_tab.py:314-321).
"""
from __future__ import annotations

from typing import List

import numpy as np

import lut_manager
import vsb_engine as E
import vsb_params as P
from config import XYBlendOperation


def _gaussian_dev(img: E.DevImage, op) -> E.DevImage:
    kx, ky = P.odd_ksize(op.gaussian_ksize_x), P.odd_ksize(op.gaussian_ksize_y)
    sx, sy = float(op.gaussian_sigma_x), float(op.gaussian_sigma_y)
    if sy <= 0:
        sy = sx
    return E.gaussian(img, P.gaussian_taps_q88(kx, sx), P.gaussian_taps_q88(ky, sy))


def apply_gaussian_blur(image: np.ndarray, op: XYBlendOperation) -> np.ndarray:
    """cv2.GaussianBlur on uint8, reference :27-33"""
    return _gaussian_dev(E.DevImage.from_numpy(image), op).numpy()


def apply_bilateral_filter(image: np.ndarray, op: XYBlendOperation) -> np.ndarray:
    """cv2.bilateralFilter on uint8, reference :35-40"""
    return E.bilateral(E.DevImage.from_numpy(image), int(op.bilateral_d), float(op.bilateral_sigma_color),
                       float(op.bilateral_sigma_space)).numpy()


def apply_median_blur(image: np.ndarray, op: XYBlendOperation) -> np.ndarray:
    """cv2.medianBlur on uint8; ksize <= 1 returns the input object itself, reference :42-46"""
    if op.median_ksize <= 1:
        return image
    return E.median(E.DevImage.from_numpy(image), P.odd_ksize(int(op.median_ksize))).numpy()


def _unsharp_dev(img: E.DevImage, op) -> E.DevImage:
    return E.unsharp(img, float(op.unsharp_amount), int(op.unsharp_threshold), int(op.unsharp_blur_ksize),
                     float(op.unsharp_blur_sigma))


def apply_unsharp_mask(image: np.ndarray, op: XYBlendOperation) -> np.ndarray:
    """GaussianBlur + addWeighted + absdiff/threshold select, reference :48-67"""
    return _unsharp_dev(E.DevImage.from_numpy(image), op).numpy()


def _resize_dev(img: E.DevImage, op):
    """None when the op leaves the image alone (reference :73, :82)."""
    dims = P.resize_dims(img.W, img.H, op.resize_width, op.resize_height)
    if dims is None:
        return None
    if dims[0] < 1 or dims[1] < 1:
        raise ValueError(f"resize to {dims[0]}x{dims[1]}")
    return E.resize(img, dims[0], dims[1], E.N.INTERP_BY_NAME.get(str(op.resample_mode).upper(), E.N.INTER_LANCZOS4))


def apply_resize(image: np.ndarray, op: XYBlendOperation) -> np.ndarray:
    """cv2.resize with the reference's size rule and mode table; returns the input object itself when there is
    nothing to do, reference :69-90"""
    if P.resize_dims(image.shape[1], image.shape[0], op.resize_width, op.resize_height) is None:
        return image
    return _resize_dev(E.DevImage.from_numpy(image), op).numpy()


def _lut_for(op) -> np.ndarray:
    return lut_manager.lut_from_params(op.lut_params)


def apply_lut_operation(image: np.ndarray, op: XYBlendOperation) -> np.ndarray:
    """table from op.lut_params (identity on any failure), gather on the device, reference :92-132"""
    return lut_manager.apply_z_lut(image, _lut_for(op))


def process_xy_pipeline_dev(dev: E.DevImage, pipeline_ops: List[XYBlendOperation]):
    """The chain on a device image: (result, touched).  Consecutive apply_lut ops are composed into one table."""
    pending_lut = None

    def flush(d):
        nonlocal pending_lut
        if pending_lut is not None:
            d = E.lut_apply(d, pending_lut)
            pending_lut = None
        return d

    touched = False
    for op in pipeline_ops or []:
        t = op.type
        if t == "apply_lut":
            table = _lut_for(op)
            pending_lut = table if pending_lut is None else table[pending_lut]
            touched = True
        elif t == "gaussian_blur":
            dev = _gaussian_dev(flush(dev), op); touched = True
        elif t == "bilateral_filter":
            dev = E.bilateral(flush(dev), int(op.bilateral_d), float(op.bilateral_sigma_color),
                              float(op.bilateral_sigma_space)); touched = True
        elif t == "median_blur":
            if op.median_ksize > 1:
                dev = E.median(flush(dev), P.odd_ksize(int(op.median_ksize))); touched = True
        elif t == "unsharp_mask":
            dev = _unsharp_dev(flush(dev), op); touched = True
        elif t == "resize":
            if P.resize_dims(dev.W, dev.H, op.resize_width, op.resize_height) is not None:
                dev = _resize_dev(flush(dev), op); touched = True
        elif t != "none":
            print(f"Warning: Unknown XY blend operation type '{t}'. Skipping.")
    return flush(dev), touched


def process_xy_pipeline(image: np.ndarray, pipeline_ops: List[XYBlendOperation]) -> np.ndarray:
    """reference :135-164: copy, dispatch by op.type, unknown types warn and are skipped."""
    if not pipeline_ops:
        return image.copy()
    dev, touched = process_xy_pipeline_dev(E.DevImage.from_numpy(image), pipeline_ops)
    if not touched:
        return image.copy()
    return dev.numpy()
