"""
processing_core -- the reference's Z-blend entry points (processing_core.py:28-460) with identical
signatures, argument meaning and "nothing to do -> zeros" behaviour, executed by the sm_100a kernels of
libvsb_b200.so.  Arrays are numpy uint8 [H, W] in and out (one H2D + one D2H per call); the stack-level
engine (vsb_engine.StackEngine, driven by processing_pipeline) keeps everything device-resident instead.

Masks are the {0, 255} images load_image() produces (README.md:25,68-70 requires non-anti-aliased input).
No CPU fallback: without the CUDA library / a device every function raises.
"""
from __future__ import annotations

import os

import cv2
import numpy as np

import vsb_engine as E
import vsb_native as N
import vsb_params as P
from config import ProcessingMode


def load_image(filepath):
    """(binary, gray) or (None, None).  PNG decode stays on the host (north_star); the threshold
    g > 127 -> 255 runs on the device (K1).  reference :28-45"""
    img = cv2.imread(filepath, cv2.IMREAD_GRAYSCALE)
    if img is None:
        print(f"Warning: Could not load image at {filepath}")
        return None, None
    return E.threshold_np(img), img


def _bits(mask: np.ndarray) -> E.DevBits:
    return E.pack(E.DevImage.from_numpy(mask))


def find_prior_combined_white_mask(prior_images_list):
    """OR of the prior masks, None for an empty list.  reference :47-67"""
    if not prior_images_list:
        return None
    return E.unpack(E.or_bits([_bits(m) for m in prior_images_list])).numpy()


def identify_rois(binary_image, min_size=100):
    """8-connected components of at least min_size pixels, in the library's label order, each with label / area /
    bbox / centroid / mask.  reference :69-105 (components + statistics on the device, masks cut on the host)."""
    return E.Components(_bits(binary_image)).rois(int(min_size), with_masks=True)


def _zeros_like(mask):
    return np.zeros_like(mask, dtype=np.uint8)


def _debug_dump(debug_info, suffix, image):
    if debug_info:
        cv2.imwrite(os.path.join(debug_info["output_folder"], f"{debug_info['base_filename']}_{suffix}.png"), image)


def _calculate_receding_gradient_field_fixed_fade(current_white_mask, prior_white_combined_mask, config, debug_info=None):
    """reference :107-153: fixed normalisation, or the global min-max variant (:141-145) when
    config.use_fixed_fade_receding is False (the dataclass default)."""
    if prior_white_combined_mask is None:
        return _zeros_like(current_white_mask)
    cur, prior = _bits(current_white_mask), _bits(prior_white_combined_mask)
    if debug_info:
        _debug_dump(debug_info, "debug_03a_receding_white_areas",
                    np.bitwise_and(prior_white_combined_mask, np.bitwise_not(current_white_mask)))
        _debug_dump(debug_info, "debug_03b_dist_src_for_transform", np.bitwise_not(current_white_mask))
    if E.count_receding(cur, [prior]) == 0:
        return _zeros_like(current_white_mask)
    z = P.zparams_from_config(_as_mode(config, ProcessingMode.FIXED_FADE), getattr(config, "distance_metric", "chamfer5"))
    if z.mode in (N.Z_MINMAX, N.Z_FIXED_FAR):
        return E.zblend_unbounded(None, cur, [prior], z).numpy()
    return E.zblend(None, cur, [prior], z).numpy()


def _calculate_receding_gradient_field_roi_fade(current_white_mask, prior_white_combined_mask, config, classified_rois,
                                                debug_info=None):
    """reference :155-220.  The ROI masks are taken as given (a label map is assembled from them on the host, later
    masks winning where masks overlap); like in the reference's pipeline they are subsets of the current mask."""
    if prior_white_combined_mask is None or not classified_rois:
        return _zeros_like(current_white_mask)
    cur, prior = _bits(current_white_mask), _bits(prior_white_combined_mask)
    if debug_info:
        _debug_dump(debug_info, "debug_03a_global_receding_areas",
                    np.bitwise_and(prior_white_combined_mask, np.bitwise_not(current_white_mask)))
    if E.count_receding(cur, [prior]) == 0:
        return _zeros_like(current_white_mask)
    skip = bool(config.roi_params.enable_raft_support_handling)
    cid = np.zeros(current_white_mask.shape, np.int32)
    eligible = np.zeros(len(classified_rois), np.uint8)
    for i, roi in enumerate(classified_rois):
        if skip and roi["classification"] in ("raft", "support"):
            continue
        eligible[i] = 1
        cid[roi["mask"] > 0] = i + 1
    t = E.torch()
    return E.roi_fade(None, cur, [prior], t.from_numpy(cid).cuda(), len(classified_rois), eligible,
                      float(config.fixed_fade_distance_receding), bool(config.use_fixed_fade_receding)).numpy()


def _calculate_weighted_receding_gradient_field(current_white_mask, prior_binary_masks, config, debug_info=None):
    """reference :222-284; prior_binary_masks newest first."""
    weights = config.manual_weights
    if not prior_binary_masks or not weights:
        return _zeros_like(current_white_mask)
    k = min(len(prior_binary_masks), len(weights))
    if k == 0 or float(sum(weights[:k])) <= 0:
        return _zeros_like(current_white_mask)
    if k > N.VSB_MAX_PRIORS:
        raise P.UnsupportedOnDevice(f"weighted stack over more than {N.VSB_MAX_PRIORS} layers")
    z = P.zparams_from_config(_as_mode(config, ProcessingMode.WEIGHTED_STACK), getattr(config, "distance_metric", "chamfer5"),
                              weighted_k=k)
    z.n_priors = k
    z.total_weight = np.float32(float(sum(weights[:k])))
    cur = _bits(current_white_mask)
    if z.mode == N.Z_WEIGHTED_FAR:
        return E.weighted_unbounded(None, cur, [_bits(m) for m in prior_binary_masks[:k]], z).numpy()
    return E.zblend(None, cur, [_bits(m) for m in prior_binary_masks[:k]], z).numpy()


def _calculate_receding_gradient_field_enhanced_edt_scipy(receding_distance_map, labels, num_labels, fade_distance_limit):
    """float32 flavour, reference :287-303 (label 0 yields 255; the dispatcher masks it afterwards)."""
    return E.enhanced_from_labels_np(receding_distance_map, labels, num_labels, fade_distance_limit, 0)


def _calculate_receding_gradient_field_enhanced_edt_numba(receding_distance_map, labels, num_labels, fade_distance_limit):
    """float64 flavour, reference :305-336."""
    return E.enhanced_from_labels_np(receding_distance_map, labels, num_labels, fade_distance_limit, 1)


def _calculate_receding_gradient_field_enhanced_edt(current_white_mask, prior_binary_masks, config, debug_info=None):
    """reference :338-404, including the anisotropic resize round trip of the distance map (:357-377)."""
    if not prior_binary_masks:
        return _zeros_like(current_white_mask)
    z = P.zparams_from_config(_as_mode(config, ProcessingMode.ENHANCED_EDT), getattr(config, "distance_metric", "chamfer5"),
                              current_white_mask.shape[1], current_white_mask.shape[0])
    cur = _bits(current_white_mask)
    priors = [_bits(m) for m in prior_binary_masks]
    if len(priors) > N.VSB_MAX_PRIORS:
        priors = [E.or_bits(priors)]
    if E.count_receding(cur, priors) == 0:
        return _zeros_like(current_white_mask)
    if config.fixed_fade_distance_receding <= 0:
        # den = min(max_label, F<=0): SciPy flavour -> den 1, d >= 1 on every receding pixel -> 0; Numba -> 0
        return _zeros_like(current_white_mask)
    return E.enhanced(None, cur, priors, z).numpy()


class _ModeView:
    """config with blending_mode overridden (the private per-mode functions ignore config.blending_mode)."""

    def __init__(self, cfg, mode):
        self._cfg, self.blending_mode = cfg, mode

    def __getattr__(self, name):
        return getattr(self._cfg, name)


def _as_mode(cfg, mode):
    return cfg if getattr(cfg, "blending_mode", None) == mode else _ModeView(cfg, mode)


def process_z_blending(current_white_mask, prior_masks, config, classified_rois, debug_info=None):
    """Mode dispatch, reference :407-440.  prior_masks: combined mask (FIXED/ROI) or newest-first list."""
    mode = config.blending_mode
    if mode == ProcessingMode.ROI_FADE:
        return _calculate_receding_gradient_field_roi_fade(current_white_mask, prior_masks, config, classified_rois, debug_info)
    if mode == ProcessingMode.WEIGHTED_STACK:
        return _calculate_weighted_receding_gradient_field(current_white_mask, prior_masks, config, debug_info)
    if mode == ProcessingMode.ENHANCED_EDT:
        return _calculate_receding_gradient_field_enhanced_edt(current_white_mask, prior_masks, config, debug_info)
    return _calculate_receding_gradient_field_fixed_fade(current_white_mask, prior_masks, config, debug_info)


def merge_to_output(original_current_image, receding_gradient):
    """np.maximum(original, gradient), reference :442-460."""
    return E.merge_np(original_current_image, receding_gradient)
