#!/usr/bin/env python
"""
Golden fixtures for ROI_FADE: identify_rois, ROITracker.update_and_classify and
_calculate_receding_gradient_field_roi_fade of the UNMODIFIED reference, driven the way the producer loop drives
them (processing_pipeline.py:163-164, :205-223) on crops of the 40 test slices.

    python tests/golden/make_golden_roi.py        # rewrites tests/golden/ref_roi.npz
"""
import collections
import glob
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, _enter_reference  # noqa: E402

WINS = {"b": (900, 1284, 1800, 2312), "w": (200, 1000, 300, 1900)}
CONFIGS = {
    # name: (window, fixed?, F, N, roi_params)
    "fix10":   ("b", True, 10.0, 3, {"min_size": 100, "enable_raft_support_handling": False}),
    "fix40":   ("w", True, 40.0, 4, {"min_size": 50, "enable_raft_support_handling": True, "raft_layer_count": 13,
                                    "raft_min_size": 20000, "support_max_size": 900, "support_max_layer": 30,
                                    "support_max_growth": 1.3}),
    "minmax":  ("w", False, 10.0, 4, {"min_size": 30, "enable_raft_support_handling": True, "raft_layer_count": 5,
                                     "raft_min_size": 10000, "support_max_size": 400, "support_max_layer": 25,
                                     "support_max_growth": 2.5}),
    "fix7p5":  ("b", True, 7.5, 2, {"min_size": 1, "enable_raft_support_handling": False}),
}
KEEP = (13, 14, 22, 30, 31, 36)


def main():
    _enter_reference()
    import cv2
    import numpy as np
    import processing_core as core
    from roi_tracker import ROITracker
    from config import Config, ProcessingMode, RoiParameters

    files = sorted(glob.glob(os.path.join(REF, "test-slices", "*.png")))
    out = {}

    # ---- identify_rois on whole slices: areas, boxes, centroids in the library's label order -----------------------
    for z in (5, 14, 30):
        binary, _ = core.load_image(files[z])
        for ms in (100, 10):
            rois = core.identify_rois(binary, ms)
            out[f"rois_{z:03d}_{ms}_area"] = np.array([r["area"] for r in rois], np.int64)
            out[f"rois_{z:03d}_{ms}_bbox"] = np.array([r["bbox"] for r in rois], np.int64).reshape(-1, 4)
            out[f"rois_{z:03d}_{ms}_centroid"] = np.array([r["centroid"] for r in rois], np.float64).reshape(-1, 2)
            out[f"rois_{z:03d}_{ms}_label"] = np.array([r["label"] for r in rois], np.int64)
            out[f"rois_{z:03d}_{ms}_masksum"] = np.array([int(r["mask"].sum()) for r in rois], np.int64)

    # ---- the stack loop in ROI_FADE mode ---------------------------------------------------------------------------
    for name, (wn, fixed, F, N, rp) in CONFIGS.items():
        y0, y1, x0, x1 = WINS[wn]
        cfg = Config()
        cfg.blending_mode = ProcessingMode.ROI_FADE
        cfg.use_fixed_fade_receding = fixed
        cfg.fixed_fade_distance_receding = F
        cfg.receding_layers = N
        cfg.roi_params = RoiParameters(**rp)
        tracker = ROITracker()
        window = collections.deque(maxlen=N)
        shas = []
        track = []
        for z, f in enumerate(files):
            binary, gray = core.load_image(f)
            binary = binary[y0:y1, x0:x1].copy(); gray = gray[y0:y1, x0:x1].copy()
            rois = core.identify_rois(binary, cfg.roi_params.min_size)
            classified = tracker.update_and_classify(rois, z, cfg)
            prior = core.find_prior_combined_white_mask(list(reversed(window)))
            grad = core.process_z_blending(binary, prior, cfg, classified)
            merged = core.merge_to_output(gray, grad)
            shas.append(hashlib.sha1(np.ascontiguousarray(merged).tobytes()).hexdigest())
            track.append([[int(r["id"]), str(r["classification"]), int(r["area"]), [int(v) for v in r["bbox"]]] for r in classified])
            if z in KEEP:
                out[f"{name}_grad_{z:03d}"] = grad
            window.append(binary)
        out[f"{name}_sha1"] = np.array(shas)
        out[f"{name}_track"] = np.array(json.dumps(track))
        print(name, "classes:", collections.Counter(c for layer in track for _, c, _, _ in layer))

    np.savez_compressed(os.path.join(HERE, "ref_roi.npz"), wins=json.dumps(WINS), configs=json.dumps(CONFIGS),
                        keep=np.array(KEEP), **out)
    print("ref_roi.npz", os.path.getsize(os.path.join(HERE, "ref_roi.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
