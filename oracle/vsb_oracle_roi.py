"""
vsb_oracle_roi -- TEST INFRASTRUCTURE ONLY (CPU oracle, SURVEY.md section 8(f) rank 3: ROI_FADE).

Restates, in NumPy + the oracle's C helpers,
    identify_rois                                     processing_core.py:69-105
    ROITracker.update_and_classify / calculate_iou     roi_tracker.py:3-109
    _calculate_receding_gradient_field_roi_fade        processing_core.py:155-220
and the way the producer loop feeds them (processing_pipeline.py:163-164, :205-223).  Pinned by
tests/golden/ref_roi.npz (outputs of the unmodified reference; tests/golden/make_golden_roi.py).

Component order: cv2.connectedComponentsWithStats numbers 8-connected components by the first 2x2 block (block
raster order) that touches them -- measured, see the golden test -- which matters because the tracker breaks score
ties by list position.  Same import rule as vsb_oracle: tests/, smoke() and bench.py's CPU legs only.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

import vsb_oracle as O

f32 = np.float32


# --------------------------------------------------------------------------------------
# identify_rois                                                processing_core.py:69-105
# --------------------------------------------------------------------------------------
def components(binary: np.ndarray):
    """(labels int32 [H,W] numbered 1.. in the library's order, stats dict of arrays indexed by label-1)."""
    n, lab = O.ccl8(binary)
    H, W = binary.shape
    ys, xs = np.nonzero(lab)
    l = lab[ys, xs]
    k = n - 1
    first_blk = np.full(k + 1, 1 << 62, np.int64)
    np.minimum.at(first_blk, l, (ys >> 1) * ((W + 1) >> 1) + (xs >> 1))
    order = np.argsort(first_blk[1:], kind="stable") + 1          # old label of new label i+1
    remap = np.zeros(k + 1, np.int64)
    remap[order] = np.arange(1, k + 1)
    lab = remap[lab].astype(np.int32)
    l = lab[ys, xs]
    area = np.bincount(l, minlength=k + 1)[1:]
    x0 = np.full(k + 1, 1 << 30, np.int64); y0 = x0.copy(); x1 = np.full(k + 1, -1, np.int64); y1 = x1.copy()
    np.minimum.at(x0, l, xs); np.minimum.at(y0, l, ys); np.maximum.at(x1, l, xs); np.maximum.at(y1, l, ys)
    sx = np.bincount(l, weights=xs.astype(np.float64), minlength=k + 1)[1:]
    sy = np.bincount(l, weights=ys.astype(np.float64), minlength=k + 1)[1:]
    stats = {"area": area, "x": x0[1:], "y": y0[1:], "w": x1[1:] - x0[1:] + 1, "h": y1[1:] - y0[1:] + 1,
             "cx": sx / np.maximum(area, 1), "cy": sy / np.maximum(area, 1)}
    return lab, stats


def identify_rois(binary: np.ndarray, min_size: int = 100, with_masks: bool = True) -> List[dict]:
    lab, st = components(binary)
    rois = []
    for i in range(len(st["area"])):
        if st["area"][i] >= min_size:
            r = {"label": i + 1, "area": int(st["area"][i]),
                 "bbox": (int(st["x"][i]), int(st["y"][i]), int(st["w"][i]), int(st["h"][i])),
                 "centroid": np.array([st["cx"][i], st["cy"][i]])}
            if with_masks:
                r["mask"] = (lab == i + 1).astype(np.uint8) * 255
            rois.append(r)
    return rois


# --------------------------------------------------------------------------------------
# tracker                                                            roi_tracker.py:3-109
# --------------------------------------------------------------------------------------
def overlap_over_min(a, b) -> float:
    """calculate_iou: intersection over the smaller box area (boxes are x, y, w, h)."""
    ix = max(0, min(a[0] + a[2], b[0] + b[2]) - max(a[0], b[0]))
    iy = max(0, min(a[1] + a[3], b[1] + b[3]) - max(a[1], b[1]))
    m = min(a[2] * a[3], b[2] * b[3])
    return 0.0 if m == 0 else (ix * iy) / float(m)


class Tracker:
    def __init__(self):
        self.next_id = 0
        self.prev: Dict[int, dict] = {}

    @staticmethod
    def _new_class(roi, layer_index, rp) -> str:
        if not rp["enable_raft_support_handling"]:
            return "model"
        if layer_index < rp["raft_layer_count"] and roi["area"] >= rp["raft_min_size"]:
            return "raft"
        if roi["area"] <= rp["support_max_size"] and layer_index < rp["support_max_layer"]:
            return "support"
        return "model"

    def update(self, rois: List[dict], layer_index: int, rp: dict) -> List[dict]:
        def fresh(r):
            r["id"] = self.next_id
            self.next_id += 1
            r["classification"] = self._new_class(r, layer_index, rp)
            return r

        if not self.prev:
            self.prev = {}
            for r in rois:
                fresh(r)
                self.prev[r["id"]] = r
            return list(self.prev.values())
        ids = list(self.prev.keys())
        score = np.zeros((len(rois), len(ids)))
        for i, r in enumerate(rois):
            for j, pid in enumerate(ids):
                score[i, j] = overlap_over_min(r["bbox"], self.prev[pid]["bbox"])
        matched: Dict[int, int] = {}
        used = set()
        for flat in np.argsort(-score, axis=None):      # descending score, ties in flattened order
            i, j = np.unravel_index(flat, score.shape)
            if score[i, j] < 0.5:
                break
            if i not in matched and ids[j] not in used:
                matched[int(i)] = ids[j]
                used.add(ids[j])
        nxt: Dict[int, dict] = {}
        for i, pid in matched.items():
            r, p = rois[i], self.prev[pid]
            r["id"] = pid
            if p["classification"] == "support":
                growth = r["area"] / p["area"] if p["area"] > 0 else float("inf")
                r["classification"] = "model" if growth > rp["support_max_growth"] else "support"
            else:
                r["classification"] = p["classification"]
            nxt[pid] = r
        for i, r in enumerate(rois):
            if i not in matched:
                fresh(r)
                nxt[r["id"]] = r
        self.prev = nxt
        return list(nxt.values())


# --------------------------------------------------------------------------------------
# gradient                                                    processing_core.py:155-220
# --------------------------------------------------------------------------------------
def _dilate_box(mask: np.ndarray, k: int) -> np.ndarray:
    """cv2.dilate with a k x k box of ones (anchor at the centre, border ignored)."""
    r = k // 2
    H, W = mask.shape
    a = mask.copy()
    for d in range(1, min(r, H - 1) + 1):
        a[d:] = np.maximum(a[d:], mask[:-d])
        a[:-d] = np.maximum(a[:-d], mask[d:])
    b = a.copy()
    for d in range(1, min(r, W - 1) + 1):
        b[:, d:] = np.maximum(b[:, d:], a[:, :-d])
        b[:, :-d] = np.maximum(b[:, :-d], a[:, d:])
    return b


def roi_fade(cur: np.ndarray, prior_combined: Optional[np.ndarray], F: float, use_fixed: bool,
             rois: Sequence[dict], skip_raft_support: bool) -> np.ndarray:
    if prior_combined is None or not rois:
        return np.zeros_like(cur)
    rec = prior_combined & ~cur
    if not rec.any():
        return np.zeros_like(cur)
    final = np.zeros_like(cur)
    k = min(51, int(F) * 2 + 1)
    for roi in rois:
        if skip_raft_support and roi["classification"] in ("raft", "support"):
            continue
        m = roi["mask"]
        zone = rec & _dilate_box(m, k) if k > 1 else rec & m
        if not zone.any():
            continue
        # distance to the nearest pixel of this ROI: reach of the dilation suffices for the min-max variant, reach
        # of F for the fixed one; farther pixels are clipped / outside the zone
        reach = max(F, 1.4 * (k // 2) + 1.0)
        d = O.distance(m, reach)
        d = np.where(np.isfinite(d), d, f32(np.finfo(np.float32).max))
        rd = np.where(zone > 0, d, f32(0)).astype(np.float32)
        if rd.max() == 0:
            continue
        if use_fixed:
            Ff = f32(F)
            den = Ff if F > 0 else f32(1.0)
            g = O._fade_u8(np.minimum(np.maximum(rd, f32(0)), Ff), den)
        else:
            vals = rd[zone > 0]
            mn, mx = float(vals.min()), float(vals.max())
            if mx <= mn:
                continue
            n = (rd - f32(mn)) / f32(mx - mn)
            with np.errstate(invalid="ignore"):
                g = (f32(255.0) * (f32(1.0) - n)).astype(np.uint8)
        final = np.maximum(final, np.where(zone > 0, g, np.uint8(0)).astype(np.uint8))
    return final


DEFAULT_ROI_PARAMS = {"min_size": 100, "enable_raft_support_handling": False, "raft_layer_count": 5,
                      "raft_min_size": 10000, "support_max_size": 500, "support_max_layer": 1000,
                      "support_max_growth": 2.5}


def run_stack_roi(layers: Sequence[np.ndarray], cfg: dict, layer_indices: Optional[Sequence[int]] = None):
    """The producer loop in ROI_FADE mode: (outputs, per-layer tracker state [(id, class, area, bbox)])."""
    import collections
    rp = dict(DEFAULT_ROI_PARAMS); rp.update(cfg.get("roi_params") or {})
    N = int(cfg.get("receding_layers", 4))
    F = float(cfg.get("fixed_fade_distance_receding", 10.0))
    window = collections.deque(maxlen=N)
    tr = Tracker()
    outs, states = [], []
    for z, g in enumerate(layers):
        cur = O.threshold(g)
        rois = identify_rois(cur, rp["min_size"])
        cl = tr.update(rois, layer_indices[z] if layer_indices is not None else z, rp)
        grad = roi_fade(cur, O.combine_priors(list(reversed(window))), F, bool(cfg.get("use_fixed_fade_receding", False)),
                        cl, rp["enable_raft_support_handling"])
        outs.append(O.xy_pipeline(O.merge(g, grad), cfg.get("xy_blend_pipeline", [])))
        states.append([[int(r["id"]), r["classification"], int(r["area"]), [int(v) for v in r["bbox"]]] for r in cl])
        window.append(cur)
    return outs, states
