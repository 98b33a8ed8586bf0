"""
roi_tracker -- frame-to-frame identity and raft / support / model classification of the ROIs of consecutive
layers (reference roi_tracker.py:3-109, same public names).  Pure host logic on bounding boxes and areas: the
per-layer component statistics come from the device (vsb_roi_components), the decision which ROIs take part in
the fade goes back as a one-byte-per-component table (vsb_roi_fade).
"""
from __future__ import annotations

import numpy as np


def calculate_iou(boxA, boxB):
    """Overlap of two (x, y, w, h) boxes as intersection over the SMALLER box area (so containment scores 1),
    0.0 when either box is empty.  reference :3-23"""
    left, top = max(boxA[0], boxB[0]), max(boxA[1], boxB[1])
    right, bottom = min(boxA[0] + boxA[2], boxB[0] + boxB[2]), min(boxA[1] + boxA[3], boxB[1] + boxB[3])
    inter = max(0, right - left) * max(0, bottom - top)
    smaller = min(boxA[2] * boxA[3], boxB[2] * boxB[3])
    return inter / float(smaller) if smaller != 0 else 0.0


def _score_pairs(cur_boxes, prev_boxes):
    """The pairs (current index, previous index) whose boxes overlap in x, and calculate_iou of each: the same integer
    intersections and the same float64 division as the scalar function.  A stack with thousands of supports makes
    millions of pairs per layer, nearly all of them disjoint (score 0); every pair that is NOT returned scores 0.
    Previous boxes sorted by left edge, one binary search per current box for the run of candidates."""
    n, m = len(cur_boxes), len(prev_boxes)
    none = np.zeros(0, np.int64)
    if n == 0 or m == 0:
        return none, none, np.zeros(0)
    c = np.asarray(cur_boxes, np.int64).reshape(n, 4)
    q = np.asarray(prev_boxes, np.int64).reshape(m, 4)
    order = np.argsort(q[:, 0], kind="stable")
    q_left = q[order, 0]
    # a previous box can only reach over c's left edge if it starts within max(width) of it, and has to start
    # before c's right edge; everything outside that run has a non-positive x overlap, i.e. score 0
    lo = np.searchsorted(q_left, c[:, 0] - max(int(q[:, 2].max()), 0), side="right")
    hi = np.searchsorted(q_left, c[:, 0] + c[:, 2], side="left")
    cnt = np.maximum(hi - lo, 0)
    total = int(cnt.sum())
    if total == 0:
        return none, none, np.zeros(0)
    ci = np.repeat(np.arange(n), cnt)
    pj = order[np.repeat(lo, cnt) + (np.arange(total) - np.repeat(np.cumsum(cnt) - cnt, cnt))]
    a, b = c[ci], q[pj]
    iw = np.minimum(a[:, 0] + a[:, 2], b[:, 0] + b[:, 2]) - np.maximum(a[:, 0], b[:, 0])
    ih = np.minimum(a[:, 1] + a[:, 3], b[:, 1] + b[:, 3]) - np.maximum(a[:, 1], b[:, 1])
    inter = np.maximum(iw, 0) * np.maximum(ih, 0)
    smaller = np.minimum(a[:, 2] * a[:, 3], b[:, 2] * b[:, 3])
    val = np.zeros(total)
    np.divide(inter, smaller.astype(np.float64), out=val, where=smaller != 0)
    return ci, pj, val


def _score_matrix(cur_boxes, prev_boxes):
    """calculate_iou for every (current, previous) pair as the reference's dense matrix, element for element."""
    out = np.zeros((len(cur_boxes), len(prev_boxes)))
    ci, pj, val = _score_pairs(cur_boxes, prev_boxes)
    out[ci, pj] = val
    return out


class ROITracker:
    """reference :25-109.  State: next_roi_id and previous_rois {id: roi dict}."""

    MATCH_THRESHOLD = 0.5

    def __init__(self):
        self.next_roi_id = 0
        self.previous_rois = {}
        self._neg = np.full(0, -0.0)   # scratch of update_and_classify

    def _classify_new_roi(self, roi, layer_index, config):
        rp = config.roi_params
        if not rp.enable_raft_support_handling:
            return "model"
        if layer_index < rp.raft_layer_count and roi["area"] >= rp.raft_min_size:
            return "raft"
        if roi["area"] <= rp.support_max_size and layer_index < rp.support_max_layer:
            return "support"
        return "model"

    def _admit(self, roi, layer_index, config, into):
        roi["id"] = self.next_roi_id
        self.next_roi_id += 1
        roi["classification"] = self._classify_new_roi(roi, layer_index, config)
        into[roi["id"]] = roi

    def update_and_classify(self, current_rois_raw, layer_index, config):
        tracked = {}
        if not self.previous_rois:                       # nothing to match against: everything is new
            for roi in current_rois_raw:
                self._admit(roi, layer_index, config, tracked)
            self.previous_rois = tracked
            return list(tracked.values())

        prev_ids = list(self.previous_rois)
        n, width = len(current_rois_raw), len(prev_ids)
        pair_c, pair_p, score = _score_pairs([r["bbox"] for r in current_rois_raw],
                                             [self.previous_rois[i]["bbox"] for i in prev_ids])

        # greedy one-to-one assignment, best score first: the reference's own call -- argsort of the negated dense matrix,
        # so that tied scores (identical boxes score 1.0 by the hundred) come out in the order its sort leaves them.  The
        # matrix (8 bytes per pair, nearly all of them -0.0) lives in a buffer that stays -0.0 between layers: only the
        # candidate pairs are written, and cleared again afterwards.
        if self._neg.size < n * width:             # with headroom: the ROI count drifts from layer to layer
            self._neg = np.full(n * width + (n * width >> 2), -0.0)
        neg = self._neg[: n * width].reshape(n, width)
        assigned, taken = {}, set()
        try:
            neg[pair_c, pair_p] = -score
            order = np.argsort(neg, axis=None) if neg.size else np.zeros(0, np.int64)
        finally:
            neg[pair_c, pair_p] = -0.0
        # scores at or above the threshold come first; the reference stops at the first one below it
        n_hits = int(np.count_nonzero(score >= self.MATCH_THRESHOLD))
        for flat in order[:n_hits].tolist():
            ci, pj = divmod(flat, width)
            pid = prev_ids[pj]
            if ci in assigned or pid in taken:
                continue
            assigned[ci] = pid
            taken.add(pid)

        for ci, pid in assigned.items():
            roi, before = current_rois_raw[ci], self.previous_rois[pid]
            roi["id"] = pid
            if before["classification"] == "support":
                growth = roi["area"] / before["area"] if before["area"] > 0 else float("inf")
                roi["classification"] = "model" if growth > config.roi_params.support_max_growth else "support"
            else:
                roi["classification"] = before["classification"]
            tracked[pid] = roi
        for ci, roi in enumerate(current_rois_raw):
            if ci not in assigned:
                self._admit(roi, layer_index, config, tracked)
        self.previous_rois = tracked
        return list(tracked.values())
