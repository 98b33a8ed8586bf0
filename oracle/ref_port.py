"""
ref_port -- TEST / MEASUREMENT INFRASTRUCTURE ONLY: the reference's per-layer call sequence on the CPU with the
same third-party wheels the reference itself delegates to (cv2, numpy, scipy, numba when importable).  This is
what `bench.py --impl reference` and bench.py's `cpu_baseline` leg time: the reference is pure Python and
cannot travel to the GPU box (/root/reference does not exist there), so its hot path is restated here op for
op -- same library calls, same float32 temporaries, same thread-pool fan-out with a sliding window of prior
masks -- and checked against the committed reference outputs in tests/test_ref_port.py.

Follows: processing_core.py:44 (threshold), :47-67 (OR), :107-153 (fixed fade), :222-284 (weighted),
:338-404 (enhanced), :442-460 (merge); xy_blend_processor.py:27-46,92-164; processing_pipeline.py:74-113,
:163-223 (deque window, newest-first snapshot, ThreadPoolExecutor with 2x back-pressure).
Never imported by the product.
"""
from __future__ import annotations

import collections
import concurrent.futures
from typing import Sequence

import numpy as np

try:
    import cv2
    HAVE_CV2 = True
except Exception:  # pragma: no cover
    cv2 = None
    HAVE_CV2 = False

import vsb_oracle as O


def threshold(gray):
    return cv2.threshold(gray, 127, 255, cv2.THRESH_BINARY)[1]


def combine(priors):
    if not priors:
        return None
    acc = priors[0].copy()
    for p in priors[1:]:
        acc = cv2.bitwise_or(acc, p)
    return acc


def fixed_fade(cur, comb, F):
    if comb is None:
        return np.zeros_like(cur)
    rec = cv2.bitwise_and(comb, cv2.bitwise_not(cur))
    if cv2.countNonZero(rec) == 0:
        return np.zeros_like(cur)
    dist = cv2.distanceTransform(cv2.bitwise_not(cur), cv2.DIST_L2, 5)
    rd = cv2.bitwise_and(dist, dist, mask=rec)
    if np.max(rd) == 0:
        return np.zeros_like(cur)
    norm = np.clip(rd, 0, F) / (F if F > 0 else 1.0)
    grad = (255 * (1.0 - norm)).astype(np.uint8)
    return cv2.bitwise_and(grad, grad, mask=rec)


def weighted_fade(cur, priors, weights, fades, default_fade):
    if not priors or not weights:
        return np.zeros_like(cur)
    k = min(len(priors), len(weights))
    priors, weights = priors[:k], weights[:k]
    total = float(sum(weights))
    if k == 0 or total <= 0:
        return np.zeros_like(cur)
    acc = np.zeros(cur.shape, np.float32)
    ncur = cv2.bitwise_not(cur)
    dist = cv2.distanceTransform(ncur, cv2.DIST_L2, 5)
    union = np.zeros_like(cur)
    for i, (p, w) in enumerate(zip(priors, weights)):
        rec = cv2.bitwise_and(p, ncur)
        union = cv2.bitwise_or(union, rec)
        if w <= 0 or cv2.countNonZero(rec) == 0:
            continue
        Fi = fades[i] if i < len(fades) else default_fade
        rd = cv2.bitwise_and(dist, dist, mask=rec)
        contrib = (1.0 - np.clip(rd, 0, Fi) / (Fi if Fi > 0 else 1.0)) * w
        acc += np.where(rec > 0, contrib, 0)
    grad = (255 * (acc / total)).astype(np.uint8)
    return cv2.bitwise_and(grad, grad, mask=union)


def enhanced_fade(cur, priors, F, numba_flavour):
    if not priors:
        return np.zeros_like(cur)
    rec = cv2.bitwise_and(combine(priors), cv2.bitwise_not(cur))
    if cv2.countNonZero(rec) == 0:
        return np.zeros_like(cur)
    dist = cv2.distanceTransform(cv2.bitwise_not(cur), cv2.DIST_L2, 5)
    rd = cv2.bitwise_and(dist, dist, mask=rec)
    n, labels = cv2.connectedComponents(rec)
    if n <= 1:
        return np.zeros_like(cur)
    grad = O.enhanced_from_distance(rd, labels.astype(np.int32), n, F, "numba" if numba_flavour else "scipy")
    return cv2.bitwise_and(grad, grad, mask=rec)


def xy_chain(img, ops):
    out = img.copy()
    for op in ops or []:
        t = op.get("type", "none")
        if t == "apply_lut":
            out = O.lut_from_params(op.get("lut_params", {}))[out]
        elif t == "gaussian_blur":
            out = cv2.GaussianBlur(out, (op.get("gaussian_ksize_x", 3), op.get("gaussian_ksize_y", 3)),
                                   sigmaX=op.get("gaussian_sigma_x", 0.0), sigmaY=op.get("gaussian_sigma_y", 0.0))
        elif t == "bilateral_filter":
            out = cv2.bilateralFilter(out, op.get("bilateral_d", 9), op.get("bilateral_sigma_color", 75.0),
                                      op.get("bilateral_sigma_space", 75.0))
        elif t == "median_blur" and op.get("median_ksize", 5) > 1:
            out = cv2.medianBlur(out, op.get("median_ksize", 5))
    return out


def process_layer(gray, cur, priors_newest_first, cfg):
    mode = cfg.get("blending_mode", "fixed_fade")
    F = cfg.get("fixed_fade_distance_receding", 10.0)
    if mode == "weighted_stack":
        grad = weighted_fade(cur, list(priors_newest_first), cfg.get("manual_weights", [100, 75, 50, 25]),
                             cfg.get("fade_distances_receding", [10.0] * 4), F)
    elif mode == "enhanced_edt":
        grad = enhanced_fade(cur, list(priors_newest_first), F, cfg.get("use_numba_jit", False))
    else:
        grad = fixed_fade(cur, combine(list(priors_newest_first)), F)
    return xy_chain(np.maximum(gray, grad), cfg.get("xy_blend_pipeline", []))


def run_stack(layers: Sequence[np.ndarray], cfg: dict, workers: int = 1, halo: Sequence[np.ndarray] = (), keep: bool = True):
    """The reference's producer loop: serial threshold on the producer, per-layer tasks on a thread pool
    (cv2 / numpy release the GIL), at most 2*workers tasks in flight."""
    if not HAVE_CV2:
        return O.run_stack(layers, cfg, halo)
    window = collections.deque(maxlen=int(cfg.get("receding_layers", 4)))
    for h in halo:
        window.append(threshold(h))
    outs = [None] * len(layers)
    with concurrent.futures.ThreadPoolExecutor(max(1, workers)) as pool:
        active = {}
        for i, g in enumerate(layers):
            while len(active) >= 2 * max(1, workers):
                done, _ = concurrent.futures.wait(active, return_when=concurrent.futures.FIRST_COMPLETED)
                for f in done:
                    j = active.pop(f)
                    r = f.result()
                    outs[j] = r if keep else True
            cur = threshold(g)
            active[pool.submit(process_layer, g, cur, list(reversed(window)), cfg)] = i
            window.append(cur)
        for f in concurrent.futures.as_completed(list(active)):
            j = active.pop(f)
            r = f.result()
            outs[j] = r if keep else True
    return outs
