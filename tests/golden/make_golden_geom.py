#!/usr/bin/env python
"""
Golden fixtures for the SURVEY.md section 8(f) rank-2 operations (unsharp mask, resize, anisotropic Enhanced EDT),
produced by the UNMODIFIED reference functions imported read-only from /root/reference (same provenance rules as
make_golden.py; /root/reference does not exist on the GPU box, so the outputs are committed).

    python tests/golden/make_golden_geom.py       # rewrites tests/golden/ref_geom.npz
"""
import glob
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, _enter_reference  # noqa: E402


def main():
    _enter_reference()
    import cv2
    import numpy as np
    import processing_core as core
    import xy_blend_processor as xy
    from config import Config, ProcessingMode, XYBlendOperation

    files = sorted(glob.glob(os.path.join(REF, "test-slices", "*.png")))
    out = {}

    # ---- inputs: a gray-level crop (slice 30 blurred by the reference's own Gaussian) and seeded noise -------------
    g30 = cv2.imread(files[30], cv2.IMREAD_GRAYSCALE)[900:1133, 1800:2101].copy()          # 233 x 301
    soft = xy.apply_gaussian_blur(g30, XYBlendOperation(type="gaussian_blur", gaussian_ksize_x=9, gaussian_ksize_y=9,
                                                         gaussian_sigma_x=2.0, gaussian_sigma_y=2.0))
    noise = np.random.default_rng(7).integers(0, 256, (97, 131), dtype=np.uint8)
    out["in_soft"] = soft
    out["in_noise"] = noise

    # ---- apply_unsharp_mask ------------------------------------------------------------------------------------------
    unsharp_sets = [(1.0, 0, 5, 0.0), (0.3, 0, 3, 0.8), (2.7, 4, 7, 1.5), (0.7, 12, 5, 0.0), (1.5, 1, 9, 2.2)]
    for i, (amount, thr, k, sig) in enumerate(unsharp_sets):
        op = XYBlendOperation(type="unsharp_mask", unsharp_amount=amount, unsharp_threshold=thr, unsharp_blur_ksize=k,
                              unsharp_blur_sigma=sig)
        for name, img in (("soft", soft), ("noise", noise)):
            out[f"unsharp_{i}_{name}"] = xy.apply_unsharp_mask(img, op)

    # ---- apply_resize ------------------------------------------------------------------------------------------------
    # (width, height) requests incl. one-sided ones; integer and fractional factors, enlarging, shrinking and mixed
    resize_sets = [(602, 466), (150, None), (None, 100), (100, 77), (301, 466), (450, 120), (43, 233), (303, 240),
                   (60, 233), (301, 233), (None, None), (700, 50)]
    modes = ["NEAREST", "BILINEAR", "BICUBIC", "LANCZOS4", "AREA", "bogus"]
    for i, (w, h) in enumerate(resize_sets):
        for m in modes:
            op = XYBlendOperation(type="resize", resize_width=w, resize_height=h, resample_mode=m)
            for name, img in (("soft", soft), ("noise", noise)):
                if name == "noise" and i not in (0, 3, 5, 8):
                    continue
                out[f"resize_{i}_{m}_{name}"] = xy.apply_resize(img, op)
                if m == "BICUBIC":   # OpenCV's own kernel (the default build routes uint8 cubic through IPP)
                    cv2.ipp.setUseIPP(False)
                    out[f"resize_{i}_{m}_{name}_noipp"] = xy.apply_resize(img, op)
                    cv2.ipp.setUseIPP(True)

    # ---- anisotropic Enhanced EDT ------------------------------------------------------------------------------------
    def masks(z, N, win):
        y0, y1, x0, x1 = win
        cur = core.load_image(files[z])[0][y0:y1, x0:x1].copy()
        pri = [core.load_image(files[k])[0][y0:y1, x0:x1].copy() for k in range(max(0, z - N), z)]
        return cur, list(reversed(pri))

    wins = {"a": (250, 634, 300, 812), "b": (900, 1284, 1800, 2312)}
    aniso_sets = [(1.5, 0.75), (2.0, 1.0), (0.5, 0.5), (1.0, 3.0), (20.0, 0.01)]
    for wn, win in wins.items():
        for z in (14, 30):
            cur, pri = masks(z, 4, win)
            for j, (xf, yf) in enumerate(aniso_sets):
                for F in (10.0, 40.0):
                    for nb in (False, True):
                        if j == 4 and (F != 10.0 or wn != "a"):
                            continue   # the clamped 10x case is big; once is enough
                        c = Config(); c.blending_mode = ProcessingMode.ENHANCED_EDT
                        c.fixed_fade_distance_receding = F; c.use_numba_jit = nb
                        c.anisotropic_params.enabled = True
                        c.anisotropic_params.x_factor = xf; c.anisotropic_params.y_factor = yf
                        for ipp in (True, False):
                            cv2.ipp.setUseIPP(ipp)
                            out[f"aniso_{wn}_{z:03d}_{j}_F{F:g}_{'numba' if nb else 'scipy'}{'' if ipp else '_noipp'}"] = \
                                core.process_z_blending(cur, pri, c, [])
                        cv2.ipp.setUseIPP(True)

    np.savez_compressed(os.path.join(HERE, "ref_geom.npz"), unsharp_sets=json.dumps(unsharp_sets),
                        resize_sets=json.dumps(resize_sets), resize_modes=json.dumps(modes),
                        aniso_sets=json.dumps(aniso_sets), wins=json.dumps(wins), **out)
    print("ref_geom.npz", os.path.getsize(os.path.join(HERE, "ref_geom.npz")) / 1e6, "MB,", len(out), "arrays")


if __name__ == "__main__":
    main()
