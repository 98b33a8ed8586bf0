#!/usr/bin/env python
"""
Generates the golden fixtures in tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported read-only) in the build container.  /root/reference does not exist on
the GPU box, so the outputs are committed; this script is the provenance.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

Library versions the fixtures were produced with are recorded in versions.json (the reference
pins other versions in requirements.txt).
"""
import glob
import hashlib
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _enter_reference():
    """cwd -> temp dir (config.py:303-305 writes app_config.json into cwd), PySide6 stub
     first on sys.path, then the reference itself."""
    tmp = tempfile.mkdtemp(prefix="vsb_golden_")
    stub = os.path.join(tmp, "stub", "PySide6")
    os.makedirs(stub)
    open(os.path.join(stub, "__init__.py"), "w").close()
    with open(os.path.join(stub, "QtCore.py"), "w") as fh:
        fh.write(
            "class Signal:\n"
            "    def __init__(self,*a): self._s=[]\n"
            "    def connect(self,f): self._s.append(f)\n"
            "    def emit(self,*a):\n"
            "        for f in self._s: f(*a)\n"
            "class QThread:\n"
            "    def __init__(self,*a,**k): pass\n"
            "    def start(self): self.run()\n"
            "    def isRunning(self): return False\n"
            "    def wait(self,*a): return True\n")
    os.chdir(tmp)
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(tmp, "stub"))
    return tmp


def sha(a):
    import numpy as np
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    tmp = _enter_reference()
    import cv2
    import numpy as np
    import scipy
    import numba
    import processing_core as core
    import xy_blend_processor as xy
    import lut_manager as lm
    import processing_pipeline as pp
    from config import Config, ProcessingMode, XYBlendOperation, LutParameters

    with open(os.path.join(HERE, "versions.json"), "w") as fh:
        json.dump({"cv2": cv2.__version__, "numpy": np.__version__, "scipy": scipy.__version__,
                   "numba": numba.__version__, "cv2_ipp": bool(cv2.ipp.useIPP()) if hasattr(cv2, "ipp") else None},
                  fh, indent=1)

    files = sorted(glob.glob(os.path.join(REF, "test-slices", "*.png")))
    grays = [cv2.imread(f, cv2.IMREAD_GRAYSCALE) for f in files]
    assert all(set(np.unique(g)) <= {0, 255} for g in grays)
    H, W = grays[0].shape

    # ---- inputs: the 40 test slices, bit-packed (strictly {0,255}) ---------------------------
    bits = np.stack([np.packbits(g > 127, axis=1) for g in grays])
    np.savez_compressed(os.path.join(HERE, "test_slices_bits.npz"), bits=bits, shape=np.array([H, W]),
                        names=np.array([os.path.basename(f) for f in files]))

    # ---- full stack through the reference's own ProcessingPipelineThread.run() -----------------
    def run_stack(cfg, tag):
        out_dir = os.path.join(tmp, "out_" + tag)
        os.makedirs(out_dir)
        cfg.input_mode = "folder"
        cfg.input_folder = os.path.join(REF, "test-slices")
        cfg.output_folder = out_dir
        cfg.start_index, cfg.stop_index = 0, None
        th = pp.ProcessingPipelineThread(cfg, max_workers=4)
        errs = []
        th.error_signal.connect(errs.append)
        th.run()
        assert not errs, errs
        outs = [cv2.imread(os.path.join(out_dir, os.path.basename(f)), cv2.IMREAD_GRAYSCALE) for f in files]
        return outs

    lut_path = os.path.join(REF, "saved_luts", "EXP(LUT)-upShifted.json")
    cfg1 = Config()
    cfg1.blending_mode = ProcessingMode.FIXED_FADE
    cfg1.receding_layers = 3
    cfg1.use_fixed_fade_receding = True
    cfg1.fixed_fade_distance_receding = 10.0
    cfg1.xy_blend_pipeline = [XYBlendOperation(type="apply_lut",
                                               lut_params=LutParameters(lut_source="file", fixed_lut_path=lut_path))]
    outs1 = run_stack(cfg1, "cfg1")

    with open(os.path.join(REF, "presets", "Preset-Double-LUT.json")) as fh:
        preset = json.load(fh)
    preset["xy_blend_pipeline"][0]["lut_params"]["fixed_lut_path"] = lut_path    # SURVEY finding 5
    cfgp = Config.from_dict(preset)
    outsp = run_stack(cfgp, "preset")

    keep = [0, 13, 14, 20, 30, 31, 36, 39]
    np.savez_compressed(
        os.path.join(HERE, "ref_stack.npz"),
        cfg1_sha1=np.array([sha(o) for o in outs1]), preset_sha1=np.array([sha(o) for o in outsp]),
        keep=np.array(keep),
        **{f"cfg1_{z:03d}": outs1[z] for z in keep}, **{f"preset_{z:03d}": outsp[z] for z in keep})
    print("cfg1 stack sha1", hashlib.sha1(b"".join(o.tobytes() for o in outs1)).hexdigest())
    print("preset stack sha1", hashlib.sha1(b"".join(o.tobytes() for o in outsp)).hexdigest())

    # ---- per-mode outputs of process_z_blending on crops ---------------------------------------
    def masks(z, N, win):
        y0, y1, x0, x1 = win
        cur = core.load_image(files[z])[0][y0:y1, x0:x1].copy()
        pri = [core.load_image(files[k])[0][y0:y1, x0:x1].copy() for k in range(max(0, z - N), z)]
        return cur, list(reversed(pri))

    wins = {"a": (250, 634, 300, 812), "b": (900, 1284, 1800, 2312), "c": (1500, 1884, 3300, 3812)}
    modes = {}
    meta = []
    for wn, win in wins.items():
        for z in (13, 14, 22, 30, 36):
            cur, pri = masks(z, 4, win)
            key = f"{wn}_{z:03d}"
            for F, N in ((10.0, 3), (40.0, 4), (7.5, 2)):
                c = Config(); c.blending_mode = ProcessingMode.FIXED_FADE
                c.use_fixed_fade_receding = True; c.fixed_fade_distance_receding = F
                comb = core.find_prior_combined_white_mask(pri[:N])
                modes[f"fixed_{key}_F{F:g}_N{N}"] = core.process_z_blending(cur, comb, c, [])
            c = Config(); c.blending_mode = ProcessingMode.FIXED_FADE; c.use_fixed_fade_receding = False
            modes[f"minmax_{key}_N4"] = core.process_z_blending(cur, core.find_prior_combined_white_mask(pri), c, [])
            c = Config(); c.blending_mode = ProcessingMode.WEIGHTED_STACK
            c.manual_weights = [100, 75, 50, 25]; c.fade_distances_receding = [40.0, 30.0, 20.0, 10.0]
            modes[f"weighted_{key}"] = core.process_z_blending(cur, pri, c, [])
            c.manual_weights = [10, 0, 30]; c.fade_distances_receding = [12.5]; c.fixed_fade_distance_receding = 25.0
            modes[f"weighted2_{key}"] = core.process_z_blending(cur, pri, c, [])
            for F in (10.0, 40.0, 64.0):
                for nb in (False, True):
                    c = Config(); c.blending_mode = ProcessingMode.ENHANCED_EDT
                    c.fixed_fade_distance_receding = F; c.use_numba_jit = nb
                    modes[f"enh_{'numba' if nb else 'scipy'}_{key}_F{F:g}"] = core.process_z_blending(cur, pri, c, [])
            meta.append((wn, z))
    np.savez_compressed(os.path.join(HERE, "ref_modes.npz"),
                        wins=json.dumps(wins), **modes)

    # ---- raw distance maps of the library call (5x5 chamfer and precise) -------------------------
    # 192x256 windows picked where the full-layer 5x5 map has the most pixels with 0 < d < 24.
    dist = {}
    dwins = {}
    for z in (14, 30, 36):
        cur_full = core.load_image(files[z])[0]
        d_full = cv2.distanceTransform(cv2.bitwise_not(cur_full), cv2.DIST_L2, 5)
        score = ((d_full > 0) & (d_full < 24)).astype(np.float64)
        ii = np.pad(score.cumsum(0).cumsum(1), ((1, 0), (1, 0)))
        best, by, bx = -1, 0, 0
        for y0 in range(0, H - 192, 64):
            for x0 in range(0, W - 256, 64):
                v = ii[y0 + 192, x0 + 256] - ii[y0, x0 + 256] - ii[y0 + 192, x0] + ii[y0, x0]
                if v > best:
                    best, by, bx = v, y0, x0
        dwins[f"{z:03d}"] = (by, by + 192, bx, bx + 256)
        cur = cur_full[by:by + 192, bx:bx + 256].copy()
        src = cv2.bitwise_not(cur)
        dist[f"d5_{z:03d}"] = cv2.distanceTransform(src, cv2.DIST_L2, 5)
        dist[f"dp_{z:03d}"] = cv2.distanceTransform(src, cv2.DIST_L2, cv2.DIST_MASK_PRECISE)
    np.savez_compressed(os.path.join(HERE, "ref_dist.npz"), wins=json.dumps(dwins), **dist)

    # ---- the scenes of the reference's own tests/test_processing.py -----------------------------
    scenes = {}
    cur = np.zeros((100, 100), np.uint8); cv2.rectangle(cur, (45, 45), (55, 55), 255, -1)
    old = np.zeros_like(cur); cv2.rectangle(old, (35, 35), (65, 65), 255, -1)
    new = np.zeros_like(cur); cv2.rectangle(new, (40, 40), (60, 60), 255, -1)
    c = Config(); c.blending_mode = ProcessingMode.WEIGHTED_STACK; c.fixed_fade_distance_receding = 20.0
    c.manual_weights = [100, 50]; c.fade_distances_receding = [20.0, 10.0]
    scenes["weighted_cur"], scenes["weighted_new"], scenes["weighted_old"] = cur, new, old
    scenes["weighted_out"] = core._calculate_weighted_receding_gradient_field(cur, [new, old], c)
    prior = np.zeros_like(cur); cv2.rectangle(prior, (25, 40), (45, 60), 255, -1); cv2.rectangle(prior, (55, 40), (60, 60), 255, -1)
    scenes["enh_cur"], scenes["enh_prior"] = cur, prior
    for nb in (False, True):
        c = Config(); c.blending_mode = ProcessingMode.ENHANCED_EDT; c.fixed_fade_distance_receding = 10.0; c.use_numba_jit = nb
        scenes["enh_out_" + ("numba" if nb else "scipy")] = core._calculate_receding_gradient_field_enhanced_edt(cur, [prior], c)
    np.savez_compressed(os.path.join(HERE, "ref_unit_scenes.npz"), **scenes)

    # ---- XY ops and LUT tables -----------------------------------------------------------------
    rng = np.random.default_rng(0)
    Hs, Ws = 150, 211
    yy, xx = np.mgrid[0:Hs, 0:Ws]
    noise = rng.integers(0, 256, (Hs, Ws), dtype=np.uint8)
    smooth = ((np.sin(xx / 17.0) + np.cos(yy / 11.0)) * 60 + 128).clip(0, 255).astype(np.uint8)
    disk = (((xx - 100) ** 2 + (yy - 70) ** 2) < 50 ** 2).astype(np.uint8) * 255
    soft = cv2.GaussianBlur(disk, (9, 9), 2.5)
    zb = outs1[30][600:750, 1000:1211].copy()            # a real Z-blended + LUT'ed patch
    xyd = {"in_noise": noise, "in_smooth": smooth, "in_soft": soft, "in_zblend": zb}
    gauss_sets = [(3, 3, 0.8, 0.0), (3, 3, 0.0, 0.0), (5, 5, 0.0, 0.0), (5, 3, 1.2, 0.7), (7, 7, 2.0, 0.0),
                  (9, 9, 0.0, 0.0), (1, 5, 0.0, 1.0), (11, 1, 3.0, 0.0), (15, 15, 4.0, 0.0), (31, 31, 6.0, 0.0),
                  (13, 7, 0.0, 0.0), (21, 21, 0.5, 0.0)]
    bil_sets = [(7, 60.0, 60.0), (9, 75.0, 75.0), (0, 30.0, 2.0), (11, 10.0, 5.0), (7, 0.0, 0.0), (5, 50.0, 50.0),
                (3, 20.0, 20.0)]
    med_sets = [1, 3, 5, 7, 9, 15]
    for nm in ("noise", "smooth", "soft", "zblend"):
        im = xyd["in_" + nm]
        for (kx, ky, sx, sy) in gauss_sets:
            op = XYBlendOperation(type="gaussian_blur", gaussian_ksize_x=kx, gaussian_ksize_y=ky,
                                  gaussian_sigma_x=sx, gaussian_sigma_y=sy)
            xyd[f"gauss_{nm}_{kx}_{ky}_{sx:g}_{sy:g}"] = xy.apply_gaussian_blur(im, op)
        for (d, sc, ss) in bil_sets:
            op = XYBlendOperation(type="bilateral_filter", bilateral_d=d, bilateral_sigma_color=sc, bilateral_sigma_space=ss)
            xyd[f"bil_{nm}_{d}_{sc:g}_{ss:g}"] = xy.apply_bilateral_filter(im, op)
        for k in med_sets:
            op = XYBlendOperation(type="median_blur", median_ksize=k)
            xyd[f"med_{nm}_{k}"] = xy.apply_median_blur(im, op)
    # the preset's XY chain applied through process_xy_pipeline
    xyd["chain_zblend_preset"] = xy.process_xy_pipeline(zb, cfgp.xy_blend_pipeline)
    xyd["chain_noise_preset"] = xy.process_xy_pipeline(noise, cfgp.xy_blend_pipeline)
    np.savez_compressed(os.path.join(HERE, "ref_xy.npz"), gauss_sets=json.dumps(gauss_sets),
                        bil_sets=json.dumps(bil_sets), med_sets=json.dumps(med_sets), **xyd)

    luts = {}
    lut_specs = []
    for rngs in ((200, 255, 200, 255), (10, 240, 30, 220), (0, 255, 0, 255), (128, 128, 0, 255), (0, 255, 255, 255)):
        for kind, params in (("linear", [{}]), ("gamma", [{"gamma_value": 0.5}, {"gamma_value": 2.2}]),
                             ("s_curve", [{"s_curve_contrast": 0.0}, {"s_curve_contrast": 0.5}, {"s_curve_contrast": 1.0}]),
                             ("log", [{"log_param": 10.0}, {"log_param": 0.5}]),
                             ("exp", [{"exp_param": 2.0}, {"exp_param": 0.7}]),
                             ("sqrt", [{"sqrt_param": 2.0}, {"sqrt_param": 3.0}]),
                             ("rodbard", [{"rodbard_param": 1.0}, {"rodbard_param": 0.4}]),
                             ("spline", [{"spline_points": [[0, 0], [255, 255]]},
                                         {"spline_points": [[0, 0], [64, 30], [128, 140], [255, 255]]},
                                         {"spline_points": [[30, 10], [200, 250]]}])):
            for p in params:
                lp = LutParameters(lut_source="generated", lut_generation_type=kind, input_min=rngs[0],
                                   input_max=rngs[1], output_min=rngs[2], output_max=rngs[3], **p)
                op = XYBlendOperation(type="apply_lut", lut_params=lp)
                ramp = np.arange(256, dtype=np.uint8).reshape(1, 256)
                spec = {"lut_generation_type": kind, "input_min": rngs[0], "input_max": rngs[1],
                        "output_min": rngs[2], "output_max": rngs[3], **p}
                luts[f"lut_{len(lut_specs):03d}"] = xy.apply_lut_operation(ramp, op).ravel()
                lut_specs.append(spec)
    for f in sorted(glob.glob(os.path.join(REF, "saved_luts", "*.json"))):
        luts["file_" + os.path.basename(f)[:-5]] = lm.load_lut(f)
    np.savez_compressed(os.path.join(HERE, "ref_luts.npz"), specs=json.dumps(lut_specs), **luts)
    print("done; fixtures in", HERE)
    for f in sorted(glob.glob(os.path.join(HERE, "*.npz"))):
        print(f"  {os.path.basename(f):28s} {os.path.getsize(f)/1e6:.2f} MB")


if __name__ == "__main__":
    main()
