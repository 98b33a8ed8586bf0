"""CPU: pins the oracle (oracle/vsb_oracle.py + oracle/c) against outputs of the unmodified
reference (tests/golden/*.npz, produced by tests/golden/make_golden.py) and against the numeric
assertions of the reference's own tests/test_processing.py."""
import hashlib
import json

import numpy as np
import pytest

import vsb_oracle as O
from conftest import preset_ops


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def _crop(slices, z, N, win):
    y0, y1, x0, x1 = win
    cur = O.threshold(slices[z][y0:y1, x0:x1])
    pri = [O.threshold(slices[k][y0:y1, x0:x1]) for k in range(max(0, z - N), z)]
    return np.ascontiguousarray(cur), [np.ascontiguousarray(p) for p in reversed(pri)]


def test_chamfer_closed_form_equals_raster_algorithm():
    rng = np.random.default_rng(3)
    for shape, p in (((40, 57), 0.02), ((33, 31), 0.3), ((64, 64), 0.002)):
        seed = (rng.random(shape) < p).astype(np.uint8) * 255
        seed[shape[0] // 2, shape[1] // 2] = 255
        R = max(shape)
        a = O.chamfer5(seed, R)
        b = O.chamfer5(seed, R, brute=True)
        c = O.chamfer5_twopass(seed)
        assert np.array_equal(a, b)
        assert np.abs(a - c).max() < 1e-4


def test_distance_maps_vs_library(slices, ref_dist):
    z, wins = ref_dist
    for lz, win in wins.items():
        cur, _ = _crop(slices, int(lz), 0, win)
        d = O.chamfer5(cur, 64)
        ref = z["d5_" + lz]
        m = np.isfinite(d)
        assert m.sum() > 10000
        assert np.abs(d[m] - ref[m]).max() < 1e-4                  # north_star: distances within 1e-4 px
        assert (ref[~m] > 64).all()
        dp = O.edt_exact(cur)
        assert np.abs(dp - z["dp_" + lz]).max() < 1e-4


def test_modes_on_crops(slices, ref_modes):
    z, wins = ref_modes
    checked = 0
    for key in z.files:
        if key == "wins":
            continue
        parts = key.split("_")
        kind = parts[0]
        if kind == "enh":
            flavour, wn, lz, Fs = parts[1], parts[2], parts[3], parts[4]
            cur, pri = _crop(slices, int(lz), 4, wins[wn])
            got = O.enhanced_fade(cur, pri, float(Fs[1:]), flavour)
        elif kind == "fixed":
            wn, lz, Fs, Ns = parts[1], parts[2], parts[3], parts[4]
            cur, pri = _crop(slices, int(lz), 4, wins[wn])
            got = O.fixed_fade(cur, O.combine_priors(pri[:int(Ns[1:])]), float(Fs[1:]), True)
        elif kind == "minmax":
            wn, lz = parts[1], parts[2]
            if wn != "a" or lz not in ("014", "022"):
                continue                                     # unbounded reach is slow on the CPU oracle
            cur, pri = _crop(slices, int(lz), 4, wins[wn])
            got = O.fixed_fade(cur, O.combine_priors(pri), 10.0, False)
            # unbounded reach uses exact-integer chamfer costs rounded once (a few ulps from the library's raster sums):
            # +-1 LSB on a handful of pixels (5 of 542,867 over all min-max goldens)
            d = np.abs(got.astype(int) - z[key].astype(int))
            assert d.max() <= 1 and (d > 0).sum() <= 5, key
            checked += 1
            continue
        elif kind == "weighted":
            wn, lz = parts[1], parts[2]
            cur, pri = _crop(slices, int(lz), 4, wins[wn])
            got = O.weighted_fade(cur, pri, [100, 75, 50, 25], [40.0, 30.0, 20.0, 10.0], 10.0)
        elif kind == "weighted2":
            wn, lz = parts[1], parts[2]
            cur, pri = _crop(slices, int(lz), 4, wins[wn])
            got = O.weighted_fade(cur, pri, [10, 0, 30], [12.5], 25.0)
        else:
            raise AssertionError(key)
        assert np.array_equal(got, z[key]), key
        checked += 1
    assert checked > 150


def test_reference_unit_test_scenes(ref_scenes):
    s = ref_scenes
    g = O.weighted_fade(s["weighted_cur"], [s["weighted_new"], s["weighted_old"]], [100, 50], [20.0, 10.0], 20.0)
    assert np.array_equal(g, s["weighted_out"])
    assert 200 < g[42, 50] < 210 and 20 < g[38, 50] < 30          # tests/test_processing.py:75-76
    for flavour in ("scipy", "numba"):
        e = O.enhanced_fade(s["enh_cur"], [s["enh_prior"]], 10.0, flavour)
        assert np.array_equal(e, s["enh_out_" + flavour])
        assert e[50, 35] == 0 and e[50, 25] == 0                   # :140,148
        assert 180 < e[50, 57] < 185 and 108 < e[50, 59] < 112     # :144,152
    # test_empty_inputs (:78-90)
    cur = np.zeros((100, 100), np.uint8)
    assert O.weighted_fade(cur, [], [100], [10.0], 10.0).sum() == 0
    assert O.weighted_fade(cur, [np.full_like(cur, 255)], [], [10.0], 10.0).sum() == 0


def test_luts(ref_luts):
    z, specs = ref_luts
    for i, spec in enumerate(specs):
        got = O.lut_from_params(dict(spec, lut_source="generated"))
        assert np.array_equal(got, z[f"lut_{i:03d}"]), spec


def test_xy_ops(ref_xy):
    z = ref_xy
    for nm in ("noise", "smooth", "soft", "zblend"):
        im = z["in_" + nm]
        for kx, ky, sx, sy in json.loads(str(z["gauss_sets"])):
            assert np.array_equal(O.gaussian_blur(im, kx, ky, sx, sy), z[f"gauss_{nm}_{kx}_{ky}_{sx:g}_{sy:g}"])
        for d, sc, ss in json.loads(str(z["bil_sets"])):
            got = O.bilateral(im, d, sc, ss)
            ref = z[f"bil_{nm}_{d}_{sc:g}_{ss:g}"]
            diff = np.abs(got.astype(int) - ref.astype(int))
            assert diff.max() <= 1                          # float path: +-1 LSB (north_star)
            if d >= 7 or d <= 0:
                assert (diff != 0).sum() <= 2               # generic OpenCV path reproduces bit for bit
        for k in json.loads(str(z["med_sets"])):
            assert np.array_equal(O.median_blur(im, k), z[f"med_{nm}_{k}"])


def test_xy_chain_preset(ref_xy, lut_file):
    ops = preset_ops(lut_file)
    for nm in ("zblend", "noise"):
        got = O.xy_pipeline(ref_xy["in_" + nm], ops)
        assert np.array_equal(got, ref_xy[f"chain_{nm}_preset"])


@pytest.mark.parametrize("which", ["cfg1", "preset"])
def test_full_stack_subset(slices, ref_stack, lut_file, which):
    """Layers of the 40-slice stack through the oracle's stack loop vs ProcessingPipelineThread.run()."""
    if which == "cfg1":
        cfg = {"blending_mode": "fixed_fade", "receding_layers": 3, "use_fixed_fade_receding": True,
               "fixed_fade_distance_receding": 10.0,
               "xy_blend_pipeline": [{"type": "apply_lut", "lut_params": {"lut_source": "file", "fixed_lut_path": lut_file}}]}
        zs = [0, 13, 14, 20, 30, 31, 36, 39]
    else:
        cfg = {"blending_mode": "fixed_fade", "receding_layers": 4, "use_fixed_fade_receding": True,
               "fixed_fade_distance_receding": 40.0, "xy_blend_pipeline": preset_ops(lut_file)}
        zs = [13, 30, 36]
    N = cfg["receding_layers"]
    for zi in zs:
        lo = max(0, zi - N)
        out = O.run_stack([slices[zi]], cfg, halo=slices[lo:zi])[0]
        assert sha(out) == str(ref_stack[which + "_sha1"][zi]), zi
        assert np.array_equal(out, ref_stack[f"{which}_{zi:03d}"])
