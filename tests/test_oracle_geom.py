"""Oracle (CPU) vs the reference's own outputs for the section-8(f) rank-2 operations (tests/golden/ref_geom.npz,
written by tests/golden/make_golden_geom.py from the unmodified reference)."""
import json
import os

import numpy as np
import pytest

import vsb_oracle as O
import vsb_oracle_geom as G

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def geom():
    return np.load(os.path.join(HERE, "golden", "ref_geom.npz"))


def test_unsharp_mask_bit_exact(geom):
    for i, (amount, thr, k, sig) in enumerate(json.loads(str(geom["unsharp_sets"]))):
        for name in ("soft", "noise"):
            got = G.unsharp_mask(geom[f"in_{name}"], amount, thr, k, sig, O.gaussian_blur)
            assert np.array_equal(got, geom[f"unsharp_{i}_{name}"]), (i, name)


def test_resize_all_modes(geom):
    sets = json.loads(str(geom["resize_sets"]))
    modes = json.loads(str(geom["resize_modes"]))
    checked = 0
    for i, (w, h) in enumerate(sets):
        for m in modes:
            for name in ("soft", "noise"):
                key = f"resize_{i}_{m}_{name}"
                if key not in geom.files:
                    continue
                src = geom[f"in_{name}"]
                got = G.apply_resize(src, w, h, m)
                ref = geom[key]
                assert got.shape == ref.shape, (key, got.shape, ref.shape)
                if m == "BICUBIC" and got.shape != src.shape:
                    # bit-exact against OpenCV's own kernel; the IPP build of cv2 is within one grey level of it
                    assert np.array_equal(got, geom[key + "_noipp"]), key
                    assert np.abs(got.astype(int) - ref.astype(int)).max() <= 1, key
                else:
                    assert np.array_equal(got, ref), key
                checked += 1
    assert checked >= 90


def test_resize_returns_input_object_when_nothing_to_do(geom):
    src = geom["in_soft"]
    assert G.apply_resize(src, None, None, "AREA") is src           # xy_blend_processor.py:73
    assert G.apply_resize(src, src.shape[1], src.shape[0], "AREA") is src   # :82


def _masks(bits_fixture, z, N, win):
    y0, y1, x0, x1 = win
    cur = bits_fixture[z][y0:y1, x0:x1]
    pri = [bits_fixture[k][y0:y1, x0:x1] for k in range(max(0, z - N), z)]
    return cur, list(reversed(pri))


def test_anisotropic_enhanced(geom, slices):
    sets = json.loads(str(geom["aniso_sets"]))
    wins = json.loads(str(geom["wins"]))
    worst = 0
    n = 0
    for wn, win in wins.items():
        for z in (14, 30):
            cur, pri = _masks(slices, z, 4, win)
            for j, (xf, yf) in enumerate(sets):
                for F in (10.0, 40.0):
                    for fl in ("scipy", "numba"):
                        key = f"aniso_{wn}_{z:03d}_{j}_F{F:g}_{fl}"
                        if key not in geom.files:
                            continue
                        got = O.enhanced_fade(cur, pri, F, fl, "chamfer5", (xf, yf))
                        # bit-exact against the reference as it runs by default (cv2 with IPP); the IPP-less
                        # code path of cv2 differs from that by one grey level on ~1e-4 of the pixels
                        assert np.array_equal(got, geom[key]), key
                        d2 = np.abs(got.astype(int) - geom[key + "_noipp"].astype(int))
                        assert d2.max() <= 1 and (d2 > 0).mean() < 2e-3, key
                        worst = max(worst, float((d2 > 0).mean()))
                        n += 1
    assert n >= 30
