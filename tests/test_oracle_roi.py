"""Oracle (CPU) vs the reference's own ROI_FADE outputs (tests/golden/ref_roi.npz, written by
tests/golden/make_golden_roi.py from the unmodified reference)."""
import hashlib
import json
import os

import numpy as np
import pytest

import vsb_oracle as O
import vsb_oracle_roi as R

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def roi():
    return np.load(os.path.join(HERE, "golden", "ref_roi.npz"))


def test_identify_rois_matches_library_order(roi, slices):
    for z in (5, 14, 30):
        binary = O.threshold(slices[z])
        for ms in (100, 10):
            got = R.identify_rois(binary, ms)
            k = f"rois_{z:03d}_{ms}_"
            assert [r["area"] for r in got] == roi[k + "area"].tolist()
            assert [list(r["bbox"]) for r in got] == roi[k + "bbox"].tolist()
            assert [r["label"] for r in got] == roi[k + "label"].tolist()
            assert np.allclose(np.array([r["centroid"] for r in got]).reshape(-1, 2), roi[k + "centroid"], rtol=0, atol=1e-9)
            assert [int(r["mask"].sum()) for r in got] == roi[k + "masksum"].tolist()


@pytest.mark.parametrize("name", ["fix10", "fix40", "minmax", "fix7p5"])
def test_roi_stack_vs_reference(roi, slices, name):
    wn, fixed, F, N, rp = json.loads(str(roi["configs"]))[name]
    y0, y1, x0, x1 = json.loads(str(roi["wins"]))[wn]
    layers = [np.ascontiguousarray(s[y0:y1, x0:x1]) for s in slices]
    cfg = {"blending_mode": "roi_fade", "use_fixed_fade_receding": fixed, "fixed_fade_distance_receding": F,
           "receding_layers": N, "roi_params": rp}
    outs, states = R.run_stack_roi(layers, cfg)
    assert states == json.loads(str(roi[f"{name}_track"]))
    shas = [hashlib.sha1(np.ascontiguousarray(o).tobytes()).hexdigest() for o in outs]
    bad = [i for i, s in enumerate(shas) if s != str(roi[f"{name}_sha1"][i])]
    assert not bad, bad
