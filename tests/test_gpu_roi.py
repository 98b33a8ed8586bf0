"""GPU (-m gpu): ROI_FADE -- device components with statistics, the host tracker and the per-ROI fade kernel against
the outputs of the unmodified reference (tests/golden/ref_roi.npz) and the CPU oracle."""
import hashlib
import json
import os

import numpy as np
import pytest

import vsb_oracle as O
import vsb_oracle_roi as R

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def E():
    import vsb_engine
    vsb_engine.torch()
    return vsb_engine


@pytest.fixture(scope="module")
def roi():
    return np.load(os.path.join(HERE, "golden", "ref_roi.npz"))


def test_identify_rois_vs_reference(E, roi, slices):
    import processing_core as core
    for z in (5, 14, 30):
        binary = O.threshold(slices[z])
        for ms in (100, 10):
            got = core.identify_rois(binary, ms)
            k = f"rois_{z:03d}_{ms}_"
            assert [r["area"] for r in got] == roi[k + "area"].tolist()
            assert [list(r["bbox"]) for r in got] == roi[k + "bbox"].tolist()
            assert [r["label"] for r in got] == roi[k + "label"].tolist()
            assert np.allclose(np.array([r["centroid"] for r in got]).reshape(-1, 2), roi[k + "centroid"], rtol=0, atol=1e-9)
            assert [int(r["mask"].sum()) for r in got] == roi[k + "masksum"].tolist()


@pytest.mark.parametrize("shape,dens", [((37, 41), 0.3), ((64, 200), 0.5), ((129, 517), 0.62), ((5, 3), 0.5), ((1, 1), 1.0)])
def test_components_random_vs_oracle(E, shape, dens):
    rng = np.random.default_rng(shape[1])
    img = (rng.random(shape) < dens).astype(np.uint8) * 255
    import processing_core as core
    got = core.identify_rois(img, 1)
    want = R.identify_rois(img, 1)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a["area"] == b["area"] and tuple(a["bbox"]) == tuple(b["bbox"]) and a["label"] == b["label"]
        assert np.array_equal(a["mask"], b["mask"])


def _cfg(fixed, F, N, rp):
    from config import Config, ProcessingMode, RoiParameters
    c = Config()
    c.blending_mode = ProcessingMode.ROI_FADE
    c.use_fixed_fade_receding = fixed
    c.fixed_fade_distance_receding = F
    c.receding_layers = N
    c.roi_params = RoiParameters(**rp)
    return c


@pytest.mark.parametrize("name", ["fix10", "fix40", "minmax", "fix7p5"])
def test_roi_fade_functions_vs_reference(E, roi, slices, name):
    """identify_rois -> ROITracker -> process_z_blending -> merge_to_output, called like the reference's producer
    loop calls them: tracker state and every layer's pixels equal the reference's."""
    import collections
    import processing_core as core
    from roi_tracker import ROITracker
    wn, fixed, F, N, rp = json.loads(str(roi["configs"]))[name]
    y0, y1, x0, x1 = json.loads(str(roi["wins"]))[wn]
    cfg = _cfg(fixed, F, N, rp)
    tracker = ROITracker()
    window = collections.deque(maxlen=N)
    want_track = json.loads(str(roi[f"{name}_track"]))
    zs = range(40) if name in ("fix10", "minmax") else range(22)
    for z in zs:
        gray = np.ascontiguousarray(slices[z][y0:y1, x0:x1])
        binary = O.threshold(gray)
        rois = core.identify_rois(binary, cfg.roi_params.min_size)
        classified = tracker.update_and_classify(rois, z, cfg)
        state = [[int(r["id"]), str(r["classification"]), int(r["area"]), [int(v) for v in r["bbox"]]] for r in classified]
        assert state == want_track[z], (name, z)
        prior = core.find_prior_combined_white_mask(list(reversed(window)))
        grad = core.process_z_blending(binary, prior, cfg, classified)
        if f"{name}_grad_{z:03d}" in roi.files:
            ref = roi[f"{name}_grad_{z:03d}"]
            assert np.array_equal(grad, ref), (name, z, int((grad != ref).sum()))
        merged = core.merge_to_output(gray, grad)
        assert hashlib.sha1(np.ascontiguousarray(merged).tobytes()).hexdigest() == str(roi[f"{name}_sha1"][z]), (name, z)
        window.append(binary)


def test_roi_fade_many_components_per_window(E):
    """More eligible components around one receding pixel than the kernel tracks per sweep."""
    import processing_core as core
    from roi_tracker import ROITracker
    rng = np.random.default_rng(5)
    H, W = 96, 160
    prev = np.zeros((H, W), np.uint8); prev[8:88, 8:152] = 255
    cur = np.zeros((H, W), np.uint8)
    for y in range(10, 86, 4):
        for x in range(10, 150, 4):
            if rng.random() < 0.8:
                cur[y:y + 2, x:x + 2] = 255          # dozens of tiny components within every 21x21 window
    for fixed, F in ((True, 10.0), (False, 10.0), (True, 30.0)):
        cfg = _cfg(fixed, F, 1, {"min_size": 1, "enable_raft_support_handling": False})
        rois = core.identify_rois(cur, 1)
        classified = ROITracker().update_and_classify(rois, 0, cfg)
        got = core.process_z_blending(cur, prev, cfg, classified)
        orois = R.identify_rois(cur, 1)
        for r in orois:
            r["classification"] = "model"
        want = R.roi_fade(cur, prev, F, fixed, orois, False)
        assert np.array_equal(got, want), (fixed, F, int((got != want).sum()))


def test_pipeline_thread_roi_mode(E, roi, slices, tmp_path):
    """ProcessingPipelineThread.run() in ROI_FADE mode on PNG files (the device stack loop with the host tracker in
    the middle): same pixels as the reference's producer loop; layer_index comes from the file name."""
    import cv2
    import processing_pipeline as pp
    name = "fix40"
    wn, fixed, F, N, rp = json.loads(str(roi["configs"]))[name]
    y0, y1, x0, x1 = json.loads(str(roi["wins"]))[wn]
    src, dst = tmp_path / "in", tmp_path / "out"
    src.mkdir(); dst.mkdir()
    os.chdir(tmp_path)
    for z in range(20):
        cv2.imwrite(str(src / f"layer{z:03d}.png"), np.ascontiguousarray(slices[z][y0:y1, x0:x1]))
    cfg = _cfg(fixed, F, N, rp)
    cfg.input_mode = "folder"; cfg.input_folder = str(src); cfg.output_folder = str(dst)
    cfg.xy_blend_pipeline = []
    th = pp.ProcessingPipelineThread(cfg, max_workers=3)
    errors = []
    th.error_signal.connect(errors.append)
    th.run()
    assert not errors and not th.error_occurred, errors
    for z in range(20):
        got = cv2.imread(str(dst / f"layer{z:03d}.png"), cv2.IMREAD_GRAYSCALE)
        assert hashlib.sha1(np.ascontiguousarray(got).tobytes()).hexdigest() == str(roi[f"{name}_sha1"][z]), z
