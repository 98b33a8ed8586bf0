import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "voxel-stack-blender_b200")
GOLD = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, PKG, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    os.environ.setdefault("VSB_NO_APP_CONFIG_WRITE", "1")


def _npz(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


@pytest.fixture(scope="session")
def slices():
    """The reference's 40 test slices (4000x2000, strictly {0,255}), unpacked from the bit fixture."""
    z = _npz("test_slices_bits.npz")
    H, W = [int(v) for v in z["shape"]]
    bits = z["bits"]
    return [np.unpackbits(b, axis=1)[:, :W] * np.uint8(255) for b in bits]


@pytest.fixture(scope="session")
def ref_stack():
    return _npz("ref_stack.npz")


@pytest.fixture(scope="session")
def ref_modes():
    z = _npz("ref_modes.npz")
    return z, json.loads(str(z["wins"]))


@pytest.fixture(scope="session")
def ref_dist():
    z = _npz("ref_dist.npz")
    return z, json.loads(str(z["wins"]))


@pytest.fixture(scope="session")
def ref_xy():
    return _npz("ref_xy.npz")


@pytest.fixture(scope="session")
def ref_luts():
    z = _npz("ref_luts.npz")
    return z, json.loads(str(z["specs"]))


@pytest.fixture(scope="session")
def ref_scenes():
    return _npz("ref_unit_scenes.npz")


@pytest.fixture(scope="session")
def lut_file(tmp_path_factory, ref_luts):
    """saved_luts/EXP(LUT)-upShifted.json re-materialised from the fixture (SURVEY finding 5)."""
    z, _ = ref_luts
    p = tmp_path_factory.mktemp("luts") / "EXP(LUT)-upShifted.json"
    p.write_text(json.dumps([int(v) for v in z["file_EXP(LUT)-upShifted"]]))
    return str(p)


def preset_ops(lut_path):
    """presets/Preset-Double-LUT.json's XY chain (first LUT re-pointed), as plain dicts."""
    return [
        {"type": "apply_lut", "lut_params": {"lut_source": "file", "fixed_lut_path": lut_path}},
        {"type": "bilateral_filter", "bilateral_d": 7, "bilateral_sigma_color": 60.0, "bilateral_sigma_space": 60.0},
        {"type": "gaussian_blur", "gaussian_ksize_x": 3, "gaussian_ksize_y": 3, "gaussian_sigma_x": 0.8,
         "gaussian_sigma_y": 0.0},
        {"type": "apply_lut", "lut_params": {"lut_source": "generated", "lut_generation_type": "exp",
                                             "input_min": 200, "input_max": 255, "output_min": 200,
                                             "output_max": 255, "exp_param": 2.0}},
    ]
