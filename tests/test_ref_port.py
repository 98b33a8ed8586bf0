"""CPU: oracle/ref_port.py (the cv2/numpy call sequence bench.py times as the CPU baseline) reproduces the
committed outputs of the unmodified reference, so the baseline it measures is the reference's arithmetic."""
import hashlib

import numpy as np
import pytest

import ref_port as RP
from conftest import preset_ops

pytestmark = pytest.mark.skipif(not RP.HAVE_CV2, reason="cv2 not importable on this box")


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_ref_port_matches_reference_stack(slices, ref_stack, lut_file):
    cfg = {"blending_mode": "fixed_fade", "receding_layers": 4, "use_fixed_fade_receding": True,
           "fixed_fade_distance_receding": 40.0, "xy_blend_pipeline": preset_ops(lut_file)}
    outs = RP.run_stack(slices[26:32], cfg, workers=3, halo=slices[22:26])
    for zi, o in zip(range(26, 32), outs):
        assert sha(o) == str(ref_stack["preset_sha1"][zi]), zi


def test_ref_port_modes_match_oracle(slices):
    import vsb_oracle as O
    y0, y1, x0, x1 = 250, 634, 300, 812
    cur = np.ascontiguousarray(O.threshold(slices[14][y0:y1, x0:x1]))
    pri = [np.ascontiguousarray(O.threshold(slices[k][y0:y1, x0:x1])) for k in (13, 12, 11, 10)]
    assert np.array_equal(RP.weighted_fade(cur, pri, [100, 75, 50, 25], [40.0, 30.0, 20.0, 10.0], 10.0),
                          O.weighted_fade(cur, pri, [100, 75, 50, 25], [40.0, 30.0, 20.0, 10.0], 10.0))
    for nb in (False, True):
        assert np.array_equal(RP.enhanced_fade(cur, pri, 40.0, nb), O.enhanced_fade(cur, pri, 40.0, "numba" if nb else "scipy"))
