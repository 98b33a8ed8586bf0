"""GPU (-m gpu): BASELINE.json's full 16K size (15120x6230).  The CPU oracle cannot run whole 16K stacks in test time,
so parity at this size is checked through size-independent properties plus oracle spot checks on crops:
  * the fused path and the unfused kernel chain give identical layers (two independent implementations),
  * a Z-slab with its N-layer halo reproduces the whole-stack run (checksum of per-layer checksums),
  * re-running is bit-identical (no data race / ordering dependence),
  * crops (with a margin wider than reach + XY halo) equal the oracle run on the same crops."""
import hashlib
import os

import numpy as np
import pytest

import vsb_oracle as O
from conftest import preset_ops

pytestmark = pytest.mark.gpu
W, H, L, N, F = 15120, 6230, 7, 4, 40.0


@pytest.fixture(scope="module")
def env(lut_file):
    import vsb_engine as E
    import vsb_synth as S
    from config import Config
    E.torch()
    stack = S.DeviceStack(W, H, L, seed=0, z0=9, total_layers=9 + L)       # z = 9..15: the raft disappears at z = 12
    cfg = Config.from_dict({"blending_mode": "fixed_fade", "receding_layers": N, "use_fixed_fade_receding": True,
                            "fixed_fade_distance_receding": F, "xy_blend_pipeline": preset_ops(lut_file)})
    return E, stack, cfg


def make_engine(E, cfg):
    import lut_manager, vsb_params as P
    ops = P.xy_ops_from_pipeline(cfg.xy_blend_pipeline, lambda op: lut_manager.lut_from_params(op.lut_params))
    return E.StackEngine(W, H, cfg, ops)


def run(E, stack, cfg, first=0, halo=0):
    t = E.torch()
    eng = make_engine(E, cfg)
    out = t.empty((L - first, H, stack.pitch), dtype=t.uint8, device="cuda")
    scratch = t.empty((1, H, stack.pitch), dtype=t.uint8, device="cuda")
    if halo:
        eng.process_device_batch(stack.ptr(first - halo), stack.pitch, stack.layer_stride, scratch.data_ptr(), stack.pitch,
                                 H * stack.pitch, 1, halo)      # outputs of halo layers are discarded
    eng.process_device_batch(stack.ptr(first), stack.pitch, stack.layer_stride, out.data_ptr(), stack.pitch, H * stack.pitch,
                             L - first, L - first)
    eng.sync()
    eng.close()
    return out[:, :, :W]


def digest(layer):
    return hashlib.sha1(layer.contiguous().cpu().numpy().tobytes()).hexdigest()


def test_fullsize_properties(env, monkeypatch):
    E, stack, cfg = env
    a = run(E, stack, cfg)
    da = [digest(a[i]) for i in range(L)]
    assert len(set(da)) == L
    b = run(E, stack, cfg)
    assert [digest(b[i]) for i in range(L)] == da                          # idempotent / race free
    del b
    monkeypatch.setenv("VSB_NO_FUSED", "1")                                # the separate kernels (K1, K3, K4 chain)
    c = run(E, stack, cfg)
    monkeypatch.delenv("VSB_NO_FUSED")
    assert [digest(c[i]) for i in range(L)] == da
    del c
    s = run(E, stack, cfg, first=4, halo=N)                                # Z-slab [4, L) + 4-layer halo
    assert [digest(s[i]) for i in range(L - 4)] == da[4:]
    gray = stack.layers[:, :, :W]
    assert bool((a >= 0).all())
    # merge keeps every white input pixel at the LUT image of 255 after the chain's LUTs only if it is interior;
    # the cheap global invariant: the output is 0 wherever the input and all N priors are 0 far from any edge --
    # checked exactly by the crop comparison below.
    wins = [(0, 0), (2900, 7000), (H - 700, W - 900), (1200, 40), (40, 9000)]
    m = 64
    ocfg = {"blending_mode": "fixed_fade", "receding_layers": N, "use_fixed_fade_receding": True,
            "fixed_fade_distance_receding": F, "xy_blend_pipeline": [
                {k: (dict(v.__dict__) if hasattr(v, "__dict__") else v) for k, v in op.__dict__.items()} for op in cfg.xy_blend_pipeline]}
    for (y0, x0) in wins:
        y1, x1 = min(H, y0 + 640), min(W, x0 + 832)
        crops = [gray[i, y0:y1, x0:x1].contiguous().cpu().numpy() for i in range(L)]
        want = O.run_stack(crops, ocfg)
        for zi in (4, 5, 6):
            got = a[zi, y0:y1, x0:x1].contiguous().cpu().numpy()
            ya, yb = (0 if y0 == 0 else m), (y1 - y0 if y1 == H else y1 - y0 - m)   # the crop's own borders are not the image's
            xa, xb = (0 if x0 == 0 else m), (x1 - x0 if x1 == W else x1 - x0 - m)
            d = np.abs(got[ya:yb, xa:xb].astype(int) - want[zi][ya:yb, xa:xb].astype(int))
            assert d.max() == 0, ((y0, x0), zi, int((d > 0).sum()), int(d.max()))


def test_fullsize_lut_only_patch_path(env, monkeypatch, lut_file):
    """BASELINE config 1's shape at 16K (fade + one LUT, no XY filter): the pack-with-LUT + patch launches equal the
    single fused kernel (VSB_NO_PATCH) and the unfused chain (VSB_NO_FUSED), layer by layer, over the raft cliff."""
    from config import Config
    E, stack, _ = env
    cfg = Config.from_dict({"blending_mode": "fixed_fade", "receding_layers": N, "use_fixed_fade_receding": True,
                            "fixed_fade_distance_receding": F,
                            "xy_blend_pipeline": [{"type": "apply_lut", "lut_params": {"lut_source": "file", "fixed_lut_path": lut_file}}]})
    a = run(E, stack, cfg)
    da = [digest(a[i]) for i in range(L)]
    del a
    monkeypatch.setenv("VSB_NO_PATCH", "1")
    b = run(E, stack, cfg)
    monkeypatch.delenv("VSB_NO_PATCH")
    assert [digest(b[i]) for i in range(L)] == da
    del b
    monkeypatch.setenv("VSB_NO_FUSED", "1")
    c = run(E, stack, cfg)
    monkeypatch.delenv("VSB_NO_FUSED")
    assert [digest(c[i]) for i in range(L)] == da


def host_mask(stack, i):
    return np.ascontiguousarray(O.threshold(stack.layers[i][:, :W].contiguous().cpu().numpy()))


def test_fullsize_exact_edt(env):
    """BASELINE config 5 at its full size: vsb_edt_exact on whole 16K masks == the oracle's Felzenszwalb transform (integer
    d2, and sqrtf of it), with the raft (z = 10: short distances) and right after it (z = 13: distances of hundreds of pixels)."""
    E, stack, _ = env
    for zi in (1, 4):
        cur = host_mask(stack, zi)
        d2o = O.edt_exact_sq(cur)
        d2, dist = E.edt_exact_np(cur)
        assert np.array_equal(d2, d2o), (zi, int((d2 != d2o).sum()))
        assert np.array_equal(dist, np.sqrt(d2o.astype(np.float32)))
        del d2, dist, d2o


@pytest.mark.parametrize("flavour", ["scipy", "numba"])
def test_fullsize_enhanced_edt_cliff(env, flavour):
    """BASELINE config 3's Z-blend (Enhanced EDT, max 64 px) on the whole 16K layer right after the raft ends -- the global
    union-find over all 23,205 tiles, components of tens of millions of pixels, the atomicMax of float bits -- against the
    oracle on the whole image, both arithmetic flavours (processing_core.py:287-303 float32, :305-336 float64)."""
    import processing_core as core
    from test_gpu_parity import make_cfg
    E, stack, _ = env
    cur = host_mask(stack, 3)                                  # z = 12: the first layer without the raft
    pri = [host_mask(stack, 2), host_mask(stack, 1)]           # z = 11, 10
    cfg = make_cfg(blending_mode="enhanced_edt", fixed_fade_distance_receding=64.0, use_numba_jit=(flavour == "numba"))
    got = core.process_z_blending(cur, pri, cfg, [])
    want = O.enhanced_fade(cur, pri, 64.0, flavour)
    assert (want > 0).mean() > 0.2                              # the cliff: a large share of the plate recedes
    assert np.array_equal(got, want), int((got != want).sum())


def test_fullsize_cfg3_fused_equals_unfused(env, monkeypatch):
    """BASELINE config 3 as a stack (Enhanced EDT F = 64, Numba arithmetic + bilateral + piecewise LUT) over the cliff: the
    fused layer kernel and the separate kernels give identical layers, and crops equal the oracle."""
    from config import Config
    E, stack, _ = env
    explut = {"type": "apply_lut", "lut_params": {"lut_source": "generated", "lut_generation_type": "exp", "input_min": 200,
                                                  "input_max": 255, "output_min": 200, "output_max": 255, "exp_param": 2.0}}
    bil = {"type": "bilateral_filter", "bilateral_d": 7, "bilateral_sigma_color": 60.0, "bilateral_sigma_space": 60.0}
    d = {"blending_mode": "enhanced_edt", "receding_layers": N, "fixed_fade_distance_receding": 64.0, "use_numba_jit": True,
         "xy_blend_pipeline": [bil, explut]}
    cfg = Config.from_dict(d)
    a = run(E, stack, cfg)
    da = [digest(a[i]) for i in range(L)]
    monkeypatch.setenv("VSB_NO_FUSED", "1")
    b = run(E, stack, cfg)
    monkeypatch.delenv("VSB_NO_FUSED")
    assert [digest(b[i]) for i in range(L)] == da
    del b
    gray = stack.layers[:, :, :W]
    import scipy.ndimage as ndi
    m = 80
    for (y0, x0) in ((40, 9000), (5560, 2000)):                 # the margins outside the raft: support cones and blob rims
        y1, x1 = y0 + 640, x0 + 832
        crops = [gray[i, y0:y1, x0:x1].contiguous().cpu().numpy() for i in range(L)]
        want = O.run_stack(crops, d)
        for zi in (2, 5):
            got = a[zi, y0:y1, x0:x1].contiguous().cpu().numpy()
            rec = O.combine_priors([O.threshold(c) for c in crops[max(0, zi - N):zi]])
            rec = rec & ~O.threshold(crops[zi])
            n, labels = O.ccl8(rec)
            edge = np.unique(np.concatenate([labels[0], labels[-1], labels[:, 0], labels[:, -1]]))
            ok = ~np.isin(labels, edge[edge > 0])               # components cut by the crop have another maximum there
            ok[:m] = ok[-m:] = False
            ok[:, :m] = ok[:, -m:] = False
            ok = ndi.binary_erosion(ok, iterations=4)           # the bilateral window of a compared pixel is comparable too
            diff = np.abs(got.astype(int) - want[zi].astype(int))[ok]
            assert ok.mean() > 0.3 and diff.max() == 0, ((y0, x0), zi, float(ok.mean()), int((diff > 0).sum()))


def test_fullsize_cfg2_12k_crops(lut_file):
    """BASELINE config 2's size (11520x5120): the preset chain on a 12K stack, crops against the oracle."""
    import vsb_engine as E
    import vsb_synth as S
    import lut_manager, vsb_params as P
    from config import Config
    t = E.torch()
    W2, H2, L2 = 11520, 5120, 6
    st = S.DeviceStack(W2, H2, L2, seed=3, z0=20, total_layers=64)
    d = {"blending_mode": "fixed_fade", "receding_layers": N, "use_fixed_fade_receding": True, "fixed_fade_distance_receding": F,
         "xy_blend_pipeline": preset_ops(lut_file)}
    cfg = Config.from_dict(d)
    ops = P.xy_ops_from_pipeline(cfg.xy_blend_pipeline, lambda op: lut_manager.lut_from_params(op.lut_params))
    eng = E.StackEngine(W2, H2, cfg, ops)
    out = t.empty((L2, H2, st.pitch), dtype=t.uint8, device="cuda")
    eng.process_device_batch(st.ptr(0), st.pitch, st.layer_stride, out.data_ptr(), st.pitch, H2 * st.pitch, L2, L2)
    eng.sync()
    eng.close()
    m = 64
    for (y0, x0) in ((0, 0), (2200, 5000), (H2 - 640, W2 - 832)):
        y1, x1 = min(H2, y0 + 640), min(W2, x0 + 832)
        crops = [st.layers[i, y0:y1, x0:x1].contiguous().cpu().numpy() for i in range(L2)]
        want = O.run_stack(crops, d)
        for zi in (4, 5):
            got = out[zi, y0:y1, x0:x1].contiguous().cpu().numpy()
            ya, yb = (0 if y0 == 0 else m), (y1 - y0 if y1 == H2 else y1 - y0 - m)
            xa, xb = (0 if x0 == 0 else m), (x1 - x0 if x1 == W2 else x1 - x0 - m)
            dd = np.abs(got[ya:yb, xa:xb].astype(int) - want[zi][ya:yb, xa:xb].astype(int))
            assert dd.max() == 0, ((y0, x0), zi, int((dd > 0).sum()))
