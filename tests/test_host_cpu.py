"""CPU: host-side logic of the package -- no kernel is launched here.
  * the C-ABI library loads and exports every symbol include/vsb.h declares,
  * host parameter preparation (chamfer table, Gaussian taps, bilateral tables, LUTs, fade parameters)
    equals the oracle / golden values,
  * Config JSON round trip incl. the reference preset's layout, file ordering rules, Z-shard plan,
  * world_size-2 gloo run of the Z-shard partition logic."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import vsb_oracle as O
from conftest import ROOT, PKG


def test_library_exports_every_declared_symbol():
    import vsb_native as N
    lib = N.load()
    hdr = open(os.path.join(ROOT, "include", "vsb.h")).read()
    declared = set(re.findall(r"\b(vsb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"vsb_chamfer_table(R)"}
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert set(N.exported_symbols()) == declared
    assert lib.vsb_version() >= 100


def test_library_fails_loudly_when_missing(tmp_path, monkeypatch):
    import vsb_native as N
    monkeypatch.setattr(N, "_lib", None)
    monkeypatch.setattr(N, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(N.VsbError):
        N.load()


def test_product_never_imports_oracle():
    for fn in os.listdir(PKG):
        if fn.endswith(".py"):
            src = open(os.path.join(PKG, fn)).read()
            assert "vsb_oracle" not in src and "import oracle" not in src, fn
    for fn in os.listdir(os.path.join(PKG, "csrc")):
        assert "oracle/" not in open(os.path.join(PKG, "csrc", fn)).read().replace("oracle/c/vsb_oracle.c orc_chamfer5", "").replace(
            "oracle/c/vsb_oracle.c", ""), fn


def test_chamfer_table_equals_oracle_table():
    import vsb_native as N
    for R in (0, 1, 9, 39, 63, 200):
        assert np.array_equal(N.chamfer_table(R), O.chamfer_table(R))
    T = N.chamfer_table(63)
    assert T[1, 0] == np.float32(1.0) and T[1, 1] == np.float32(1.4) and T[2, 1] == np.float32(2.1969)
    assert (np.diff(T, axis=1)[np.triu_indices(63, 0)] > 0).all()      # monotone in the minor offset


def test_gaussian_taps_equal_oracle():
    import vsb_params as P
    assert list(P.gaussian_taps_q88(3, 0.8)) == [61, 134, 61]
    assert list(P.gaussian_taps_q88(3, 0.0)) == [64, 128, 64]
    for k in range(1, 33, 2):
        for s in (0.0, 0.3, 0.5, 0.8, 1.0, 1.7, 2.5, 4.0, 6.0, 11.0):
            a, b = P.gaussian_taps_q88(k, s), O.gaussian_kernel_q88(k, s)
            assert np.array_equal(a, b), (k, s)
            assert a.sum() == 256


def test_bilateral_tables():
    import vsb_params as P
    r, dy, dx, sw, cw = P.bilateral_tables(7, 60.0, 60.0)
    assert r == 3 and len(sw) == 29 == len(dy) == len(dx)              # 29 taps inside the radius-3 circle (d = 7)
    assert cw[0] == 1.0 and sw[len(sw) // 2] == 1.0
    r, *_ = P.bilateral_tables(0, 30.0, 2.0)
    assert r == 3
    r, *_ = P.bilateral_tables(0, 30.0, 0.2)
    assert r == 1


def test_lut_tables_equal_reference(ref_luts):
    import lut_manager
    from config import LutParameters
    z, specs = ref_luts
    for i, spec in enumerate(specs):
        got = lut_manager.lut_from_params(LutParameters(lut_source="generated", **spec))
        assert np.array_equal(got, z[f"lut_{i:03d}"]), spec
    # missing file and broken parameters degrade to identity (xy_blend_processor.py:120-130)
    ident = np.arange(256, dtype=np.uint8)
    assert np.array_equal(lut_manager.lut_from_params(LutParameters(lut_source="file", fixed_lut_path="/nope.json")), ident)
    with pytest.raises(TypeError):
        lut_manager.apply_z_lut(np.zeros((2, 2), np.float32), ident)
    with pytest.raises(ValueError):
        lut_manager.apply_z_lut(np.zeros((2, 2), np.uint8), ident[:10])


def test_lut_file_roundtrip(tmp_path):
    import lut_manager
    t = lut_manager.generate_gamma_lut(2.2, 0, 255, 0, 255)
    p = str(tmp_path / "l.json")
    lut_manager.save_lut(p, t)
    assert np.array_equal(lut_manager.load_lut(p), t)
    (tmp_path / "bad.json").write_text("[1,2,3]")
    with pytest.raises(IOError):
        lut_manager.load_lut(str(tmp_path / "bad.json"))
    with pytest.raises(FileNotFoundError):
        lut_manager.load_lut(str(tmp_path / "missing.json"))


PRESET = {
    "input_mode": "uvtools", "receding_layers": 4, "use_fixed_fade_receding": True,
    "fixed_fade_distance_receding": 40.0, "thread_count": 12, "some_future_key": 1,
    "xy_blend_pipeline": [
        {"type": "apply_lut", "lut_params": {"lut_source": "file", "fixed_lut_path": "C:/x.json", "sqrt_param": 1.0}},
        {"type": "bilateral_filter", "bilateral_d": 7, "bilateral_sigma_color": 60.0, "bilateral_sigma_space": 60.0},
        {"type": "gaussian_blur", "gaussian_ksize_x": 4, "gaussian_ksize_y": 3, "gaussian_sigma_x": 0.8},
        {"type": "apply_lut", "lut_params": {"lut_generation_type": "exp", "input_min": 255, "input_max": 200,
                                             "output_min": 200, "output_max": 255, "exp_param": 99.0}}]}


def test_config_semantics(tmp_path, capsys):
    from config import Config, ProcessingMode, XYBlendOperation, LutParameters
    c = Config()
    assert c.blending_mode == ProcessingMode.FIXED_FADE and c.receding_layers == 4          # config.py:182-183
    assert c.use_fixed_fade_receding is False and c.fixed_fade_distance_receding == 10.0     # :184-185
    assert c.manual_weights == [100, 75, 50, 25] and c.use_numba_jit is False
    p = Config.from_dict(PRESET)
    assert "Unrecognized config key 'some_future_key'" in capsys.readouterr().out
    assert p.blending_mode == ProcessingMode.FIXED_FADE                                       # SURVEY finding 5
    ops = p.xy_blend_pipeline
    assert [o.type for o in ops] == ["apply_lut", "bilateral_filter", "gaussian_blur", "apply_lut"]
    assert ops[2].gaussian_ksize_x == 5                                                       # forced odd (:127-129)
    lp = ops[3].lut_params
    assert (lp.input_min, lp.input_max) == (200, 255) and lp.exp_param == 10.0                # swapped / clamped
    path = str(tmp_path / "c.json")
    p.save(path)
    q = Config.load(path)
    assert q.to_dict() == p.to_dict()
    assert Config.from_dict({"blending_mode": "bogus"}).blending_mode == ProcessingMode.FIXED_FADE
    assert Config.from_dict({"blending_mode": "enhanced_edt"}).blending_mode == ProcessingMode.ENHANCED_EDT
    assert Config.from_dict({"debug_save": "true"}).debug_save is True
    assert XYBlendOperation(type="MEDIAN_BLUR", median_ksize=4).median_ksize == 5
    assert LutParameters(lut_source="weird").lut_source == "generated"


def test_zparams_from_config():
    import vsb_native as N
    import vsb_params as P
    from config import Config, ProcessingMode
    c = Config(); c.use_fixed_fade_receding = True; c.fixed_fade_distance_receding = 40.0
    z = P.zparams_from_config(c)
    assert (z.mode, z.reach, z.fade, z.den) == (N.Z_FIXED, 39, 40.0, 40.0)
    c.fixed_fade_distance_receding = 10.5
    assert P.zparams_from_config(c).reach == 10
    c.fixed_fade_distance_receding = 0.0
    z = P.zparams_from_config(c)
    assert (z.mode, z.const_val) == (N.Z_CONST, 255)
    c.fixed_fade_distance_receding = 200.0
    assert P.zparams_from_config(c).mode == N.Z_FIXED_FAR        # beyond the tile reach: unbounded path
    c = Config()
    assert P.zparams_from_config(c).mode == N.Z_MINMAX           # dataclass default = global min-max
    c.blending_mode = ProcessingMode.WEIGHTED_STACK
    c.manual_weights = [100, 0, 50]; c.fade_distances_receding = [20.0, 90.0]; c.fixed_fade_distance_receding = 12.0
    z = P.zparams_from_config(c)
    assert z.mode == N.Z_WEIGHTED and z.n_priors == 3 and z.reach == 19
    assert list(z.fades)[:3] == [20.0, 90.0, 12.0] and z.total_weight == 150.0
    c.fade_distances_receding = [20.0, 90.0, 300.0]
    z = P.zparams_from_config(c)
    assert z.mode == N.Z_WEIGHTED_FAR and z.reach == 0 and z.fade == 300.0   # a fade beyond the tile reach
    c.blending_mode = ProcessingMode.ENHANCED_EDT; c.fixed_fade_distance_receding = 100.0
    assert P.zparams_from_config(c).mode == N.Z_ENHANCED_FAR
    c.fixed_fade_distance_receding = 40.0
    c.anisotropic_params.enabled = True; c.anisotropic_params.x_factor = 2.0; c.anisotropic_params.y_factor = 0.05
    z = P.zparams_from_config(c, "chamfer5", 400, 200)
    assert (z.mode, z.aniso_w, z.aniso_h, z.aniso_reach) == (N.Z_ENHANCED, 800, 20, 42)   # y factor clamps to 0.1
    with pytest.raises(ValueError):
        P.zparams_from_config(c)                                   # the resized-mask size needs the layer size
    c.blending_mode = ProcessingMode.ROI_FADE
    with pytest.raises(P.UnsupportedOnDevice):
        P.zparams_from_config(c)                                   # ROI_FADE runs through RoiStackEngine instead


def test_xy_ops_geometry_descriptors():
    import vsb_native as N
    import vsb_params as P
    from config import XYBlendOperation
    ops = [XYBlendOperation(type="unsharp_mask", unsharp_amount=0.5, unsharp_threshold=3, unsharp_blur_ksize=4),
           XYBlendOperation(type="resize", resize_width=100, resample_mode="area"),
           XYBlendOperation(type="resize"),                                     # no target: dropped
           XYBlendOperation(type="resize", resize_width=100, resize_height=50),  # already that size: dropped
           XYBlendOperation(type="resize", resize_height=25, resample_mode="nonsense")]
    x = P.xy_ops_from_pipeline(ops, None, 400, 200)
    assert [o.type for o in x] == [N.OP_UNSHARP, N.OP_RESIZE, N.OP_RESIZE]
    assert (x[0].nkx, x[0].nky, x[0].unsharp_threshold) == (5, 5, 3) and abs(x[0].unsharp_amount - 0.5) < 1e-7
    assert (x[1].resize_w, x[1].resize_h, x[1].interp) == (100, 50, N.INTER_AREA)
    assert (x[2].resize_w, x[2].resize_h, x[2].interp) == (50, 25, N.INTER_LANCZOS4)
    with pytest.raises(ValueError):
        P.xy_ops_from_pipeline(ops, None)                                       # a resize needs the layer size


def test_file_ordering_rules(tmp_path):
    import processing_pipeline as pp
    for n in ("layer10.png", "layer2.png", "layer1.PNG", "notes.txt", "x.bmp", "slice_0003.tif", "layer7.jpg"):
        (tmp_path / n).write_bytes(b"")
    assert pp.list_layer_files(str(tmp_path)) == ["layer1.PNG", "layer2.png", "slice_0003.tif", "layer10.png", "x.bmp"]
    assert pp.list_layer_files(str(tmp_path), 2, 9) == ["layer2.png", "slice_0003.tif"]


def test_roi_tracker_equals_scalar_oracle_tracker():
    """ROITracker (vectorised, candidate-pair score matrix) against the oracle's scalar restatement of
    roi_tracker.py:3-109 on a drifting population of boxes: contained boxes (score 1.0), splits, births and deaths,
    zero-area boxes, one box spanning the plate.  Ids, classes and list order per layer are identical."""
    import roi_tracker
    import vsb_oracle_roi as R
    from config import Config, RoiParameters
    rng = np.random.default_rng(11)
    for a, b in [((0, 0, 10, 10), (5, 5, 10, 10)), ((0, 0, 0, 5), (0, 0, 3, 3)), ((2, 2, 4, 4), (0, 0, 20, 20)), ((0, 0, 3, 3), (3, 0, 3, 3))]:
        assert roi_tracker.calculate_iou(a, b) == R.overlap_over_min(a, b)
    for n, m, big in [(0, 4, 0), (4, 0, 0), (1, 1, 0), (120, 150, 0), (150, 120, 2)]:
        def boxes(k):
            out = [(int(rng.integers(0, 900)), int(rng.integers(0, 700)), int(rng.integers(0, 60)), int(rng.integers(0, 60))) for _ in range(k)]
            for i in range(min(big, k)):
                out[i] = (0, 0, 960, 760)
            return out
        c, q = boxes(n), boxes(m)
        want = np.array([[roi_tracker.calculate_iou(x, y) for y in q] for x in c]).reshape(n, m)
        assert np.array_equal(roi_tracker._score_matrix(c, q), want)

    rp = dict(min_size=1, enable_raft_support_handling=True, raft_layer_count=3, raft_min_size=900, support_max_size=400,
              support_max_layer=30, support_max_growth=1.5)
    cfg = Config()
    cfg.roi_params = RoiParameters(**rp)
    mine, theirs = roi_tracker.ROITracker(), R.Tracker()
    pop = [[int(rng.integers(0, 900)), int(rng.integers(0, 700)), int(rng.integers(4, 50)), int(rng.integers(4, 50))] for _ in range(90)]
    for z in range(12):
        nxt = []
        for x, y, w, h in pop:
            if rng.random() < 0.05:
                continue                                                  # dies
            x, y = x + int(rng.integers(-3, 4)), y + int(rng.integers(-3, 4))
            w, h = max(1, w + int(rng.integers(-2, 5))), max(1, h + int(rng.integers(-2, 5)))
            nxt.append([x, y, w, h])
            if rng.random() < 0.08:
                nxt.append([x + 1, y + 1, max(1, w // 2), max(1, h // 2)])   # a contained sibling: scores 1.0 twice
        for _ in range(int(rng.integers(0, 6))):
            nxt.append([int(rng.integers(0, 900)), int(rng.integers(0, 700)), int(rng.integers(4, 50)), int(rng.integers(4, 50))])
        if z == 5:
            nxt.append([0, 0, 960, 760])
        pop = nxt
        order = rng.permutation(len(pop))
        raw = [{"label": i + 1, "area": int(pop[k][2] * pop[k][3] * 0.7) + 1, "bbox": tuple(pop[k])} for i, k in enumerate(order)]
        got = mine.update_and_classify([dict(r) for r in raw], z, cfg)
        want = theirs.update([dict(r) for r in raw], z, rp)
        assert [(r["id"], r["classification"], r["bbox"]) for r in got] == [(r["id"], r["classification"], r["bbox"]) for r in want], z
        assert not np.any(mine._neg.view(np.uint64) != np.float64(-0.0).view(np.uint64))   # the scratch matrix is left clear
    assert mine.next_roi_id == theirs.next_id


def test_zshard_plan():
    import vsb_engine as E
    plan = E.zshard_plan(2000, 8, 4)
    assert [p["z1"] - p["z0"] for p in plan] == [250] * 8
    assert plan[0]["halo0"] == 0 and plan[3]["halo0"] == 746 and plan[3]["z0"] == 750
    plan = E.zshard_plan(10, 4, 4)
    assert [(p["z0"], p["z1"]) for p in plan] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert sum(p["z1"] - p["z0"] for p in plan) == 10


def test_synth_table_is_seeded_and_binary():
    import vsb_synth as S
    a, ra = S.make_table(640, 480, 30, seed=1, n_blobs=8, grid=90, r_range=(15, 60), r_turn=90)
    b, _ = S.make_table(640, 480, 30, seed=1, n_blobs=8, grid=90, r_range=(15, 60), r_turn=90)
    c, _ = S.make_table(640, 480, 30, seed=2, n_blobs=8, grid=90, r_range=(15, 60), r_turn=90)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    imgs = [S.raster_numpy(a, ra, z, 640, 480) for z in (0, 11, 12, 20)]
    assert all(set(np.unique(i)) <= {0, 255} for i in imgs)
    assert (imgs[1] > 0).mean() > (imgs[2] > 0).mean() + 0.2        # the raft disappears at z = 12
    rec = (imgs[2] > 0) & ~(imgs[3] > 0)
    assert rec.any()


GLOO_WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
os.environ["VSB_NO_APP_CONFIG_WRITE"] = "1"
import numpy as np, torch, torch.distributed as dist
import vsb_engine as E, vsb_synth as S, vsb_oracle as O
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[3]}", rank=int(sys.argv[4]), world_size=2)
rank = dist.get_rank()
W, H, L, N = 320, 200, 12, 3
table, raft = S.make_table(W, H, L, seed=3, n_blobs=6, grid=70, raft=False, r_range=(12, 50), r_turn=70)
layers = [S.raster_numpy(table, raft, z, W, H) for z in range(L)]
cfg = {"blending_mode": "fixed_fade", "receding_layers": N, "use_fixed_fade_receding": True, "fixed_fade_distance_receding": 8.0}
me = E.zshard_plan(L, 2, N)[rank]
outs = O.run_stack(layers[me["z0"]:me["z1"]], cfg, halo=layers[me["halo0"]:me["z0"]])
digest = torch.tensor([int(np.sum([o.astype(np.int64).sum() for o in outs])), me["z1"] - me["z0"]])
gathered = [torch.zeros_like(digest) for _ in range(2)]
dist.all_gather(gathered, digest)
if rank == 0:
    whole = O.run_stack(layers, cfg)
    ok = int(sum(int(g[0]) for g in gathered)) == int(np.sum([o.astype(np.int64).sum() for o in whole])) and sum(int(g[1]) for g in gathered) == L
    # slab outputs are identical to the corresponding layers of the single-process run
    ok = ok and all(np.array_equal(a, b) for a, b in zip(outs, whole[me["z0"]:me["z1"]]))
    print("GLOO_OK" if ok else "GLOO_MISMATCH")
dist.barrier(); dist.destroy_process_group()
'''


def test_zshard_partition_world2_gloo(tmp_path):
    """Z-slabs with an N-layer halo reproduce the single-process stack exactly (checked with the CPU oracle
    over a 2-rank gloo group; the device engine takes the same plan in test_gpu_parity.py)."""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), PKG, os.path.join(ROOT, "oracle"), port, str(r)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GLOO_OK" in outs[0], outs
