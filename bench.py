#!/usr/bin/env python
"""
bench.py -- headline benchmark: 16K slice layers/sec through the Z-blend -> LUT -> XY hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ...] [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json metric, configs[3] per GPU): synthetic 16K slices 15120x6230 (seeded generator,
vsb_synth), FIXED_FADE with N=4 receding layers and F=40 px on the reference's chamfer5 metric, followed by the
Preset-Double-LUT XY chain [file LUT, bilateral d7 60/60, Gaussian 3x3 sigma 0.8, exp LUT 200-255].  One "step" = one
pass of the hot path over a batch of --batch layers per GPU.  The stack is Z-sharded: every rank owns its own
contiguous slab (weak scaling, no data-path collective); `--scaling strong` splits ONE fixed stack over the ranks
instead, each rank ingesting the 4 layers below its slab as a mask-only halo, and checks the concatenated per-layer
digests against a single-GPU pass of the whole stack.

  value : layers/s, whole job, inputs resident in HBM, CUDA events on the compute stream, max over ranks
  e2e   : the same metric through the host-buffer C-ABI (pinned host layers in, host layers out, H2D + D2H
          inside the timed region), next to `pcie`: a plain concurrent pinned H2D+D2H copy of the same bytes
  roofline / stages : per-kernel CUDA-event times of an extra profiled pass of the same K steps
  zblend_lut_only, cfg3, cfg2, edt_only, cliff, pipeline_png (N=1 only): the other BASELINE.json configurations,
          each with ms/layer, per-stage times, roofline fraction and the clocks sampled while it ran
  cpu_baseline : oracle/ref_port.py (the reference's cv2/numpy call sequence) on this box's host cores,
          bounded sample, rank 0 at N=1 only
  --impl reference : only that CPU arm, as its own JSON line
  --workload zlut|cfg2|cfg3|edt : make that configuration the main line instead of the headline
"""
import argparse
import contextlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "voxel-stack-blender_b200")
os.environ.setdefault("VSB_NO_APP_CONFIG_WRITE", "1")
sys.path.insert(0, PKG)

import numpy as np  # noqa: E402

W16, H16 = 15120, 6230
W12, H12 = 11520, 5120
METRIC = "16K slice layers/sec (Z-EDT blend+LUT+XY)"
UNIT = "layers/s"
TABLE_LAYERS = 8192     # the synthetic primitive table is always drawn for this many layers, so that every arm (device
                        # stack, CPU arm, any rank's slab) rasterises the SAME stack whatever part of it it walks
SEED = 1
Z_STEADY = 20           # first layer of rank 0's steady-state slab (the raft ends at z = 12: see the `cliff` block)

BIL = {"type": "bilateral_filter", "bilateral_d": 7, "bilateral_sigma_color": 60.0, "bilateral_sigma_space": 60.0}
GAUSS = {"type": "gaussian_blur", "gaussian_ksize_x": 3, "gaussian_ksize_y": 3, "gaussian_sigma_x": 0.8, "gaussian_sigma_y": 0.0}
EXPLUT = {"type": "apply_lut", "lut_params": {"lut_source": "generated", "lut_generation_type": "exp", "input_min": 200,
                                              "input_max": 255, "output_min": 200, "output_max": 255, "exp_param": 2.0}}


def preset_luts():
    """EXP(LUT)-upShifted table of the reference preset (committed fixture, SURVEY finding 5)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_luts.npz"))
    return np.array(z["file_EXP(LUT)-upShifted"], np.uint8)


def write_lut_file():
    path = os.path.join(tempfile.mkdtemp(prefix="vsb_bench_"), "EXP(LUT)-upShifted.json")
    with open(path, "w") as fh:
        json.dump([int(v) for v in preset_luts()], fh)
    return path


def workload_dict(name, lut_json_path):
    """BASELINE.json configurations as reference Config dictionaries."""
    filelut = {"type": "apply_lut", "lut_params": {"lut_source": "file", "fixed_lut_path": lut_json_path}}
    fixed = {"blending_mode": "fixed_fade", "receding_layers": 4, "use_fixed_fade_receding": True,
             "fixed_fade_distance_receding": 40.0}
    if name in ("preset", "cfg2"):          # cfg4 per GPU (16K) / cfg2 (12K): Z-blend + Preset-Double-LUT chain
        return dict(fixed, xy_blend_pipeline=[filelut, BIL, GAUSS, EXPLUT])
    if name == "zlut":                      # north_star's "fused EDT-blend-LUT pipeline": Z-blend + one file LUT
        return dict(fixed, xy_blend_pipeline=[filelut])
    if name == "cfg3":                      # Enhanced EDT (max 64 px) + bilateral + piecewise LUT
        return {"blending_mode": "enhanced_edt", "receding_layers": 4, "fixed_fade_distance_receding": 64.0,
                "use_numba_jit": True, "xy_blend_pipeline": [BIL, EXPLUT]}
    raise ValueError(name)


def workload_config(lut_json_path, name="preset"):
    """the same as a reference Config object (tools/*.py)"""
    from config import Config
    return Config.from_dict(workload_dict(name, lut_json_path))


WORKLOAD_TEXT = {
    "preset": "synthetic 16K slices 15120x6230, FIXED_FADE chamfer5 N=4 F=40 + Preset-Double-LUT XY chain "
              "(file LUT, bilateral d7 60/60, gaussian 3x3 s0.8, exp LUT 200-255)",
    "zlut": "synthetic 16K slices 15120x6230, FIXED_FADE chamfer5 N=4 F=40 + file LUT (no XY filter)",
    "cfg2": "synthetic 12K slices 11520x5120, FIXED_FADE chamfer5 N=4 F=40 + Preset-Double-LUT XY chain",
    "cfg3": "synthetic 16K slices 15120x6230, ENHANCED_EDT (Numba float64 arithmetic) N=4 max fade 64 px + bilateral d7 "
            "60/60 + piecewise exp LUT 200-255",
    "edt": "EDT-only: 16K binary masks (raft + dense supports + blobs), exact and chamfer5 distance maps",
}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed regions run (B200_PROFILING.md recipe); every sample is
    stamped on arrival so that each bench block can report the clocks seen during its own interval."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [c.strip() for c in line.split(",")]))
        except Exception:
            pass

    def window(self, t0, t1):
        rows = [r for ts, r in self.rows if t0 - 0.05 <= ts <= t1 + 0.15]
        if not rows and self.rows:                       # a block shorter than the sampling period: nearest sample
            rows = [min(self.rows, key=lambda e: abs(e[0] - 0.5 * (t0 + t1)))[1]]
        sm = sorted(int(float(r[1])) for r in rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(rows)}

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)


def triangle(L):
    """0,1,..,L-1,L-2,..,1 -- walking a short host slab up and down keeps consecutive layers one Z step apart."""
    return list(range(L)) + list(range(L - 2, 0, -1))


def host_workers():
    n = os.cpu_count() or 1
    try:
        import psutil
        n = min(n, max(1, int(psutil.virtual_memory().available / (5 << 30))))   # ~4-5 GB of float32 temporaries per 16K task
    except Exception:
        pass
    return max(1, min(n, 64))


def host_layers_numpy(W, H, z_first, count, threads):
    """`count` consecutive layers of the bench stack rasterised on the host (same table, same float32 arithmetic as
    the device rasteriser: tests/test_host_cpu.py), in parallel -- numpy releases the GIL."""
    import concurrent.futures
    import vsb_synth as S
    table, raft = S.make_table(W, H, TABLE_LAYERS, seed=SEED)
    with concurrent.futures.ThreadPoolExecutor(max(1, threads)) as pool:
        return list(pool.map(lambda z: S.raster_numpy(table, raft, z, W, H), range(z_first, z_first + count)))


def cpu_reference_arm(layers, halo, cfg_dict, min_seconds, workers):
    """layers/s of the reference's CPU call sequence over a bounded sample; also the single-worker / one cv2 thread
    figure SURVEY.md 8d asks for (on a quarter of the sample)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_port as RP
    try:
        import cv2
        cv2_threads = cv2.getNumThreads()
    except Exception:
        cv2, cv2_threads = None, None
    n, t0 = 0, time.perf_counter()
    while True:
        RP.run_stack(layers, cfg_dict, workers=workers, halo=halo, keep=False)
        n += len(layers)
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    single = None
    if cv2 is not None:
        few = layers[:max(1, min(2, len(layers)))]
        cv2.setNumThreads(1)
        t1 = time.perf_counter()
        RP.run_stack(few, cfg_dict, workers=1, halo=halo, keep=False)
        single = len(few) / (time.perf_counter() - t1)
        cv2.setNumThreads(cv2_threads if cv2_threads and cv2_threads > 0 else 0)
    in_flight = min(len(layers), 2 * workers)
    return n / dt, {"kind": "port", "cores": workers, "os_cpu_count": os.cpu_count(), "cv2_threads": cv2_threads,
                    "have_cv2": RP.HAVE_CV2, "tasks_in_flight": in_flight, "layers_per_pass": len(layers),
                    "single_worker_one_cv2_thread_layers_per_s": round(single, 4) if single else None,
                    "sample": f"{len(layers)} consecutive layers (+{len(halo)} halo) of the bench stack from z = {Z_STEADY} x "
                              f"{n // len(layers)} passes, {dt:.1f} s, ThreadPool({workers}) with 2x back-pressure over "
                              f"oracle/ref_port.py (cv2/numpy call sequence of the reference)"}


def pin_rank_to_gpu_numa(local_rank):
    """Run this rank's host threads (and therefore its pinned allocations, first touch) on the CPUs next to its GPU."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as fh:
            txt = fh.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        node = open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip()
        return {"pci": bus, "numa_node": int(node), "cpus": len(cpus)}
    except Exception as exc:
        return {"error": str(exc)[:80]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="layers per step per GPU")
    ap.add_argument("--resident", type=int, default=0,
                    help="device-resident layers per GPU (0 = as many as the run walks through, at most 640 = 61 GB at 16K)")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the zblend_lut_only / cfg3 / cfg2 / edt_only / cliff / pipeline_png blocks")
    ap.add_argument("--workload", default="preset", choices=["preset", "zlut", "cfg2", "cfg3", "edt"],
                    help="preset = headline (Z-blend + Preset-Double-LUT chain at 16K); zlut = Z-blend + file LUT only; "
                         "cfg2 = the preset chain at 12K; cfg3 = Enhanced EDT + bilateral + piecewise LUT; edt = EDT-only microbench")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--strong-layers", type=int, default=512, help="--scaling strong: layers of the one stack that is split")
    ap.add_argument("--png-layers", type=int, default=32)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W, H = (W12, H12) if args.workload == "cfg2" else (W16, H16)
    W, H = (args.width or W), (args.height or H)
    lut_path = write_lut_file()
    main_wl = "preset" if args.workload == "edt" else args.workload
    cfg_dict = workload_dict(main_wl, lut_path)
    wl = WORKLOAD_TEXT[args.workload]
    base_config = {"workload": wl if (args.width == 0 and args.height == 0) else wl + f" [size overridden: {W}x{H}]",
                   "width": W, "height": H, "batch_layers_per_gpu": args.batch, "resident_layers_per_gpu": 0,
                   "z_shard": f"{world} contiguous slabs, 4-layer mask halo, no collective", "metric_variant": "chamfer5",
                   "l2": f"each step streams {args.batch} x {W * H / 1e6:.0f} MB of input per GPU (>> 126 MB L2)",
                   "slab": f"layers z = {Z_STEADY} + 1000*rank upwards of one {TABLE_LAYERS}-layer synthetic stack (seed {SEED}), walked once in "
                           "print order (supports + model blobs); the layers right after the raft ends are timed in the `cliff` block"}

    # ------------------------------------------------------------------ reference arm (CPU only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        workers = host_workers()
        n_sample = max(8, 2 * workers)                        # the reference keeps 2 x workers tasks in flight (:169)
        layers = host_layers_numpy(W, H, Z_STEADY, 4 + n_sample, workers)
        vals = []
        info = {}
        for i in range(args.warmup + args.steps):
            v, info = cpu_reference_arm(layers[4:], layers[:4], cfg_dict, 0.0, workers)
            if i >= args.warmup:
                vals.append(v)
        v = float(np.mean(vals)) if vals else 0.0
        info["value"] = v
        info["unit"] = UNIT
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * n_sample / v if v else None,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                          "data": "synthetic", "config": base_config,
                          "cpu_baseline": info,
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import vsb_engine as E
    import vsb_native as N
    import vsb_params as P
    import vsb_synth as S
    import lut_manager
    from config import Config

    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = pin_rank_to_gpu_numa(local_rank)                  # pinned staging buffers land next to this rank's GPU
    if world > 1:
        # NCCL's init lines (rank / nranks) go to stderr; stdout stays the one JSON line
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):   # the image presets VERSION
            os.environ["NCCL_DEBUG"] = "INFO"
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist = None
    B, K, Wm = args.batch, args.steps, max(3, args.warmup)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    OUT_RING = 4

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def all_max(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def make_engine(cfgd, w, h):
        cfg = Config.from_dict(cfgd)
        ops = P.xy_ops_from_pipeline(cfg.xy_blend_pipeline, lambda op: lut_manager.lut_from_params(op.lut_params), w, h)
        return E.StackEngine(w, h, cfg, ops)

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)

    def stage_table(stages, alg_bytes):
        tot_ms = sum(s[1] for s in stages) or 1.0
        tbl = {}
        for name, sms, ln in stages:
            per = sms / max(1, ln if name != "ccl8_max" else ln // 3)
            tbl[name] = {"ms_per_layer": round(per, 5), "share": round(sms / tot_ms, 4),
                         "gbs_algorithmic": round(alg_bytes / (per * 1e-3) / 1e9, 1) if per > 0 else None}
        return tbl, tot_ms

    def run_block(name, cfgd, stack, first, n_fill, n_timed, w, h):
        """Device-resident per-layer time of one configuration on layers first .. first+n_fill+n_timed-1 of `stack`:
        n_fill layers fill the window (untimed), then n_timed layers timed as ONE batch with CUDA events on the engine's
        stream, then the same layers again with per-stage events."""
        t_begin = time.time()
        eng = make_engine(cfgd, w, h)
        outs = torch.zeros((OUT_RING, eng.out_H, E.pitch_for(eng.out_W)), dtype=torch.uint8, device="cuda")
        st = torch.cuda.ExternalStream(eng.stream)
        op, ostride = E.pitch_for(eng.out_W), eng.out_H * E.pitch_for(eng.out_W)

        def go(a, n):
            eng.process_device_batch(stack.ptr(a), stack.pitch, stack.layer_stride, outs.data_ptr(), op, ostride, OUT_RING, n)
        alg = 2.0 * w * h
        for rep in range(2):                                   # first repetition = warm-up (module load, table upload)
            eng.reset()
            go(first, n_fill)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            go(first + n_fill, n_timed)
            b.record(st)
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / n_timed
        eng.reset()
        go(first, n_fill)
        eng.profile(True)
        go(first + n_fill, n_timed)
        stages = eng.profile_read()
        eng.profile(False)
        eng.close()
        tbl, _ = stage_table(stages, alg)
        return {"workload": WORKLOAD_TEXT.get(name, name), "layers_timed": n_timed,
                "slab_layers": [first + n_fill, first + n_fill + n_timed - 1], "ms_per_layer": round(ms, 5),
                "layers_per_s_per_gpu": round(1000.0 / ms, 1), "hbm_gbs_algorithmic": round(alg / (ms * 1e-3) / 1e9, 1),
                "frac_of_measured_peak": round(alg / (ms * 1e-3) / 1e9 / peak, 4),
                "stage_ms": {k: v["ms_per_layer"] for k, v in tbl.items()},
                "stage_ms_note": "per-stage events of a serialised extra pass; ms_per_layer is the overlapped batch",
                "clocks": sampler.window(t_begin, time.time())}

    # ---------------------------------------------------------------- EDT-only microbench (BASELINE config 5)
    def edt_block():
        t_begin = time.time()
        lib = N.load()
        w, h = W16, H16
        stack5 = S.DeviceStack(w, h, 4, seed=2, z0=8, total_layers=TABLE_LAYERS)   # z = 8..11: raft + supports + blobs
        dist_t = torch.empty((h, w), dtype=torch.float32, device="cuda")
        d2 = torch.empty((h, w), dtype=torch.int32, device="cuda")
        g16 = torch.empty((h, w), dtype=torch.int16, device="cuda")
        bits = []
        for i in range(4):
            img = E.DevImage(h, w)
            img.t.copy_(stack5.layers[i])
            bits.append(E.pack(img))
        wpr = bits[0].wpr
        gmin = torch.empty(int(lib.vsb_coldist_scratch_elems(w, h, wpr)), dtype=torch.int16, device="cuda")
        T63 = E.table_dev(63)
        alg = 5.0 * w * h

        def timed(fn, n=8):
            for i in range(2):
                fn(i)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(n):
                fn(i)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n

        res = {}
        if hasattr(lib, "vsb_edt_exact"):
            scratch = torch.empty(int(lib.vsb_edt_exact_scratch_bytes(w, h)), dtype=torch.uint8, device="cuda")
            res["exact_unbounded"] = timed(lambda i: N.check(lib.vsb_edt_exact(bits[i % 4].ptr, wpr, w, h, scratch.data_ptr(),
                                                                                 d2.data_ptr(), dist_t.data_ptr(), w, 0)))
            res["exact_unbounded_dist_only"] = timed(lambda i: N.check(lib.vsb_edt_exact(bits[i % 4].ptr, wpr, w, h, scratch.data_ptr(),
                                                                                           None, dist_t.data_ptr(), w, 0)))
        res["exact_unbounded_r01_row_search"] = timed(lambda i: N.check(lib.vsb_dt_unbounded(
            bits[i % 4].ptr, wpr, w, h, 1, g16.data_ptr(), w, gmin.data_ptr(), dist_t.data_ptr(), d2.data_ptr(), w, 0)), n=4)
        res["chamfer5_unbounded"] = timed(lambda i: N.check(lib.vsb_dt_unbounded(
            bits[i % 4].ptr, wpr, w, h, 0, g16.data_ptr(), w, gmin.data_ptr(), dist_t.data_ptr(), None, w, 0)), n=4)
        res["chamfer5_windowed_R63"] = timed(lambda i: N.check(lib.vsb_dt_chamfer5(bits[i % 4].ptr, wpr, w, h, 63, T63.data_ptr(), 64.0,
                                                                                    dist_t.data_ptr(), w, 0)))
        res["exact_windowed_R63"] = timed(lambda i: N.check(lib.vsb_dt_exact_windowed(bits[i % 4].ptr, wpr, w, h, 63, 64.0,
                                                                                       dist_t.data_ptr(), w, 0)))
        out = {k: {"ms_per_image": round(v, 4), "images_per_s": round(1000 / v, 1), "gbs_algorithmic": round(alg / (v * 1e-3) / 1e9, 1),
                   "frac_of_measured_peak": round(alg / (v * 1e-3) / 1e9 / peak, 4)} for k, v in res.items()}
        out["algorithmic_bytes_per_image"] = alg
        out["white_fraction"] = round(float((stack5.layers[0][:, :w] > 127).float().mean()), 4)
        out["workload"] = WORKLOAD_TEXT["edt"] + " (z = 8..11 of seed 2); bytes = 1 B/px mask in + 4 B/px float32 out"
        try:
            import cv2
            m = stack5.layers[0][:, :w].contiguous().cpu().numpy()
            src = cv2.bitwise_not(cv2.threshold(m, 127, 255, cv2.THRESH_BINARY)[1])
            t0 = time.perf_counter()
            cv2.distanceTransform(src, cv2.DIST_L2, 5)
            t1 = time.perf_counter()
            cv2.distanceTransform(src, cv2.DIST_L2, cv2.DIST_MASK_PRECISE)
            t2 = time.perf_counter()
            out["cpu_cv2"] = {"mask5_ms": round(1000 * (t1 - t0), 1), "precise_ms": round(1000 * (t2 - t1), 1),
                              "threads": cv2.getNumThreads()}
        except Exception as exc:
            out["cpu_cv2"] = {"error": str(exc)[:100]}
        out["clocks"] = sampler.window(t_begin, time.time())
        return out

    if args.workload == "edt":
        blk = edt_block()
        best = blk.get("exact_unbounded", blk["exact_unbounded_r01_row_search"])
        sampler.stop()
        if rank == 0:
            print(json.dumps({"metric": "EDT-only 16K images/sec (exact Euclidean distance map)", "value": best["images_per_s"],
                              "unit": "images/s", "n_gpus": 1, "steps": 8, "warmup": 2, "ms_per_step": best["ms_per_image"],
                              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/f32",
                              "data": "synthetic", "config": {"workload": WORKLOAD_TEXT["edt"]}, "impl": "b200",
                              "clocks": blk["clocks"], "edt_only": blk,
                              "roofline": {"bound": "hbm", "kernel": "exact EDT (k_edt_col* + k_edt_rowdc)",
                                           "achieved": best["gbs_algorithmic"], "peak": peak, "unit": "GB/s",
                                           "frac": best["frac_of_measured_peak"], "traffic": None, "peak_source": peak_src}}))
        return 0

    # ---------------------------------------------------------------- strong scaling: one stack, Z-sharded with halo ingest
    if args.scaling == "strong":
        Ls, halo = args.strong_layers, 4
        plan = E.zshard_plan(Ls, world, halo)[rank]
        z0, z1, h0 = plan["z0"], plan["z1"], plan["halo0"]
        stack = S.DeviceStack(W, H, z1 - h0, seed=SEED, z0=Z_STEADY + h0, total_layers=TABLE_LAYERS)
        eng = make_engine(cfg_dict, W, H)
        outs = torch.zeros((z1 - z0, H, stack.pitch), dtype=torch.uint8, device="cuda")
        st = torch.cuda.ExternalStream(eng.stream)

        def shard_pass():
            eng.reset()
            for i in range(z0 - h0):                                          # mask-only halo: the layers below the slab
                eng.push_halo_device(stack.ptr(i), stack.pitch)
            eng.process_device_batch(stack.ptr(z0 - h0), stack.pitch, stack.layer_stride, outs.data_ptr(), stack.pitch,
                                     H * stack.pitch, z1 - z0, z1 - z0)
        for _ in range(max(1, min(Wm, 2))):
            shard_pass()
        barrier()
        t_begin = time.time()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(st)
        for _ in range(K):
            shard_pass()
        ev1.record(st)
        barrier()
        ms = all_max(ev0.elapsed_time(ev1))
        clocks = sampler.window(t_begin, time.time())

        gen = torch.Generator(device="cuda")
        gen.manual_seed(1234)
        mult = torch.randint(-(1 << 62), 1 << 62, (H * stack.pitch // 8,), dtype=torch.int64, device="cuda", generator=gen) | 1

        def digest(layer):                                                     # wrap-around multiply-sum over 8-byte words
            return int((layer.reshape(-1).view(torch.int64) * mult).sum().item())
        mine = torch.tensor([digest(outs[i]) for i in range(z1 - z0)], dtype=torch.int64, device="cuda")
        per = -(-Ls // world)
        padded = torch.zeros(per, dtype=torch.int64, device="cuda")
        padded[: z1 - z0] = mine
        if dist is not None:
            gathered = [torch.zeros_like(padded) for _ in range(world)]
            dist.all_gather(gathered, padded)
        else:
            gathered = [padded]
        check = None
        if rank == 0:
            got = torch.cat(gathered)[:Ls].cpu().tolist()
            # single-GPU pass of the whole stack, generated and processed chunk by chunk (outside the timed region)
            eng.reset()
            want, C = [], 32
            obuf = torch.zeros((C, H, stack.pitch), dtype=torch.uint8, device="cuda")
            for c0 in range(0, Ls, C):
                n = min(C, Ls - c0)
                chunk = S.DeviceStack(W, H, n, seed=SEED, z0=Z_STEADY + c0, total_layers=TABLE_LAYERS)
                eng.process_device_batch(chunk.ptr(0), chunk.pitch, chunk.layer_stride, obuf.data_ptr(), chunk.pitch,
                                         H * chunk.pitch, C, n)
                eng.sync()
                want += [digest(obuf[i]) for i in range(n)]
                del chunk
            bad = [i for i, (a, b) in enumerate(zip(got, want)) if a != b]
            check = {"layers": Ls, "mismatching_layers": len(bad), "first_bad": bad[:4], "ok": not bad,
                     "what": "per-layer 64-bit multiply-sum digests of the N ranks' slabs (all_gather) == a single-GPU pass of the whole stack"}
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": round(Ls * K / (ms / 1000.0), 2), "unit": UNIT, "n_gpus": world, "steps": K,
                              "warmup": Wm, "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": "strong",
                              "vs_baseline": None, "dtype": "u8", "data": "synthetic", "impl": "b200", "clocks": clocks,
                              "config": dict(base_config, strong_layers=Ls, halo_layers=halo,
                                             z_shard=f"one {Ls}-layer stack in {world} contiguous slabs, each rank ingests the {halo} layers below its slab "
                                                     "as a mask-only halo (vsb_stack_push_halo_device), no collective on the data path"),
                              "gpu_launches": None, "halo_check": check, "numa": numa}))
        eng.close()
        sampler.stop()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0 if (check is None or check["ok"]) else 1

    # ---------------------------------------------------------------- weak scaling (default): the headline line
    # The slab is walked upwards only, like a print job.  By default it is long enough for warm-up + timed steps + the
    # profiled pass without ever wrapping; a shorter slab wraps to its first layer with a window reset (the first N
    # layers after a wrap then see fewer priors, exactly like the first layers of a stack).
    cap = 640 if W * H > 60e6 else 1024
    L = args.resident if args.resident > 0 else min(cap, (Wm + 2 * K) * B)
    base_config["resident_layers_per_gpu"] = L
    z0 = Z_STEADY + rank * 1000            # this rank's slab of the stack; its first layers double as the halo
    stack = S.DeviceStack(W, H, L, seed=SEED, z0=z0, total_layers=TABLE_LAYERS)
    eng = make_engine(cfg_dict, W, H)
    outs = torch.zeros((OUT_RING, eng.out_H, E.pitch_for(eng.out_W)), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.ExternalStream(eng.stream)
    pos = [0]

    def run_step():
        """B layers, continuing the upward walk through the slab."""
        i = 0
        while i < B:
            if pos[0] >= L:
                eng.reset()
                pos[0] = 0
            run = min(B - i, L - pos[0])
            eng.process_device_batch(stack.ptr(pos[0]), stack.pitch, stack.layer_stride, outs.data_ptr(), stack.pitch,
                                     H * stack.pitch, OUT_RING, run)
            pos[0] += run
            i += run

    for _ in range(Wm):
        run_step()
    barrier()
    t_begin = time.time()
    launches0 = N.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(K):
        run_step()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = N.launch_count() - launches0
    time.sleep(0.15)
    clocks = sampler.window(t_begin, time.time())
    ms = all_max(ms)
    value = world * B * K / (ms / 1000.0)

    # ---- per-stage profile (extra pass, same K steps, one stream) -> roofline of the dominant kernel
    eng.profile(True)
    for _ in range(K):
        run_step()
    stages = eng.profile_read()
    eng.profile(False)
    alg_bytes = 2.0 * W * H                                   # one uint8 read + one uint8 write per layer
    stage_tbl, tot_ms = stage_table(stages, alg_bytes)
    dom = max(stages, key=lambda s: s[1]) if stages else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom[0]) if dom else None
    except Exception:
        pass
    roofline = None
    if dom:
        per = dom[1] / max(1, dom[2])
        ach = alg_bytes / (per * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom[0], "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": round(per, 5),
                    "pipeline_gbs": round(alg_bytes / (ms / (B * K) * 1e-3) / 1e9, 1),
                    "pipeline_frac": round(alg_bytes / (ms / (B * K) * 1e-3) / 1e9 / peak, 4),
                    "note": "achieved = 2*W*H bytes per layer / CUDA-event time of the dominant kernel (serialised profiled pass); "
                            "pipeline_* = the same bytes over the whole per-layer time of the timed region (pack pass of layer z+1 "
                            "overlapping the layer kernel of layer z on a second stream)"}

    # ---- the other BASELINE configurations, each as its own block (single-GPU run only)
    extras = {}
    if world == 1 and not args.no_extras:
        def guarded(name, fn):
            try:
                extras[name] = fn()
            except Exception as exc:  # the headline must not die on a side measurement
                extras[name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            torch.cuda.empty_cache()
        nfill, ntime = 8, min(L - 8, 96)
        # the blocks time the layers the headline's timed region starts with (the slab's layers after the warm-up steps)
        blk0 = max(0, min(Wm * B - nfill, L - nfill - ntime))
        if main_wl != "zlut":
            guarded("zblend_lut_only", lambda: run_block("zlut", workload_dict("zlut", lut_path), stack, blk0, nfill, ntime, W, H))
        if main_wl != "cfg3" and (W, H) == (W16, H16):
            guarded("cfg3", lambda: run_block("cfg3", workload_dict("cfg3", lut_path), stack, blk0, nfill, min(ntime, 32), W, H))
        if main_wl == "preset" and (W, H) == (W16, H16):
            def cliff():
                # the raft (60 % of the plate) ends at z = 12: layers 12..15 see it recede all at once over the dense
                # support grid.  Per-layer times of z = 6..19, headline configuration.
                t0 = time.time()
                cs = S.DeviceStack(W, H, 14, seed=SEED, z0=6, total_layers=TABLE_LAYERS)
                e2 = make_engine(cfg_dict, W, H)
                st2 = torch.cuda.ExternalStream(e2.stream)
                o2 = torch.zeros((eng.out_H, stack.pitch), dtype=torch.uint8, device="cuda")
                us = {}
                for rep in range(2):
                    e2.reset()
                    for i in range(14):
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record(st2)
                        e2.process_device(cs.ptr(i), cs.pitch, o2.data_ptr(), stack.pitch)
                        b.record(st2)
                        torch.cuda.synchronize()
                        us[6 + i] = round(a.elapsed_time(b) * 1000.0, 1)
                e2.close()
                post = [us[z] for z in (12, 13, 14, 15)]
                return {"workload": "headline configuration on z = 6..19 of the same stack; the raft ends at z = 12",
                        "us_per_layer": us, "post_raft_layers_ms": [round(v / 1000.0, 3) for v in post],
                        "post_raft_mean_ms": round(sum(post) / 4000.0, 3), "steady_ms": round(us[19] / 1000.0, 3),
                        "clocks": sampler.window(t0, time.time())}
            guarded("cliff", cliff)

            def cfg2():
                s12 = S.DeviceStack(W12, H12, 8 + 64, seed=SEED, z0=Z_STEADY, total_layers=TABLE_LAYERS)
                r = run_block("cfg2", workload_dict("cfg2", lut_path), s12, 0, 8, 64, W12, H12)
                del s12
                return r
            guarded("cfg2", cfg2)
            guarded("edt_only", edt_block)

    # ---- e2e: pinned host layers in, host layers out, through vsb_stack_submit_host / vsb_stack_wait
    e2e, pcie = None, None
    if not args.no_e2e:
        NH = min(L, 10)
        Be, Ke = min(B, 16), max(1, min(K, 3))
        hin = torch.empty((NH, H, stack.pitch), dtype=torch.uint8, pin_memory=True)
        hout = torch.empty((eng.slots + 1, H, stack.pitch), dtype=torch.uint8, pin_memory=True)
        hin.copy_(stack.layers[:NH])
        torch.cuda.synchronize()
        horder = triangle(NH)
        eng.reset()

        def e2e_pass(nlayers, start):
            tickets = []
            for j in range(nlayers):
                src = hin[horder[(start + j) % len(horder)]]
                dst = hout[j % hout.shape[0]]
                tickets.append(eng.submit(src.data_ptr(), stack.pitch, dst.data_ptr(), stack.pitch))
                if len(tickets) >= eng.slots:
                    eng.wait(tickets.pop(0))
            for tk in tickets:
                eng.wait(tk)

        e2e_pass(Be, 0)                                       # warm-up
        barrier()
        t0 = time.perf_counter()
        for k in range(Ke):
            e2e_pass(Be, (k + 1) * Be)
        eng.sync()
        dt = all_max(time.perf_counter() - t0)
        e2e = {"value": round(world * Be * Ke / dt, 2), "unit": UNIT, "h2d_bytes_per_step": Be * W * H,
               "d2h_bytes_per_step": Be * W * H, "layers_per_step": Be, "steps": Ke,
               "path": f"vsb_stack_submit_host/vsb_stack_wait, pinned host buffers, {eng.slots} staging slots, H2D | kernels | D2H streams"}
        # the ceiling of that path: the same bytes as plain concurrent pinned copies, all ranks at once, no kernels
        dbuf = torch.empty((2, H, stack.pitch), dtype=torch.uint8, device="cuda")
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        nbytes = H * stack.pitch

        def copy_pass(n):
            for j in range(n):
                with torch.cuda.stream(s_up):
                    dbuf[0].copy_(hin[j % NH], non_blocking=True)
                with torch.cuda.stream(s_dn):
                    hout[j % hout.shape[0]].copy_(dbuf[1], non_blocking=True)
            s_up.synchronize()
            s_dn.synchronize()
        copy_pass(4)
        barrier()
        t0 = time.perf_counter()
        copy_pass(Be * Ke)
        dtc = all_max(time.perf_counter() - t0)
        ceil_lps = world * Be * Ke / dtc
        pcie = {"concurrent_h2d_d2h_layers_per_s": round(ceil_lps, 1), "gbs_each_way_per_gpu": round(Be * Ke * nbytes / dtc / 1e9, 1),
                "e2e_fraction_of_copy_ceiling": round(e2e["value"] / ceil_lps, 3),
                "note": "plain pinned cudaMemcpyAsync up and down at once on two streams, all ranks concurrently, no kernels: "
                        "the bound of any host-buffer path on this box", "numa": numa}
        del dbuf

    os.sched_setaffinity(0, all_cpus)                        # the host-side legs below use every core again
    # ---- PNG-inclusive pipeline: ProcessingPipelineThread.run() on synthetic 16K PNGs
    if world == 1 and not args.no_extras and main_wl == "preset" and args.png_layers > 0:
        def png_block():
            import concurrent.futures
            import cv2
            from processing_pipeline import ProcessingPipelineThread
            t0w = time.time()
            n = min(args.png_layers, L)
            tmp = tempfile.mkdtemp(prefix="vsb_png_")
            ind, outd = os.path.join(tmp, "in"), os.path.join(tmp, "out")
            os.makedirs(ind)
            os.makedirs(outd)
            hostl = [stack.layers[i][:, :W].contiguous().cpu().numpy() for i in range(n)]
            workers = max(1, (os.cpu_count() or 2) - 1)             # the reference's default, config.py:27
            t0 = time.perf_counter()
            with concurrent.futures.ThreadPoolExecutor(workers) as pool:
                list(pool.map(lambda i: cv2.imwrite(os.path.join(ind, f"layer_{i:04d}.png"), hostl[i]), range(n)))
            t_write = time.perf_counter() - t0
            del hostl
            cfgp = Config.from_dict(dict(cfg_dict, input_mode="folder", input_folder=ind, output_folder=outd))
            th = ProcessingPipelineThread(cfgp, workers)
            errs = []
            th.error_signal.connect(errs.append)
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(sys.stderr):        # run() prints one line per written file, like the reference
                th.run()
            dtp = time.perf_counter() - t0
            nout = len([f for f in os.listdir(outd) if f.endswith(".png")])
            io = getattr(th, "io_seconds", {})
            shutil.rmtree(tmp, ignore_errors=True)
            return {"layers": n, "written": nout, "workers": workers, "layers_per_s": round(n / dtp, 2), "wall_s": round(dtp, 2),
                    "decode_s_per_layer_per_thread": round(io.get("decode", 0.0) / max(1, n), 3),
                    "encode_s_per_layer_per_thread": round(io.get("encode", 0.0) / max(1, n), 3),
                    "device_wait_s_total": round(io.get("device_wait", 0.0), 3),
                    "input_png_write_s": round(t_write, 2), "errors": [e[:200] for e in errs],
                    "reference_survey_box": "0.153 layers/s (1 worker) / 0.646 layers/s (7 workers) on an 8-vCPU box, SURVEY.md section 6",
                    "what": "ProcessingPipelineThread.run(): cv2.imread pool -> pinned buffers -> H2D | kernels | D2H -> cv2.imwrite pool; "
                            "codec seconds are summed over the worker threads", "clocks": sampler.window(t0w, time.time())}
        try:
            extras["pipeline_png"] = png_block()
        except Exception as exc:
            extras["pipeline_png"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- CPU baseline beside it (rank 0, single-GPU run only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        workers = host_workers()
        n_sample = max(8, min(2 * workers, L - 4))
        host_layers = [stack.layers[i][:, :W].contiguous().cpu().numpy() for i in range(4 + n_sample)]
        v, cpu = cpu_reference_arm(host_layers[4:], host_layers[:4], cfg_dict, args.cpu_seconds, workers)
        cpu["value"] = round(v, 4)
        cpu["unit"] = UNIT

    sampler.stop()
    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic", "config": base_config, "impl": "b200", "clocks": clocks,
                "e2e": e2e, "pcie": pcie, "gpu_launches": int(launches), "roofline": roofline, "stages": stage_tbl}
        line.update(extras)
        line["cpu_baseline"] = cpu
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
