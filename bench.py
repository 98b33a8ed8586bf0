#!/usr/bin/env python
"""
bench.py -- headline benchmark: 16K slice layers/sec through the Z-blend -> LUT -> XY hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json metric, configs[3] per GPU): synthetic 16K slices 15120x6230 (seeded generator,
vsb_synth), FIXED_FADE with N=4 receding layers and F=40 px on the reference's chamfer5 metric, followed by
the Preset-Double-LUT XY chain [file LUT, bilateral d7 60/60, Gaussian 3x3 sigma 0.8, exp LUT 200-255].
One "step" = one pass of the hot path over a batch of --batch layers per GPU.  The stack is Z-sharded: every
rank owns its own contiguous slab (weak scaling, no data-path collective).

  value : layers/s, whole job, inputs resident in HBM, CUDA events on the compute stream, max over ranks
  e2e   : the same metric through the host-buffer C-ABI (pinned host layers in, host layers out, H2D + D2H
          inside the timed region)
  roofline / stages : per-kernel CUDA-event times of an extra profiled pass of the same K steps
  cpu_baseline : oracle/ref_port.py (the reference's cv2/numpy call sequence) on this box's host cores,
          bounded sample, rank 0 at N=1 only
  --impl reference : only that CPU arm, as its own JSON line
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "voxel-stack-blender_b200")
os.environ.setdefault("VSB_NO_APP_CONFIG_WRITE", "1")
sys.path.insert(0, PKG)

import numpy as np  # noqa: E402

W16, H16 = 15120, 6230
METRIC = "16K slice layers/sec (Z-EDT blend+LUT+XY)"
UNIT = "layers/s"
WORKLOAD = ("synthetic 16K slices 15120x6230, FIXED_FADE chamfer5 N=4 F=40 + Preset-Double-LUT XY chain "
            "(file LUT, bilateral d7 60/60, gaussian 3x3 s0.8, exp LUT 200-255)")


def preset_luts():
    """EXP(LUT)-upShifted table of the reference preset (committed fixture, SURVEY finding 5)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_luts.npz"))
    return np.array(z["file_EXP(LUT)-upShifted"], np.uint8)


def zlut_config(lut_json_path):
    """the fused EDT-blend-LUT pipeline of north_star: Z-blend + one file LUT, no XY filter"""
    from config import Config
    return Config.from_dict({
        "blending_mode": "fixed_fade", "receding_layers": 4, "use_fixed_fade_receding": True,
        "fixed_fade_distance_receding": 40.0,
        "xy_blend_pipeline": [{"type": "apply_lut", "lut_params": {"lut_source": "file", "fixed_lut_path": lut_json_path}}]})


def workload_config(lut_json_path):
    from config import Config
    return Config.from_dict({
        "blending_mode": "fixed_fade", "receding_layers": 4, "use_fixed_fade_receding": True,
        "fixed_fade_distance_receding": 40.0,
        "xy_blend_pipeline": [
            {"type": "apply_lut", "lut_params": {"lut_source": "file", "fixed_lut_path": lut_json_path}},
            {"type": "bilateral_filter", "bilateral_d": 7, "bilateral_sigma_color": 60.0, "bilateral_sigma_space": 60.0},
            {"type": "gaussian_blur", "gaussian_ksize_x": 3, "gaussian_ksize_y": 3, "gaussian_sigma_x": 0.8,
             "gaussian_sigma_y": 0.0},
            {"type": "apply_lut", "lut_params": {"lut_source": "generated", "lut_generation_type": "exp",
                                                 "input_min": 200, "input_max": 255, "output_min": 200,
                                                 "output_max": 255, "exp_param": 2.0}}]})


def write_lut_file():
    import tempfile
    path = os.path.join(tempfile.mkdtemp(prefix="vsb_bench_"), "EXP(LUT)-upShifted.json")
    with open(path, "w") as fh:
        json.dump([int(v) for v in preset_luts()], fh)
    return path


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def triangle(L):
    """0,1,..,L-1,L-2,..,1 -- walking the slab up and down keeps consecutive layers one Z step apart."""
    return list(range(L)) + list(range(L - 2, 0, -1))


def cpu_reference_arm(layers, halo, cfg_dict, min_seconds, workers):
    """layers/s of the reference's CPU call sequence over a bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_port as RP
    try:
        import cv2
        cv2_threads = cv2.getNumThreads()
    except Exception:
        cv2_threads = None
    n, t0 = 0, time.perf_counter()
    while True:
        RP.run_stack(layers, cfg_dict, workers=workers, halo=halo, keep=False)
        n += len(layers)
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    return n / dt, {"kind": "port", "cores": workers, "cv2_threads": cv2_threads, "have_cv2": RP.HAVE_CV2,
                    "sample": f"{len(layers)} layers (+{len(halo)} halo) of the bench stack x {n // len(layers)} passes, "
                              f"{dt:.1f} s, ThreadPool({workers}) over oracle/ref_port.py (cv2/numpy call sequence of the reference)"}


def config_as_dict(cfg, lut_path):
    d = cfg.to_dict()
    return {"blending_mode": d["blending_mode"], "receding_layers": d["receding_layers"],
            "use_fixed_fade_receding": True, "fixed_fade_distance_receding": d["fixed_fade_distance_receding"],
            "xy_blend_pipeline": d["xy_blend_pipeline"]}


def host_workers():
    n = os.cpu_count() or 1
    try:
        import psutil
        n = min(n, max(1, int(psutil.virtual_memory().available / (4 << 30))))   # ~4 GB of temporaries per 16K task
    except Exception:
        pass
    return max(1, min(n, 32))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="layers per step per GPU")
    ap.add_argument("--resident", type=int, default=0,
                    help="device-resident layers per GPU (0 = as many as the run walks through, at most 640 = 61 GB at 16K)")
    ap.add_argument("--width", type=int, default=W16)
    ap.add_argument("--height", type=int, default=H16)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="preset", choices=["preset", "zlut"],
                    help="preset = headline (Z-blend + Preset-Double-LUT chain); zlut = Z-blend + file LUT only")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W, H = args.width, args.height
    lut_path = write_lut_file()
    cfg = workload_config(lut_path) if args.workload == 'preset' else zlut_config(lut_path)
    cfg_dict = config_as_dict(cfg, lut_path)
    wl = WORKLOAD if args.workload == "preset" else WORKLOAD.split(" + ")[0] + " + file LUT (no XY filter)"
    base_config = {"workload": wl if (W, H) == (W16, H16) else wl.replace("15120x6230", f"{W}x{H}"),
                   "width": W, "height": H, "batch_layers_per_gpu": args.batch, "resident_layers_per_gpu": 0,
                   "z_shard": f"{world} contiguous slabs, 4-layer mask halo, no collective", "metric_variant": "chamfer5",
                   "l2": f"each step streams {args.batch} x {W * H / 1e6:.0f} MB of input per GPU (>> 126 MB L2)",
                   "slab": "layers z = 20 + 1000*rank upwards, walked once in print order (steady state: supports + model blobs; "
                           "the raft and the layers right after it ends, z < 16, are not in the slab)"}

    # ------------------------------------------------------------------ reference arm (CPU only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        import vsb_synth as S
        n_sample = 8
        table, raft = S.make_table(W, H, 64, seed=1)
        layers = [S.raster_numpy(table, raft, 20 + z, W, H) for z in range(4 + n_sample)]
        workers = host_workers()
        vals = []
        for i in range(args.warmup + args.steps):
            v, info = cpu_reference_arm(layers[4:], layers[:4], cfg_dict, 0.0, workers)
            if i >= args.warmup:
                vals.append(v)
        v = float(np.mean(vals))
        info["value"] = v
        info["unit"] = UNIT
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * n_sample / v,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                          "data": "synthetic", "config": dict(base_config, batch_layers_per_gpu=n_sample),
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

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line (a VERSION banner would precede it)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B, K, Wm = args.batch, args.steps, max(3, args.warmup)
    # The slab is walked upwards only, like a print job.  By default it is long enough for warm-up + timed steps + the
    # profiled pass without ever wrapping; a shorter slab wraps to its first layer with a window reset (the first N
    # layers after a wrap then see fewer priors, exactly like the first layers of a stack).
    L = args.resident if args.resident > 0 else min(640, (Wm + 2 * K) * B)
    base_config["resident_layers_per_gpu"] = L
    # this rank's slab of an (world * 1000)-layer stack; the first layers of the slab double as the halo
    z0 = 20 + rank * 1000
    stack = S.DeviceStack(W, H, L, seed=1, z0=z0, total_layers=z0 + L)
    ops = P.xy_ops_from_pipeline(cfg.xy_blend_pipeline, lambda op: lut_manager.lut_from_params(op.lut_params))
    eng = E.StackEngine(W, H, cfg, ops)
    OUT_RING = 4
    outs = torch.empty((OUT_RING, H, stack.pitch), dtype=torch.uint8, device="cuda")
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

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(Wm):
        run_step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
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
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * K / (ms / 1000.0)

    # ---- per-stage profile (extra pass, same K steps) -> roofline of the dominant kernel
    eng.profile(True)
    for _ in range(K):
        run_step()
    stages = eng.profile_read()
    eng.profile(False)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    alg_bytes = 2.0 * W * H                                   # one uint8 read + one uint8 write per layer
    tot_ms = sum(s[1] for s in stages) or 1.0
    stage_tbl = {}
    for name, sms, ln in stages:
        per = sms / max(1, ln if name != "ccl8_max" else ln // 3)
        stage_tbl[name] = {"ms_per_layer": round(per, 5), "share": round(sms / tot_ms, 4),
                           "gbs_algorithmic": round(alg_bytes / (per * 1e-3) / 1e9, 1) if per > 0 else None}
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
                    "pipeline_gbs": round(alg_bytes * (B * K) / (tot_ms * 1e-3) / 1e9, 1),
                    "pipeline_frac": round(alg_bytes * (B * K) / (tot_ms * 1e-3) / 1e9 / peak, 4),
                    "note": "achieved = 2*W*H bytes per layer / CUDA-event time of the dominant kernel; "
                            "pipeline_* = the same bytes over the sum of all kernels of a layer"}

    # ---- secondary measurement: the fused EDT-blend-LUT pipeline alone (north_star's >=60 % target)
    zlut = None
    try:
        cfg2 = zlut_config(lut_path)
        ops2 = P.xy_ops_from_pipeline(cfg2.xy_blend_pipeline, lambda op: lut_manager.lut_from_params(op.lut_params))
        eng2 = E.StackEngine(W, H, cfg2, ops2)
        stream2 = torch.cuda.ExternalStream(eng2.stream)

        L2 = min(L, 4 * B)

        def run2(first, n):                                   # upward through layers first .. first+n-1
            eng2.process_device_batch(stack.ptr(first), stack.pitch, stack.layer_stride, outs.data_ptr(), stack.pitch,
                                      H * stack.pitch, OUT_RING, n)
        run2(0, min(8, L2))                                   # fills the window of prior masks
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream2)
        run2(min(8, L2), L2 - min(8, L2))
        b.record(stream2)
        torch.cuda.synchronize()
        ms2 = a.elapsed_time(b) / max(1, L2 - min(8, L2))
        eng2.reset()
        run2(0, min(8, L2))
        eng2.profile(True)
        run2(min(8, L2), L2 - min(8, L2))
        st2 = {n: round(m / max(1, c), 5) for n, m, c in eng2.profile_read()}
        eng2.close()
        zlut = {"workload": "FIXED_FADE chamfer5 N=4 F=40 + file LUT (no XY filter), same stack", "ms_per_layer": round(ms2, 5),
                "layers_per_s_per_gpu": round(1000.0 / ms2, 1), "hbm_gbs_algorithmic": round(alg_bytes / (ms2 * 1e-3) / 1e9, 1),
                "frac_of_measured_peak": round(alg_bytes / (ms2 * 1e-3) / 1e9 / peak, 4), "stage_ms": st2}
    except Exception as exc:  # the headline must not die on the side measurement
        zlut = {"error": str(exc)}

    # ---- e2e: pinned host layers in, host layers out, through vsb_stack_submit_host / vsb_stack_wait
    e2e = None
    if not args.no_e2e:
        NH = min(L, 10)
        Be, Ke = min(B, 16), max(1, min(K, 3))
        hin = torch.empty((NH, H, stack.pitch), dtype=torch.uint8, pin_memory=True)
        hout = torch.empty((eng.slots + 1, H, stack.pitch), dtype=torch.uint8, pin_memory=True)
        hin.copy_(stack.layers[:NH])
        torch.cuda.synchronize()
        horder = triangle(NH)

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
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": round(world * Be * Ke / dt, 2), "unit": UNIT, "h2d_bytes_per_step": Be * W * H,
               "d2h_bytes_per_step": Be * W * H, "layers_per_step": Be, "steps": Ke,
               "path": "vsb_stack_submit_host/vsb_stack_wait, pinned host buffers, 3-stream H2D|kernels|D2H pipeline"}

    # ---- CPU baseline beside it (rank 0, single-GPU run only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_sample = 8
        host_layers = [stack.layers[i][:, :W].contiguous().cpu().numpy() for i in range(4 + n_sample)]
        v, cpu = cpu_reference_arm(host_layers[4:], host_layers[:4], cfg_dict, args.cpu_seconds, host_workers())
        cpu["value"] = round(v, 4)
        cpu["unit"] = UNIT

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic", "config": base_config, "impl": "b200", "clocks": clocks,
                "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "stages": stage_tbl, "zblend_lut_only": zlut, "cpu_baseline": cpu}
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
