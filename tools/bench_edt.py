"""EDT-only microbench (BASELINE.json config 5): 16K binary masks with dense support / raft geometry.
   distance maps: chamfer5 (cv2 DIST_L2 mask 5 analogue, windowed R=63), exact windowed (R=63), exact unbounded.
   Algorithmic bytes per image: 1 B/px mask in + 4 B/px float out = 5*W*H."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("VSB_NO_APP_CONFIG_WRITE", "1")
sys.path.insert(0, os.path.join(ROOT, "voxel-stack-blender_b200"))
import numpy as np, torch
import vsb_engine as E, vsb_native as N, vsb_synth as S

W, H = 15120, 6230
stack = S.DeviceStack(W, H, 4, seed=2, z0=8, total_layers=12)      # z = 8..11: raft + supports + blobs
lib = N.load()
t = torch
dist = t.empty((H, W), dtype=t.float32, device="cuda")
d2 = t.empty((H, W), dtype=t.int32, device="cuda")
col = t.empty((H, W), dtype=t.int16, device="cuda")
res = {}
alg = 5.0 * W * H
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, n=8):
    for _ in range(2):
        fn()
    t.cuda.synchronize()
    a, b = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    t.cuda.synchronize()
    return a.elapsed_time(b) / n


bits = []
for i in range(4):
    img = E.DevImage(H, W)
    img.t.copy_(stack.layers[i])
    bits.append(E.pack(img))
T = E.table_dev(63)
ms = timed(lambda i=0: N.check(lib.vsb_dt_chamfer5(bits[i % 4].ptr, bits[0].wpr, W, H, 63, T.data_ptr(), 64.0, dist.data_ptr(), W, 0)))
res["chamfer5_R63"] = ms
ms = timed(lambda i=0: N.check(lib.vsb_dt_exact_windowed(bits[i % 4].ptr, bits[0].wpr, W, H, 63, 64.0, dist.data_ptr(), W, 0)))
res["exact_R63"] = ms
ms = timed(lambda i=0: N.check(lib.vsb_dt_exact(bits[i % 4].ptr, bits[0].wpr, W, H, col.data_ptr(), d2.data_ptr(), dist.data_ptr(), W, 0)), n=3)
res["exact_unbounded_v1_bruteforce_rows"] = ms
g16 = t.empty((H, W), dtype=t.int16, device="cuda")
gmin = t.empty(int(lib.vsb_coldist_scratch_elems(W, H, bits[0].wpr)), dtype=t.int16, device="cuda")
ms = timed(lambda i=0: N.check(lib.vsb_dt_unbounded(bits[i % 4].ptr, bits[0].wpr, W, H, 1, g16.data_ptr(), W, gmin.data_ptr(), dist.data_ptr(), d2.data_ptr(), W, 0)), n=4)
res["exact_unbounded"] = ms
ms = timed(lambda i=0: N.check(lib.vsb_dt_unbounded(bits[i % 4].ptr, bits[0].wpr, W, H, 0, g16.data_ptr(), W, gmin.data_ptr(), dist.data_ptr(), None, W, 0)), n=4)
res["chamfer5_unbounded"] = ms
ms = timed(lambda i=0: N.check(lib.vsb_coldist(bits[i % 4].ptr, bits[0].wpr, W, H, g16.data_ptr(), W, gmin.data_ptr(), 0)), n=4)
res["column_pass_only"] = ms
out = {k: {"ms_per_image": round(v, 3), "images_per_s": round(1000 / v, 1), "gbs_algorithmic": round(alg / (v * 1e-3) / 1e9, 1),
           "frac_of_measured_peak": round(alg / (v * 1e-3) / 1e9 / peak, 4)} for k, v in res.items()}
try:
    import cv2
    m = stack.layers[0][:, :W].contiguous().cpu().numpy()
    src = cv2.bitwise_not(cv2.threshold(m, 127, 255, cv2.THRESH_BINARY)[1])
    t0 = time.perf_counter(); cv2.distanceTransform(src, cv2.DIST_L2, 5); t1 = time.perf_counter()
    cv2.distanceTransform(src, cv2.DIST_L2, cv2.DIST_MASK_PRECISE); t2 = time.perf_counter()
    out["cpu_cv2"] = {"mask5_ms": round(1000 * (t1 - t0), 1), "precise_ms": round(1000 * (t2 - t1), 1), "threads": cv2.getNumThreads()}
except Exception as exc:
    out["cpu_cv2"] = {"error": str(exc)}
print(json.dumps({"edt_only_16k": out, "white_fraction": float((stack.layers[0][:, :W] > 127).float().mean())}))
