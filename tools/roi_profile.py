"""Where one ROI_FADE layer at 16K spends its time (host tracker in the loop): per-stage wall clock with a device
sync after each stage."""
import os, sys, time
sys.path.insert(0, '/root/repo'); os.environ["VSB_NO_APP_CONFIG_WRITE"] = "1"
sys.path.insert(0, '/root/repo/voxel-stack-blender_b200')
import numpy as np, torch
import vsb_engine as E, vsb_synth as S
from config import Config
W, H = 15120, 6230
stack = S.DeviceStack(W, H, 8, seed=1, z0=20, total_layers=28)
cfg = Config.from_dict({"blending_mode": "roi_fade", "receding_layers": 4, "use_fixed_fade_receding": True, "fixed_fade_distance_receding": 20.0,
                        "roi_params": {"min_size": 100, "enable_raft_support_handling": True, "support_max_size": 2000}})
eng = E.RoiStackEngine(W, H, cfg)
host = [stack.layers[z][:, :W].cpu().numpy() for z in range(8)]
for z in range(4):
    eng.process(host[z], 20 + z)
acc = {}
def lap(name, t0):
    torch.cuda.synchronize(); t1 = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t1 - t0); return t1
for z in range(4, 8):
    t = time.perf_counter()
    img = E.DevImage.from_numpy(host[z]); t = lap("h2d", t)
    cur = E.pack(img); t = lap("pack", t)
    comp = E.Components(cur); t = lap("components+stats d2h", t)
    rois = comp.rois(100); t = lap("rois() dicts", t)
    cl = eng.tracker.update_and_classify(rois, 20 + z, cfg); t = lap("tracker", t)
    el = np.zeros(max(1, comp.n), np.uint8)
    for r in cl:
        if r["classification"] not in ("raft", "support"):
            el[r["component"]] = 1
    t = lap("eligible", t)
    pri = list(reversed(eng.window))
    out = E.roi_fade(img, cur, pri, comp.cid, comp.n, el, 20.0, True); t = lap("roi_fade kernel", t)
    o = out.numpy(); t = lap("d2h", t)
    eng.window.append(cur); eng.window.pop(0)
for k, v in acc.items():
    print(f"{k:24s} {v / 4 * 1e3:8.2f} ms")
print("components", comp.n, "rois", len(rois))
