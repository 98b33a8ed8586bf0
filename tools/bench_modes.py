"""Device-resident per-layer time of the other BASELINE configurations at 16K (parity-test cases, not bench lines):
config 3 (Enhanced EDT F=64 + bilateral + piecewise LUT), weighted stack, min-max, ROI_FADE, unsharp / resize chains."""
import os, sys, time
sys.path.insert(0, '/root/repo'); os.environ["VSB_NO_APP_CONFIG_WRITE"] = "1"
sys.path.insert(0, '/root/repo/voxel-stack-blender_b200')
import numpy as np, torch
import vsb_engine as E, vsb_params as P, vsb_synth as S, lut_manager
from config import Config
W, H, L = 15120, 6230, 14
stack = S.DeviceStack(W, H, L, seed=1, z0=20, total_layers=20 + L)
EXP = {"type": "apply_lut", "lut_params": {"lut_source": "generated", "lut_generation_type": "exp", "input_min": 200, "input_max": 255,
                                            "output_min": 200, "output_max": 255, "exp_param": 2.0}}
BIL = {"type": "bilateral_filter", "bilateral_d": 7, "bilateral_sigma_color": 60.0, "bilateral_sigma_space": 60.0}
CASES = {
    "cfg3 enhanced F=64 numba + bilateral + exp LUT": {"blending_mode": "enhanced_edt", "receding_layers": 4, "fixed_fade_distance_receding": 64.0,
                                                       "use_numba_jit": True, "xy_blend_pipeline": [BIL, EXP]},
    "enhanced F=64 scipy, no XY": {"blending_mode": "enhanced_edt", "receding_layers": 4, "fixed_fade_distance_receding": 64.0, "use_numba_jit": False},
    "weighted N=4 F<=40 + exp LUT": {"blending_mode": "weighted_stack", "receding_layers": 4, "manual_weights": [100, 75, 50, 25],
                                     "fade_distances_receding": [40.0, 30.0, 20.0, 10.0], "xy_blend_pipeline": [EXP]},
    "fixed min-max (dataclass default)": {"blending_mode": "fixed_fade", "receding_layers": 4, "use_fixed_fade_receding": False},
    "fixed F=200 (unbounded path)": {"blending_mode": "fixed_fade", "receding_layers": 4, "use_fixed_fade_receding": True, "fixed_fade_distance_receding": 200.0},
    "fixed F=40 + unsharp + resize 0.5 AREA": {"blending_mode": "fixed_fade", "receding_layers": 4, "use_fixed_fade_receding": True,
                                              "fixed_fade_distance_receding": 40.0,
                                              "xy_blend_pipeline": [{"type": "unsharp_mask", "unsharp_amount": 1.0, "unsharp_threshold": 2, "unsharp_blur_ksize": 5},
                                                                    {"type": "resize", "resize_width": W // 2, "resize_height": H // 2, "resample_mode": "AREA"}]},
    "anisotropic enhanced F=24 x1.25 y0.6": {"blending_mode": "enhanced_edt", "receding_layers": 4, "fixed_fade_distance_receding": 24.0,
                                             "anisotropic_params": {"enabled": True, "x_factor": 1.25, "y_factor": 0.6}},
}
for name, d in (CASES.items() if len(sys.argv) < 2 else []):
    cfg = Config.from_dict(d)
    ops = P.xy_ops_from_pipeline(cfg.xy_blend_pipeline, lambda op: lut_manager.lut_from_params(op.lut_params), W, H)
    eng = E.StackEngine(W, H, cfg, ops)
    out = torch.empty((eng.out_H, E.pitch_for(eng.out_W)), dtype=torch.uint8, device='cuda')
    for z in range(6):
        eng.process_device(stack.ptr(z), stack.pitch, out.data_ptr(), E.pitch_for(eng.out_W))
    eng.sync()
    t0 = time.perf_counter()
    for z in range(6, L):
        eng.process_device(stack.ptr(z), stack.pitch, out.data_ptr(), E.pitch_for(eng.out_W))
    eng.sync()
    dt = (time.perf_counter() - t0) / (L - 6)
    print(f"{name:48s} {dt*1e3:8.3f} ms/layer  {1/dt:8.1f} layers/s")
    eng.close()
# ROI_FADE through the Python-level engine (host tracker in the loop; host copies in and out)
d = {"blending_mode": "roi_fade", "receding_layers": 4, "use_fixed_fade_receding": True, "fixed_fade_distance_receding": 20.0,
     "roi_params": {"min_size": 100, "enable_raft_support_handling": True, "support_max_size": 2000}}
cfg = Config.from_dict(d)
eng = E.RoiStackEngine(W, H, cfg)
host = [stack.layers[z][:, :W].cpu().numpy() for z in range(8)]
for z in range(4):
    eng.process(host[z], 20 + z)
t0 = time.perf_counter()
for z in range(4, 8):
    eng.process(host[z], 20 + z)
dt = (time.perf_counter() - t0) / 4
print(f"{'ROI_FADE F=20 (host arrays in/out, tracker on host)':48s} {dt*1e3:8.3f} ms/layer  {1/dt:8.1f} layers/s  ({len(eng.last_rois)} ROIs)")
