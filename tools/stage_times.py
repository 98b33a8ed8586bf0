"""Per-stage CUDA-event times of one configuration at 16K (vsb_stack_profile_*).  usage: stage_times.py enhanced|weighted|minmax"""
import os, sys
sys.path.insert(0, '/root/repo'); os.environ["VSB_NO_APP_CONFIG_WRITE"] = "1"
sys.path.insert(0, '/root/repo/voxel-stack-blender_b200')
import torch
import vsb_engine as E, vsb_params as P, vsb_synth as S, lut_manager
from config import Config
W, H, L = 15120, 6230, 16
BIL = {"type": "bilateral_filter", "bilateral_d": 7, "bilateral_sigma_color": 60.0, "bilateral_sigma_space": 60.0}
EXP = {"type": "apply_lut", "lut_params": {"lut_source": "generated", "lut_generation_type": "exp", "input_min": 200, "input_max": 255,
                                            "output_min": 200, "output_max": 255, "exp_param": 2.0}}
CASES = {"enhanced": {"blending_mode": "enhanced_edt", "receding_layers": 4, "fixed_fade_distance_receding": 64.0, "use_numba_jit": True,
                      "xy_blend_pipeline": [BIL, EXP]},
         "weighted": {"blending_mode": "weighted_stack", "receding_layers": 4, "manual_weights": [100, 75, 50, 25],
                      "fade_distances_receding": [40.0, 30.0, 20.0, 10.0], "xy_blend_pipeline": [BIL, EXP]},
         "minmax": {"blending_mode": "fixed_fade", "receding_layers": 4, "use_fixed_fade_receding": False}}
d = CASES[sys.argv[1] if len(sys.argv) > 1 else "enhanced"]
cfg = Config.from_dict(d)
stack = S.DeviceStack(W, H, L, seed=1, z0=20, total_layers=436)
ops = P.xy_ops_from_pipeline(cfg.xy_blend_pipeline, lambda op: lut_manager.lut_from_params(op.lut_params), W, H)
eng = E.StackEngine(W, H, cfg, ops)
out = torch.empty((H, stack.pitch), dtype=torch.uint8, device='cuda')
for z in range(6): eng.process_device(stack.ptr(z), stack.pitch, out.data_ptr(), stack.pitch)
eng.sync(); eng.profile(True)
n = L - 6
for z in range(6, L): eng.process_device(stack.ptr(z), stack.pitch, out.data_ptr(), stack.pitch)
tot = 0
for name, ms, launches in eng.profile_read():
    print(f"{name:18s} {ms/n*1000:8.1f} us/layer  ({launches/n:.0f} launches)"); tot += ms / n
print(f"{'total':18s} {tot*1000:8.1f} us/layer")
