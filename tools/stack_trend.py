import os, sys
sys.path.insert(0, '/root/repo'); os.environ["VSB_NO_APP_CONFIG_WRITE"] = "1"
sys.path.insert(0, '/root/repo/voxel-stack-blender_b200')
import numpy as np, torch, ctypes
import bench, vsb_engine as E, vsb_native as N, vsb_params as P, vsb_synth as S, lut_manager
W, H, L = 15120, 6230, 416
cfg = bench.workload_config(bench.write_lut_file())
stack = S.DeviceStack(W, H, L, seed=1, z0=20, total_layers=20 + L)
ops = P.xy_ops_from_pipeline(cfg.xy_blend_pipeline, lambda op: lut_manager.lut_from_params(op.lut_params))
eng = E.StackEngine(W, H, cfg, ops)
out = torch.empty((H, stack.pitch), dtype=torch.uint8, device='cuda')
st = torch.cuda.ExternalStream(eng.stream)
for z0 in (0, 40, 100, 200, 300, 400):
    eng.reset()
    for z in range(z0, z0 + 5): eng.process_device(stack.ptr(z), stack.pitch, out.data_ptr(), stack.pitch)
    N.load().vsb_layer_stats(1, None)
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(st)
    for z in range(z0 + 5, z0 + 13): eng.process_device(stack.ptr(z), stack.pitch, out.data_ptr(), stack.pitch)
    b.record(st); torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)(); N.load().vsb_layer_stats(2, buf)
    v = [x / 8 for x in buf]
    g = stack.layers[z0 + 8][:, :W]
    print(f"z={20+z0+8:4d} white {float((g>127).float().mean()):.3f}  {a.elapsed_time(b)*1000/8:6.0f} us/layer  worked tiles {v[1]:.0f} rec px {v[3]:.0f} bilateral px {v[4]:.0f}")
