import os, sys, time
sys.path.insert(0, '/root/repo'); os.environ["VSB_NO_APP_CONFIG_WRITE"] = "1"
sys.path.insert(0, '/root/repo/voxel-stack-blender_b200')
import numpy as np, torch
import bench, vsb_engine as E, vsb_params as P, vsb_synth as S, lut_manager
W, H, L = 15120, 6230, 40
cfg = bench.workload_config(bench.write_lut_file())
stack = S.DeviceStack(W, H, L, seed=1, z0=20, total_layers=20 + L)
ops = P.xy_ops_from_pipeline(cfg.xy_blend_pipeline, lambda op: lut_manager.lut_from_params(op.lut_params))
eng = E.StackEngine(W, H, cfg, ops)
outs = torch.empty((4, H, stack.pitch), dtype=torch.uint8, device='cuda')
st = torch.cuda.ExternalStream(eng.stream)
def run(order, ring):
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(st)
    for i, z in enumerate(order):
        eng.process_device(stack.ptr(z), stack.pitch, outs[i % ring].data_ptr(), stack.pitch)
    b.record(st)
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1000 / len(order)
up = list(range(0, 40)); down = list(range(39, -1, -1))
for name, order, ring in (("up, 1 out", up, 1), ("up, 4 outs", up, 4), ("down, 4 outs", down, 4), ("up", up, 4), ("down", down, 4)):
    eng.reset()
    for z in order[:5]: eng.process_device(stack.ptr(z), stack.pitch, outs[0].data_ptr(), stack.pitch)
    print(f"{name:14s} {run(order[5:], ring):7.1f} us/layer")
