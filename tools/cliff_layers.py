"""A bench configuration (argv[1], default the headline one) on z = 8..13 of the bench stack (the raft ends at z = 12); under ncu, `-k regex:k_layer2 -s 4 -c 1`
captures the first post-raft layer."""
import os, sys
sys.path.insert(0, '/root/repo'); os.environ["VSB_NO_APP_CONFIG_WRITE"] = "1"
sys.path.insert(0, '/root/repo/voxel-stack-blender_b200')
import torch
import bench, vsb_engine as E, vsb_params as P, vsb_synth as S, lut_manager
W, H, L = 15120, 6230, 6
cfg = bench.workload_config(bench.write_lut_file(), sys.argv[1] if len(sys.argv) > 1 else "preset")  # preset | zlut | cfg3
stack = S.DeviceStack(W, H, L, seed=1, z0=8, total_layers=8192)
ops = P.xy_ops_from_pipeline(cfg.xy_blend_pipeline, lambda op: lut_manager.lut_from_params(op.lut_params))
eng = E.StackEngine(W, H, cfg, ops)
out = torch.empty((H, stack.pitch), dtype=torch.uint8, device='cuda')
st = torch.cuda.ExternalStream(eng.stream)
for z in range(L):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(st)
    eng.process_device(stack.ptr(z), stack.pitch, out.data_ptr(), stack.pitch)
    b.record(st)
    torch.cuda.synchronize()
    print(8 + z, round(a.elapsed_time(b) * 1000), "us")
