"""Per-layer device time of the bench stack (which layers dominate the average)."""
import os, sys
sys.path.insert(0, '/root/repo'); os.environ["VSB_NO_APP_CONFIG_WRITE"] = "1"
sys.path.insert(0, '/root/repo/voxel-stack-blender_b200')
import numpy as np, torch
import bench, vsb_engine as E, vsb_params as P, vsb_synth as S, lut_manager
W, H, L = 15120, 6230, 40
cfg = bench.workload_config(bench.write_lut_file())
if len(sys.argv) > 1 and sys.argv[1] == "zlut":
    cfg.xy_blend_pipeline = cfg.xy_blend_pipeline[:1]
Z0 = int(os.environ.get('Z0', '0'))
stack = S.DeviceStack(W, H, L, seed=1, z0=Z0, total_layers=Z0 + L)
ops = P.xy_ops_from_pipeline(cfg.xy_blend_pipeline, lambda op: lut_manager.lut_from_params(op.lut_params))
eng = E.StackEngine(W, H, cfg, ops)
out = torch.empty((H, stack.pitch), dtype=torch.uint8, device='cuda')
for rep in range(2):
    eng.reset()
    ts = []
    for z in range(L):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        st = torch.cuda.ExternalStream(eng.stream)
        a.record(st)
        eng.process_device(stack.ptr(z), stack.pitch, out.data_ptr(), stack.pitch)
        b.record(st)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1000)
print("us per layer:", " ".join(f"{z}:{t:.0f}" for z, t in enumerate(ts)))
print("mean %.0f  median %.0f  max %.0f" % (np.mean(ts), np.median(ts), np.max(ts)))
