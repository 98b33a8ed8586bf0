"""vsb_edt_exact alone on one 16K mask of the EDT-only bench stack (run under ncu for per-kernel times)."""
import os, sys
sys.path.insert(0, '/root/repo'); os.environ["VSB_NO_APP_CONFIG_WRITE"] = "1"
sys.path.insert(0, '/root/repo/voxel-stack-blender_b200')
import torch
import vsb_engine as E, vsb_native as N, vsb_synth as S
W, H = 15120, 6230
Z0 = int(os.environ.get('Z0', '8'))
lib = N.load()
stack = S.DeviceStack(W, H, 1, seed=2, z0=Z0, total_layers=8192)
img = E.DevImage(H, W); img.t.copy_(stack.layers[0]); b = E.pack(img)
scratch = torch.empty(int(lib.vsb_edt_exact_scratch_bytes(W, H)), dtype=torch.uint8, device="cuda")
dist = torch.empty((H, W), dtype=torch.float32, device="cuda")
for rep in range(int(os.environ.get('REPS', '3'))):
    N.check(lib.vsb_edt_exact(b.ptr, b.wpr, W, H, scratch.data_ptr(), None, dist.data_ptr(), W, 0))
torch.cuda.synchronize()
print("max dist", float(dist.max()), "white", float((stack.layers[0][:, :W] > 127).float().mean()))
