"""Receding-pixel statistics of the bench stack: listed tiles, pixels per 128x32 tile and per 8-row band."""
import os, sys
sys.path.insert(0, '/root/repo'); os.environ["VSB_NO_APP_CONFIG_WRITE"] = "1"
sys.path.insert(0, '/root/repo/voxel-stack-blender_b200')
import torch, numpy as np
import vsb_synth as S
W, H = 15120, 6230
Z0 = int(os.environ.get('Z0', '40'))
st = S.DeviceStack(W, H, 5, seed=1, z0=Z0, total_layers=8192)
m = st.layers[:, :, :W] > 127
rec = (m[0] | m[1] | m[2] | m[3]) & ~m[4]
print("receding px", int(rec.sum()), "frac", float(rec.float().mean()))
Hp, Wp = (H + 31) // 32 * 32, (W + 127) // 128 * 128
r = torch.zeros((Hp, Wp), dtype=torch.int32, device="cuda"); r[:H, :W] = rec.int()
t = r.view(Hp // 32, 32, Wp // 128, 128).sum(dim=(1, 3)).flatten().cpu().numpy()
b = r.view(Hp // 8, 8, Wp // 128, 128).sum(dim=(1, 3)).flatten().cpu().numpy()
for name, a in (("tile", t), ("band", b)):
    nz = a[a > 0]
    print(name, "units", len(a), "nonzero", len(nz), "mean", nz.mean(), "pcts 50/90/99/max", np.percentile(nz, [50, 90, 99]), nz.max(),
          "hist", np.histogram(nz, bins=[1, 8, 32, 64, 128, 256, 512, 1024, 4097])[0])
