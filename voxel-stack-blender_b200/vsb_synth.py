"""
vsb_synth -- seeded synthetic slice stacks (measurement utility).

A stack is a table of disk segments  (cx, cy, r_start, slope, z_start, z_end): at layer z in
[z_start, z_end) the disk has radius r_start + slope*(z - z_start); a pixel is 255 iff it lies inside any
active disk or inside the raft (a rounded rectangle over 60 % of the plate for z < 12).  Families:
  supports: cones on a jittered 150-px grid, r0 in U[6,20], shrinking 0.002..0.02 px/layer, re-seeded at 0;
  model blobs: 400 disks, r0 in U[20,300], slope in U[-3,+1.5] px/layer; a growing blob turns around at
               r = 400, a vanished one is re-seeded elsewhere -> every layer has growing and receding rims
               0.5-3 px wide, values strictly {0,255} (the reference requires non-anti-aliased input).
The host draws the table with numpy.random.default_rng(seed); layers are rasterised on the device
(vsb_synth_layer) or, for small sizes and tests, by `raster_numpy` with the same float32 arithmetic.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def make_table(W: int, H: int, L: int, seed: int = 0, n_blobs: int = 400, grid: int = 150, raft: bool = True,
               r_range=(20.0, 300.0), r_turn: float = 400.0):
    rng = np.random.default_rng(seed)
    segs = []
    # supports
    for gy in range(grid // 2, H, grid):
        for gx in range(grid // 2, W, grid):
            z = 0
            cx, cy = gx + rng.uniform(-40, 40), gy + rng.uniform(-40, 40)
            while z < L:
                r0, s = rng.uniform(6, 20), rng.uniform(0.002, 0.02)
                life = int(np.ceil(r0 / s))
                segs.append((cx, cy, r0, -s, z, min(L, z + life)))
                z += life
    # model blobs
    for _ in range(n_blobs):
        z = 0
        while z < L:
            cx, cy = rng.uniform(0, W), rng.uniform(0, H)
            r, slope = rng.uniform(*r_range), rng.uniform(-3.0, 1.5)
            while z < L:
                if slope >= 0:
                    life = max(1, int(np.ceil((r_turn - r) / max(slope, 1e-3)))) if r < r_turn else 1
                    life = min(life, L - z)
                    segs.append((cx, cy, r, slope, z, z + life))
                    r = r + slope * life
                    z += life
                    slope = -rng.uniform(0.5, 3.0)
                else:
                    life = min(max(1, int(np.ceil(r / -slope))), L - z)
                    segs.append((cx, cy, r, slope, z, z + life))
                    z += life
                    break  # vanished: re-seed elsewhere
    table = np.array(segs, np.float32).reshape(-1, 6)
    raft_rec = None
    if raft:
        s = np.sqrt(0.6)
        raft_rec = np.array([W * (1 - s) / 2, H * (1 - s) / 2, W * (1 + s) / 2, H * (1 + s) / 2, 60.0, 12.0], np.float32)
    return table, raft_rec


def active(table: np.ndarray, z: int) -> np.ndarray:
    """segments alive at layer z (host-side cull; the device kernel culls per tile)."""
    m = (table[:, 4] <= z) & (z < table[:, 5])
    return np.ascontiguousarray(table[m])


def raster_numpy(table: np.ndarray, raft, z: int, W: int, H: int) -> np.ndarray:
    """float32 restatement of csrc/vsb_synth.cu for small sizes."""
    out = np.zeros((H, W), bool)
    fx = np.arange(W, dtype=np.float32)[None, :]
    fy = np.arange(H, dtype=np.float32)[:, None]
    zf = f32(z)
    if raft is not None and zf < raft[5]:
        x0, y0, x1, y1, cr = [f32(v) for v in raft[:5]]
        ex = np.maximum(np.maximum((x0 + cr) - fx, fx - (x1 - cr)), f32(0))
        ey = np.maximum(np.maximum((y0 + cr) - fy, fy - (y1 - cr)), f32(0))
        inside = (fx >= x0) & (fx <= x1) & (fy >= y0) & (fy <= y1) & (ex * ex + ey * ey <= cr * cr)
        out |= inside
    for cx, cy, r0, slope, zs, ze in active(table, z):
        r = f32(r0) + f32(slope) * (zf - f32(zs))
        if r <= 0:
            continue
        xa, xb = max(0, int(np.floor(cx - r)) - 1), min(W, int(np.ceil(cx + r)) + 2)
        ya, yb = max(0, int(np.floor(cy - r)) - 1), min(H, int(np.ceil(cy + r)) + 2)
        if xa >= xb or ya >= yb:
            continue
        dx = fx[:, xa:xb] - f32(cx)
        dy = fy[ya:yb, :] - f32(cy)
        out[ya:yb, xa:xb] |= (dx * dx + dy * dy) <= r * r
    return out.astype(np.uint8) * np.uint8(255)


class DeviceStack:
    """L device-resident synthetic layers, uint8 [L, H, pitch] (pitch multiple of 128)."""

    def __init__(self, W: int, H: int, L: int, seed: int = 0, z0: int = 0, total_layers: int = None):
        import vsb_engine as E
        t = E.torch()
        self.W, self.H, self.L, self.pitch = W, H, L, E.pitch_for(W)
        table, raft = make_table(W, H, total_layers or (z0 + L), seed)
        self.layers = t.empty((L, H, self.pitch), dtype=t.uint8, device="cuda")
        self.layers.zero_()
        for i in range(L):
            z = z0 + i
            act = active(table, z)
            dev = t.from_numpy(act).cuda() if len(act) else None
            E.synth_layer(dev if dev is not None else t.empty(0, device="cuda"), len(act), raft, z, W, H,
                          self.layers[i].data_ptr(), self.pitch)
        t.cuda.synchronize()

    def ptr(self, i: int) -> int:
        return self.layers[i].data_ptr()

    @property
    def layer_stride(self) -> int:
        return self.H * self.pitch
