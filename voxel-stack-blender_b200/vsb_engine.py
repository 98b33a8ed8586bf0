"""
vsb_engine -- device-side plumbing above the C-ABI: PyTorch owns device buffers and streams, every
arithmetic step is a kernel of libvsb_b200.so.  Two layers:

  * per-call helpers (numpy in -> numpy out) used by the re-authored processing_core /
    xy_blend_processor / lut_manager entry points -- one H2D, the kernels, one D2H;
  * StackEngine: the stack-level engine (C++ runtime in csrc/vsb_stack.cu) that keeps the sliding window
    of packed prior masks on the device and pipelines H2D / kernels / D2H -- what
    ProcessingPipelineThread.run() and bench.py drive.

No CPU fallback: without a CUDA device every entry point raises.
"""
from __future__ import annotations

import ctypes
import math
import threading
from ctypes import c_void_p
from typing import List, Optional, Sequence

import numpy as np

import vsb_native as N
import vsb_params as P

_torch = None
_lock = threading.Lock()


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        if not _t.cuda.is_available():
            raise N.VsbError("no CUDA device visible: the B200 path has no CPU fallback")
        _torch = _t
    return _torch


def _stream() -> int:
    return torch().cuda.current_stream().cuda_stream


def wpr_for(W: int) -> int:
    return ((W + 127) // 128) * 4


def pitch_for(W: int) -> int:
    return (W + 127) & ~127


class DevImage:
    """uint8 [H, pitch] device buffer (pitch multiple of 128 so every row is 16-byte aligned)."""

    def __init__(self, H: int, W: int, zero: bool = False):
        t = torch()
        self.H, self.W, self.pitch = H, W, pitch_for(W)
        self.t = (t.zeros if zero else t.empty)((H, self.pitch), dtype=t.uint8, device="cuda")

    @property
    def ptr(self) -> int:
        return self.t.data_ptr()

    @classmethod
    def from_numpy(cls, a: np.ndarray) -> "DevImage":
        a = np.ascontiguousarray(a)
        if a.dtype != np.uint8 or a.ndim != 2:
            raise TypeError("expected a 2-D uint8 array")
        d = cls(a.shape[0], a.shape[1])
        d.t[:, : d.W].copy_(torch().from_numpy(a))
        return d

    def numpy(self) -> np.ndarray:
        t = torch()
        if self.H * self.W < (1 << 22):
            return self.t[:, : self.W].contiguous().cpu().numpy()
        # large layers: into page-locked memory (blocks come from torch's caching host allocator, so only the first
        # layer of a size pays for the registration); a pageable copy of a 16K layer runs at a tenth of the link
        host = t.empty((self.H, self.W), dtype=t.uint8, pin_memory=True)
        host.copy_(self.t[:, : self.W])
        return host.numpy()


class DevBits:
    """bit-packed mask: uint32 [H, wpr]."""

    def __init__(self, H: int, W: int):
        t = torch()
        self.H, self.W, self.wpr = H, W, wpr_for(W)
        self.t = t.empty((H, self.wpr), dtype=t.int32, device="cuda")

    @property
    def ptr(self) -> int:
        return self.t.data_ptr()


_tables = {}
_luts = {}


def table_dev(R: int):
    key = (torch().cuda.current_device(), R)
    if key not in _tables:
        _tables[key] = torch().from_numpy(N.chamfer_table(R)).cuda()
    return _tables[key]


def lut_dev(lut: np.ndarray):
    lut = np.ascontiguousarray(lut)
    if lut.dtype != np.uint8 or lut.shape != (256,):
        raise ValueError("Provided lut_array must be a 256-entry NumPy array of dtype uint8.")
    return torch().from_numpy(lut.copy()).cuda()


# ---- per-call helpers -------------------------------------------------------------------------------------
def pack(img: DevImage, thresh: int = 127) -> DevBits:
    b = DevBits(img.H, img.W)
    N.check(N.load().vsb_threshold_pack(img.ptr, img.pitch, img.W, img.H, thresh, b.ptr, b.wpr, _stream()),
            "vsb_threshold_pack")
    return b


def unpack(b: DevBits) -> DevImage:
    o = DevImage(b.H, b.W)
    N.check(N.load().vsb_bits_unpack(b.ptr, b.wpr, b.W, b.H, o.ptr, o.pitch, _stream()), "vsb_bits_unpack")
    return o


def threshold_np(gray: np.ndarray, thresh: int = 127) -> np.ndarray:
    return unpack(pack(DevImage.from_numpy(gray), thresh)).numpy()


def or_bits(masks: Sequence[DevBits]) -> DevBits:
    out = DevBits(masks[0].H, masks[0].W)
    cur = list(masks)
    first = True
    while cur:
        chunk, cur = cur[: N.VSB_MAX_PRIORS - (0 if first else 1)], cur[N.VSB_MAX_PRIORS - (0 if first else 1):]
        srcs = [m.ptr for m in chunk] + ([] if first else [out.ptr])
        N.check(N.load().vsb_bits_or(N.ptr_array(srcs), len(srcs), out.wpr, out.H, out.ptr, _stream()), "vsb_bits_or")
        first = False
    return out


def count_receding(cur: DevBits, priors: Sequence[DevBits]) -> int:
    t = torch()
    cnt = t.zeros(1, dtype=t.int64, device="cuda")
    N.check(N.load().vsb_maskdiff(cur.ptr, N.ptr_array([p.ptr for p in priors]), len(priors), cur.wpr, cur.H, None,
                                  cnt.data_ptr(), _stream()), "vsb_maskdiff")
    return int(cnt.item())


def zblend(gray: Optional[DevImage], cur: DevBits, priors: Sequence[DevBits], z: N.ZParams,
           lut: Optional[np.ndarray] = None) -> DevImage:
    """One fused Z-blend launch; gray=None -> the bare gradient (merge with zeros)."""
    H, W = cur.H, cur.W
    if gray is None:
        gray = DevImage(H, W, zero=True)
    out = DevImage(H, W)
    zz = N.ZParams.from_buffer_copy(z)
    zz.n_priors = min(len(priors), N.VSB_MAX_PRIORS) if z.mode != N.Z_WEIGHTED else z.n_priors
    T = table_dev(zz.reach)
    ld = lut_dev(lut) if lut is not None else None
    N.check(N.load().vsb_zblend_fused(gray.ptr, gray.pitch, W, H, cur.ptr, N.ptr_array([p.ptr for p in priors]),
                                      cur.wpr, ctypes.byref(zz), T.data_ptr(), ld.data_ptr() if ld is not None else None,
                                      out.ptr, out.pitch, None, 0, _stream()), "vsb_zblend_fused")
    return out


def zblend_unbounded(gray: Optional[DevImage], cur: DevBits, priors: Sequence[DevBits], z: N.ZParams,
                     lut: Optional[np.ndarray] = None) -> DevImage:
    """Fixed fade with any F, or the global min-max normalisation, through the unbounded distance path."""
    t = torch()
    H, W = cur.H, cur.W
    if gray is None:
        gray = DevImage(H, W, zero=True)
    out = DevImage(H, W)
    g = t.empty((H, W), dtype=t.int16, device="cuda")
    gmin = t.empty(int(N.load().vsb_coldist_scratch_elems(W, H, cur.wpr)), dtype=t.int16, device="cuda")
    dist = t.empty((H, W), dtype=t.float32, device="cuda")
    mm = t.empty(2, dtype=t.int32, device="cuda")
    ld = lut_dev(lut) if lut is not None else None
    N.check(N.load().vsb_zblend_unbounded(gray.ptr, gray.pitch, W, H, cur.ptr, N.ptr_array([p.ptr for p in priors]),
                                          len(priors), cur.wpr, int(z.metric), 1 if z.mode == N.Z_MINMAX else 0,
                                          float(z.fade), float(z.den), g.data_ptr(), W, gmin.data_ptr(), dist.data_ptr(), W,
                                          mm.data_ptr(), ld.data_ptr() if ld is not None else None, out.ptr, out.pitch,
                                          _stream()), "vsb_zblend_unbounded")
    return out


def _unbounded_scratch(H: int, W: int, wpr: int):
    t = torch()
    g = t.empty((H, W), dtype=t.int16, device="cuda")
    gmin = t.empty(int(N.load().vsb_coldist_scratch_elems(W, H, wpr)), dtype=t.int16, device="cuda")
    mm = t.empty(2, dtype=t.int32, device="cuda")
    return g, gmin, mm


def weighted_unbounded(gray: Optional[DevImage], cur: DevBits, priors: Sequence[DevBits], z: N.ZParams,
                       lut: Optional[np.ndarray] = None) -> DevImage:
    """Weighted stack with a fade distance beyond the tile kernels' reach (unbounded distances)."""
    t = torch()
    H, W = cur.H, cur.W
    if gray is None:
        gray = DevImage(H, W, zero=True)
    out = DevImage(H, W)
    g, gmin, mm = _unbounded_scratch(H, W, cur.wpr)
    dist = t.empty((H, W), dtype=t.float32, device="cuda")
    ld = lut_dev(lut) if lut is not None else None
    zz = N.ZParams.from_buffer_copy(z)
    N.check(N.load().vsb_weighted_unbounded(gray.ptr, gray.pitch, W, H, cur.ptr, N.ptr_array([p.ptr for p in priors]), cur.wpr,
                                            ctypes.byref(zz), g.data_ptr(), W, gmin.data_ptr(), dist.data_ptr(), W,
                                            mm.data_ptr(), ld.data_ptr() if ld is not None else None, out.ptr, out.pitch,
                                            _stream()), "vsb_weighted_unbounded")
    return out


def dt_unbounded_np(cur_white: np.ndarray, metric: str = "chamfer5"):
    """Dense unbounded distance map (float32) and, for the exact metric, the integer squared distances."""
    t = torch()
    img = DevImage.from_numpy(cur_white)
    b = pack(img)
    H, W = img.H, img.W
    g = t.empty((H, W), dtype=t.int16, device="cuda")
    gmin = t.empty(int(N.load().vsb_coldist_scratch_elems(W, H, b.wpr)), dtype=t.int16, device="cuda")
    dist = t.empty((H, W), dtype=t.float32, device="cuda")
    d2 = t.empty((H, W), dtype=t.int32, device="cuda") if metric == "exact" else None
    N.check(N.load().vsb_dt_unbounded(b.ptr, b.wpr, W, H, N.METRIC_EXACT if metric == "exact" else N.METRIC_CHAMFER5,
                                      g.data_ptr(), W, gmin.data_ptr(), dist.data_ptr(),
                                      d2.data_ptr() if d2 is not None else None, W, _stream()), "vsb_dt_unbounded")
    return dist.cpu().numpy(), (d2.cpu().numpy() if d2 is not None else None)


def enhanced(gray: Optional[DevImage], cur: DevBits, priors: Sequence[DevBits], z: N.ZParams,
             lut: Optional[np.ndarray] = None) -> DevImage:
    """Enhanced EDT: rec mask -> clipped distance at rec pixels -> CCL + per-component max -> fade."""
    t = torch()
    H, W = cur.H, cur.W
    if gray is None:
        gray = DevImage(H, W, zero=True)
    out = DevImage(H, W)
    lib = N.load()
    pri = N.ptr_array([p.ptr for p in priors])
    rec = DevBits(H, W)
    N.check(lib.vsb_maskdiff(cur.ptr, pri, len(priors), cur.wpr, H, rec.ptr, None, _stream()), "vsb_maskdiff")
    dist = t.empty((H, W), dtype=t.float32, device="cuda")
    labels = t.empty((H, W), dtype=t.int32, device="cuda")
    cmax = t.empty((H, W), dtype=t.int32, device="cuda")
    if z.mode == N.Z_ENHANCED_FAR:
        g, gmin, mm = _unbounded_scratch(H, W, cur.wpr)
        N.check(lib.vsb_rec_dist_unbounded(cur.ptr, pri, len(priors), cur.wpr, W, H, int(z.metric), g.data_ptr(), W,
                                           gmin.data_ptr(), dist.data_ptr(), W, mm.data_ptr(), _stream()), "vsb_rec_dist_unbounded")
    elif z.aniso_w > 0:
        # anisotropic branch (processing_core.py:357-377): nearest-resize the mask, transform there, resize back
        aw, ah = int(z.aniso_w), int(z.aniso_h)
        near = ResizePlan(W, H, aw, ah, N.INTER_NEAREST)
        lin = ResizePlan(aw, ah, W, H, N.INTER_LINEAR_F32)
        small = DevBits(ah, aw)
        N.check(lib.vsb_resize_bits_nearest(near.h, cur.ptr, cur.wpr, small.ptr, small.wpr, _stream()), "vsb_resize_bits_nearest")
        sd = t.empty((ah, aw), dtype=t.float32, device="cuda")
        far = float(max(float(z.fade_limit), 0.0) + 4.0)
        N.check(lib.vsb_dt_chamfer5(small.ptr, small.wpr, aw, ah, int(z.aniso_reach), table_dev(int(z.aniso_reach)).data_ptr(),
                                    far, sd.data_ptr(), aw, _stream()), "vsb_dt_chamfer5")
        N.check(lib.vsb_resize_f32_linear(lin.h, sd.data_ptr(), aw, dist.data_ptr(), W, _stream()), "vsb_resize_f32_linear")
        t.cuda.current_stream().synchronize()
        near.close(); lin.close()
    else:
        zd = N.ZParams.from_buffer_copy(z)
        zd.mode, zd.n_priors = N.Z_DIST, len(priors)
        T = table_dev(zd.reach)
        N.check(lib.vsb_zblend_fused(None, 0, W, H, cur.ptr, pri, cur.wpr, ctypes.byref(zd), T.data_ptr(), None, None, 0,
                                     dist.data_ptr(), W, _stream()), "vsb_zblend_fused(dist)")
    N.check(lib.vsb_ccl8_max(rec.ptr, rec.wpr, W, H, dist.data_ptr(), W, labels.data_ptr(), cmax.data_ptr(), _stream()),
            "vsb_ccl8_max")
    ld = lut_dev(lut) if lut is not None else None
    N.check(lib.vsb_enhanced_fade(gray.ptr, gray.pitch, W, H, rec.ptr, rec.wpr, dist.data_ptr(), W, labels.data_ptr(),
                                  cmax.data_ptr(), float(z.fade_limit), int(z.enhanced_arith),
                                  ld.data_ptr() if ld is not None else None, out.ptr, out.pitch, _stream()),
            "vsb_enhanced_fade")
    return out


def enhanced_from_labels_np(rd: np.ndarray, labels: np.ndarray, num_labels: int, F: float, arith: int) -> np.ndarray:
    t = torch()
    rd = np.ascontiguousarray(rd, np.float32)
    labels = np.ascontiguousarray(labels, np.int32)
    H, W = rd.shape
    d = t.from_numpy(rd).cuda()
    l = t.from_numpy(labels).cuda()
    cmax = t.empty(max(1, int(num_labels)), dtype=t.int32, device="cuda")
    out = t.empty((H, W), dtype=t.uint8, device="cuda")
    N.check(N.load().vsb_enhanced_from_labels(d.data_ptr(), l.data_ptr(), W, H, int(num_labels), float(F), int(arith),
                                              cmax.data_ptr(), out.data_ptr(), _stream()), "vsb_enhanced_from_labels")
    return out.cpu().numpy()


def distance_np(cur_white: np.ndarray, R: int, clip: float, metric: str = "chamfer5") -> np.ndarray:
    """Dense clipped distance map to the nearest cur_white>127 pixel (K2)."""
    t = torch()
    img = DevImage.from_numpy(cur_white)
    b = pack(img)
    H, W = img.H, img.W
    dist = t.empty((H, W), dtype=t.float32, device="cuda")
    if metric == "chamfer5":
        N.check(N.load().vsb_dt_chamfer5(b.ptr, b.wpr, W, H, R, table_dev(R).data_ptr(), float(clip), dist.data_ptr(), W,
                                         _stream()), "vsb_dt_chamfer5")
    else:
        N.check(N.load().vsb_dt_exact_windowed(b.ptr, b.wpr, W, H, R, float(clip), dist.data_ptr(), W, _stream()),
                "vsb_dt_exact_windowed")
    return dist.cpu().numpy()


def edt_exact_np(cur_white: np.ndarray):
    """Unbounded exact EDT (separable lower-envelope formulation, csrc/vsb_edt.cu): (squared distances int32, distances float32)."""
    t = torch()
    img = DevImage.from_numpy(cur_white)
    b = pack(img)
    H, W = img.H, img.W
    scratch = t.empty(int(N.load().vsb_edt_exact_scratch_bytes(W, H)), dtype=t.uint8, device="cuda")
    d2 = t.empty((H, W), dtype=t.int32, device="cuda")
    dist = t.empty((H, W), dtype=t.float32, device="cuda")
    N.check(N.load().vsb_edt_exact(b.ptr, b.wpr, W, H, scratch.data_ptr(), d2.data_ptr(), dist.data_ptr(), W, _stream()),
            "vsb_edt_exact")
    return d2.cpu().numpy(), dist.cpu().numpy()


def merge_np(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    da, db = DevImage.from_numpy(a), DevImage.from_numpy(b)
    o = DevImage(da.H, da.W)
    N.check(N.load().vsb_merge_max(da.ptr, da.pitch, db.ptr, db.pitch, da.W, da.H, o.ptr, o.pitch, _stream()),
            "vsb_merge_max")
    return o.numpy()


def lut_apply(img: DevImage, lut: np.ndarray) -> DevImage:
    o = DevImage(img.H, img.W)
    ld = lut_dev(lut)
    N.check(N.load().vsb_lut_apply(img.ptr, img.pitch, img.W, img.H, ld.data_ptr(), o.ptr, o.pitch, _stream()),
            "vsb_lut_apply")
    return o


def gaussian(img: DevImage, qx: np.ndarray, qy: np.ndarray) -> DevImage:
    o = DevImage(img.H, img.W)
    qx, qy = np.ascontiguousarray(qx, np.int32), np.ascontiguousarray(qy, np.int32)
    N.check(N.load().vsb_gaussian_q88(img.ptr, img.pitch, img.W, img.H, qx.ctypes.data_as(c_void_p), len(qx),
                                      qy.ctypes.data_as(c_void_p), len(qy), o.ptr, o.pitch, _stream()), "vsb_gaussian_q88")
    return o


def bilateral(img: DevImage, d: int, sigma_color: float, sigma_space: float) -> DevImage:
    radius, dy, dx, sw, cw = P.bilateral_tables(d, sigma_color, sigma_space)
    o = DevImage(img.H, img.W)
    N.check(N.load().vsb_bilateral(img.ptr, img.pitch, img.W, img.H, radius, len(sw), dy.ctypes.data_as(c_void_p),
                                   dx.ctypes.data_as(c_void_p), sw.ctypes.data_as(c_void_p), cw.ctypes.data_as(c_void_p),
                                   o.ptr, o.pitch, _stream()), "vsb_bilateral")
    return o


def median(img: DevImage, ksize: int) -> DevImage:
    o = DevImage(img.H, img.W)
    N.check(N.load().vsb_median(img.ptr, img.pitch, img.W, img.H, int(ksize), o.ptr, o.pitch, _stream()), "vsb_median")
    return o


# ---- ROI_FADE -----------------------------------------------------------------------------
class Components:
    """8-connected components of a packed mask with cv2.connectedComponentsWithStats' statistics and label order.
    `cid` (device int32 [H, W]) holds compact component id + 1; `order[i]` = compact id of the library's label i+1."""

    def __init__(self, bits: DevBits):
        t = torch()
        H, W = bits.H, bits.W
        self.H, self.W = H, W
        labels = t.empty((H, W), dtype=t.int32, device="cuda")
        self.cid = t.empty((H, W), dtype=t.int32, device="cuda")
        count = t.zeros(1, dtype=t.int32, device="cuda")
        cap = 1 << 14
        while True:
            stats = t.empty(cap * ctypes.sizeof(N.RoiStat), dtype=t.uint8, device="cuda")
            N.check(N.load().vsb_roi_components(bits.ptr, bits.wpr, W, H, labels.data_ptr(), self.cid.data_ptr(),
                                                stats.data_ptr(), count.data_ptr(), cap, _stream()), "vsb_roi_components")
            n = int(count.item())
            if n <= cap:
                break
            cap = 1 << int(n - 1).bit_length()
        raw = stats[: n * ctypes.sizeof(N.RoiStat)].cpu().numpy()
        dt = np.dtype([("root", "<i4"), ("area", "<i4"), ("x0", "<i4"), ("y0", "<i4"), ("x1", "<i4"), ("y1", "<i4"),
                       ("first_block", "<i4"), ("pad", "<i4"), ("sum_x", "<u8"), ("sum_y", "<u8")])
        self.stats = raw.view(dt) if n else np.zeros(0, dt)
        self.n = n
        self.order = np.argsort(self.stats["first_block"], kind="stable") if n else np.zeros(0, np.int64)

    def rois(self, min_size: int, with_masks: bool = False) -> List[dict]:
        """identify_rois' list (processing_core.py:82-103): label, area, bbox, centroid (+ mask on request)."""
        cid_host = self.cid.cpu().numpy() if with_masks else None
        st = self.stats[self.order]                      # the library's label order
        keep = np.nonzero(st["area"] >= min_size)[0]
        st = st[keep]
        area = st["area"].astype(np.int64)
        cx, cy = st["sum_x"] / area, st["sum_y"] / area   # uint64 / int64 -> float64, as the per-ROI division did
        cols = (keep.tolist(), self.order[keep].tolist(), area.tolist(), st["x0"].tolist(), st["y0"].tolist(),
                (st["x1"] - st["x0"] + 1).tolist(), (st["y1"] - st["y0"] + 1).tolist(), cx.tolist(), cy.tolist())
        out = []
        for rank, c, a, x0, y0, w, h, mx, my in zip(*cols):
            r = {"label": rank + 1, "area": a, "bbox": (x0, y0, w, h), "centroid": np.array([mx, my]), "component": c}
            if with_masks:
                r["mask"] = (cid_host == c + 1).astype(np.uint8) * 255
            out.append(r)
        return out


def roi_fade(gray: Optional[DevImage], cur: DevBits, priors: Sequence[DevBits], cid, n_components: int,
             eligible: np.ndarray, F: float, use_fixed: bool, lut: Optional[np.ndarray] = None) -> DevImage:
    """_calculate_receding_gradient_field_roi_fade + merge (+ leading LUT) on the device.  `cid`: device int32 map of
    component id + 1, `eligible`: uint8 per component."""
    t = torch()
    H, W = cur.H, cur.W
    out = DevImage(H, W)
    K = min(51, int(F) * 2 + 1) // 2 if int(F) >= 0 else 0
    K = max(K, 0)
    near_reach = int(math.floor(1.4 * K)) + 1
    R = min(P.chamfer_reach(F), near_reach) if use_fixed else near_reach
    R = max(0, min(R, N.VSB_FAST_REACH))
    el = t.from_numpy(np.ascontiguousarray(eligible, np.uint8)).cuda() if n_components else None
    mm = t.empty(max(1, 2 * n_components), dtype=t.int32, device="cuda")
    pri = N.ptr_array([p.ptr for p in priors])
    ld = lut_dev(lut) if lut is not None else None
    den = float(F) if F > 0 else 1.0
    N.check(N.load().vsb_roi_fade(gray.ptr if gray is not None else None, gray.pitch if gray is not None else 0, W, H,
                                  cur.ptr, pri, len(priors), cur.wpr, cid.data_ptr(),
                                  el.data_ptr() if el is not None else None, int(n_components), K, R,
                                  table_dev(R).data_ptr(), 1 if use_fixed else 0, float(F), den, mm.data_ptr(),
                                  ld.data_ptr() if ld is not None else None, out.ptr, out.pitch, _stream()), "vsb_roi_fade")
    return out


class RoiStackEngine:
    """ROI_FADE stack loop (processing_pipeline.py:163-164, :205-223): per layer the device labels the current mask
    and returns the component statistics, the host tracker classifies them, the device fades per ROI, merges and
    runs the XY chain.  One host round trip per layer (a few KB of statistics down, one byte per component up)."""

    def __init__(self, W: int, H: int, cfg, xy_chain=None, device: Optional[int] = None):
        import roi_tracker
        t = torch()
        if device is not None:
            t.cuda.set_device(device)
        self.W, self.H, self.cfg = int(W), int(H), cfg
        self.tracker = roi_tracker.ROITracker()
        self.window: List[DevBits] = []
        self.n_priors = max(0, int(cfg.receding_layers))
        if self.n_priors > N.VSB_MAX_WINDOW:
            raise P.UnsupportedOnDevice(f"more than {N.VSB_MAX_WINDOW} receding layers")
        self.xy_chain = xy_chain          # callable DevImage -> DevImage, or None
        self.last_rois: List[dict] = []

    def reset(self):
        import roi_tracker
        self.tracker = roi_tracker.ROITracker()
        self.window = []

    def process(self, gray: np.ndarray, layer_index: int) -> np.ndarray:
        gray = np.ascontiguousarray(gray, np.uint8)
        if gray.shape != (self.H, self.W):
            raise ValueError(f"layer shape {gray.shape} != engine shape {(self.H, self.W)}")
        cfg = self.cfg
        img = DevImage.from_numpy(gray)
        cur = pack(img)
        comp = Components(cur)
        rois = comp.rois(int(cfg.roi_params.min_size))
        classified = self.tracker.update_and_classify(rois, layer_index, cfg)
        self.last_rois = classified
        eligible = np.zeros(max(1, comp.n), np.uint8)
        skip = bool(cfg.roi_params.enable_raft_support_handling)
        for r in classified:
            if not (skip and r["classification"] in ("raft", "support")):
                eligible[r["component"]] = 1
        priors = list(reversed(self.window))     # newest first (only their OR matters here)
        if len(priors) > N.VSB_MAX_PRIORS:
            priors = [or_bits(priors)]
        out = roi_fade(img, cur, priors, comp.cid, comp.n if (classified and priors) else 0, eligible,
                       float(cfg.fixed_fade_distance_receding), bool(cfg.use_fixed_fade_receding))
        if self.xy_chain is not None:
            out = self.xy_chain(out)
        if self.n_priors > 0:
            self.window.append(cur)
            if len(self.window) > self.n_priors:
                self.window.pop(0)
        return out.numpy()

    def close(self):
        self.window = []


class ResizePlan:
    """Device tables of one cv2.resize geometry (csrc/vsb_geom.cu)."""

    def __init__(self, sw: int, sh: int, dw: int, dh: int, interp: int):
        h = c_void_p()
        N.check(N.load().vsb_resize_plan_create(int(sw), int(sh), int(dw), int(dh), int(interp), ctypes.byref(h)),
                "vsb_resize_plan_create")
        self.h, self.dw, self.dh = h, int(dw), int(dh)

    def close(self):
        if getattr(self, "h", None):
            N.load().vsb_resize_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def unsharp(img: DevImage, amount: float, threshold: int, ksize: int, sigma: float) -> DevImage:
    """apply_unsharp_mask (xy_blend_processor.py:48-67): Gaussian, addWeighted, optional threshold select."""
    q = P.gaussian_taps_q88(P.odd_ksize(int(ksize)), float(sigma))
    blurred = gaussian(img, q, q)
    o = DevImage(img.H, img.W)
    N.check(N.load().vsb_unsharp_combine(img.ptr, img.pitch, blurred.ptr, blurred.pitch, img.W, img.H, float(amount),
                                         int(threshold), o.ptr, o.pitch, _stream()), "vsb_unsharp_combine")
    return o


def resize(img: DevImage, dw: int, dh: int, interp: int) -> DevImage:
    """cv2.resize(image, (dw, dh), interpolation) on uint8 (xy_blend_processor.py:90)."""
    plan = ResizePlan(img.W, img.H, dw, dh, interp)
    o = DevImage(int(dh), int(dw))
    N.check(N.load().vsb_resize_u8(plan.h, img.ptr, img.pitch, o.ptr, o.pitch, _stream()), "vsb_resize_u8")
    torch().cuda.current_stream().synchronize()
    plan.close()
    return o


def resize_f32_linear_np(a: np.ndarray, dw: int, dh: int) -> np.ndarray:
    t = torch()
    a = np.ascontiguousarray(a, np.float32)
    sh, sw = a.shape
    plan = ResizePlan(sw, sh, dw, dh, N.INTER_LINEAR_F32)
    d = t.from_numpy(a).cuda()
    o = t.empty((dh, dw), dtype=t.float32, device="cuda")
    N.check(N.load().vsb_resize_f32_linear(plan.h, d.data_ptr(), sw, o.data_ptr(), dw, _stream()), "vsb_resize_f32_linear")
    r = o.cpu().numpy()
    plan.close()
    return r


# ---- synthetic stacks (measurement utility) -----------------------------------------------------------------
def synth_layer(prims_dev, n_prims: int, raft: Optional[np.ndarray], z: int, W: int, H: int, out_ptr: int, pitch: int):
    r = np.ascontiguousarray(raft, np.float32) if raft is not None else None
    N.check(N.load().vsb_synth_layer(prims_dev.data_ptr() if n_prims else None, n_prims,
                                     r.ctypes.data_as(c_void_p) if r is not None else None, int(z), W, H, out_ptr, pitch,
                                     _stream()), "vsb_synth_layer")


# ---- the stack-level engine -------------------------------------------------------------------------------------
class StackEngine:
    """Device-resident per-stack engine (csrc/vsb_stack.cu).  Layers must be fed in Z order; the window of
    the last `receding_layers` thresholded masks lives on the device (processing_pipeline.py:163,216-223)."""

    def __init__(self, W: int, H: int, cfg, xy_ops: Optional[List[N.XYOp]] = None, metric: str = "chamfer5",
                 device: Optional[int] = None):
        t = torch()
        if device is not None:
            t.cuda.set_device(device)
        self.device = t.cuda.current_device()
        self.W, self.H = int(W), int(H)
        self.pitch = pitch_for(W)
        z = P.zparams_from_config(cfg, metric, self.W, self.H)
        n_priors = max(0, int(cfg.receding_layers))
        if n_priors > N.VSB_MAX_WINDOW:   # the GUI offers 0..100 (ui_components.py:168)
            raise P.UnsupportedOnDevice(f"more than {N.VSB_MAX_WINDOW} receding layers")
        ops = list(xy_ops or [])
        if len(ops) > N.VSB_MAX_XY_OPS:
            raise P.UnsupportedOnDevice(f"more than {N.VSB_MAX_XY_OPS} XY operations")
        arr = (N.XYOp * max(1, len(ops)))(*ops)
        h = c_void_p()
        N.check(N.load().vsb_stack_create(self.W, self.H, n_priors, ctypes.byref(z), arr, len(ops), ctypes.byref(h)),
                "vsb_stack_create")
        self._h = h
        self.slots = N.load().vsb_stack_slots(h)
        ow, oh = ctypes.c_int(), ctypes.c_int()
        N.check(N.load().vsb_stack_out_size(h, ctypes.byref(ow), ctypes.byref(oh)), "vsb_stack_out_size")
        self.out_W, self.out_H = ow.value, oh.value   # differs from (W, H) when the XY chain holds a resize

    def close(self):
        if getattr(self, "_h", None):
            N.load().vsb_stack_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        N.check(N.load().vsb_stack_reset(self._h), "vsb_stack_reset")

    def push_halo(self, gray: np.ndarray):
        gray = np.ascontiguousarray(gray, np.uint8)
        N.check(N.load().vsb_stack_push_halo(self._h, gray.ctypes.data_as(c_void_p), gray.strides[0]), "vsb_stack_push_halo")

    def push_halo_device(self, gray_ptr: int, pitch: int):
        """Halo layer already in HBM (a Z-shard of a device-resident stack): thresholded into the window, no output."""
        N.check(N.load().vsb_stack_push_halo_device(self._h, gray_ptr, pitch), "vsb_stack_push_halo_device")

    def process(self, gray: np.ndarray) -> np.ndarray:
        gray = np.ascontiguousarray(gray, np.uint8)
        if gray.shape != (self.H, self.W):
            raise ValueError(f"layer shape {gray.shape} != engine shape {(self.H, self.W)}")
        out = np.empty((self.out_H, self.out_W), np.uint8)
        N.check(N.load().vsb_stack_process_host(self._h, gray.ctypes.data_as(c_void_p), gray.strides[0],
                                                out.ctypes.data_as(c_void_p), out.strides[0]), "vsb_stack_process_host")
        return out

    def submit(self, gray_ptr: int, in_pitch: int, out_ptr: int, out_pitch: int) -> int:
        """Pipelined host path on raw (ideally pinned) host pointers; returns a ticket for wait()."""
        return N.check(N.load().vsb_stack_submit_host(self._h, gray_ptr, in_pitch, out_ptr, out_pitch),
                       "vsb_stack_submit_host")

    def wait(self, ticket: int):
        N.check(N.load().vsb_stack_wait(self._h, ticket), "vsb_stack_wait")

    def process_device(self, gray_ptr: int, in_pitch: int, out_ptr: int, out_pitch: int):
        N.check(N.load().vsb_stack_process_device(self._h, gray_ptr, in_pitch, out_ptr, out_pitch),
                "vsb_stack_process_device")

    def process_device_batch(self, gray_ptr: int, in_pitch: int, in_stride: int, out_ptr: int, out_pitch: int,
                             out_stride: int, out_ring: int, n: int):
        N.check(N.load().vsb_stack_process_device_batch(self._h, gray_ptr, in_pitch, in_stride, out_ptr, out_pitch,
                                                        out_stride, out_ring, n), "vsb_stack_process_device_batch")

    def sync(self):
        N.check(N.load().vsb_stack_sync(self._h), "vsb_stack_sync")

    def profile(self, on: bool):
        N.check(N.load().vsb_stack_profile_enable(self._h, 1 if on else 0), "vsb_stack_profile_enable")

    def profile_read(self):
        """[(stage name, summed ms, kernel launches)] since profile(True)."""
        ms = (ctypes.c_float * 32)()
        ln = (ctypes.c_int * 32)()
        names = ctypes.create_string_buffer(32 * 32)
        n = N.check(N.load().vsb_stack_profile_read(self._h, 32, ms, ln, names), "vsb_stack_profile_read")
        return [(names.raw[i * 32:(i + 1) * 32].split(b"\0")[0].decode(), float(ms[i]), int(ln[i])) for i in range(n)
                if ln[i] > 0]

    @property
    def stream(self) -> int:
        return N.load().vsb_stack_stream(self._h) or 0


def zshard_plan(n_layers: int, world: int, halo: int):
    """Contiguous Z-slabs [z0,z1) per rank plus the read-only halo [max(0,z0-halo), z0) each slab re-ingests
    (masks only).  No data-path collective."""
    per = -(-n_layers // world)
    plan = []
    for r in range(world):
        z0, z1 = min(n_layers, r * per), min(n_layers, (r + 1) * per)
        plan.append({"rank": r, "z0": z0, "z1": z1, "halo0": max(0, z0 - halo)})
    return plan
