"""
vsb_oracle_geom -- TEST INFRASTRUCTURE ONLY (CPU oracle, SURVEY.md section 8(f) rank 2 operations).

NumPy restatement of what the reference delegates to OpenCV for
    apply_unsharp_mask   xy_blend_processor.py:48-67   (GaussianBlur + addWeighted + absdiff/threshold select)
    apply_resize         xy_blend_processor.py:69-90   (cv2.resize, 5 interpolation modes, uint8)
    anisotropic branch   processing_core.py:357-377    (INTER_NEAREST of the mask, INTER_LINEAR of the float32 map)

The arithmetic below was fitted to opencv-python 4.13.0 (the reference pins 4.12.0.88, requirements.txt:25) and is
pinned by tests/golden/ref_geom.npz (outputs of the unmodified reference functions):
  * addWeighted u8:  rint(fma(a, alpha, fl32(b * beta)))  in float32, saturated          -- bit-exact
  * NEAREST / BILINEAR / LANCZOS4 / AREA (all three of its code paths) on uint8           -- bit-exact
  * BICUBIC uint8: OpenCV's own kernel (float vertical pass on whole 8-lane groups, fixed point on the row tail)
    -- bit-exact against cv2 with IPP disabled; the IPP build of cv2 replaces it by ippiResizeCubic, which
    differs by +-1 LSB on ~1 % of the pixels (same situation as bilateral d=3,5)
  * INTER_LINEAR float32: (S0*a0 + S1*a1)*b0 + (..)*b1, no FMA -- bit-exact against cv2 without IPP; the IPP
    build differs by <= 2.2e-4 absolute.
Same import rule as vsb_oracle: tests/, smoke() and bench.py's CPU legs only.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

f32 = np.float32
COEF_BITS = 11
COEF_ONE = 1 << COEF_BITS


# --------------------------------------------------------------------------------------
# addWeighted / unsharp mask                                  xy_blend_processor.py:48-67
# --------------------------------------------------------------------------------------
def add_weighted_u8(a: np.ndarray, alpha: float, b: np.ndarray, beta: float) -> np.ndarray:
    """cv2.addWeighted(a, alpha, b, beta, 0) on uint8: float32 scalars, fma(a, alpha, b*beta), round half even."""
    al, be = f32(alpha), f32(beta)
    p2 = (b.astype(np.float32) * be).astype(np.float32)
    # float32 FMA through float64: the product of a float32 and an 8-bit integer is exact in float64 and the sum
    # with a float32 stays within 53 bits here, so one final rounding equals the hardware FMA
    t = (a.astype(np.float64) * np.float64(al) + p2.astype(np.float64)).astype(np.float32)
    return np.clip(np.rint(t), 0, 255).astype(np.uint8)


def unsharp_mask(img: np.ndarray, amount: float, threshold: int, ksize: int, sigma: float, gaussian_blur) -> np.ndarray:
    """apply_unsharp_mask (xy_blend_processor.py:48-67).  `gaussian_blur` = vsb_oracle.gaussian_blur."""
    blurred = gaussian_blur(img, ksize, ksize, sigma, sigma)
    sharp = add_weighted_u8(img, 1.0 + amount, blurred, -amount)
    if threshold > 0:
        diff = np.abs(img.astype(np.int16) - sharp.astype(np.int16))
        return np.where(diff > threshold, sharp, img).astype(np.uint8)   # THRESH_BINARY_INV mask==255 keeps the original
    return sharp


# --------------------------------------------------------------------------------------
# resize                                                       xy_blend_processor.py:69-90
# --------------------------------------------------------------------------------------
def resize_dims(cur_w: int, cur_h: int, width: Optional[int], height: Optional[int]) -> Optional[Tuple[int, int]]:
    """Target size rule of apply_resize (:71-82); None = return the input unchanged."""
    if width is None and height is None:
        return None
    if width is None:
        width = int(cur_w * (height / cur_h))
    elif height is None:
        height = int(cur_h * (width / cur_w))
    if width == cur_w and height == cur_h:
        return None
    return int(width), int(height)


def _rnd_short(c: np.ndarray) -> np.ndarray:
    return np.clip(np.rint(c), -32768, 32767).astype(np.int64)


def _scale(ssz: int, dsz: int) -> Tuple[float, float]:
    inv = dsz / ssz               # inv_scale = (double)dsize / ssize
    return inv, 1.0 / inv         # scale = 1. / inv_scale


def nearest_tab(ssz: int, dsz: int) -> np.ndarray:
    _, scale = _scale(ssz, dsz)
    return np.minimum(np.floor(np.arange(dsz) * scale).astype(np.int64), ssz - 1)


def _frac_tab(ssz: int, dsz: int, area_up: bool):
    inv, scale = _scale(ssz, dsz)
    d = np.arange(dsz)
    if not area_up:
        f = ((d + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int64)
        f = (f - s.astype(np.float32)).astype(np.float32)
    else:   # INTER_AREA when enlarging: linear interpolation with the area-style coordinate mapping
        s = np.floor(d * scale).astype(np.int64)
        f = ((d + 1) - (s + 1) * inv).astype(np.float32)
        f = np.where(f <= 0, f32(0), f - np.floor(f)).astype(np.float32)
    return s, f


def linear_tabs(ssz: int, dsz: int, axis: str, area_up: bool = False):
    """2-tap tables.  The x table clamps the fraction at both image edges; the y table keeps the fraction and
    clamps the row indices instead (that asymmetry is OpenCV's and is visible in the first / last output rows)."""
    s, f = _frac_tab(ssz, dsz, area_up)
    if axis == "x":
        lo = s < 0
        f = np.where(lo, f32(0), f); s = np.where(lo, 0, s)
        hi = s >= ssz - 1
        f = np.where(hi, f32(0), f); s = np.where(hi, ssz - 1, s)
    i0 = np.clip(s, 0, ssz - 1); i1 = np.clip(s + 1, 0, ssz - 1)
    w0 = (f32(1) - f).astype(np.float32); w1 = f.astype(np.float32)
    return i0, i1, w0, w1


def cubic_coeffs(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.float32); A = f32(-0.75); one = f32(1)
    x1 = (x + one).astype(np.float32)
    c0 = ((A * x1 - f32(5) * A) * x1 + f32(8) * A) * x1 - f32(4) * A
    c1 = ((A + f32(2)) * x - (A + f32(3))) * x * x + one
    xm = (one - x).astype(np.float32)
    c2 = ((A + f32(2)) * xm - (A + f32(3))) * xm * xm + one
    c3 = one - c0 - c1 - c2
    return np.stack([c0, c1, c2, c3], -1).astype(np.float32)


def lanczos4_coeffs(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, np.float32)
    s45 = 0.70710678118654752440084436210485
    cs = ((1, 0), (-s45, -s45), (0, 1), (s45, -s45), (-1, 0), (s45, s45), (0, -1), (-s45, s45))
    y0 = -(x + f32(3)).astype(np.float32).astype(np.float64) * np.pi * 0.25
    s0, c0 = np.sin(y0), np.cos(y0)
    co = np.zeros(x.shape + (8,), np.float32)
    sm = np.zeros(x.shape, np.float32)
    for i in range(8):
        y0_ = (x + f32(3) - f32(i)).astype(np.float32)
        y = -y0_.astype(np.float64) * np.pi * 0.25
        with np.errstate(divide="ignore", invalid="ignore"):
            v = ((cs[i][0] * s0 + cs[i][1] * c0) / (y * y)).astype(np.float32)
        v = np.where(np.abs(y0_) >= f32(1e-6), v, f32(1e30)).astype(np.float32)
        co[..., i] = v
        sm = (sm + v).astype(np.float32)
    sm = (f32(1) / sm).astype(np.float32)
    return (co * sm[..., None]).astype(np.float32)


def kernel_tabs(ssz: int, dsz: int, ks: int):
    """ks-tap (4 cubic / 8 Lanczos) index + Q11 coefficient tables; indices clamp to the image (replicate)."""
    s, f = _frac_tab(ssz, dsz, False)
    co = _rnd_short((cubic_coeffs(f) if ks == 4 else lanczos4_coeffs(f)) * f32(COEF_ONE))
    idx = np.clip(s[:, None] - (ks // 2 - 1) + np.arange(ks)[None, :], 0, ssz - 1)
    return idx, co


def area_tab(ssz: int, dsz: int):
    """computeResizeAreaTab: (dst index, src index, float32 weight) triples in generation order."""
    _, scale = _scale(ssz, dsz)
    di, si, al = [], [], []
    for dx in range(dsz):
        fsx1 = dx * scale; fsx2 = fsx1 + scale; cw = min(scale, ssz - fsx1)
        sx1 = int(np.ceil(fsx1)); sx2 = int(np.floor(fsx2))
        sx2 = min(sx2, ssz - 1); sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            di.append(dx); si.append(sx1 - 1); al.append(f32((sx1 - fsx1) / cw))
        for sx in range(sx1, sx2):
            di.append(dx); si.append(sx); al.append(f32(1.0 / cw))
        if fsx2 - sx2 > 1e-3:
            di.append(dx); si.append(sx2); al.append(f32(min(min(fsx2 - sx2, 1.0), cw) / cw))
    return np.array(di, np.int64), np.array(si, np.int64), np.array(al, np.float32)


def _csr(di: np.ndarray, n: int) -> np.ndarray:
    start = np.zeros(n + 1, np.int64)
    np.add.at(start, di + 1, 1)
    return np.cumsum(start)


def _sat_u8(v) -> np.ndarray:
    return np.clip(v, 0, 255).astype(np.uint8)


def resize_nearest(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    sh, sw = src.shape
    return src[nearest_tab(sh, dh)][:, nearest_tab(sw, dw)]


def resize_linear_u8(src: np.ndarray, dw: int, dh: int, area_up: bool = False) -> np.ndarray:
    sh, sw = src.shape
    x0, x1, a0, a1 = linear_tabs(sw, dw, "x", area_up)
    y0, y1, b0, b1 = linear_tabs(sh, dh, "y", area_up)
    ia0, ia1 = _rnd_short(a0 * f32(COEF_ONE)), _rnd_short(a1 * f32(COEF_ONE))
    ib0, ib1 = _rnd_short(b0 * f32(COEF_ONE)), _rnd_short(b1 * f32(COEF_ONE))
    S = src.astype(np.int64)
    Hr = S[:, x0] * ia0[None, :] + S[:, x1] * ia1[None, :]
    out = (((ib0[:, None] * (Hr[y0] >> 4)) >> 16) + ((ib1[:, None] * (Hr[y1] >> 4)) >> 16) + 2) >> 2
    return _sat_u8(out)


def resize_linear_f32(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    sh, sw = src.shape
    x0, x1, a0, a1 = linear_tabs(sw, dw, "x")
    y0, y1, b0, b1 = linear_tabs(sh, dh, "y")
    S = src.astype(np.float32)
    Hr = ((S[:, x0] * a0[None, :]).astype(np.float32) + (S[:, x1] * a1[None, :]).astype(np.float32)).astype(np.float32)
    return ((Hr[y0] * b0[:, None]).astype(np.float32) + (Hr[y1] * b1[:, None]).astype(np.float32)).astype(np.float32)


def _wrap32(v: np.ndarray) -> np.ndarray:
    return ((v + (1 << 31)) % (1 << 32)) - (1 << 31)   # int arithmetic of the C code wraps


def resize_kernel_u8(src: np.ndarray, dw: int, dh: int, ks: int) -> np.ndarray:
    sh, sw = src.shape
    xi, xa = kernel_tabs(sw, dw, ks)
    yi, yb = kernel_tabs(sh, dh, ks)
    S = src.astype(np.int64)
    Hr = sum(S[:, xi[:, k]] * xa[None, :, k] for k in range(ks))
    acc = np.zeros((dh, dw), np.int64)
    for k in range(ks):
        acc = _wrap32(acc + _wrap32(Hr[yi[:, k]] * yb[:, k][:, None]))
    fixed = _sat_u8(_wrap32(acc + (1 << 21)) >> 22)
    if ks != 4:
        return fixed
    # cubic: the vertical pass of whole 8-pixel groups runs in float32 (beta / 2^22, nested multiply-adds without
    # FMA, round half even); the remaining dw % 8 columns use the fixed-point expression above
    scale = f32(1.0) / f32(COEF_ONE * COEF_ONE)
    b = (yb.astype(np.float32) * scale).astype(np.float32)
    Sf = [Hr[yi[:, k]].astype(np.float32) for k in range(4)]
    t = (Sf[3] * b[:, 3][:, None]).astype(np.float32)
    for k in (2, 1, 0):
        t = ((Sf[k] * b[:, k][:, None]).astype(np.float32) + t).astype(np.float32)
    out = _sat_u8(np.rint(t))
    n = dw - dw % 8
    out[:, n:] = fixed[:, n:]
    return out


def resize_area_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    sh, sw = src.shape
    if dw > sw or dh > sh:   # enlarging along either axis: 2-tap interpolation with area coordinates
        return resize_linear_u8(src, dw, dh, area_up=True)
    if sw % dw == 0 and sh % dh == 0:   # integer factors: box sums
        ix, iy = sw // dw, sh // dh
        S = src.astype(np.int64).reshape(dh, iy, dw, ix).sum((1, 3))
        if ix == 2 and iy == 2:
            return ((S + 2) >> 2).astype(np.uint8)
        sc = f32(1.0) / f32(ix * iy)
        return _sat_u8(np.rint((S.astype(np.float32) * sc).astype(np.float32)))
    xd, xs, xa = area_tab(sw, dw)
    yd, ys, yb = area_tab(sh, dh)
    xstart = _csr(xd, dw)
    nmax = int((xstart[1:] - xstart[:-1]).max())
    S = src.astype(np.float32)
    out = np.zeros((dh, dw), np.uint8)
    acc = None
    prev = -1
    cols = np.arange(dw)
    for j in range(len(yd)):
        row = S[ys[j]]
        buf = np.zeros(dw, np.float32)
        for k in range(nmax):
            e = xstart[:-1] + k
            ok = e < xstart[1:]
            e = np.where(ok, e, 0)
            term = (row[xs[e]] * xa[e]).astype(np.float32)
            buf = np.where(ok, (buf + term).astype(np.float32), buf)
        if yd[j] != prev:
            if prev >= 0:
                out[prev] = _sat_u8(np.rint(acc))
            acc = (yb[j] * buf).astype(np.float32)
            prev = int(yd[j])
        else:
            acc = (acc + (yb[j] * buf).astype(np.float32)).astype(np.float32)
    if prev >= 0:
        out[prev] = _sat_u8(np.rint(acc))
    _ = cols
    return out


RESAMPLE_MODES = ("NEAREST", "BILINEAR", "BICUBIC", "LANCZOS4", "AREA")


def resize_u8(src: np.ndarray, dw: int, dh: int, mode: str) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=flags.get(mode.upper(), INTER_LANCZOS4)) on uint8 (:84-90)."""
    m = mode.upper()
    if m == "NEAREST":
        return resize_nearest(src, dw, dh)
    if m == "BILINEAR":
        return resize_linear_u8(src, dw, dh)
    if m == "BICUBIC":
        return resize_kernel_u8(src, dw, dh, 4)
    if m == "AREA":
        return resize_area_u8(src, dw, dh)
    return resize_kernel_u8(src, dw, dh, 8)   # LANCZOS4 and the fallback for unknown names


def apply_resize(img: np.ndarray, width: Optional[int], height: Optional[int], mode: str) -> np.ndarray:
    dims = resize_dims(img.shape[1], img.shape[0], width, height)
    if dims is None:
        return img
    return resize_u8(img, dims[0], dims[1], mode)


# --------------------------------------------------------------------------------------
# anisotropic correction of the Enhanced-EDT distance map          processing_core.py:357-377
# --------------------------------------------------------------------------------------
def anisotropic_dims(w: int, h: int, x_factor: float, y_factor: float) -> Optional[Tuple[int, int]]:
    xf = max(0.1, min(10.0, x_factor)); yf = max(0.1, min(10.0, y_factor))
    if xf == 1.0 and yf == 1.0:
        return None
    return int(w * xf), int(h * yf)


def anisotropic_distance(cur_white: np.ndarray, x_factor: float, y_factor: float, dist_fn) -> Optional[np.ndarray]:
    """Distance map of the anisotropic branch: nearest-resize the mask, transform, bilinear-resize the float map
    back.  `dist_fn(mask_u8) -> float32 map` is the isotropic transform (vsb_oracle.chamfer5_unbounded or a reach
    clipped variant).  None when both clamped factors are 1 (the caller takes the plain path)."""
    h, w = cur_white.shape
    dims = anisotropic_dims(w, h, x_factor, y_factor)
    if dims is None:
        return None
    nw, nh = dims
    small = resize_nearest(cur_white, nw, nh)
    return resize_linear_f32(dist_fn(small), w, h)
