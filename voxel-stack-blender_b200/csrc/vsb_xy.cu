// vsb_xy.cu -- K4: the XY operations of xy_blend_processor.py on uint8 layers.
//
//   vsb_lut_apply     lut[img]                                   lut_manager.py:57-63
//   vsb_merge_max     np.maximum(orig, grad)                     processing_core.py:458
//   vsb_gaussian_q88  cv2.GaussianBlur  (Q8.8 separable, exact)  xy_blend_processor.py:27-33
//   vsb_bilateral     cv2.bilateralFilter (float32, FMA order)   xy_blend_processor.py:35-40
//   vsb_median        cv2.medianBlur (exact, replicate border)   xy_blend_processor.py:42-46
//
// All are 128x32 tiles with the halo staged in shared memory.  Bilateral and median are bound by
// FP32/LDS issue, not HBM, so both skip every pixel whose whole (2r+1)^2 window is one value -- the
// filter output is then exactly that value (sum v*w / sum w == v; median of equal values) -- and run
// the real filter on a compacted list of the remaining pixels.
#include "vsb_common.cuh"

// ------------------------------------------------------------------------------------------------
// elementwise
// ------------------------------------------------------------------------------------------------
template <bool ALIGNED>
__global__ void __launch_bounds__(256) k_lut_apply(const uint8_t *__restrict__ in, size_t ipitch, int W, int H,
                                                   const uint8_t *__restrict__ lut, uint8_t *__restrict__ out,
                                                   size_t opitch)
{
    __shared__ __align__(4) uint8_t s_lut[256];
    VsbLut L = vsb_lut_stage(lut, s_lut);
    __syncthreads();
    const int cx = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int x = cx * 16;
    if (y >= H || x >= W) return;
    const int n = min(16, W - x);
    const uint8_t *src = in + (size_t)y * ipitch + x;
    uint8_t *dst = out + (size_t)y * opitch + x;
    if (ALIGNED && n == 16) vsb_st16_stream(dst, L.map16(vsb_ld16_stream(src)));
    else vsb_st_bytes(dst, L.map16(vsb_ld_bytes(src, n)), n);
}

template <bool ALIGNED>
__global__ void __launch_bounds__(256) k_merge_max(const uint8_t *__restrict__ a, size_t apitch,
                                                   const uint8_t *__restrict__ b, size_t bpitch, int W, int H,
                                                   uint8_t *__restrict__ out, size_t opitch)
{
    const int cx = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int x = cx * 16;
    if (y >= H || x >= W) return;
    const int n = min(16, W - x);
    uint4 va, vb;
    if (ALIGNED && n == 16) {
        va = vsb_ld16_stream(a + (size_t)y * apitch + x);
        vb = vsb_ld16_stream(b + (size_t)y * bpitch + x);
    } else {
        va = vsb_ld_bytes(a + (size_t)y * apitch + x, n);
        vb = vsb_ld_bytes(b + (size_t)y * bpitch + x, n);
    }
    uint4 m = make_uint4(__vmaxu4(va.x, vb.x), __vmaxu4(va.y, vb.y), __vmaxu4(va.z, vb.z), __vmaxu4(va.w, vb.w));
    uint8_t *dst = out + (size_t)y * opitch + x;
    if (ALIGNED && n == 16) vsb_st16_stream(dst, m);
    else vsb_st_bytes(dst, m, n);
}

static dim3 grid16(int W, int H) { return dim3(((W + 15) / 16 + 63) / 64, (H + 3) / 4); }

extern "C" int vsb_lut_apply(const uint8_t *in, size_t in_pitch, int W, int H, const uint8_t *lut256_dev,
                             uint8_t *out, size_t out_pitch, vsb_stream_t stream)
{
    VSB_REQUIRE(in && out && lut256_dev, "vsb_lut_apply: null pointer");
    VSB_REQUIRE(W > 0 && H > 0 && in_pitch >= (size_t)W && out_pitch >= (size_t)W, "vsb_lut_apply: bad geometry");
    cudaStream_t s = (cudaStream_t)stream;
    if (vsb_aligned16(in, in_pitch) && vsb_aligned16(out, out_pitch))
        k_lut_apply<true><<<grid16(W, H), 256, 0, s>>>(in, in_pitch, W, H, lut256_dev, out, out_pitch);
    else
        k_lut_apply<false><<<grid16(W, H), 256, 0, s>>>(in, in_pitch, W, H, lut256_dev, out, out_pitch);
    return vsb_check_launch("k_lut_apply");
}

extern "C" int vsb_merge_max(const uint8_t *a, size_t a_pitch, const uint8_t *b, size_t b_pitch, int W, int H,
                             uint8_t *out, size_t out_pitch, vsb_stream_t stream)
{
    VSB_REQUIRE(a && b && out, "vsb_merge_max: null pointer");
    VSB_REQUIRE(W > 0 && H > 0 && a_pitch >= (size_t)W && b_pitch >= (size_t)W && out_pitch >= (size_t)W,
                "vsb_merge_max: bad geometry");
    cudaStream_t s = (cudaStream_t)stream;
    if (vsb_aligned16(a, a_pitch) && vsb_aligned16(b, b_pitch) && vsb_aligned16(out, out_pitch))
        k_merge_max<true><<<grid16(W, H), 256, 0, s>>>(a, a_pitch, b, b_pitch, W, H, out, out_pitch);
    else
        k_merge_max<false><<<grid16(W, H), 256, 0, s>>>(a, a_pitch, b, b_pitch, W, H, out, out_pitch);
    return vsb_check_launch("k_merge_max");
}

// ------------------------------------------------------------------------------------------------
// halo tile staging: rows y0-ry .. y0+TH-1+ry, columns x0-hl .. x0+TW-1+hl (hl = halo rounded up to a
// multiple of 4 so the staged row starts word-aligned), border REFLECT_101 (mode 0) or REPLICATE (1).
// s_tile pitch = sp bytes (multiple of 4).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_u8_tile(uint8_t *s_tile, int sp, const uint8_t *__restrict__ in, size_t pitch,
                                              int W, int H, int x0, int y0, int hl, int ry, int border, bool word_ok)
{
    const int rows = VSB_TH + 2 * ry;
    const int cols = VSB_TW + 2 * hl;
    const int xs = x0 - hl, ys = y0 - ry;
    const bool inside = xs >= 0 && ys >= 0 && xs + cols <= W && ys + rows <= H;
    if (inside && word_ok) {
        const int wcols = cols >> 2;
        for (int i = threadIdx.x; i < rows * wcols; i += blockDim.x) {
            const int rr = i / wcols, wc = i - rr * wcols;
            ((uint32_t *)(s_tile + rr * sp))[wc] = *(const uint32_t *)(in + (size_t)(ys + rr) * pitch + xs + 4 * wc);
        }
    } else {
        for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) {
            const int rr = i / cols, cc = i - rr * cols;
            int yy = ys + rr, xx = xs + cc;
            if (border == 0) { yy = vsb_reflect101(yy, H); xx = vsb_reflect101(xx, W); }
            else { yy = vsb_clampi(yy, H); xx = vsb_clampi(xx, W); }
            s_tile[rr * sp + cc] = in[(size_t)yy * pitch + xx];
        }
    }
}

// write the 128x32 result tile held in shared memory (pitch VSB_TW) to the image
__device__ __forceinline__ void store_u8_tile(const uint8_t *s_out, uint8_t *__restrict__ out, size_t opitch, int W,
                                              int H, int x0, int y0, bool aligned)
{
    const int t = threadIdx.x;
    const int r = t >> 3, c = t & 7;
    const int y = y0 + r, x = x0 + c * 16;
    if (y >= H || x >= W) return;
    const int n = min(16, W - x);
    const uint4 v = ((const uint4 *)s_out)[t];
    uint8_t *dst = out + (size_t)y * opitch + x;
    if (aligned && n == 16) vsb_st16_stream(dst, v);
    else vsb_st_bytes(dst, v, n);
}

// ------------------------------------------------------------------------------------------------
// Gaussian, Q8.8 separable
// ------------------------------------------------------------------------------------------------
struct GaussParams {
    int32_t kx[63], ky[63];
    int nkx, nky;
};

__global__ void __launch_bounds__(VSB_THREADS) k_gaussian(const uint8_t *__restrict__ in, size_t ipitch, int W, int H,
                                                          const __grid_constant__ GaussParams g,
                                                          uint8_t *__restrict__ out, size_t opitch, int word_ok,
                                                          int aligned)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int rx = g.nkx >> 1, ry = g.nky >> 1;
    const int hl = (rx + 3) & ~3;
    const int sp = VSB_TW + 2 * hl;
    const int rows = VSB_TH + 2 * ry;
    uint8_t *s_tile = smem;                                              // rows x sp
    uint16_t *s_h = (uint16_t *)(smem + (((size_t)rows * sp + 15) & ~(size_t)15)); // rows x TW
    uint8_t *s_out = (uint8_t *)(s_h + (size_t)rows * VSB_TW);           // TH x TW
    const int x0 = blockIdx.x * VSB_TW, y0 = blockIdx.y * VSB_TH;
    stage_u8_tile(s_tile, sp, in, ipitch, W, H, x0, y0, hl, ry, 0, word_ok != 0);
    __syncthreads();
    for (int i = threadIdx.x; i < rows * VSB_TW; i += VSB_THREADS) {
        const int rr = i >> 7, cc = i & 127;
        const uint8_t *p = s_tile + rr * sp + cc + hl - rx;
        uint32_t s = 0;
        for (int k = 0; k < g.nkx; ++k) s += (uint32_t)g.kx[k] * p[k];
        s_h[i] = (uint16_t)s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < VSB_TH * VSB_TW; i += VSB_THREADS) {
        const int rr = i >> 7, cc = i & 127;
        uint32_t s = 0;
        for (int k = 0; k < g.nky; ++k) s += (uint32_t)g.ky[k] * s_h[(rr + k) * VSB_TW + cc];
        const uint32_t v = (s + 32768u) >> 16;
        s_out[i] = (uint8_t)min(v, 255u);
    }
    __syncthreads();
    store_u8_tile(s_out, out, opitch, W, H, x0, y0, aligned != 0);
}

extern "C" int vsb_gaussian_q88(const uint8_t *in, size_t in_pitch, int W, int H, const int32_t *kx, int nkx,
                                const int32_t *ky, int nky, uint8_t *out, size_t out_pitch, vsb_stream_t stream)
{
    VSB_REQUIRE(in && out && kx && ky && in != out, "vsb_gaussian_q88: null pointer / in-place");
    VSB_REQUIRE(W > 0 && H > 0 && in_pitch >= (size_t)W && out_pitch >= (size_t)W, "vsb_gaussian_q88: bad geometry");
    VSB_REQUIRE(nkx >= 1 && nkx <= 63 && (nkx & 1) && nky >= 1 && nky <= 63 && (nky & 1),
                "vsb_gaussian_q88: kernel sizes must be odd and <= 63");
    GaussParams g;
    int sx = 0, sy = 0;
    for (int i = 0; i < 63; ++i) {
        g.kx[i] = i < nkx ? kx[i] : 0;
        g.ky[i] = i < nky ? ky[i] : 0;
        sx += g.kx[i]; sy += g.ky[i];
    }
    VSB_REQUIRE(sx == 256 && sy == 256, "vsb_gaussian_q88: Q8.8 taps must sum to 256 (got %d, %d)", sx, sy);
    g.nkx = nkx; g.nky = nky;
    const int rx = nkx >> 1, ry = nky >> 1, hl = (rx + 3) & ~3;
    const int rows = VSB_TH + 2 * ry, sp = VSB_TW + 2 * hl;
    const size_t smem = (((size_t)rows * sp + 15) & ~(size_t)15) + (size_t)rows * VSB_TW * 2 + VSB_TH * VSB_TW;
    cudaStream_t s = (cudaStream_t)stream;
    if (smem > 48 * 1024)
        VSB_CUDA(cudaFuncSetAttribute(k_gaussian, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((W + VSB_TW - 1) / VSB_TW, (H + VSB_TH - 1) / VSB_TH);
    const int word_ok = ((((uintptr_t)in) & 3u) == 0 && (in_pitch & 3u) == 0) ? 1 : 0;
    k_gaussian<<<grid, VSB_THREADS, smem, s>>>(in, in_pitch, W, H, g, out, out_pitch, word_ok,
                                                vsb_aligned16(out, out_pitch) ? 1 : 0);
    return vsb_check_launch("k_gaussian");
}

// ------------------------------------------------------------------------------------------------
// window filters with the uniform-neighbourhood bypass (bilateral, median)
// ------------------------------------------------------------------------------------------------
struct BilTaps {
    int8_t dy[256], dx[256];
    float sw[256];
    float cw[256];
};

struct WinParams {
    const uint8_t *in;
    size_t ipitch;
    uint8_t *out;
    size_t opitch;
    int W, H;
    int r;       // window radius
    int ntaps;   // bilateral
    int border;  // 0 reflect101, 1 replicate
    int word_ok, aligned;
    const int8_t *tdy, *tdx; // bilateral taps in global memory when ntaps > 256
    const float *tsw;
};

// Shared staging + bypass detection.  Returns the number of pixels that need the real filter; their
// codes (row<<7|col) are in s_list.  s_out already holds the centre value for every bypassed pixel.
__device__ __forceinline__ int window_prepare(const WinParams &p, uint8_t *s_tile, int sp, int hl, uint16_t *s_wf,
                                              uint16_t *s_vf, uint8_t *s_out, uint16_t *s_list, int *s_cnt, int x0,
                                              int y0)
{
    const int r = p.r;
    const int rows = VSB_TH + 2 * r;
    const int wcols = sp >> 2;
    stage_u8_tile(s_tile, sp, p.in, p.ipitch, p.W, p.H, x0, y0, hl, r, p.border, p.word_ok != 0);
    if (threadIdx.x == 0) *s_cnt = 0;
    __syncthreads();
    // per staged 4-byte word: its value if the 4 bytes are equal, else 0x100
    for (int i = threadIdx.x; i < rows * wcols; i += VSB_THREADS) {
        const int rr = i / wcols, wc = i - rr * wcols;
        const uint32_t w = ((const uint32_t *)(s_tile + rr * sp))[wc];
        const uint32_t b0 = w & 255u;
        s_wf[i] = (w == b0 * 0x01010101u) ? (uint16_t)b0 : (uint16_t)0x100;
    }
    __syncthreads();
    // per tile row and word column: the value if the word is uniform over rows -r..+r, else 0x100
    for (int i = threadIdx.x; i < VSB_TH * wcols; i += VSB_THREADS) {
        const int rr = i / wcols, wc = i - rr * wcols;
        uint32_t v = s_wf[rr * wcols + wc];
        for (int k = 1; k <= 2 * r; ++k) v = (s_wf[(rr + k) * wcols + wc] == v) ? v : 0x100u;
        s_vf[i] = (uint16_t)v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < VSB_TH * VSB_TW; i += VSB_THREADS) {
        const int pr = i >> 7, pc = i & 127;
        const int c = pc + hl;
        const uint32_t v0 = s_tile[(pr + r) * sp + c];
        bool same = true;
        for (int wc = (c - r) >> 2; wc <= (c + r) >> 2; ++wc) same = same && (s_vf[pr * wcols + wc] == v0);
        s_out[i] = (uint8_t)v0;
        if (!same) {
            const int slot = atomicAdd(s_cnt, 1);
            s_list[slot] = (uint16_t)i;
        }
    }
    __syncthreads();
    return *s_cnt;
}

template <bool TAPS_IN_PARAM>
__global__ void __launch_bounds__(VSB_THREADS) k_bilateral(const __grid_constant__ WinParams p,
                                                           const __grid_constant__ BilTaps taps)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_cnt;
    __shared__ float s_cw[256];
    const int r = p.r, hl = (r + 3) & ~3, sp = VSB_TW + 2 * hl, rows = VSB_TH + 2 * r, wcols = sp >> 2;
    uint8_t *s_tile = smem;
    size_t off = ((size_t)rows * sp + 15) & ~(size_t)15;
    uint8_t *s_out = smem + off;            off += VSB_TH * VSB_TW;
    uint16_t *s_list = (uint16_t *)(smem + off); off += VSB_TH * VSB_TW * 2;
    uint16_t *s_wf = (uint16_t *)(smem + off);   off += (((size_t)rows * wcols * 2) + 15) & ~(size_t)15;
    uint16_t *s_vf = (uint16_t *)(smem + off);
    const int x0 = blockIdx.x * VSB_TW, y0 = blockIdx.y * VSB_TH;
    s_cw[threadIdx.x] = taps.cw[threadIdx.x];
    const int n = window_prepare(p, s_tile, sp, hl, s_wf, s_vf, s_out, s_list, &s_cnt, x0, y0);
    for (int i = threadIdx.x; i < n; i += VSB_THREADS) {
        const int code = s_list[i];
        const int pr = code >> 7, pc = code & 127;
        const uint8_t *c = s_tile + (pr + r) * sp + pc + hl;
        const int v0 = *c;
        float s = 0.0f, ws = 0.0f;
        for (int k = 0; k < p.ntaps; ++k) {
            int dy, dx;
            float swk;
            if (TAPS_IN_PARAM) { dy = taps.dy[k]; dx = taps.dx[k]; swk = taps.sw[k]; }
            else { dy = p.tdy[k]; dx = p.tdx[k]; swk = p.tsw[k]; }
            const int v = c[dy * sp + dx];
            const float w = __fmul_rn(swk, s_cw[abs(v - v0)]);
            s = __fmaf_rn((float)v, w, s);
            ws = __fadd_rn(ws, w);
        }
        const int o = __float2int_rn(__fdiv_rn(s, ws));
        s_out[code] = (uint8_t)min(max(o, 0), 255);
    }
    __syncthreads();
    store_u8_tile(s_out, p.out, p.opitch, p.W, p.H, x0, y0, p.aligned != 0);
}

__global__ void __launch_bounds__(VSB_THREADS) k_median(const __grid_constant__ WinParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_cnt;
    const int r = p.r, hl = (r + 3) & ~3, sp = VSB_TW + 2 * hl, rows = VSB_TH + 2 * r, wcols = sp >> 2;
    uint8_t *s_tile = smem;
    size_t off = ((size_t)rows * sp + 15) & ~(size_t)15;
    uint8_t *s_out = smem + off;            off += VSB_TH * VSB_TW;
    uint16_t *s_list = (uint16_t *)(smem + off); off += VSB_TH * VSB_TW * 2;
    uint16_t *s_wf = (uint16_t *)(smem + off);   off += (((size_t)rows * wcols * 2) + 15) & ~(size_t)15;
    uint16_t *s_vf = (uint16_t *)(smem + off);
    const int x0 = blockIdx.x * VSB_TW, y0 = blockIdx.y * VSB_TH;
    const int n = window_prepare(p, s_tile, sp, hl, s_wf, s_vf, s_out, s_list, &s_cnt, x0, y0);
    const int k = 2 * r + 1, need = (k * k) >> 1; // rank of the median, 0-based
    for (int i = threadIdx.x; i < n; i += VSB_THREADS) {
        const int code = s_list[i];
        const int pr = code >> 7, pc = code & 127;
        const uint8_t *c = s_tile + pr * sp + pc + hl - r; // top-left of the window
        // smallest v with count(window <= v) > need: bisection over the 8 value bits
        int lo = 0, hi = 255;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            int cnt = 0;
            for (int dy = 0; dy < k; ++dy) {
                const uint8_t *row = c + dy * sp;
                for (int dx = 0; dx < k; ++dx) cnt += (row[dx] <= mid);
            }
            if (cnt > need) hi = mid; else lo = mid + 1;
        }
        s_out[code] = (uint8_t)lo;
    }
    __syncthreads();
    store_u8_tile(s_out, p.out, p.opitch, p.W, p.H, x0, y0, p.aligned != 0);
}

static size_t window_smem(int r)
{
    const int hl = (r + 3) & ~3, sp = VSB_TW + 2 * hl, rows = VSB_TH + 2 * r, wcols = sp >> 2;
    size_t off = ((size_t)rows * sp + 15) & ~(size_t)15;
    off += VSB_TH * VSB_TW;
    off += VSB_TH * VSB_TW * 2;
    off += (((size_t)rows * wcols * 2) + 15) & ~(size_t)15;
    off += (size_t)VSB_TH * wcols * 2 + 16;
    return off;
}

static int fill_win(WinParams &p, const uint8_t *in, size_t in_pitch, int W, int H, uint8_t *out, size_t out_pitch,
                    int r, int border)
{
    VSB_REQUIRE(in && out && in != out, "window filter: null pointer / in-place");
    VSB_REQUIRE(W > 0 && H > 0 && in_pitch >= (size_t)W && out_pitch >= (size_t)W, "window filter: bad geometry");
    VSB_REQUIRE(r >= 1 && window_smem(r) <= 200 * 1024, "window filter: radius %d not supported", r);
    p.in = in; p.ipitch = in_pitch; p.out = out; p.opitch = out_pitch; p.W = W; p.H = H; p.r = r;
    p.ntaps = 0; p.border = border;
    p.word_ok = ((((uintptr_t)in) & 3u) == 0 && (in_pitch & 3u) == 0) ? 1 : 0;
    p.aligned = vsb_aligned16(out, out_pitch) ? 1 : 0;
    p.tdy = p.tdx = nullptr; p.tsw = nullptr;
    return VSB_OK;
}

extern "C" int vsb_bilateral(const uint8_t *in, size_t in_pitch, int W, int H, int radius, int ntaps,
                             const int8_t *dy, const int8_t *dx, const float *sw, const float *cw256, uint8_t *out,
                             size_t out_pitch, vsb_stream_t stream)
{
    WinParams p;
    int rc = fill_win(p, in, in_pitch, W, H, out, out_pitch, radius, 0);
    if (rc) return rc;
    VSB_REQUIRE(dy && dx && sw && cw256 && ntaps >= 1, "vsb_bilateral: tap tables missing");
    VSB_REQUIRE(radius <= 127, "vsb_bilateral: radius too large");
    p.ntaps = ntaps;
    BilTaps local; // passed by value as a kernel parameter
    BilTaps *tp = &local;
    for (int i = 0; i < 256; ++i) tp->cw[i] = cw256[i];
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = window_smem(radius);
    dim3 grid((W + VSB_TW - 1) / VSB_TW, (H + VSB_TH - 1) / VSB_TH);
    if (ntaps <= 256) {
        for (int i = 0; i < 256; ++i) {
            tp->dy[i] = i < ntaps ? dy[i] : 0;
            tp->dx[i] = i < ntaps ? dx[i] : 0;
            tp->sw[i] = i < ntaps ? sw[i] : 0.0f;
        }
        if (smem > 48 * 1024)
            VSB_CUDA(cudaFuncSetAttribute(k_bilateral<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_bilateral<true><<<grid, VSB_THREADS, smem, s>>>(p, *tp);
        return vsb_check_launch("k_bilateral");
    }
    // large tap sets: stage the tables in stream-ordered device memory
    for (int i = 0; i < 256; ++i) { tp->dy[i] = 0; tp->dx[i] = 0; tp->sw[i] = 0.0f; }
    int8_t *d_dy = nullptr, *d_dx = nullptr;
    float *d_sw = nullptr;
    VSB_CUDA(cudaMallocAsync((void **)&d_dy, ntaps, s));
    VSB_CUDA(cudaMallocAsync((void **)&d_dx, ntaps, s));
    VSB_CUDA(cudaMallocAsync((void **)&d_sw, (size_t)ntaps * 4, s));
    VSB_CUDA(cudaMemcpyAsync(d_dy, dy, ntaps, cudaMemcpyHostToDevice, s));
    VSB_CUDA(cudaMemcpyAsync(d_dx, dx, ntaps, cudaMemcpyHostToDevice, s));
    VSB_CUDA(cudaMemcpyAsync(d_sw, sw, (size_t)ntaps * 4, cudaMemcpyHostToDevice, s));
    VSB_CUDA(cudaStreamSynchronize(s)); // the host tables may be freed by the caller on return
    p.tdy = d_dy; p.tdx = d_dx; p.tsw = d_sw;
    if (smem > 48 * 1024)
        VSB_CUDA(cudaFuncSetAttribute(k_bilateral<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_bilateral<false><<<grid, VSB_THREADS, smem, s>>>(p, *tp);
    rc = vsb_check_launch("k_bilateral");
    cudaFreeAsync(d_dy, s); cudaFreeAsync(d_dx, s); cudaFreeAsync(d_sw, s);
    return rc;
}

extern "C" int vsb_median(const uint8_t *in, size_t in_pitch, int W, int H, int ksize, uint8_t *out, size_t out_pitch,
                          vsb_stream_t stream)
{
    VSB_REQUIRE(ksize >= 3 && (ksize & 1), "vsb_median: ksize must be odd and >= 3 (ksize<=1 is a no-op on the host)");
    WinParams p;
    int rc = fill_win(p, in, in_pitch, W, H, out, out_pitch, ksize >> 1, 1);
    if (rc) return rc;
    const size_t smem = window_smem(p.r);
    if (smem > 48 * 1024)
        VSB_CUDA(cudaFuncSetAttribute(k_median, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((W + VSB_TW - 1) / VSB_TW, (H + VSB_TH - 1) / VSB_TH);
    k_median<<<grid, VSB_THREADS, smem, (cudaStream_t)stream>>>(p);
    return vsb_check_launch("k_median");
}
