// vsb_tile.cuh -- staged bit-mask tile (128x32 pixels + 64-bit / R-row halo) and the row probe
// shared by the fused Z-blend kernel and the dense distance kernels.
#pragma once
#include "vsb_common.cuh"

#define HALO_WORDS 2                       // 64 mask bits either side of the tile
#define ROW_WORDS (4 + 2 * HALO_WORDS)     // 8 words = 256 bits per staged row
#define MAX_ROWS (VSB_TH + 2 * VSB_FAST_REACH)

// nearest set bit to bit position px in a staged 256-bit row, searched over px-63 .. px+63
__device__ __forceinline__ int nearest_dx(const uint32_t *rowp, int px)
{
    const int wi = px >> 5, b = px & 31; // wi in [2,5]
    const uint32_t w0 = rowp[wi - 2], w1 = rowp[wi - 1], w2 = rowp[wi], w3 = rowp[wi + 1], w4 = rowp[wi + 2];
    const uint32_t rlo = __funnelshift_r(w2, w3, b), rhi = __funnelshift_r(w3, w4, b);
    const uint32_t ltop = __funnelshift_l(w1, w2, 31 - b), lnext = __funnelshift_l(w0, w1, 31 - b);
    const int dr = rlo ? (__ffs(rlo) - 1) : (rhi ? 31 + __ffs(rhi) : 64);
    const int dl = ltop ? __clz(ltop) : (lnext ? 32 + __clz(lnext) : 64);
    return min(dl, dr);
}


// the same over px-31 .. px+31 only (three words instead of five): enough once the running minimum is <= 32, because a
// seed 32 or more columns away cannot undercut it (every metric here is >= max(|dx|, |dy|)); returns 64 when empty
__device__ __forceinline__ int nearest_dx_narrow(const uint32_t *rowp, int px)
{
    const int wi = px >> 5, b = px & 31;
    const uint32_t w1 = rowp[wi - 1], w2 = rowp[wi], w3 = rowp[wi + 1];
    const uint32_t rlo = __funnelshift_r(w2, w3, b), ltop = __funnelshift_l(w1, w2, 31 - b);
    const int dr = rlo ? (__ffs(rlo) - 1) : 64;
    const int dl = ltop ? __clz(ltop) : 64;
    return min(dl, dr);
}

// stage rows y0-R .. y0+TH-1+R, words (x0>>5)-2 .. (x0>>5)+5 of a packed mask; outside the image = 0
__device__ __forceinline__ void stage_mask_tile(uint32_t (*s_cur)[ROW_WORDS], const uint32_t *__restrict__ bits,
                                                int wpr, int H, int x0, int y0, int R)
{
    const int nrows = VSB_TH + 2 * R;
    const int wbase = (x0 >> 5) - HALO_WORDS;
    for (int i = threadIdx.x; i < nrows * ROW_WORDS; i += VSB_THREADS) {
        const int hr = i >> 3, hw = i & 7;
        const int yy = y0 - R + hr, wx = wbase + hw;
        uint32_t v = 0;
        if (yy >= 0 && yy < H && wx >= 0 && wx < wpr) v = bits[(size_t)yy * wpr + wx];
        s_cur[hr][hw] = v;
    }
}

// generic variant: `nrows` rows starting at image row ystart, RW words per row starting at word wbase
template <int RW>
__device__ __forceinline__ void stage_mask_rows(uint32_t (*s_cur)[RW], const uint32_t *__restrict__ bits, int wpr,
                                                int H, int wbase, int ystart, int nrows)
{
    const int total = nrows * RW;
    if (ystart >= 0 && ystart + nrows <= H && wbase >= 0 && wbase + RW <= wpr) { // inside the image: no checks, four requests in flight
        const uint32_t *__restrict__ src = bits + (size_t)ystart * wpr + wbase;
        uint32_t *dst = &s_cur[0][0];
        int i = threadIdx.x;
        for (; i + 3 * VSB_THREADS < total; i += 4 * VSB_THREADS) {
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = i + u * VSB_THREADS, hr = j / RW;
                v[u] = src[hr * wpr + (j - hr * RW)];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) dst[i + u * VSB_THREADS] = v[u];
        }
        for (; i < total; i += VSB_THREADS) {
            const int hr = i / RW;
            dst[i] = src[hr * wpr + (i - hr * RW)];
        }
        return;
    }
    for (int i = threadIdx.x; i < total; i += VSB_THREADS) {
        const int hr = i / RW, hw = i - hr * RW;
        const int yy = ystart + hr, wx = wbase + hw;
        uint32_t v = 0;
        if (yy >= 0 && yy < H && wx >= 0 && wx < wpr) v = bits[(size_t)yy * wpr + wx];
        s_cur[hr][hw] = v;
    }
}

template <int RW>
__device__ __forceinline__ void stage_rowany_rows(uint32_t *s_rowany, const uint32_t (*s_cur)[RW], int nrows)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nmask = (nrows + 31) / 32;
    if (warp < nmask) {
        const int hr = warp * 32 + lane;
        uint32_t v = 0;
        if (hr < nrows) {
#pragma unroll
            for (int k = 0; k < RW; ++k) v |= s_cur[hr][k];
        }
        const uint32_t m = __ballot_sync(0xffffffffu, v != 0);
        if (lane == 0) s_rowany[warp] = m;
    }
}

// distance from staged-row position (hr0, px) to the nearest set bit of the staged mask, either metric;
// returns a value >= 3e38 when nothing lies within reach R.  T monotone in |dx| per row => the nearest
// set bit of each row is that row's minimum; rows are visited outwards while |dy| < best.
// TRI: the table is stored as its lower triangle, entry (M, m) at M*(M+1)/2 + m; `stride`: floats per table row if not R + 1
template <int METRIC, int RW, bool TRI = false>
__device__ __forceinline__ float tile_search(const uint32_t (*s_cur)[RW], const uint32_t *s_rowany, const float *s_T,
                                             int R, int hr0, int px, int stride = 0)
{
    const int n1 = stride ? stride : R + 1; // floats per table row
    if (METRIC == VSB_METRIC_CHAMFER5) {
        float best = 3.0e38f;
        {
            const int dx = nearest_dx(s_cur[hr0], px);
            if (dx <= R) best = TRI ? s_T[(dx * (dx + 1)) >> 1] : s_T[dx * n1];
        }
        for (int k = 1; k <= R && (float)k < best; ++k) {
            const int ha = hr0 - k, hb = hr0 + k;
            if ((s_rowany[ha >> 5] >> (ha & 31)) & 1u) {
                const int dx = best <= 32.0f ? nearest_dx_narrow(s_cur[ha], px) : nearest_dx(s_cur[ha], px);
                if (dx <= R) { const int M = max(k, dx), m = min(k, dx); best = fminf(best, TRI ? s_T[((M * (M + 1)) >> 1) + m] : s_T[M * n1 + m]); }
            }
            if ((s_rowany[hb >> 5] >> (hb & 31)) & 1u) {
                const int dx = best <= 32.0f ? nearest_dx_narrow(s_cur[hb], px) : nearest_dx(s_cur[hb], px);
                if (dx <= R) { const int M = max(k, dx), m = min(k, dx); best = fminf(best, TRI ? s_T[((M * (M + 1)) >> 1) + m] : s_T[M * n1 + m]); }
            }
        }
        return best;
    } else {
        int best2 = 0x7fffffff;
        {
            const int dx = nearest_dx(s_cur[hr0], px);
            if (dx <= R) best2 = dx * dx;
        }
        for (int k = 1; k <= R && k * k < best2; ++k) {
            const int ha = hr0 - k, hb = hr0 + k;
            if ((s_rowany[ha >> 5] >> (ha & 31)) & 1u) {
                const int dx = nearest_dx(s_cur[ha], px);
                if (dx <= R) best2 = min(best2, k * k + dx * dx);
            }
            if ((s_rowany[hb >> 5] >> (hb & 31)) & 1u) {
                const int dx = nearest_dx(s_cur[hb], px);
                if (dx <= R) best2 = min(best2, k * k + dx * dx);
            }
        }
        return best2 == 0x7fffffff ? 3.0e38f : __fsqrt_rn((float)best2);
    }
}

// one bit per staged row: does the row hold any set bit (call after a __syncthreads())
__device__ __forceinline__ void stage_rowany(uint32_t *s_rowany, const uint32_t (*s_cur)[ROW_WORDS], int nrows)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nmask = (nrows + 31) / 32;
    if (warp < nmask) {
        const int hr = warp * 32 + lane;
        uint32_t v = 0;
        if (hr < nrows) {
#pragma unroll
            for (int k = 0; k < ROW_WORDS; ++k) v |= s_cur[hr][k];
        }
        const uint32_t m = __ballot_sync(0xffffffffu, v != 0);
        if (lane == 0) s_rowany[warp] = m;
    }
}

