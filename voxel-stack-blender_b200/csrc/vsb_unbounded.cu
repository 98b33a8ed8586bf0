// vsb_unbounded.cu -- distance transforms with unbounded reach, and the two consumers that need them:
//   * global min-max normalisation of the fade (use_fixed_fade_receding = False, the dataclass default):
//     processing_core.py:141-145  (cv2.minMaxLoc over the receding pixels, then (d-min)/(max-min))
//   * fixed fade distances above the 63-pixel reach of the tile kernels (the GUI allows F up to 1000)
//   * the dense exact Euclidean transform of the EDT-only configuration.
//
// Separable: (1) column pass -- vertical distance g(x,y) to the nearest white pixel of column x (uint16,
// 0xFFFF = none), one warp per 32-column word and 128-row band, the carry across bands found by walking the
// packed mask; (2) per-row minima of g over 64-column blocks; (3) row search D(x,y) = min_x' N(|x-x'|, g(x',y)),
// walking outwards from x and skipping whole 64-column blocks whose lower bound N(k, gmin) cannot improve.
//
// Chamfer values beyond the tabulated range are exact integers in units of 2^-23 (a = 2^23, b = 1.4f, c = 2.1969f
// as exact binary fractions); the float result is the correctly rounded value (oracle: orc_chamfer5_unbounded).
#include "vsb_common.cuh"

#define G_NONE 0xFFFFu
#define UB_BAND 128
#define CH_A 8388608LL                 // 1.0   * 2^23
#define CH_B 11744051LL                // 1.4f  * 2^23 (exact)
#define CH_C 18428932LL                // 2.1969f * 2^23 (exact: 0x400C9A02 -> mantissa 9214466 * 2^-22)
#define COST_INF 0x3fffffffffffffffLL

__device__ __forceinline__ long long chamfer_cost(int k, int g)
{
    const int M = max(k, g), m = min(k, g);
    return (2 * m <= M) ? (long long)(M - 2 * m) * CH_A + (long long)m * CH_C
                        : (long long)(2 * m - M) * CH_B + (long long)(M - m) * CH_C;
}

// ---- (1) column pass ------------------------------------------------------------------------------------
// first: per 128-row band and column, the offset of the first and of the last white row of the band (0xFFFF = none),
// so the carry across bands is a walk over at most H/128 band records instead of over rows
__global__ void __launch_bounds__(128) k_bandseeds(const uint32_t *__restrict__ bits, int wpr, int H,
                                                   uint16_t *__restrict__ btop, uint16_t *__restrict__ bbot)
{
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const int wx = blockIdx.x * 4 + wq;
    const int y0 = blockIdx.y * UB_BAND;
    if (wx >= wpr) return;
    uint32_t top = G_NONE, bot = G_NONE;
#pragma unroll
    for (int q = 0; q < UB_BAND / 32; ++q) {
        const int yy = y0 + 32 * q + lane;
        const uint32_t mine = yy < H ? bits[(size_t)yy * wpr + wx] : 0u;
        for (int j = 0; j < 32; ++j) {
            const uint32_t w = __shfl_sync(0xffffffffu, mine, j);
            if ((w >> lane) & 1u) { if (top == G_NONE) top = 32 * q + j; bot = 32 * q + j; }
        }
    }
    const size_t o = (size_t)blockIdx.y * wpr * 32 + (size_t)wx * 32 + lane;
    btop[o] = (uint16_t)top;
    bbot[o] = (uint16_t)bot;
}

__global__ void __launch_bounds__(128) k_coldist(const uint32_t *__restrict__ bits, int wpr, int W, int H,
                                                 const uint16_t *__restrict__ btop, const uint16_t *__restrict__ bbot,
                                                 int nbands, uint16_t *__restrict__ g, size_t gpitch)
{
    __shared__ uint16_t s_down[4][UB_BAND][32];
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const int wx = blockIdx.x * 4 + wq;
    const int band = blockIdx.y;
    const int y0 = band * UB_BAND;
    if (wx >= wpr) return;
    const int x = wx * 32 + lane;
    const int y1 = min(H, y0 + UB_BAND);
    const size_t bstride = (size_t)wpr * 32;
    // carry from above: distance from row y0-1 to the nearest white pixel at or above it
    uint32_t run = G_NONE;
    for (int bb = band - 1; bb >= 0; --bb) {
        const uint32_t v = bbot[(size_t)bb * bstride + x];
        if (v != G_NONE) { run = (uint32_t)min(y0 - 1 - (bb * UB_BAND + (int)v), 0xFFFD); break; }
    }
    uint32_t wv[UB_BAND / 32];
#pragma unroll
    for (int q = 0; q < UB_BAND / 32; ++q) {
        const int yy = y0 + 32 * q + lane;
        wv[q] = yy < H ? bits[(size_t)yy * wpr + wx] : 0u;
    }
    for (int yy = y0; yy < y1; ++yy) { // downward: distance to the nearest white pixel at or above row yy
        const int q = (yy - y0) >> 5;
        const uint32_t src = q == 0 ? wv[0] : (q == 1 ? wv[1] : (q == 2 ? wv[2] : wv[3]));
        const uint32_t w = __shfl_sync(0xffffffffu, src, (yy - y0) & 31);
        run = ((w >> lane) & 1u) ? 0u : (run == G_NONE ? G_NONE : min(run + 1u, 0xFFFEu));
        s_down[wq][yy - y0][lane] = (uint16_t)run;
    }
    run = G_NONE; // carry from below: distance from row y1 to the nearest white pixel at or below it
    for (int bb = band + 1; bb < nbands; ++bb) {
        const uint32_t v = btop[(size_t)bb * bstride + x];
        if (v != G_NONE) { run = (uint32_t)min(bb * UB_BAND + (int)v - y1, 0xFFFD); break; }
    }
    for (int yy = y1 - 1; yy >= y0; --yy) { // upward, combined with the downward result
        const int q = (yy - y0) >> 5;
        const uint32_t src = q == 0 ? wv[0] : (q == 1 ? wv[1] : (q == 2 ? wv[2] : wv[3]));
        const uint32_t w = __shfl_sync(0xffffffffu, src, (yy - y0) & 31);
        run = ((w >> lane) & 1u) ? 0u : (run == G_NONE ? G_NONE : min(run + 1u, 0xFFFEu));
        const uint32_t v = min(run, (uint32_t)s_down[wq][yy - y0][lane]);
        if (x < W) g[(size_t)yy * gpitch + x] = (uint16_t)v;
    }
}

// ---- (2) block minima -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_blockmin(const uint16_t *__restrict__ g, size_t gpitch, int W, int H,
                                                  uint16_t *__restrict__ gmin, int nblk)
{
    const int y = blockIdx.y;
    const int blk = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (blk >= nblk) return;
    const int x = blk * 64 + lane * 2;
    uint32_t v = G_NONE;
    if (x < W) v = g[(size_t)y * gpitch + x];
    if (x + 1 < W) v = min(v, (uint32_t)g[(size_t)y * gpitch + x + 1]);
    for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) gmin[(size_t)y * nblk + blk] = (uint16_t)v;
}

// ---- (3) row search ---------------------------------------------------------------------------------------------
// METRIC 0: chamfer cost in 2^-23 units; METRIC 1: squared Euclidean distance.  Returns COST_INF if no white pixel.
// exact metric in 32-bit arithmetic (d^2 < 2^31 for any image the uint16 column distances allow up to 46340 px)
__device__ __forceinline__ long long row_search_exact32(const uint16_t *__restrict__ grow, const uint16_t *__restrict__ gmrow,
                                                        int W, int x)
{
    const uint32_t INF32 = 0x7fffffffu;
    uint32_t best = INF32;
    {
        const uint32_t g0 = grow[x];
        if (g0 != G_NONE) best = min(g0, 46340u) * min(g0, 46340u);
    }
    for (int k = 1; x - k >= 0 && (uint32_t)k * (uint32_t)k < best && k < 46340;) {
        const int xx = x - k;
        if ((xx & 63) == 63) {
            const uint32_t gm = gmrow[xx >> 6];
            if (gm == G_NONE || gm >= 46340u || (uint32_t)k * k + gm * gm >= best) { k += 64; continue; }
        }
        const uint32_t gg = grow[xx];
        if (gg < 46340u) best = min(best, (uint32_t)k * k + gg * gg);
        ++k;
    }
    for (int k = 1; x + k < W && (uint32_t)k * (uint32_t)k < best && k < 46340;) {
        const int xx = x + k;
        if ((xx & 63) == 0) {
            const uint32_t gm = gmrow[xx >> 6];
            if (gm == G_NONE || gm >= 46340u || (uint32_t)k * k + gm * gm >= best) { k += 64; continue; }
        }
        const uint32_t gg = grow[xx];
        if (gg < 46340u) best = min(best, (uint32_t)k * k + gg * gg);
        ++k;
    }
    return best >= INF32 ? COST_INF : (long long)best;
}

template <int METRIC>
__device__ __forceinline__ long long row_search(const uint16_t *__restrict__ grow, const uint16_t *__restrict__ gmrow,
                                                int W, int x, long long cap)
{
    if (METRIC == 1 && W < 32768) return row_search_exact32(grow, gmrow, W, x);
    auto cost = [](int k, int gg) -> long long {
        return METRIC == 0 ? chamfer_cost(k, gg) : (long long)k * k + (long long)gg * gg;
    };
    auto lower = [](int k, int gg) -> long long { // bound for everything at horizontal distance >= k with g >= gg
        return METRIC == 0 ? (long long)max(k, gg) * CH_A : (long long)k * k + (long long)gg * gg;
    };
    long long best = cap;
    {
        const uint32_t g0 = grow[x];
        if (g0 != G_NONE) best = min(best, cost(0, (int)g0));
    }
    for (int k = 1; x - k >= 0 && lower(k, 0) < best;) { // leftwards
        const int xx = x - k;
        if ((xx & 63) == 63) {
            const uint32_t gm = gmrow[xx >> 6];
            if (gm == G_NONE || lower(k, (int)gm) >= best) { k += 64; continue; }
        }
        const uint32_t gg = grow[xx];
        if (gg != G_NONE) best = min(best, cost(k, (int)gg));
        ++k;
    }
    for (int k = 1; x + k < W && lower(k, 0) < best;) { // rightwards
        const int xx = x + k;
        if ((xx & 63) == 0) {
            const uint32_t gm = gmrow[xx >> 6];
            if (gm == G_NONE || lower(k, (int)gm) >= best) { k += 64; continue; }
        }
        const uint32_t gg = grow[xx];
        if (gg != G_NONE) best = min(best, cost(k, (int)gg));
        ++k;
    }
    return best;
}

template <int METRIC>
__device__ __forceinline__ float cost_to_float(long long c)
{
    if (c >= COST_INF) return 3.0e38f;
    return METRIC == 0 ? (float)((double)c * (1.0 / 8388608.0)) : __fsqrt_rn((float)c);
}

template <int METRIC>
__global__ void __launch_bounds__(256) k_dense_dist(const uint16_t *__restrict__ g, size_t gpitch,
                                                    const uint16_t *__restrict__ gmin, int nblk, int W, int H,
                                                    float *__restrict__ dist, int32_t *__restrict__ d2, size_t dpitch)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    const long long c = row_search<METRIC>(g + (size_t)y * gpitch, gmin + (size_t)y * nblk, W, x, COST_INF);
    if (dist) dist[(size_t)y * dpitch + x] = cost_to_float<METRIC>(c);
    if (d2) d2[(size_t)y * dpitch + x] = c >= 0x7fffffffLL ? 0x7fffffff : (int32_t)c;
}

struct PriorPtrsU {
    const uint32_t *p[VSB_MAX_PRIORS];
};

// distance at the receding pixels only (one CTA per 128x32 tile, compacted list); float bits of min and max over the
// receding pixels are accumulated in mm[0], mm[1] (caller initialises them to 0xFFFFFFFF / 0)
template <int METRIC>
__global__ void __launch_bounds__(256) k_rec_dist(const uint32_t *__restrict__ cur, PriorPtrsU pri, int np, int wpr,
                                                  const uint16_t *__restrict__ g, size_t gpitch,
                                                  const uint16_t *__restrict__ gmin, int nblk, int W, int H,
                                                  float *__restrict__ dist, size_t dpitch, uint32_t *__restrict__ mm)
{
    __shared__ uint16_t s_list[VSB_TH * VSB_TW];
    __shared__ int s_cnt;
    __shared__ uint32_t s_mn, s_mx;
    const int t = threadIdx.x;
    const int x0 = blockIdx.x * VSB_TW, y0 = blockIdx.y * VSB_TH;
    if (t == 0) { s_cnt = 0; s_mn = 0xFFFFFFFFu; s_mx = 0u; }
    __syncthreads();
    if (t < 128) {
        const int rr = t >> 2, wi = t & 3;
        const int yy = y0 + rr, wx = (x0 >> 5) + wi;
        if (yy < H && wx < wpr) {
            const size_t o = (size_t)yy * wpr + wx;
            uint32_t acc = 0;
            for (int i = 0; i < np; ++i) acc |= pri.p[i][o];
            uint32_t rw = acc & ~cur[o];
            if (rw) {
                int base = atomicAdd(&s_cnt, __popc(rw));
                while (rw) {
                    const int b = __ffs(rw) - 1;
                    rw &= rw - 1;
                    s_list[base++] = (uint16_t)((rr << 7) | (wi << 5) | b);
                }
            }
        }
    }
    __syncthreads();
    const int total = s_cnt;
    if (total == 0) return;
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
    for (int i = t; i < total; i += 256) {
        const int code = s_list[i];
        const int y = y0 + (code >> 7), x = x0 + (code & 127);
        const long long c = row_search<METRIC>(g + (size_t)y * gpitch, gmin + (size_t)y * nblk, W, x, COST_INF);
        const float d = cost_to_float<METRIC>(c);
        dist[(size_t)y * dpitch + x] = d;
        const uint32_t u = __float_as_uint(d);
        mn = min(mn, u); mx = max(mx, u);
    }
    for (int o = 16; o; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((t & 31) == 0) { atomicMin(&s_mn, mn); atomicMax(&s_mx, mx); }
    __syncthreads();
    if (t == 0 && mm) { atomicMin(&mm[0], s_mn); atomicMax(&mm[1], s_mx); }
}

// gradient from the distances at the receding pixels, merge with gray, LUT.  norm 0: fixed fade
// g = trunc(255*(1 - min(d,F)/den)); norm 1: global min-max g = trunc(255*(1 - (d-min)/(max-min))), 0 if max<=min;
// norm 2: weighted stack, acc += (1 - min(d,F_j)/den_j)*w_j over the priors the pixel recedes from, g = trunc(255*acc/W).
struct UFParams {
    const uint8_t *gray; size_t gpitch;
    uint8_t *out; size_t opitch;
    const uint32_t *cur; PriorPtrsU pri; int np, wpr;
    const float *dist; size_t dpitch;
    const uint32_t *mm;
    const uint8_t *lut;
    float F, den;
    int norm, W, H, aligned;
    float w[VSB_MAX_PRIORS], wf[VSB_MAX_PRIORS], wden[VSB_MAX_PRIORS], wtotal; // norm 2: weighted stack
};

__global__ void __launch_bounds__(256) k_fade_unbounded(const __grid_constant__ UFParams p)
{
    __shared__ __align__(4) uint8_t s_lut[256];
    VsbLut Lt = vsb_lut_stage(p.lut, s_lut);
    __syncthreads();
    const int cx = blockIdx.x * 64 + (threadIdx.x & 63);
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    const int x = cx * 16;
    if (y >= p.H || x >= p.W) return;
    const int n = min(16, p.W - x);
    const uint8_t *src = p.gray + (size_t)y * p.gpitch + x;
    uint4 gv = (p.aligned && n == 16) ? vsb_ld16_stream(src) : vsb_ld_bytes(src, n);
    const size_t o = (size_t)y * p.wpr + (x >> 5);
    uint32_t acc = 0;
    for (int i = 0; i < p.np; ++i) acc |= p.pri.p[i][o];
    uint32_t m = ((acc & ~p.cur[o]) >> (x & 16)) & 0xFFFFu;
    if (m) {
        float mn = 0.0f, den = p.den;
        bool dead = false;
        if (p.norm == 1) {
            const float fmn = __uint_as_float(p.mm[0]), fmx = __uint_as_float(p.mm[1]);
            dead = !(fmx > fmn);                       // max_val <= min_val -> zeros (processing_core.py:143-144)
            mn = fmn;
            den = (float)((double)fmx - (double)fmn);  // Python-float difference, then float32 array division
        }
        uint32_t gw[4] = {0, 0, 0, 0};
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const float d = p.dist[(size_t)y * p.dpitch + x + b];
            uint32_t g;
            if (p.norm == 2) {
                float a = 0.0f;
                const uint32_t ncur = ~p.cur[o];
                const int bit = (x & 16) + b;
                for (int j = 0; j < p.np; ++j)
                    if (p.w[j] > 0.0f && (((p.pri.p[j][o] & ncur) >> bit) & 1u))
                        a = __fadd_rn(a, __fmul_rn(__fsub_rn(1.0f, __fdiv_rn(fminf(d, p.wf[j]), p.wden[j])), p.w[j]));
                g = (uint32_t)(int)__fmul_rn(255.0f, __fdiv_rn(a, p.wtotal));
            } else {
                float nrm;
                if (p.norm == 1) nrm = __fdiv_rn(__fsub_rn(d, mn), den);
                else nrm = __fdiv_rn(fminf(d, p.F), den);
                g = dead ? 0u : (uint32_t)(int)__fmul_rn(255.0f, __fsub_rn(1.0f, nrm));
            }
            gw[b >> 2] |= (g & 255u) << ((b & 3) * 8);
        }
        gv = make_uint4(__vmaxu4(gv.x, gw[0]), __vmaxu4(gv.y, gw[1]), __vmaxu4(gv.z, gw[2]), __vmaxu4(gv.w, gw[3]));
    }
    gv = Lt.map16(gv);
    uint8_t *dst = p.out + (size_t)y * p.opitch + x;
    if (p.aligned && n == 16) vsb_st16_stream(dst, gv);
    else vsb_st_bytes(dst, gv, n);
}

// ---- C-ABI --------------------------------------------------------------------------------------------------------
extern "C" size_t vsb_coldist_scratch_elems(int W, int H, int wpr)
{
    const size_t nbands = (size_t)(H + UB_BAND - 1) / UB_BAND;
    return (size_t)H * ((W + 63) / 64) + 2 * nbands * (size_t)wpr * 32;
}

extern "C" int vsb_coldist(const uint32_t *seed_bits, int wpr, int W, int H, uint16_t *g, size_t g_pitch_elems,
                           uint16_t *gmin64, vsb_stream_t stream)
{
    VSB_REQUIRE(seed_bits && g && gmin64, "vsb_coldist: null pointer");
    VSB_REQUIRE(W > 0 && H > 0 && H < 65534 && wpr >= (W + 31) / 32 && g_pitch_elems >= (size_t)W, "vsb_coldist: bad geometry");
    cudaStream_t s = (cudaStream_t)stream;
    const int nbands = (H + UB_BAND - 1) / UB_BAND;
    const int nblk = (W + 63) / 64;
    // band records live behind the block minima in the same scratch buffer (see vsb_coldist_scratch_elems)
    uint16_t *btop = gmin64 + (size_t)H * nblk;
    uint16_t *bbot = btop + (size_t)nbands * wpr * 32;
    dim3 grid((wpr + 3) / 4, nbands);
    k_bandseeds<<<grid, 128, 0, s>>>(seed_bits, wpr, H, btop, bbot);
    int rc = vsb_check_launch("k_bandseeds");
    if (rc) return rc;
    k_coldist<<<grid, 128, 0, s>>>(seed_bits, wpr, W, H, btop, bbot, nbands, g, g_pitch_elems);
    rc = vsb_check_launch("k_coldist");
    if (rc) return rc;
    dim3 grid2((nblk + 7) / 8, H);
    k_blockmin<<<grid2, 256, 0, s>>>(g, g_pitch_elems, W, H, gmin64, nblk);
    return vsb_check_launch("k_blockmin");
}

extern "C" int vsb_dt_unbounded(const uint32_t *seed_bits, int wpr, int W, int H, int metric, uint16_t *g,
                                size_t g_pitch_elems, uint16_t *gmin64, float *dist_out, int32_t *d2_out,
                                size_t pitch_elems, vsb_stream_t stream)
{
    VSB_REQUIRE(dist_out || d2_out, "vsb_dt_unbounded: no output");
    VSB_REQUIRE(metric == VSB_METRIC_CHAMFER5 || metric == VSB_METRIC_EXACT, "vsb_dt_unbounded: unknown metric");
    VSB_REQUIRE(metric == VSB_METRIC_EXACT || !d2_out, "vsb_dt_unbounded: d2_out needs the exact metric");
    VSB_REQUIRE(pitch_elems >= (size_t)W, "vsb_dt_unbounded: pitch");
    int rc = vsb_coldist(seed_bits, wpr, W, H, g, g_pitch_elems, gmin64, stream);
    if (rc) return rc;
    const int nblk = (W + 63) / 64;
    dim3 grid((W + 255) / 256, H);
    cudaStream_t s = (cudaStream_t)stream;
    if (metric == VSB_METRIC_CHAMFER5) k_dense_dist<0><<<grid, 256, 0, s>>>(g, g_pitch_elems, gmin64, nblk, W, H, dist_out, nullptr, pitch_elems);
    else k_dense_dist<1><<<grid, 256, 0, s>>>(g, g_pitch_elems, gmin64, nblk, W, H, dist_out, d2_out, pitch_elems);
    return vsb_check_launch("k_dense_dist");
}

extern "C" int vsb_rec_dist_unbounded(const uint32_t *cur_bits, const uint32_t *const *prior_bits, int n_priors, int wpr,
                                      int W, int H, int metric, uint16_t *g, size_t g_pitch_elems, uint16_t *gmin64,
                                      float *dist, size_t dist_pitch_elems, uint32_t *minmax2, vsb_stream_t stream)
{
    VSB_REQUIRE(cur_bits && g && gmin64 && dist && minmax2, "vsb_rec_dist_unbounded: null pointer");
    VSB_REQUIRE(n_priors >= 0 && n_priors <= VSB_MAX_PRIORS && (n_priors == 0 || prior_bits), "vsb_rec_dist_unbounded: priors");
    VSB_REQUIRE(W > 0 && H > 0 && dist_pitch_elems >= (size_t)W, "vsb_rec_dist_unbounded: bad geometry");
    cudaStream_t s = (cudaStream_t)stream;
    int rc = vsb_coldist(cur_bits, wpr, W, H, g, g_pitch_elems, gmin64, stream);
    if (rc) return rc;
    static const uint32_t init[2] = {0xFFFFFFFFu, 0u};
    VSB_CUDA(cudaMemcpyAsync(minmax2, init, sizeof init, cudaMemcpyHostToDevice, s));
    PriorPtrsU pp;
    for (int i = 0; i < VSB_MAX_PRIORS; ++i) pp.p[i] = i < n_priors ? prior_bits[i] : nullptr;
    const int nblk = (W + 63) / 64;
    dim3 tgrid((W + VSB_TW - 1) / VSB_TW, (H + VSB_TH - 1) / VSB_TH);
    if (metric == VSB_METRIC_EXACT)
        k_rec_dist<1><<<tgrid, 256, 0, s>>>(cur_bits, pp, n_priors, wpr, g, g_pitch_elems, gmin64, nblk, W, H, dist, dist_pitch_elems, minmax2);
    else
        k_rec_dist<0><<<tgrid, 256, 0, s>>>(cur_bits, pp, n_priors, wpr, g, g_pitch_elems, gmin64, nblk, W, H, dist, dist_pitch_elems, minmax2);
    return vsb_check_launch("k_rec_dist");
}

static int fade_from_dist(const uint8_t *gray, size_t gray_pitch, int W, int H, const uint32_t *cur_bits,
                          const uint32_t *const *prior_bits, int n_priors, int wpr, int norm, float fade, float den,
                          const vsb_zparams *zp, const float *dist, size_t dist_pitch_elems, const uint32_t *minmax2,
                          const uint8_t *lut256_dev, uint8_t *out, size_t out_pitch, vsb_stream_t stream)
{
    UFParams p;
    memset(&p, 0, sizeof p);
    for (int i = 0; i < VSB_MAX_PRIORS; ++i) p.pri.p[i] = i < n_priors ? prior_bits[i] : nullptr;
    p.gray = gray; p.gpitch = gray_pitch; p.out = out; p.opitch = out_pitch; p.cur = cur_bits; p.np = n_priors;
    p.wpr = wpr; p.dist = dist; p.dpitch = dist_pitch_elems; p.mm = minmax2; p.lut = lut256_dev; p.F = fade; p.den = den;
    p.norm = norm; p.W = W; p.H = H;
    if (norm == 2) {
        for (int i = 0; i < n_priors; ++i) { p.w[i] = zp->weights[i]; p.wf[i] = zp->fades[i]; p.wden[i] = zp->dens[i]; }
        p.wtotal = zp->total_weight;
    }
    p.aligned = (vsb_aligned16(gray, gray_pitch) && vsb_aligned16(out, out_pitch)) ? 1 : 0;
    dim3 grid(((W + 15) / 16 + 63) / 64, (H + 3) / 4);
    k_fade_unbounded<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    return vsb_check_launch("k_fade_unbounded");
}

extern "C" int vsb_zblend_unbounded(const uint8_t *gray, size_t gray_pitch, int W, int H, const uint32_t *cur_bits,
                                    const uint32_t *const *prior_bits, int n_priors, int wpr, int metric, int norm,
                                    float fade, float den, uint16_t *g, size_t g_pitch_elems, uint16_t *gmin64,
                                    float *dist, size_t dist_pitch_elems, uint32_t *minmax2, const uint8_t *lut256_dev,
                                    uint8_t *out, size_t out_pitch, vsb_stream_t stream)
{
    VSB_REQUIRE(gray && out, "vsb_zblend_unbounded: null pointer");
    VSB_REQUIRE(norm == 0 || norm == 1, "vsb_zblend_unbounded: norm must be 0 (fixed) or 1 (min-max)");
    VSB_REQUIRE(gray_pitch >= (size_t)W && out_pitch >= (size_t)W, "vsb_zblend_unbounded: bad geometry");
    int rc = vsb_rec_dist_unbounded(cur_bits, prior_bits, n_priors, wpr, W, H, metric, g, g_pitch_elems, gmin64, dist,
                                    dist_pitch_elems, minmax2, stream);
    if (rc) return rc;
    return fade_from_dist(gray, gray_pitch, W, H, cur_bits, prior_bits, n_priors, wpr, norm, fade, den, nullptr, dist,
                          dist_pitch_elems, minmax2, lut256_dev, out, out_pitch, stream);
}

extern "C" int vsb_weighted_unbounded(const uint8_t *gray, size_t gray_pitch, int W, int H, const uint32_t *cur_bits,
                                      const uint32_t *const *prior_bits, int wpr, const vsb_zparams *zp, uint16_t *g,
                                      size_t g_pitch_elems, uint16_t *gmin64, float *dist, size_t dist_pitch_elems,
                                      uint32_t *minmax2, const uint8_t *lut256_dev, uint8_t *out, size_t out_pitch,
                                      vsb_stream_t stream)
{
    VSB_REQUIRE(gray && out && zp, "vsb_weighted_unbounded: null pointer");
    VSB_REQUIRE(gray_pitch >= (size_t)W && out_pitch >= (size_t)W, "vsb_weighted_unbounded: bad geometry");
    VSB_REQUIRE(zp->n_priors >= 0 && zp->n_priors <= VSB_MAX_PRIORS, "vsb_weighted_unbounded: priors");
    int rc = vsb_rec_dist_unbounded(cur_bits, prior_bits, zp->n_priors, wpr, W, H, zp->metric, g, g_pitch_elems, gmin64, dist,
                                    dist_pitch_elems, minmax2, stream);
    if (rc) return rc;
    return fade_from_dist(gray, gray_pitch, W, H, cur_bits, prior_bits, zp->n_priors, wpr, 2, 0.0f, 1.0f, zp, dist,
                          dist_pitch_elems, minmax2, lut256_dev, out, out_pitch, stream);
}
