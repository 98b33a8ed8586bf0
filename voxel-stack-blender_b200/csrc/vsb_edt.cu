// vsb_edt.cu -- K2': exact Euclidean distance transform, unbounded reach, cost independent of the distances.
//
//   replaces cv2.distanceTransform(src, DIST_L2, DIST_MASK_PRECISE)  (the `exact` metric; the reference's own call sites
//   processing_core.py:129,250,379 use mask size 5 -- SURVEY.md finding 1)
//
// Separable lower-envelope formulation (Felzenszwalb & Huttenlocher / Meijster et al.):
//
//     g(x,y)  = distance from (x,y) to the nearest white pixel of COLUMN x                       (column pass)
//     D2(x,y) = min over x' of (x - x')^2 + g(x',y)^2   -- the lower envelope of W parabolas     (row pass)
//
// Both passes are laid out for the GPU instead of for a scan line:
//
//   column pass (k_edt_colsum, k_edt_colscan, k_edt_coldist).  The packed mask is cut into blocks of 32 rows; a CTA brings a
//   32-row x 8-word tile in with one bulk tensor load (cp.async.bulk.tensor.2d + mbarrier), a warp takes one 32 x 32 pixel
//   square of it and transposes it with five shuffle butterflies so that every lane holds ITS column as one 32-bit word.
//   A first launch records per (block, column) the first and last white row (ffs / clz), a second one carries them down
//   and up the (few) blocks, the third runs the two distance recurrences over the lane's word in registers and writes g
//   as uint16 rows (EDT_GNONE where the column holds no white pixel at all).
//
//   row pass (k_edt_rowdc).  One CTA per image row, the row of g staged in shared memory with one bulk copy.  The stack of
//   Meijster's scan is a chain of dependent steps as long as the row; here the envelope is evaluated by bisection instead:
//   the column A(x) of the parabola that rules at x is non-decreasing in x (the cost (x-x')^2 + g(x')^2 is a Monge array), so
//   once A is known at the multiples of 2S, every pixel x = S (2k+1) only has to look at the candidates A(x-S) .. A(x+S).
//   Levels S = 2^k, ..., 2, 1 -- all pixels of a level are independent, the ranges of a level add up to about W, and a
//   candidate further away than the best distance known so far is never looked at, so pixels close to white cost a few
//   evaluations and nothing costs more than O(log W) per pixel whatever the distances are.  Short ranges are scanned by the
//   pixel's own lane; long ones (the coarse levels, and pixels where A jumps) by the pixel's whole warp, 32 candidates per
//   step, merged with redux.sync.  The last level (the odd pixels) is fused with the output: every lane stores a pair of
//   results (d2 and / or sqrtf), 8 bytes wide and coalesced; pairs of white pixels skip the arithmetic altogether.
//
// d2 is the exact integer squared distance (INT32_MAX where the image holds no white pixel), dist = sqrtf(d2) correctly
// rounded (3e38 there).  Limits: W, H <= 32767 (all costs fit 32 bits: 46341^2 + 32766^2 < 2^32).
#include "vsb_tma.cuh"

#define EDT_GNONE 46341u // g of a column without white pixels: its square exceeds INT32_MAX, every real candidate beats it
#define EDT_BIG (1 << 20)

// ---- column pass ------------------------------------------------------------------------------------------------------
// The warp's 32 x 32 square, transposed: bit r of the returned word = pixel (32 wc + lane, y0 + r).  Five butterfly stages:
// lanes k and k ^ j swap the off-diagonal j x j blocks of the bit matrix.
template <int NW>
__device__ __forceinline__ uint32_t edt_column_word(const CUtensorMap *tm, int wc0, int y0, uint32_t (*s_sq)[NW], uint64_t *s_bar)
{
    // the CTA's eight squares side by side are a 32-row x 32-byte tile of the packed mask: one bulk tensor load (rows and
    // words outside the mask arrive as zeros), then every warp picks up its word column
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) {
        vsb_mbar_init(s_bar, 1);
        vsb_fence_barrier_init();
    }
    __syncthreads();
    if (t == 0) {
        vsb_mbar_expect_tx(s_bar, 32 * NW * 4);
        vsb_tma_load_2d(s_sq, tm, wc0, y0, s_bar);
    }
    vsb_mbar_wait(s_bar, 0);
    uint32_t v = s_sq[lane][warp];
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t m = j == 16 ? 0x0000FFFFu : (j == 8 ? 0x00FF00FFu : (j == 4 ? 0x0F0F0F0Fu : (j == 2 ? 0x33333333u : 0x55555555u)));
        const uint32_t o = __shfl_xor_sync(0xffffffffu, v, j);
        v = (lane & j) ? ((v & ~m) | ((o >> j) & m)) : ((v & m) | ((o << j) & ~m));
    }
    return v;
}

// first / last white row of every (block of 32 rows, column): 0..31, 0xFF = none
__global__ void __launch_bounds__(256) k_edt_colsum(const __grid_constant__ CUtensorMap tm_bits, int wpr, uint8_t *__restrict__ first8,
                                                    uint8_t *__restrict__ last8)
{
    __shared__ __align__(128) uint32_t s_sq[32][8];
    __shared__ __align__(8) uint64_t s_bar;
    const int lane = threadIdx.x & 31, wc = blockIdx.x * 8 + (threadIdx.x >> 5);
    const uint32_t col = edt_column_word<8>(&tm_bits, blockIdx.x * 8, blockIdx.y * 32, s_sq, &s_bar);
    if (wc >= wpr) return;
    const size_t o = (size_t)blockIdx.y * (32 * (size_t)wpr) + 32 * wc + lane;
    first8[o] = col ? (uint8_t)(__ffs(col) - 1) : (uint8_t)0xFF;
    last8[o] = col ? (uint8_t)(31 - __clz(col)) : (uint8_t)0xFF;
}

// blockIdx.y == 0: up[b][x] = last white row above block b; == 1: dn[b][x] = first white row below block b (0xFFFF = none).
// One thread per column; the block records are requested 16 at a time, the running value is the only dependent chain.
#define EDT_SCAN_BATCH 16
__global__ void __launch_bounds__(64) k_edt_colscan(const uint8_t *__restrict__ first8, const uint8_t *__restrict__ last8, int Wp, int nb,
                                                    uint16_t *__restrict__ up, uint16_t *__restrict__ dn)
{
    const int x = blockIdx.x * 64 + threadIdx.x;
    if (x >= Wp) return;
    uint32_t run = 0xFFFFu;
    const bool down = blockIdx.y == 0;
    const uint8_t *src = down ? last8 : first8;
    uint16_t *dst = down ? up : dn;
    for (int i0 = 0; i0 < nb; i0 += EDT_SCAN_BATCH) {
        uint32_t v[EDT_SCAN_BATCH];
#pragma unroll
        for (int i = 0; i < EDT_SCAN_BATCH; ++i) {
            const int b = down ? i0 + i : nb - 1 - (i0 + i);
            v[i] = (i0 + i < nb) ? (uint32_t)src[(size_t)b * Wp + x] : 0xFFu;
        }
#pragma unroll
        for (int i = 0; i < EDT_SCAN_BATCH; ++i) {
            const int b = down ? i0 + i : nb - 1 - (i0 + i);
            if (i0 + i < nb) {
                dst[(size_t)b * Wp + x] = (uint16_t)run;
                if (v[i] != 0xFFu) run = 32u * b + v[i];
            }
        }
    }
}

#define EDT_CD_WARPS 8 // word columns per CTA of k_edt_coldist (wider CTAs were measured slower: 94 vs 79 us)
__global__ void __launch_bounds__(EDT_CD_WARPS * 32, 6) k_edt_coldist(const __grid_constant__ CUtensorMap tm_bits, int wpr, int H,
                                                                      const uint16_t *__restrict__ up, const uint16_t *__restrict__ dn,
                                                                      uint16_t *__restrict__ g, size_t gpitch)
{
    __shared__ __align__(128) uint32_t s_sq[32][EDT_CD_WARPS];
    __shared__ __align__(8) uint64_t s_bar;
    const int lane = threadIdx.x & 31, wc = blockIdx.x * EDT_CD_WARPS + (threadIdx.x >> 5);
    const int y0 = blockIdx.y * 32;
    const uint32_t col = edt_column_word<EDT_CD_WARPS>(&tm_bits, blockIdx.x * EDT_CD_WARPS, y0, s_sq, &s_bar);
    if (wc >= wpr) return;
    const size_t o = (size_t)blockIdx.y * (32 * (size_t)wpr) + 32 * wc + lane;
    const uint32_t u = up[o], d = dn[o];
    if ((size_t)(32 * wc) >= gpitch) return; // (gpitch is a multiple of 64: a word column is inside or outside as a whole)
    // per row: the nearest white pixel at or above it from a clz of the word's low part (or the block above), at or below it
    // from a running count walked upwards; the minimum leaves
    const int up1 = u != 0xFFFFu ? y0 - (int)u : EDT_BIG;       // from row y0 to the nearest white row above the block
    int run = d != 0xFFFFu ? (int)d - (y0 + 32) : EDT_BIG;      // at "row 32" of the block
    const int rows = min(32, H - y0);
    // two rows per trip; neighbouring lanes swap one value each so that every lane stores a pair of columns (32 bits) of one row
    uint32_t *gw = (uint32_t *)(g + (size_t)y0 * gpitch + 32 * wc + (lane & ~1));
    auto dist_at = [&](const int r) -> uint32_t {
        run = (col >> r) & 1u ? 0 : run + 1;
        const uint32_t low = col & (0xFFFFFFFFu >> (31 - r)); // rows 0 .. r
        const int du = low ? r - (31 - __clz(low)) : up1 + r;
        return (uint32_t)min(min(du, run), (int)EDT_GNONE);
    };
#pragma unroll
    for (int r = 31; r >= 1; r -= 2) {
        const uint32_t hi_row = dist_at(r), lo_row = dist_at(r - 1);
        const uint32_t got = __shfl_xor_sync(0xffffffffu, (lane & 1) ? lo_row : hi_row, 1);
        const int row = (lane & 1) ? r : r - 1; // odd lanes store row r, even lanes row r - 1
        const uint32_t word = (lane & 1) ? (got | (hi_row << 16)) : (lo_row | (got << 16));
        if (row < rows) gw[(size_t)row * (gpitch / 2)] = word;
    }
}

// ---- row pass ---------------------------------------------------------------------------------------------------------
#define DC_THREADS 256
#define DC_INLINE 24 // candidates a pixel's own lane scans; a longer range is scanned by the pixel's whole warp

__device__ __forceinline__ void edt_one(const uint16_t *__restrict__ s_g, int x, int xp, uint32_t &best, int &arg)
{
    const uint32_t gg = s_g[xp];
    const int d = x - xp;
    const uint32_t c = (uint32_t)(d * d) + gg * gg;
    if (c < best) { best = c; arg = xp; }
}
// candidates a, a + STEP, ... <= b of pixel x, two per trip (an index past b repeats b)
template <int STEP>
__device__ __forceinline__ void edt_scan(const uint16_t *__restrict__ s_g, int x, int a, int b, uint32_t &best, int &arg)
{
    for (int xp = a; xp <= b; xp += 2 * STEP) {
        const int x1 = min(xp + STEP, b);
        const uint32_t g0 = s_g[xp], g1 = s_g[x1];
        const int d0 = x - xp, d1 = x - x1;
        const uint32_t c0 = (uint32_t)(d0 * d0) + g0 * g0, c1 = (uint32_t)(d1 * d1) + g1 * g1;
        if (c0 < best) { best = c0; arg = xp; }
        if (c1 < best) { best = c1; arg = x1; }
    }
}

// One pixel per lane (x < 0: the lane has none).  Pixel x lies between two solved pixels whose ruling candidates lo <= hi
// bracket its own.  Start from the pixel's own column and the two bracket ends; only candidates closer than the square root
// of that value can do better, which leaves the range [a, b].  Short ranges are scanned by the lane, long ones (coarse
// levels; pixels where the ruling candidate jumps) by the whole warp, one after the other.  Returns the cost, sets arg.
__device__ __forceinline__ uint32_t edt_solve(const uint16_t *__restrict__ s_g, int lane, int x, int lo, int hi, int &arg)
{
    uint32_t best = 0u;
    int a = 1, b = 0;
    arg = x;
    if (x >= 0) {
        const uint32_t gx = s_g[x];
        if (gx != 0u) {
            best = gx * gx;
            edt_one(s_g, x, lo, best, arg);
            edt_one(s_g, x, hi, best, arg);
            if (hi - lo > 1) {
                const int r = (int)__fsqrt_rn((float)best) + 1; // an over-estimate is fine
                a = max(lo + 1, x - r);
                b = min(hi - 1, x + r);
                if (b - a < DC_INLINE) {
                    edt_scan<1>(s_g, x, a, b, best, arg);
                    b = a - 1;
                }
            }
        }
    }
    uint32_t pend = __ballot_sync(0xffffffffu, b >= a);
    while (pend) {
        const int src = __ffs(pend) - 1;
        pend &= pend - 1u;
        const int X = __shfl_sync(0xffffffffu, x, src), A = __shfl_sync(0xffffffffu, a, src), B = __shfl_sync(0xffffffffu, b, src);
        uint32_t wb = 0xFFFFFFFFu;
        int wa = 0x7FFF;
        edt_scan<32>(s_g, X, A + lane, B, wb, wa);
        const uint32_t m = __reduce_min_sync(0xffffffffu, wb);
        if (m < __shfl_sync(0xffffffffu, best, src)) {                       // warp-uniform
            const uint32_t am = __reduce_min_sync(0xffffffffu, wb == m ? (uint32_t)wa : 0x7FFFu);
            if (lane == src) { best = m; arg = (int)am; }
        }
    }
    return best;
}

__global__ void __launch_bounds__(DC_THREADS, 5) k_edt_rowdc(const uint16_t *__restrict__ g, size_t gpitch, int W, int Wa, int top_log2,
                                                             int32_t *__restrict__ d2_out, float *__restrict__ dist_out, size_t opitch)
{
    extern __shared__ __align__(128) uint8_t dsm[];
    __shared__ __align__(8) uint64_t s_bar;
    uint16_t *s_g = (uint16_t *)dsm; // the row of column distances
    uint16_t *s_a = s_g + Wa;        // ruling candidate of the EVEN pixels (index x / 2)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int y = blockIdx.x;
    if (tid == 0) {
        vsb_mbar_init(&s_bar, 1);
        vsb_fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t row_bytes = (uint32_t)(((size_t)W * 2 + 15) & ~(size_t)15);
        vsb_mbar_expect_tx(&s_bar, row_bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(vsb_smem_u32(s_g)),
                     "l"(g + (size_t)y * gpitch), "r"(row_bytes), "r"(vsb_smem_u32(&s_bar))
                     : "memory");
    }
    vsb_mbar_wait(&s_bar, 0);

    // ---- pixel 0 on its own (every other pixel has a solved neighbour on its left)
    if (warp == 0) {
        int arg;
        edt_solve(s_g, lane, lane == 0 ? 0 : -1, 0, W - 1, arg);
        if (lane == 0) s_a[0] = (uint16_t)arg;
    }
    __syncthreads();
    // ---- levels S = 2^top_log2 .. 2: the pixels x = S (2k+1).  Level S = 1 (the odd pixels) is fused with the output below.
    for (int lv = top_log2; lv >= 1; --lv) {
        const int Sx = 1 << lv;
        const int n = (W + Sx - 1) >> (lv + 1);
        // coarse levels hold few pixels and nearly all of them need their whole warp: deal them out in small groups so that
        // all warps get some (C pixels per warp and trip, the lanes above C idle)
        const int C = n >= 512 ? 32 : 1 << (31 - __clz(max(n >> 4, 1)));
        for (int k0 = warp * C; k0 < n; k0 += (DC_THREADS / 32) * C) {
            const int k = k0 + lane;
            int x = -1, lo = 0, hi = 0;
            if (lane < C && k < n) {
                x = Sx * (2 * k + 1);
                lo = s_a[(x - Sx) >> 1];
                hi = x + Sx < W ? (int)s_a[(x + Sx) >> 1] : W - 1;
            }
            int arg;
            edt_solve(s_g, lane, x, lo, hi, arg);
            if (x >= 0) s_a[x >> 1] = (uint16_t)arg;
        }
        __syncthreads();
    }

    // ---- level S = 1 and the output: lane k takes the pixel pair (2k, 2k+1) -- the even one is solved, the odd one sits
    //      between two solved pixels -- and stores both results side by side.
    int32_t *d2_row = d2_out ? d2_out + (size_t)y * opitch : nullptr;
    float *dist_row = dist_out ? dist_out + (size_t)y * opitch : nullptr;
    const bool pair_d2 = (((uintptr_t)d2_row) & 7u) == 0, pair_dist = (((uintptr_t)dist_row) & 7u) == 0;
    for (int k0 = warp * 32; 2 * k0 < W; k0 += DC_THREADS) {
        const int k = k0 + lane, xe = 2 * k, xo = xe + 1;
        const bool ve = xe < W, vo = xo < W;
        const int ae = ve ? (int)s_a[k] : 0;
        uint32_t ce = 0u, co = 0u;
        const uint32_t gpair = ve ? *(const uint32_t *)(s_g + xe) : 0u;      // g of both pixels (the row buffer is padded)
        if (__any_sync(0xffffffffu, gpair != 0u)) {
            int arg;
            co = edt_solve(s_g, lane, vo ? xo : -1, ae, xo + 1 < W ? (int)s_a[k + 1] : W - 1, arg);
            if (ve) {
                const uint32_t gg = s_g[ae];
                const int d = xe - ae;
                ce = (uint32_t)(d * d) + gg * gg;
            }
        }
        if (d2_row) {
            const int32_t ie = ce >= 0x80000000u ? 0x7fffffff : (int32_t)ce, io = co >= 0x80000000u ? 0x7fffffff : (int32_t)co;
            if (pair_d2 && vo) *(int2 *)(d2_row + xe) = make_int2(ie, io);
            else {
                if (ve) d2_row[xe] = ie;
                if (vo) d2_row[xo] = io;
            }
        }
        if (dist_row) {
            const float fe = ce == 0u ? 0.0f : (ce >= 0x80000000u ? 3.0e38f : __fsqrt_rn((float)ce));
            const float fo = co == 0u ? 0.0f : (co >= 0x80000000u ? 3.0e38f : __fsqrt_rn((float)co));
            if (pair_dist && vo) *(float2 *)(dist_row + xe) = make_float2(fe, fo);
            else {
                if (ve) dist_row[xe] = fe;
                if (vo) dist_row[xo] = fo;
            }
        }
    }
}

// ---- C-ABI ----------------------------------------------------------------------------------------------------------------
static size_t edt_gpitch(int W) { return ((size_t)W + 63) & ~(size_t)63; }
static size_t edt_align(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" size_t vsb_edt_exact_scratch_bytes(int W, int H)
{
    if (W <= 0 || H <= 0) return 0;
    const size_t Wp = edt_gpitch(W) + 128, nb = ((size_t)H + 31) / 32; // Wp >= 32 * wpr for wpr = 4 * ceil(W / 128)
    return edt_align(edt_gpitch(W) * (size_t)H * 2) + 2 * edt_align(nb * Wp) + 2 * edt_align(nb * Wp * 2) + 256;
}

extern "C" int vsb_edt_exact(const uint32_t *seed_bits, int wpr, int W, int H, void *scratch, int32_t *d2_out, float *dist_out,
                             size_t pitch_elems, vsb_stream_t stream)
{
    VSB_REQUIRE(seed_bits && scratch && (d2_out || dist_out), "vsb_edt_exact: null pointer");
    VSB_REQUIRE(W > 0 && H > 0 && W <= 32767 && H <= 32767 && wpr >= (W + 31) / 32 && (wpr & 3) == 0 && pitch_elems >= (size_t)W,
                "vsb_edt_exact: bad geometry (W, H <= 32767; wpr a multiple of 4 words)");
    VSB_REQUIRE((((uintptr_t)seed_bits) & 15u) == 0 && (((uintptr_t)scratch) & 127u) == 0, "vsb_edt_exact: mask / scratch alignment");
    const size_t gpitch = edt_gpitch(W), Wp = (size_t)32 * wpr, nb = ((size_t)H + 31) / 32;
    VSB_REQUIRE(Wp <= gpitch + 128, "vsb_edt_exact: wpr = %d is wider than the scratch layout allows for W = %d", wpr, W);
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t *p = (uint8_t *)scratch;
    uint16_t *g = (uint16_t *)p;             p += edt_align(gpitch * (size_t)H * 2);
    uint8_t *first8 = p;                     p += edt_align(nb * Wp);
    uint8_t *last8 = p;                      p += edt_align(nb * Wp);
    uint16_t *up = (uint16_t *)p;            p += edt_align(nb * Wp * 2);
    uint16_t *dn = (uint16_t *)p;

    const dim3 gsq((wpr + 7) / 8, (unsigned)nb);
    CUtensorMap tm_bits; // the packed mask as a 2-D tensor of words: boxes of 8 words x 32 rows = the eight squares of a CTA
    int rc = vsb_make_tmap_2d(&tm_bits, seed_bits, 4, (size_t)wpr, (size_t)H, (size_t)wpr * 4, 8, 32);
    if (rc) return rc;
    k_edt_colsum<<<gsq, 256, 0, s>>>(tm_bits, wpr, first8, last8);
    if ((rc = vsb_check_launch("k_edt_colsum"))) return rc;
    k_edt_colscan<<<dim3((unsigned)((Wp + 63) / 64), 2), 64, 0, s>>>(first8, last8, (int)Wp, (int)nb, up, dn);
    if ((rc = vsb_check_launch("k_edt_colscan"))) return rc;
    CUtensorMap tm_wide;
    if ((rc = vsb_make_tmap_2d(&tm_wide, seed_bits, 4, (size_t)wpr, (size_t)H, (size_t)wpr * 4, EDT_CD_WARPS, 32))) return rc;
    k_edt_coldist<<<dim3((wpr + EDT_CD_WARPS - 1) / EDT_CD_WARPS, (unsigned)nb), EDT_CD_WARPS * 32, 0, s>>>(tm_wide, wpr, H, up, dn, g, gpitch);
    if ((rc = vsb_check_launch("k_edt_coldist"))) return rc;

    const int Wa = (W + 15) & ~15; // row buffer: 2 Wa bytes of g + Wa bytes of A; at 16K five CTAs share an SM's shared memory
    const size_t smem = (size_t)Wa * 3;
    static bool seen[VSB_MAX_DEVICES] = {false};
    if (vsb_first_on_device(seen)) {
        VSB_CUDA(cudaFuncSetAttribute(k_edt_rowdc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        VSB_CUDA(cudaFuncSetAttribute(k_edt_rowdc, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    VSB_REQUIRE(smem <= 200 * 1024, "vsb_edt_exact: a row of %d pixels does not fit the row kernel's shared memory", W);
    int top_log2 = 0; // levels S = 2^top_log2 .. 2 (the largest power of two below W), none for W <= 2; S = 1 always runs
    while ((2 << top_log2) < W) ++top_log2;
    k_edt_rowdc<<<H, DC_THREADS, smem, s>>>(g, gpitch, W, Wa, top_log2, d2_out, dist_out, pitch_elems);
    return vsb_check_launch("k_edt_rowdc");
}
