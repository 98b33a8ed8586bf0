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
//   column pass (k_edt_colsum, k_edt_colscan, k_edt_coldist).  The packed mask is cut into blocks of 32 rows; a warp takes a
//   32 x 32 pixel square (one mask word column), transposes it with 32 ballots so that every lane holds ITS column as one
//   32-bit word, and answers "nearest white pixel above / below inside the block" with clz / ffs.  A first launch records per
//   (block, column) the first and last white row, a second one carries them down and up the (few) blocks, the third writes g
//   as uint16 rows (EDT_GNONE where the column holds no white pixel at all).
//
//   row pass (k_edt_rowdc).  One CTA per image row, the row of g staged in shared memory with one bulk copy.  The stack of
//   Meijster's scan is a chain of dependent steps as long as the row; here the envelope is evaluated by bisection instead:
//   the column A(x) of the parabola that rules at x is non-decreasing in x (the cost (x-x')^2 + g(x')^2 is a Monge array), so
//   once A is known at the multiples of 2S, every pixel x = S (2k+1) only has to look at the candidates A(x-S) .. A(x+S).
//   Levels S = 2^k, ..., 2, 1 -- all pixels of a level are independent, the ranges of a level add up to about W, and a
//   candidate further away than the best distance known so far is never looked at, so pixels close to white cost a few
//   evaluations and nothing costs more than O(log W) per pixel whatever the distances are.  Short ranges are scanned by the
//   pixel's own thread; long ones (the coarse levels, and pixels where A jumps) go to a work list in shared memory, cut
//   into chunks that the warps scan cooperatively (redux.sync min + a 64-bit shared-memory atomicMin per chunk).
//   The results leave the CTA in one coalesced pass (d2 and / or sqrtf).
//
// d2 is the exact integer squared distance (INT32_MAX where the image holds no white pixel), dist = sqrtf(d2) correctly
// rounded (3e38 there).  Limits: W, H <= 32767 (all costs fit 32 bits: 46341^2 + 32766^2 < 2^32).
#include "vsb_tma.cuh"

#define EDT_GNONE 46341u // g of a column without white pixels: its square exceeds INT32_MAX, every real candidate beats it
#define EDT_BIG (1 << 20)

// ---- column pass ------------------------------------------------------------------------------------------------------
// The warp's 32 x 32 square, transposed: bit r of the returned word = pixel (32 wc + lane, y0 + r).
__device__ __forceinline__ uint32_t edt_column_word(const uint32_t *__restrict__ bits, int wpr, int H, int wc, int y0, int lane)
{
    const int y = y0 + lane;
    const uint32_t roww = y < H ? bits[(size_t)y * wpr + wc] : 0u;
    uint32_t col = 0;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        const uint32_t b = __ballot_sync(0xffffffffu, (roww >> c) & 1u);
        if (c == lane) col = b;
    }
    return col;
}

// first / last white row of every (block of 32 rows, column): 0..31, 0xFF = none
__global__ void __launch_bounds__(256) k_edt_colsum(const uint32_t *__restrict__ bits, int wpr, int H, uint8_t *__restrict__ first8,
                                                    uint8_t *__restrict__ last8)
{
    const int lane = threadIdx.x & 31, wc = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (wc >= wpr) return;
    const uint32_t col = edt_column_word(bits, wpr, H, wc, blockIdx.y * 32, lane);
    const size_t o = (size_t)blockIdx.y * (32 * (size_t)wpr) + 32 * wc + lane;
    first8[o] = col ? (uint8_t)(__ffs(col) - 1) : (uint8_t)0xFF;
    last8[o] = col ? (uint8_t)(31 - __clz(col)) : (uint8_t)0xFF;
}

// blockIdx.y == 0: up[b][x] = last white row above block b; == 1: dn[b][x] = first white row below block b (0xFFFF = none)
__global__ void __launch_bounds__(128) k_edt_colscan(const uint8_t *__restrict__ first8, const uint8_t *__restrict__ last8, int Wp, int nb,
                                                     uint16_t *__restrict__ up, uint16_t *__restrict__ dn)
{
    const int x = blockIdx.x * 128 + threadIdx.x;
    if (x >= Wp) return;
    uint32_t run = 0xFFFFu;
    if (blockIdx.y == 0) {
#pragma unroll 8
        for (int b = 0; b < nb; ++b) {
            const uint32_t l = last8[(size_t)b * Wp + x];
            up[(size_t)b * Wp + x] = (uint16_t)run;
            if (l != 0xFFu) run = 32u * b + l;
        }
    } else {
#pragma unroll 8
        for (int b = nb - 1; b >= 0; --b) {
            const uint32_t f = first8[(size_t)b * Wp + x];
            dn[(size_t)b * Wp + x] = (uint16_t)run;
            if (f != 0xFFu) run = 32u * b + f;
        }
    }
}

__global__ void __launch_bounds__(256) k_edt_coldist(const uint32_t *__restrict__ bits, int wpr, int H, const uint16_t *__restrict__ up,
                                                     const uint16_t *__restrict__ dn, uint16_t *__restrict__ g, size_t gpitch)
{
    const int lane = threadIdx.x & 31, wc = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (wc >= wpr) return;
    const int y0 = blockIdx.y * 32;
    const size_t o = (size_t)blockIdx.y * (32 * (size_t)wpr) + 32 * wc + lane;
    const uint32_t u = up[o], d = dn[o];
    const uint32_t col = edt_column_word(bits, wpr, H, wc, y0, lane);
    const int up0 = u != 0xFFFFu ? y0 - (int)u : EDT_BIG;        // from row y0 to the nearest white row above the block
    const int dn0 = d != 0xFFFFu ? (int)d - (y0 + 31) : EDT_BIG; // from row y0 + 31 to the nearest white row below it
    if ((size_t)(32 * wc + lane) >= gpitch) return;
    uint16_t *gp = g + (size_t)y0 * gpitch + 32 * wc + lane;
    const int rows = min(32, H - y0);
#pragma unroll 8
    for (int r = 0; r < rows; ++r) {
        const uint32_t above = col << (31 - r), below = col >> r;
        const int du = above ? __clz(above) : up0 + r;
        const int dd = below ? __ffs(below) - 1 : dn0 + 31 - r;
        gp[(size_t)r * gpitch] = (uint16_t)min(min(du, dd), (int)EDT_GNONE);
    }
}

// ---- row pass ---------------------------------------------------------------------------------------------------------
#define DC_THREADS 256
#define DC_INLINE 40     // candidates a pixel's own thread scans; longer ranges go to the warps' work list
#define DC_CHUNK 512     // candidates per work-list chunk (one warp, 16 strides)
#define DC_MAXPIX 512    // work-list pixels per level  (more than fit: scanned inline -- slower, still exact)
#define DC_MAXCHUNK 768  // work-list chunks per level

struct EdtRowSmem {
    uint16_t *g, *a;           // the row of column distances; the ruling candidate per pixel
    unsigned long long *key;   // work list: best (cost << 16 | candidate) per listed pixel
    uint16_t *px;              //            the listed pixel
    uint2 *chunk;              //            {slot | x << 16, lo | hi << 16}
};

__device__ __forceinline__ void edt_scan(const uint16_t *__restrict__ s_g, int x, int a, int b, int step, uint32_t &best, int &arg)
{
    for (int xp = a; xp <= b; xp += step) {
        const uint32_t gg = s_g[xp];
        const int d = x - xp;
        const uint32_t c = (uint32_t)(d * d) + gg * gg;
        if (c < best) { best = c; arg = xp; }
    }
}

__global__ void __launch_bounds__(DC_THREADS, 3) k_edt_rowdc(const uint16_t *__restrict__ g, size_t gpitch, int W, int Wa, int top_S,
                                                             int32_t *__restrict__ d2_out, float *__restrict__ dist_out, size_t opitch)
{
    extern __shared__ __align__(128) uint8_t dsm[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_cnt[2][2]; // [level parity][pixels, chunks]
    EdtRowSmem S;
    S.g = (uint16_t *)dsm;
    S.a = S.g + Wa;
    S.key = (unsigned long long *)(S.a + Wa);
    S.chunk = (uint2 *)(S.key + DC_MAXPIX);
    S.px = (uint16_t *)(S.chunk + DC_MAXCHUNK);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int y = blockIdx.x;
    if (tid == 0) {
        vsb_mbar_init(&s_bar, 1);
        vsb_fence_barrier_init();
        s_cnt[0][0] = s_cnt[0][1] = s_cnt[1][0] = s_cnt[1][1] = 0;
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t row_bytes = (uint32_t)(((size_t)W * 2 + 15) & ~(size_t)15);
        vsb_mbar_expect_tx(&s_bar, row_bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(vsb_smem_u32(S.g)),
                     "l"(g + (size_t)y * gpitch), "r"(row_bytes), "r"(vsb_smem_u32(&s_bar))
                     : "memory");
    }
    vsb_mbar_wait(&s_bar, 0);

    int parity = 0;
    // level -1 is pixel 0 on its own (every other pixel has a solved neighbour on its left), then S = top_S .. 1
    for (int Sx = -1; Sx != 0; Sx = (Sx < 0 ? top_S : Sx >> 1)) {
        const int n = Sx < 0 ? 1 : (W - Sx + 2 * Sx - 1) / (2 * Sx); // pixels x = S (2k+1) < W
        int *cnt = s_cnt[parity];
        if (tid == 0) { s_cnt[parity ^ 1][0] = 0; s_cnt[parity ^ 1][1] = 0; }
        bool listed = false;
        for (int k = tid; k < n; k += DC_THREADS) {
            int x, lo, hi;
            if (Sx < 0) { x = 0; lo = 0; hi = W - 1; }
            else { x = Sx * (2 * k + 1); lo = S.a[x - Sx]; hi = x + Sx < W ? (int)S.a[x + Sx] : W - 1; }
            const uint32_t gx = S.g[x];
            if (gx == 0u) { S.a[x] = (uint16_t)x; continue; }
            uint32_t best = gx * gx;
            int arg = x;
            edt_scan(S.g, x, lo, lo, 1, best, arg);
            edt_scan(S.g, x, hi, hi, 1, best, arg);
            const int r = (int)__fsqrt_rn((float)best) + 1;      // no candidate further away than this can win (over-estimates are fine)
            const int a = max(lo + 1, x - r), b = min(hi - 1, x + r);
            const int len = b - a + 1;
            bool inl = len <= DC_INLINE;
            if (!inl) {
                const int nch = (len + DC_CHUNK - 1) / DC_CHUNK;
                const int slot = atomicAdd(&cnt[0], 1);
                inl = slot >= DC_MAXPIX;
                if (!inl) {
                    const int e0 = atomicAdd(&cnt[1], nch);
                    const bool fits = e0 + nch <= DC_MAXCHUNK;
                    S.px[slot] = fits ? (uint16_t)x : (uint16_t)0xFFFFu;
                    S.key[slot] = ((unsigned long long)best << 16) | (unsigned long long)arg;
                    for (int c = 0; c < nch && e0 + c < DC_MAXCHUNK; ++c) {
                        const int ca = a + c * DC_CHUNK, cb = min(b, ca + DC_CHUNK - 1);
                        S.chunk[e0 + c] = make_uint2(fits ? ((uint32_t)slot | ((uint32_t)x << 16)) : 0xFFFFu, (uint32_t)ca | ((uint32_t)cb << 16));
                    }
                    inl = !fits;
                    listed = true;
                }
            }
            if (inl) {
                edt_scan(S.g, x, a, b, 1, best, arg);
                S.a[x] = (uint16_t)arg;
            }
        }
        if (__syncthreads_or(listed)) {
            const int npix = min(cnt[0], DC_MAXPIX), nchunk = min(cnt[1], DC_MAXCHUNK);
            for (int e = warp; e < nchunk; e += DC_THREADS / 32) {
                const uint2 ch = S.chunk[e];
                const uint32_t slot = ch.x & 0xFFFFu;
                if (slot == 0xFFFFu) continue;
                uint32_t best = 0xFFFFFFFFu;
                int arg = 0xFFFF;
                edt_scan(S.g, (int)(ch.x >> 16), (int)(ch.y & 0xFFFFu) + lane, (int)(ch.y >> 16), 32, best, arg);
                const uint32_t m = __reduce_min_sync(0xffffffffu, best);
                const uint32_t am = __reduce_min_sync(0xffffffffu, best == m ? (uint32_t)arg : 0xFFFFu);
                if (lane == 0 && m != 0xFFFFFFFFu) atomicMin(&S.key[slot], ((unsigned long long)m << 16) | (unsigned long long)am);
            }
            __syncthreads();
            for (int s = tid; s < npix; s += DC_THREADS) {
                const uint32_t x = S.px[s];
                if (x != 0xFFFFu) S.a[x] = (uint16_t)(S.key[s] & 0xFFFFull);
            }
            __syncthreads();
        }
        parity ^= 1;
    }

    for (int x = tid; x < W; x += DC_THREADS) {
        const int a = S.a[x];
        const uint32_t gg = S.g[a];
        const int d = x - a;
        const uint32_t c = (uint32_t)(d * d) + gg * gg;
        const bool none = c >= 0x80000000u;
        if (d2_out) d2_out[(size_t)y * opitch + x] = none ? 0x7fffffff : (int32_t)c;
        if (dist_out) dist_out[(size_t)y * opitch + x] = none ? 3.0e38f : __fsqrt_rn((float)c);
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
    k_edt_colsum<<<gsq, 256, 0, s>>>(seed_bits, wpr, H, first8, last8);
    int rc = vsb_check_launch("k_edt_colsum");
    if (rc) return rc;
    k_edt_colscan<<<dim3((unsigned)((Wp + 127) / 128), 2), 128, 0, s>>>(first8, last8, (int)Wp, (int)nb, up, dn);
    if ((rc = vsb_check_launch("k_edt_colscan"))) return rc;
    k_edt_coldist<<<gsq, 256, 0, s>>>(seed_bits, wpr, H, up, dn, g, gpitch);
    if ((rc = vsb_check_launch("k_edt_coldist"))) return rc;

    const int Wa = (int)gpitch;
    const size_t smem = (size_t)Wa * 4 + DC_MAXPIX * 8 + DC_MAXCHUNK * 8 + DC_MAXPIX * 2;
    static bool configured = false;
    if (!configured) {
        VSB_CUDA(cudaFuncSetAttribute(k_edt_rowdc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    VSB_REQUIRE(smem <= 200 * 1024, "vsb_edt_exact: a row of %d pixels does not fit the row kernel's shared memory", W);
    int top_S = 1;
    while (top_S * 2 < W) top_S *= 2;
    if (W == 1) top_S = 0;
    k_edt_rowdc<<<H, DC_THREADS, smem, s>>>(g, gpitch, W, Wa, top_S, d2_out, dist_out, pitch_elems);
    return vsb_check_launch("k_edt_rowdc");
}
