// vsb_edt.cu -- K2': exact Euclidean distance transform, unbounded reach, cost independent of the distances.
//
//   replaces cv2.distanceTransform(src, DIST_L2, DIST_MASK_PRECISE)  (the `exact` metric; the reference's own call sites
//   processing_core.py:129,250,379 use mask size 5 -- SURVEY.md finding 1)
//
// Separable lower-envelope algorithm (Felzenszwalb & Huttenlocher; the integer formulation of Meijster, Roerdink &
// Hesselink), two launches:
//
//   k_edt_rows : h(x,y) = distance from (x,y) to the nearest white pixel of ROW y (uint16, 0xFFFF = none).  One CTA per
//                row: the row's packed mask arrives with one bulk copy (cp.async.bulk + mbarrier), a block-wide
//                prefix-max / suffix-min over the words gives every word its nearest white column on either side, each
//                lane then resolves one pixel with clz / ffs, and the finished row leaves with one bulk store.
//   k_edt_cols : D2(x,y) = min_y' (y - y')^2 + h(x,y')^2 along every column.  A white pixel ends the influence of
//                everything above it on everything below it, so columns split into independent runs of black pixels; a
//                thread owns the runs that START in its 64-row band (lanes = 32 adjacent columns: the band's mask rows are
//                transposed with warp ballots so that run starts fall out of two 32-bit words per lane, and the loads of
//                lanes working on the same rows coalesce) and for each run does Meijster's two scans: downwards, the stack of
//                parabolas on the lower envelope (integer `Sep`, the stack lives in a global scratch array indexed like
//                the image -- entry q of a run sits q rows below the run's first candidate, which never collides with
//                the run above or below); upwards, the evaluation.  Every candidate is pushed and popped at most once:
//                O(run length), whatever the distances are.  White pixels get their zeros from the thread that scans
//                over them.
//
// d2 is the exact integer squared distance (INT32_MAX where the image holds no white pixel in reach of the column's run,
// i.e. nowhere), dist = sqrtf(d2) correctly rounded.  Limits: W, H <= 32767 (d2 and the envelope arithmetic stay in int32).
#include "vsb_tma.cuh"

#define EDT_NONE 0xFFFFu
#define EDT_BAND 64
#define EDT_INF_POS 0x3fffffff

// ---- row pass -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_edt_rows(const uint32_t *__restrict__ bits, int wpr, int W, uint16_t *__restrict__ h,
                                                  size_t hpitch)
{
    extern __shared__ __align__(128) uint8_t dsm[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_wmax[8], s_wmin[8];
    uint32_t *s_bits = (uint32_t *)dsm;                                   // wpr words
    int *s_left = (int *)(dsm + (((size_t)wpr * 4 + 127) & ~(size_t)127)); // nearest white column at or before the END of word w
    int *s_right = s_left + wpr;                                          // nearest white column at or after the START of word w
    uint16_t *s_h = (uint16_t *)(s_right + wpr + ((32 - (2 * wpr) % 32) % 32)); // 128-byte aligned row of distances
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int y = blockIdx.x;
    const uint32_t row_bytes = (uint32_t)wpr * 4u;
    if (t == 0) {
        vsb_mbar_init(&s_bar, 1);
        vsb_fence_barrier_init();
    }
    __syncthreads();
    if (t == 0) {
        vsb_mbar_expect_tx(&s_bar, row_bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(vsb_smem_u32(s_bits)),
                     "l"(bits + (size_t)y * wpr), "r"(row_bytes), "r"(vsb_smem_u32(&s_bar))
                     : "memory");
    }
    vsb_mbar_wait(&s_bar, 0);

    // per word: last / first white column inside it; then an inclusive prefix-max and suffix-min over the row's words.
    // Thread t owns the contiguous chunk of words [t*per, (t+1)*per).
    const int per = (wpr + 255) / 256;
    const int w0 = t * per, w1 = min(wpr, w0 + per);
    int mx = -1, mn = EDT_INF_POS;
    for (int w = w0; w < w1; ++w) {
        const uint32_t v = s_bits[w];
        if (v) { mx = 32 * w + 31 - __clz(v); mn = min(mn, 32 * w + __ffs(v) - 1); }
    }
    int pmx = mx, smn = mn; // inclusive scans of the per-thread values across the block
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, pmx, o), b = __shfl_down_sync(0xffffffffu, smn, o);
        if (lane >= o) pmx = max(pmx, a);
        if (lane + o < 32) smn = min(smn, b);
    }
    if (lane == 31) s_wmax[warp] = pmx;
    if (lane == 0) s_wmin[warp] = smn;
    __syncthreads();
    int before = -1, after = EDT_INF_POS; // over the warps before / after mine
    for (int k = 0; k < 8; ++k) {
        if (k < warp) before = max(before, s_wmax[k]);
        if (k > warp) after = min(after, s_wmin[k]);
    }
    {   // exclusive values for my chunk, then the per-word tables
        int run = max(before, __shfl_up_sync(0xffffffffu, pmx, 1));
        if (lane == 0) run = before;
        for (int w = w0; w < w1; ++w) {
            const uint32_t v = s_bits[w];
            if (v) run = 32 * w + 31 - __clz(v);
            s_left[w] = run;
        }
        int rr = min(after, __shfl_down_sync(0xffffffffu, smn, 1));
        if (lane == 31) rr = after;
        for (int w = w1 - 1; w >= w0; --w) {
            const uint32_t v = s_bits[w];
            if (v) rr = 32 * w + __ffs(v) - 1;
            s_right[w] = rr;
        }
    }
    __syncthreads();
    for (int w = warp; w < wpr; w += 8) {
        const uint32_t v = s_bits[w];
        const int x = 32 * w + lane;
        const uint32_t le = v & (0xFFFFFFFFu >> (31 - lane)), ge = v & (0xFFFFFFFFu << lane);
        const int pl = le ? 32 * w + 31 - __clz(le) : (w > 0 ? s_left[w - 1] : -1);
        const int pr = ge ? 32 * w + __ffs(ge) - 1 : (w + 1 < wpr ? s_right[w + 1] : EDT_INF_POS);
        uint32_t d = EDT_NONE;
        if (pl >= 0) d = (uint32_t)(x - pl);
        if (pr < EDT_INF_POS) d = min(d, (uint32_t)(pr - x));
        s_h[x] = (uint16_t)min(d, 0xFFFEu + (d == EDT_NONE));
    }
    vsb_fence_async_smem();
    __syncthreads();
    if (t == 0) {
        const uint32_t out_bytes = (uint32_t)(((size_t)W * 2 + 15) & ~(size_t)15);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(h + (size_t)y * hpitch), "r"(vsb_smem_u32(s_h)),
                     "r"(out_bytes)
                     : "memory");
        vsb_tma_commit();
        vsb_tma_wait_all<0>();
    }
}

// ---- column pass ----------------------------------------------------------------------------------------------------
#define EDT_PF 8 // rows of h requested ahead of the envelope scan (independent loads; the scan itself is a dependent chain)

__global__ void __launch_bounds__(128) k_edt_cols(const uint32_t *__restrict__ bits, int wpr, int W, int H,
                                                  const uint16_t *__restrict__ h, size_t hpitch, uint2 *__restrict__ stack,
                                                  size_t spitch, int32_t *__restrict__ d2_out, float *__restrict__ dist_out,
                                                  size_t opitch)
{
    const int lane = threadIdx.x & 31;
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int wc = (blockIdx.x * 128 + (threadIdx.x & ~31)) >> 5; // the warp's mask word column (warp-uniform)
    if (wc >= wpr) return;
    const int yb0 = blockIdx.y * EDT_BAND, yb1 = min(H, yb0 + EDT_BAND);
    // ---- the band's 64 rows of this warp's mask word, transposed: every lane gets its column as two 32-bit words ----------
    uint32_t col[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int yy = yb0 + 32 * q + lane;
        const uint32_t roww = yy < yb1 ? bits[(size_t)yy * wpr + wc] : 0xFFFFFFFFu; // rows below the image end a run like white
        uint32_t mine = 0;
#pragma unroll
        for (int cidx = 0; cidx < 32; ++cidx) {
            const uint32_t b = __ballot_sync(0xffffffffu, (roww >> cidx) & 1u);
            if (cidx == lane) mine = b;
        }
        col[q] = mine;
    }
    const bool above = (yb0 == 0) ? true : ((bits[(size_t)(yb0 - 1) * wpr + wc] >> lane) & 1u); // top edge starts a run too
    // zeros for the band's white pixels (coalesced: the warp walks the rows together)
    for (int r = 0; r < yb1 - yb0; ++r) {
        const bool wht = (col[r >> 5] >> (r & 31)) & 1u;
        if (wht && x < W) {
            if (d2_out) d2_out[(size_t)(yb0 + r) * opitch + x] = 0;
            if (dist_out) dist_out[(size_t)(yb0 + r) * opitch + x] = 0.0f;
        }
    }
    if (x >= W) return;
    // run starts of my column inside the band: black pixel whose upper neighbour is white (or the image's top edge)
    const unsigned long long c64 = (unsigned long long)col[0] | ((unsigned long long)col[1] << 32);
    unsigned long long starts = ~c64 & ((c64 << 1) | (above ? 1ULL : 0ULL));
    if (yb1 - yb0 < 64) starts &= (1ULL << (yb1 - yb0)) - 1ULL;

    while (starts) {
        const int ys = yb0 + __ffsll((long long)starts) - 1;
        starts &= starts - 1ULL;
        // ---- downward scan: lower envelope of the parabolas (y - s)^2 + f_s over the candidates of this run --------------
        // candidates: the white pixel above the run (f = 0), the run's black pixels whose row holds a white pixel, and the
        // white pixel that ends the run (f = 0; h == 0 marks it)
        const int base = ys > 0 ? ys - 1 : 0;                    // row of the first possible candidate = stack slot 0
        int q = -1, s_top = 0, t_top = 0, f_top = 0;            // the stack's top entry lives in registers
        if (ys > 0) { q = 0; s_top = ys - 1; t_top = ys; f_top = 0; }
        int ye = H - 1;
        bool open = true;
        for (int u0 = ys; u0 < H && open; u0 += EDT_PF) {
            uint32_t hv[EDT_PF];
#pragma unroll
            for (int i = 0; i < EDT_PF; ++i) hv[i] = (u0 + i < H) ? (uint32_t)h[(size_t)(u0 + i) * hpitch + x] : EDT_NONE;
#pragma unroll
            for (int i = 0; i < EDT_PF; ++i) {
                const int u = u0 + i;
                if (!open || u >= H) break;
                if (hv[i] != EDT_NONE) {
                    const int fu = (int)(hv[i] * hv[i]);
                    while (q >= 0) { // pop while the new parabola is already lower where the top one starts to rule
                        const int da = t_top - s_top, db = t_top - u;
                        if (da * da + f_top <= db * db + fu) break;
                        if (--q >= 0) {
                            const uint2 e = stack[(size_t)(base + q) * spitch + x];
                            s_top = (int)(e.x & 0xFFFFu); t_top = (int)(e.x >> 16); f_top = (int)e.y;
                        }
                    }
                    if (q < 0) {
                        q = 0; s_top = u; t_top = ys; f_top = fu;
                    } else {
                        // first row where the new parabola is strictly lower: 1 + floor((u^2 - s^2 + fu - fs) / (2 (u - s)))
                        const unsigned num = (unsigned)((u * u - s_top * s_top) + (fu - f_top));
                        const int wst = 1 + (int)(num / (unsigned)(2 * (u - s_top)));
                        if (wst < H) {
                            stack[(size_t)(base + q) * spitch + x] = make_uint2((uint32_t)s_top | ((uint32_t)t_top << 16), (uint32_t)f_top);
                            ++q; s_top = u; t_top = wst; f_top = fu;
                        }
                    }
                    if (hv[i] == 0u) { ye = u - 1; open = false; } // the white pixel that ends the run
                }
            }
        }
        // ---- upward scan: evaluate ---------------------------------------------------------------------------------------
        for (int yq = ye; yq >= ys; --yq) {
            while (q > 0 && t_top > yq) {
                --q;
                const uint2 e = stack[(size_t)(base + q) * spitch + x];
                s_top = (int)(e.x & 0xFFFFu); t_top = (int)(e.x >> 16); f_top = (int)e.y;
            }
            int d2 = 0x7fffffff;
            if (q >= 0) { const int dy = yq - s_top; d2 = dy * dy + f_top; }
            if (d2_out) d2_out[(size_t)yq * opitch + x] = d2;
            if (dist_out) dist_out[(size_t)yq * opitch + x] = d2 == 0x7fffffff ? 3.0e38f : __fsqrt_rn((float)d2);
        }
    }
}

// ---- C-ABI ----------------------------------------------------------------------------------------------------------------
static size_t edt_hpitch(int W) { return ((size_t)W + 63) & ~(size_t)63; }

extern "C" size_t vsb_edt_exact_scratch_bytes(int W, int H)
{
    if (W <= 0 || H <= 0) return 0;
    return edt_hpitch(W) * (size_t)H * sizeof(uint16_t) + (size_t)W * (size_t)H * sizeof(uint2) + 256;
}

extern "C" int vsb_edt_exact(const uint32_t *seed_bits, int wpr, int W, int H, void *scratch, int32_t *d2_out, float *dist_out,
                             size_t pitch_elems, vsb_stream_t stream)
{
    VSB_REQUIRE(seed_bits && scratch && (d2_out || dist_out), "vsb_edt_exact: null pointer");
    VSB_REQUIRE(W > 0 && H > 0 && W <= 32767 && H <= 32767 && wpr >= (W + 31) / 32 && (wpr & 3) == 0 && pitch_elems >= (size_t)W,
                "vsb_edt_exact: bad geometry (W, H <= 32767; wpr a multiple of 4 words)");
    VSB_REQUIRE((((uintptr_t)seed_bits) & 15u) == 0 && (((uintptr_t)scratch) & 127u) == 0, "vsb_edt_exact: mask / scratch alignment");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t hpitch = edt_hpitch(W);
    uint16_t *h = (uint16_t *)scratch;
    uint2 *stack = (uint2 *)((uint8_t *)scratch + ((hpitch * (size_t)H * sizeof(uint16_t) + 127) & ~(size_t)127));
    const size_t hrow = hpitch > (size_t)32 * wpr ? hpitch : (size_t)32 * wpr; // the row kernel resolves every bit of every word
    const size_t smem = (((size_t)wpr * 4 + 127) & ~(size_t)127) + (size_t)2 * wpr * 4 + 128 + hrow * 2 + 128;
    static bool configured = false;
    if (!configured) {
        VSB_CUDA(cudaFuncSetAttribute(k_edt_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    VSB_REQUIRE(smem <= 200 * 1024, "vsb_edt_exact: row of %d pixels does not fit the row kernel's shared memory", W);
    k_edt_rows<<<H, 256, smem, s>>>(seed_bits, wpr, W, h, hpitch);
    int rc = vsb_check_launch("k_edt_rows");
    if (rc) return rc;
    dim3 grid((W + 127) / 128, (H + EDT_BAND - 1) / EDT_BAND);
    k_edt_cols<<<grid, 128, 0, s>>>(seed_bits, wpr, W, H, h, hpitch, stack, (size_t)W, d2_out, dist_out, pitch_elems);
    return vsb_check_launch("k_edt_cols");
}
