// vsb_ccl.cu -- K5: Enhanced-EDT support.  8-connected components of the receding mask, per-component
// maximum of the clipped distance, and the ROI-normalised fade.
//
// Replaces cv2.connectedComponents(rec) (processing_core.py:383), ndimage.maximum / the numba first
// pass (:291, :313-320) and the two arithmetic flavours of the fade (:295-302 float32, :323-334 float64),
// followed by the final mask (:404), merge (:458) and the leading LUT.
//
// Union-find over pixels (label-equivalence, atomicMin hooking).  Only rec pixels are ever touched in
// `labels` / `compmax`, so neither scratch array needs initialising; the mask itself is bit-packed.
// Horizontal adjacency inside a mask word is resolved when the forest is initialised (every pixel of a
// run points at the run's first pixel in the word), so the merge kernel only links across words and rows.
#include "vsb_common.cuh"

__device__ __forceinline__ int uf_find(volatile int32_t *L, int a)
{
    int p = L[a];
    while (p != a) { a = p; p = L[a]; }
    return a;
}

__device__ __forceinline__ void uf_union(int32_t *L, int a, int b)
{
    for (;;) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a > b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(&L[b], a); // hook the larger root under the smaller
        if (old == b) return;
        b = old;
    }
}

// thread per mask word: parent = first pixel of the horizontal run inside the word; compmax = 0
__global__ void __launch_bounds__(256) k_ccl_init(const uint32_t *__restrict__ rec, int wpr, int W, int H,
                                                  int32_t *__restrict__ L, uint32_t *__restrict__ cmax)
{
    const int wx = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (wx >= wpr) return;
    uint32_t m = rec[(size_t)y * wpr + wx];
    if (!m) return;
    const int base = y * W + (wx << 5);
    int start = -1;
    for (int b = 0; b < 32; ++b) {
        if ((m >> b) & 1u) {
            if (start < 0) start = b;
            L[base + b] = base + start;
            cmax[base + b] = 0u;
        } else start = -1;
    }
}

// thread per mask word: link run starts to the previous word, and every pixel to the row above
__global__ void __launch_bounds__(256) k_ccl_merge(const uint32_t *__restrict__ rec, int wpr, int W, int H,
                                                   int32_t *__restrict__ L)
{
    const int wx = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (wx >= wpr) return;
    const size_t o = (size_t)y * wpr + wx;
    const uint32_t m = rec[o];
    if (!m) return;
    const int base = y * W + (wx << 5);
    if ((m & 1u) && wx > 0 && (rec[o - 1] >> 31)) uf_union(L, base, base - 1);
    if (y == 0) return;
    // 34-bit window of the row above: bit i+1 of `up` = pixel (x0 + i, y-1), i = -1..32
    const uint32_t u = rec[o - wpr];
    const uint32_t ul = wx > 0 ? (rec[o - wpr - 1] >> 31) : 0u;
    const uint32_t ur = wx + 1 < wpr ? (rec[o - wpr + 1] & 1u) : 0u;
    const unsigned long long up = (unsigned long long)ul | ((unsigned long long)u << 1) | ((unsigned long long)ur << 33);
    // One union per pair of touching runs, not per pixel: the pixels of a horizontal run are already linked (inside the
    // word by k_ccl_init, across words by the union above), so linking each run of this word once to every run of the
    // row above that touches its one-pixel dilation connects exactly the 8-neighbours.
    uint32_t todo = m;
    while (todo) {
        const int a = __ffs(todo) - 1;                              // run [a, e) of this word
        const uint32_t gaps = ~m & (0xFFFFFFFFu << a);
        const int e = gaps ? __ffs(gaps) - 1 : 32;
        todo = e >= 32 ? 0u : (todo & (0xFFFFFFFFu << e));
        // pixels a-1 .. e of the row above = window bits a .. e+1
        unsigned long long touch = up & (~0ULL << a) & ((1ULL << (e + 2)) - 1ULL);
        while (touch) {
            const int q = __ffsll((long long)touch) - 1;            // a pixel of a touching run above: x0 + q - 1
            uf_union(L, base + a, base - W + q - 1);
            const unsigned long long ugaps = ~up & (~0ULL << q);    // skip the rest of that run
            const int ue = ugaps ? __ffsll((long long)ugaps) - 1 : 64;
            touch = ue >= 64 ? 0ULL : (touch & (~0ULL << ue));
        }
    }
}

// thread per mask word: flatten and accumulate the component maximum of the (non-negative) distance
__global__ void __launch_bounds__(256) k_ccl_flatten_max(const uint32_t *__restrict__ rec, int wpr, int W, int H,
                                                         const float *__restrict__ dist, size_t dpitch,
                                                         int32_t *__restrict__ L, uint32_t *__restrict__ cmax)
{
    const int wx = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (wx >= wpr) return;
    uint32_t m = rec[(size_t)y * wpr + wx];
    if (!m) return;
    const int base = y * W + (wx << 5);
    // one root look-up per horizontal run (its pixels all hang off the run's first pixel since k_ccl_init), one
    // atomicMax per group of consecutive runs with the same root
    int last_root = -1;
    uint32_t acc_max = 0;
    while (m) {
        const int a = __ffs(m) - 1;
        const uint32_t gaps = ~m & (0xFFFFFFFFu << a);
        const int e = gaps ? __ffs(gaps) - 1 : 32;
        m = e >= 32 ? 0u : (m & (0xFFFFFFFFu << e));
        const int root = uf_find(L, base + a);
        uint32_t run_max = 0;
        const float *drow = dist + (size_t)y * dpitch + (wx << 5);
        for (int b = a; b < e; ++b) {
            L[base + b] = root;
            run_max = max(run_max, __float_as_uint(drow[b]));
        }
        if (root != last_root) {
            if (last_root >= 0) atomicMax(&cmax[last_root], acc_max);
            last_root = root; acc_max = run_max;
        } else acc_max = max(acc_max, run_max);
    }
    if (last_root >= 0) atomicMax(&cmax[last_root], acc_max);
}

extern "C" int vsb_ccl8_max(const uint32_t *rec_bits, int wpr, int W, int H, const float *dist,
                            size_t dist_pitch_elems, int32_t *labels, uint32_t *compmax, vsb_stream_t stream)
{
    VSB_REQUIRE(rec_bits && dist && labels && compmax, "vsb_ccl8_max: null pointer");
    VSB_REQUIRE(W > 0 && H > 0 && wpr >= (W + 31) / 32 && dist_pitch_elems >= (size_t)W, "vsb_ccl8_max: bad geometry");
    VSB_REQUIRE((long long)W * H < 0x7fffffffLL, "vsb_ccl8_max: image too large for int32 labels");
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((wpr + 255) / 256, H);
    k_ccl_init<<<grid, 256, 0, s>>>(rec_bits, wpr, W, H, labels, compmax);
    int rc = vsb_check_launch("k_ccl_init");
    if (rc) return rc;
    k_ccl_merge<<<grid, 256, 0, s>>>(rec_bits, wpr, W, H, labels);
    rc = vsb_check_launch("k_ccl_merge");
    if (rc) return rc;
    k_ccl_flatten_max<<<grid, 256, 0, s>>>(rec_bits, wpr, W, H, dist, dist_pitch_elems, labels, compmax);
    return vsb_check_launch("k_ccl_flatten_max");
}

// ---- ROI-normalised fade + merge + LUT ----------------------------------------------------------------
struct EFParams {
    const uint8_t *gray;
    size_t gpitch;
    uint8_t *out;
    size_t opitch;
    const uint32_t *rec;
    const float *dist;
    size_t dpitch;
    const int32_t *L;
    const uint32_t *cmax;
    const uint8_t *lut;
    double F;
    int W, H, wpr, aligned;
};

template <int ARITH>
__global__ void __launch_bounds__(256) k_enhanced_fade(const __grid_constant__ EFParams p)
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
    uint32_t m = (p.rec[(size_t)y * p.wpr + (x >> 5)] >> (x & 16)) & 0xFFFFu;
    if (m) {
        uint32_t gw[4] = {0, 0, 0, 0};
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const int xx = x + b;
            const float d = p.dist[(size_t)y * p.dpitch + xx];
            const float mx = __uint_as_float(p.cmax[p.L[y * p.W + xx]]);
            uint32_t g;
            if (ARITH == 0) { // float32, processing_core.py:295-302
                float den = (float)fmin((double)mx, p.F);
                if (den <= 0.0f) den = 1.0f;
                const float nrm = __fdiv_rn(fminf(fmaxf(d, 0.0f), den), den);
                g = (uint32_t)(int)__fmul_rn(255.0f, __fsub_rn(1.0f, nrm));
            } else {          // float64, processing_core.py:329-334
                const double den = fmin((double)mx, p.F);
                g = 0;
                if (den > 0.0) {
                    const double nrm = __ddiv_rn(fmin((double)d, den), den);
                    g = (uint32_t)(long long)__dmul_rn(__dsub_rn(1.0, nrm), 255.0);
                }
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

extern "C" int vsb_enhanced_fade(const uint8_t *gray, size_t gray_pitch, int W, int H, const uint32_t *rec_bits,
                                 int wpr, const float *dist, size_t dist_pitch_elems, const int32_t *labels,
                                 const uint32_t *compmax, double fade_limit, int arith, const uint8_t *lut256_dev,
                                 uint8_t *out, size_t out_pitch, vsb_stream_t stream)
{
    VSB_REQUIRE(gray && rec_bits && dist && labels && compmax && out, "vsb_enhanced_fade: null pointer");
    VSB_REQUIRE(W > 0 && H > 0 && wpr >= (W + 31) / 32 && gray_pitch >= (size_t)W && out_pitch >= (size_t)W &&
                    dist_pitch_elems >= (size_t)W, "vsb_enhanced_fade: bad geometry");
    VSB_REQUIRE(arith == 0 || arith == 1, "vsb_enhanced_fade: arith must be 0 (float32) or 1 (float64)");
    EFParams p;
    p.gray = gray; p.gpitch = gray_pitch; p.out = out; p.opitch = out_pitch; p.rec = rec_bits;
    p.dist = dist; p.dpitch = dist_pitch_elems; p.L = labels; p.cmax = compmax; p.lut = lut256_dev;
    p.F = fade_limit; p.W = W; p.H = H; p.wpr = wpr;
    p.aligned = (vsb_aligned16(gray, gray_pitch) && vsb_aligned16(out, out_pitch)) ? 1 : 0;
    dim3 grid(((W + 15) / 16 + 63) / 64, (H + 3) / 4);
    cudaStream_t s = (cudaStream_t)stream;
    if (arith == 0) k_enhanced_fade<0><<<grid, 256, 0, s>>>(p);
    else k_enhanced_fade<1><<<grid, 256, 0, s>>>(p);
    return vsb_check_launch("k_enhanced_fade");
}

// ---- per-label fades on caller-supplied labels (the reference's two helper functions) ------------------
__global__ void __launch_bounds__(256) k_label_max(const float *__restrict__ dist, const int32_t *__restrict__ labels,
                                                   size_t n, int num_labels, uint32_t *__restrict__ cmax)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int l = labels[i];
        if (l > 0 && l < num_labels) {
            const float d = dist[i];
            if (d > 0.0f) atomicMax(&cmax[l], __float_as_uint(d));
        }
    }
}

template <int ARITH>
__global__ void __launch_bounds__(256) k_fade_from_labels(const float *__restrict__ dist,
                                                          const int32_t *__restrict__ labels, size_t n, int num_labels,
                                                          double F, const uint32_t *__restrict__ cmax,
                                                          uint8_t *__restrict__ out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int l = labels[i];
        const float d = dist[i];
        const float mx = (l > 0 && l < num_labels) ? __uint_as_float(cmax[l]) : 0.0f;
        uint32_t g = 0;
        if (ARITH == 0) {
            float den = (float)fmin((double)mx, F);
            if (den <= 0.0f) den = 1.0f;
            const float nrm = __fdiv_rn(fminf(fmaxf(d, 0.0f), den), den);
            g = (uint32_t)(int)__fmul_rn(255.0f, __fsub_rn(1.0f, nrm));
        } else if (l > 0) {
            const double den = fmin((double)mx, F);
            if (den > 0.0) {
                const double nrm = __ddiv_rn(fmin((double)d, den), den);
                g = (uint32_t)(long long)__dmul_rn(__dsub_rn(1.0, nrm), 255.0);
            }
        }
        out[i] = (uint8_t)g;
    }
}

extern "C" int vsb_enhanced_from_labels(const float *dist, const int32_t *labels, int W, int H, int num_labels,
                                        double fade_limit, int arith, uint32_t *cmax, uint8_t *out, vsb_stream_t stream)
{
    VSB_REQUIRE(dist && labels && cmax && out, "vsb_enhanced_from_labels: null pointer");
    VSB_REQUIRE(W > 0 && H > 0 && num_labels >= 1, "vsb_enhanced_from_labels: bad geometry");
    VSB_REQUIRE(arith == 0 || arith == 1, "vsb_enhanced_from_labels: arith must be 0 or 1");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)W * H;
    VSB_CUDA(cudaMemsetAsync(cmax, 0, (size_t)num_labels * sizeof(uint32_t), s));
    const int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    k_label_max<<<grid, 256, 0, s>>>(dist, labels, n, num_labels, cmax);
    int rc = vsb_check_launch("k_label_max");
    if (rc) return rc;
    if (arith == 0) k_fade_from_labels<0><<<grid, 256, 0, s>>>(dist, labels, n, num_labels, fade_limit, cmax, out);
    else k_fade_from_labels<1><<<grid, 256, 0, s>>>(dist, labels, n, num_labels, fade_limit, cmax, out);
    return vsb_check_launch("k_fade_from_labels");
}

// ---- components with statistics (ROI_FADE: identify_rois, processing_core.py:69-105) -----------------------------
// cv2.connectedComponentsWithStats(binary, 8): area, bounding box, centroid sums per component, plus the index of
// the first 2x2 block (block raster order) touching the component -- the library numbers its labels in that order.
// Components get compact ids in arrival order; `cid` holds id + 1 at every foreground pixel, 0 elsewhere.
__global__ void __launch_bounds__(256) k_roi_roots(const uint32_t *__restrict__ bits, int wpr, int W, int H,
                                                   int32_t *__restrict__ L, int32_t *__restrict__ cid,
                                                   vsb_roi_stat *__restrict__ stats, int *__restrict__ count, int max_comp)
{
    const int wx = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (wx >= wpr) return;
    uint32_t m = bits[(size_t)y * wpr + wx];
    const int base = y * W + (wx << 5);
    for (int b = 0; b < 32 && (wx << 5) + b < W; ++b) {
        const int p = base + b;
        if (!((m >> b) & 1u)) { cid[p] = 0; continue; }
        const int root = uf_find(L, p);
        L[p] = root;
        if (root == p) {
            const int id = atomicAdd(count, 1);
            if (id < max_comp) {
                vsb_roi_stat s;
                s.root = p; s.area = 0; s.x0 = W; s.y0 = H; s.x1 = -1; s.y1 = -1; s.first_block = 0x7fffffff; s.pad = 0;
                s.sum_x = 0ULL; s.sum_y = 0ULL;
                stats[id] = s;
            }
            cid[p] = id + 1;
        }
    }
}

__global__ void __launch_bounds__(256) k_roi_stats(const uint32_t *__restrict__ bits, int wpr, int W, int H,
                                                   const int32_t *__restrict__ L, int32_t *__restrict__ cid,
                                                   vsb_roi_stat *__restrict__ stats, int max_comp)
{
    const int wx = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (wx >= wpr) return;
    uint32_t m = bits[(size_t)y * wpr + wx];
    if (!m) return;
    const int base = y * W + (wx << 5);
    const int bw = (W + 1) >> 1;
    // runs of pixels with the same root are accumulated locally, one set of atomics per run
    int run_id = -1, area = 0, xmin = 0, xmax = 0, fb = 0;
    unsigned long long sx = 0;
    auto flush = [&]() {
        if (run_id < 0 || run_id >= max_comp) return;
        vsb_roi_stat *s = &stats[run_id];
        atomicAdd(&s->area, area);
        atomicMin(&s->x0, xmin); atomicMax(&s->x1, xmax);
        atomicMin(&s->y0, y); atomicMax(&s->y1, y);
        atomicMin(&s->first_block, fb);
        atomicAdd(&s->sum_x, sx);
        atomicAdd(&s->sum_y, (unsigned long long)y * (unsigned long long)area);
    };
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const int p = base + b, x = (wx << 5) + b;
        const int id = cid[L[p]] - 1; // the root's own entry was written by k_roi_roots
        cid[p] = id + 1;
        if (id != run_id) {
            flush();
            run_id = id; area = 0; xmin = x; sx = 0; fb = (y >> 1) * bw + (x >> 1);
        }
        ++area; xmax = x; sx += (unsigned long long)x;
    }
    flush();
}

extern "C" int vsb_roi_components(const uint32_t *bits, int wpr, int W, int H, int32_t *labels, int32_t *cid,
                                  vsb_roi_stat *stats_dev, int *count_dev, int max_components, vsb_stream_t stream)
{
    VSB_REQUIRE(bits && labels && cid && stats_dev && count_dev, "vsb_roi_components: null pointer");
    VSB_REQUIRE(W > 0 && H > 0 && wpr >= (W + 31) / 32 && max_components > 0, "vsb_roi_components: bad geometry");
    VSB_REQUIRE((long long)W * H < 0x7fffffffLL, "vsb_roi_components: image too large for int32 labels");
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((wpr + 255) / 256, H);
    VSB_CUDA(cudaMemsetAsync(count_dev, 0, sizeof(int), s));
    k_ccl_init<<<grid, 256, 0, s>>>(bits, wpr, W, H, labels, (uint32_t *)cid);
    int rc = vsb_check_launch("k_ccl_init");
    if (rc) return rc;
    k_ccl_merge<<<grid, 256, 0, s>>>(bits, wpr, W, H, labels);
    rc = vsb_check_launch("k_ccl_merge");
    if (rc) return rc;
    k_roi_roots<<<grid, 256, 0, s>>>(bits, wpr, W, H, labels, cid, stats_dev, count_dev, max_components);
    rc = vsb_check_launch("k_roi_roots");
    if (rc) return rc;
    k_roi_stats<<<grid, 256, 0, s>>>(bits, wpr, W, H, labels, cid, stats_dev, max_components);
    return vsb_check_launch("k_roi_stats");
}
