// vsb_layer.cu -- the whole per-layer hot path in two launches:
//
//   k_pack_flags : threshold + bit-pack (K1) and, for every 128x32 tile, a uniformity flag (the tile's single
//                  gray value, or MIXED).  Reads every input byte exactly once.
//   k_layer      : one CTA per 128x32 output tile: receding mask, windowed distance search on rec pixels,
//                  fade gradient, merge, leading LUT (stage 1), bilateral filter (stage 2), Gaussian blur,
//                  trailing LUT, store.  Intermediate images never leave shared memory.
//
// Reference call sequence replaced (processing_pipeline.py:98-107): process_z_blending -> merge_to_output ->
// process_xy_pipeline for chains of the shape [apply_lut*] [bilateral_filter] [gaussian_blur] [apply_lut*].
//
// The per-pixel filters are FP32/LDS-issue bound, not HBM bound, so the kernel never runs them where the
// result is known: (a) tiles whose 3x3 tile neighbourhood is one uniform gray value and provably free of
// receding pixels are written as a constant without even reading the input ("quiet" tiles, decided from the
// flags of k_pack_flags); (b) inside other tiles, 4-pixel words whose whole filter window is one value copy
// it (sum v*w / sum w == v exactly; Gaussian taps sum to 2^16), decided from three bit-planes per staged row
// (word not uniform / differs from the word to the right / differs from the word below) OR-ed over the
// window rows.  Only the remaining pixels run the real filters, from a compacted list.
#include <cstdlib>
#include "vsb_tile.cuh"

#define ZL_NONE 5               // stage-1 image is the input itself (XY-only use)
#define ZL_SEARCHES(Z) ((Z) != ZL_NONE && (Z) != VSB_Z_CONST && (Z) != VSB_Z_ENHANCED) // modes that run the row search
#define LH_MAX 7                // max XY halo (bilateral radius + Gaussian radius)
#define RG_W 160                // staged region: x0-16 .. x0+143
#define RG_CH 10
#define RG_ROWS_MAX (VSB_TH + 2 * LH_MAX)
#define LBW 10                  // mask words per staged row: x0-96 .. x0+223
#define LB_ROWS_MAX (RG_ROWS_MAX + 2 * VSB_FAST_REACH)
#define FLAG_MIXED 0x100
#define FLAG_BINARY 0x200 // every byte of the tile is 0 or 255

static __device__ __forceinline__ float lds_f32(uint32_t a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}

struct BilTapsL {
    int32_t off[96];   // dy * RG_W + dx
    float sw[96];
    float cw[256];
};

#ifndef VSB_LAYER_CTAS_XY
#define VSB_LAYER_CTAS_XY 4 // resident CTAs per SM the XY instantiations are register-budgeted for
#endif

struct LayerParams {
    const uint8_t *in;
    size_t ipitch;
    uint8_t *out;
    size_t opitch;
    const uint32_t *cur;
    const uint32_t *prior[VSB_MAX_PRIORS];
    const uint16_t *cur_flags;
    const uint16_t *prior_flags[VSB_MAX_PRIORS];
    const float *T;
    const uint8_t *lut1, *lut2;
    int tiles_x, tiles_y;
    int W, H, wpr, aligned;
    int halo;            // hb + max(rx, ry)
    int hb, ntaps;       // bilateral radius (0 = none) and tap count
    int nkx, nky;        // Gaussian taps (0 = none)
    int32_t kx[7], ky[7];
    alignas(16) int tap_off[96];     // bilateral taps as byte offsets in the staged region, and their space weights: read through
    alignas(16) float tap_sw[96];    // the constant bank with a warp-uniform index, so they cost no vector instruction
    unsigned long long *stats; // optional work counters (vsb_layer_stats)
    struct { int T, s1, s2, bits, list, glist; } sm; // dynamic shared memory carve (bytes / words)
    int ntiles;
    const int *rec_list; // sparse: [0] = number of tiles with receding pixels, [1..] their indices (nullptr = all tiles)
    int gen_grid;        // CTAs of the list-mode launch
    int *gen_list;       // XY: k_layer2 appends the tiles it leaves to the general kernel; k_layer then walks that list
    int *dense_list;     // XY: interior tiles with thousands of receding pixels, set aside for k_layer2_dense
    uint16_t *tile_class; // k_layer2: per tile, from k_layer2_classify
    int dense_from;      // k_layer2: receding pixels from which a tile takes the dense launch's row-distance table
    int route_dense;     // k_layer2_classify: all-receding interior tiles go straight to the dense launch (modes that search)
    int rcap;            // k_layer2: entries of the receding-pixel list (sm.glist = entries of the bilateral pixel list)
    int two_valued;      // k_layer2: a window holding only LUT1[0] / LUT1[255] with >= 3 taps equal to the centre keeps the centre
    // VSB_Z_ENHANCED: distances at the receding pixels, their component roots and the per-root maxima (vsb_ccl8_max)
    const float *enh_dist; size_t enh_dpitch; const int32_t *enh_labels; const uint32_t *enh_cmax;
    int sparse;          // no-XY variant only: `out` already holds LUT1[gray]; rewrite just the receding pixels
    vsb_zparams z;
};
#define GL_OFF 1400 // Gaussian word list starts here in s_list (bilateral word list: at most 36*38 entries)

// ------------------------------------------------------------------------------------------------------
// One CTA = two vertically adjacent 128x32 tiles; a thread owns one mask word (32 pixels, two 128-bit loads, both
// issued before first use), so there is no cross-lane combine; the uniformity flag of each tile is a warp vote per
// 8 rows plus one shared-memory AND.
struct PackPriors { const uint32_t *p[VSB_MAX_PRIORS]; int n; };

// LUTOUT: the same pass also writes lut[gray] (the layer as it leaves a LUT-only pipeline wherever nothing recedes;
// vsb_layer_patch then rewrites the receding pixels), so the gray layer is read from HBM once.
template <bool ALIGNED, bool LUTOUT>
__global__ void __launch_bounds__(VSB_THREADS) k_pack_flags(const uint8_t *__restrict__ gray, size_t pitch, int W, int H,
                                                            uint32_t thr4, uint32_t *__restrict__ bits, int wpr,
                                                            uint16_t *__restrict__ flags, int tiles_x, int tiles_y,
                                                            const uint8_t *__restrict__ lut, uint8_t *__restrict__ out,
                                                            size_t opitch, int out_aligned, PackPriors pr,
                                                            int *__restrict__ rec_list)
{
    __shared__ int s_ok[2];
    __shared__ int s_bin[2];
    __shared__ int s_rec[2];
    __shared__ __align__(4) uint8_t s_plut[256];
    if (LUTOUT && lut && threadIdx.x < 64) ((uint32_t *)s_plut)[threadIdx.x] = ((const uint32_t *)lut)[threadIdx.x];
    const int t = threadIdx.x, half = t >> 7, u = t & 127;
    const int ty = blockIdx.y * 2 + half;
    const int x0 = blockIdx.x * VSB_TW, y0 = ty * VSB_TH;
    const int r = u >> 2, wq = u & 3;
    const int y = y0 + r, x = x0 + wq * 32;
    if (t < 2) { s_ok[t] = 1; s_bin[t] = 1; s_rec[t] = 0; }
    const int n = (y < H) ? max(0, min(32, W - x)) : 0;
    uint4 a = make_uint4(0, 0, 0, 0), b = make_uint4(0, 0, 0, 0);
    uint32_t first = 0;
    if (ty < tiles_y) first = gray[(size_t)y0 * pitch + x0];
    if (n == 32 && ALIGNED) {
        const uint8_t *src = gray + (size_t)y * pitch + x;
        a = vsb_ld16_stream(src);
        b = vsb_ld16_stream(src + 16);
    } else if (n > 0) {
        const uint8_t *src = gray + (size_t)y * pitch + x;
        a = vsb_ld_bytes(src, n);
        if (n > 16) b = vsb_ld_bytes(src + 16, n - 16);
    }
    const int wx = (x0 >> 5) + wq;
    uint32_t pacc = 0; // OR of the prior masks' words: requested together with the gray chunks (one round trip)
    if (LUTOUT && rec_list && y < H && wx < wpr)
        for (int j = 0; j < pr.n; ++j) pacc |= pr.p[j][(size_t)y * wpr + wx];
    __syncthreads();
    uint32_t word = vsb_gt_bits16(a, thr4) | (vsb_gt_bits16(b, thr4) << 16);
    if (n < 32) word &= n > 0 ? ((1u << n) - 1u) : 0u;
    if (y < H && wx < wpr) bits[(size_t)y * wpr + wx] = word;
    if (LUTOUT && rec_list) { // tiles with receding pixels (prior white, current black) go on the patch kernel's work list
        if (__any_sync(0xffffffffu, (pacc & ~word) != 0u) && (t & 31) == 0) s_rec[half] = 1;
    }
    if (LUTOUT && n > 0) {
        VsbLut Lp;
        Lp.s = lut ? s_plut : nullptr;
        Lp.l0 = lut ? (uint32_t)s_plut[0] * 0x01010101u : 0u;
        Lp.l255 = lut ? (uint32_t)s_plut[255] * 0x01010101u : 0xFFFFFFFFu;
        uint8_t *dst = out + (size_t)y * opitch + x;
        const uint4 oa = Lp.map16(a), ob = Lp.map16(b);
        if (n == 32 && out_aligned) { vsb_st16_stream(dst, oa); vsb_st16_stream(dst + 16, ob); }
        else { vsb_st_bytes(dst, oa, n); if (n > 16) vsb_st_bytes(dst + 16, ob, n - 16); }
    }
    const uint32_t bc = first * 0x01010101u;
    bool ok = true;
    if (n == 32) ok = (a.x == bc) & (a.y == bc) & (a.z == bc) & (a.w == bc) & (b.x == bc) & (b.y == bc) & (b.z == bc) & (b.w == bc);
    else if (n > 0) {
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        for (int i = 0; i < n; ++i) ok = ok && (((w[i >> 2] >> ((i & 3) * 8)) & 255u) == first);
    }
    if (!__all_sync(0xffffffffu, ok) && (t & 31) == 0) s_ok[half] = 0;
    {   // strictly binary tile (every byte 0 or 255; zero-filled bytes outside the image count as 0)
        const uint32_t w8[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t nb = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) nb |= w8[i] ^ (((w8[i] >> 7) & 0x01010101u) * 0xFFu);
        if (!__all_sync(0xffffffffu, nb == 0u) && (t & 31) == 0) s_bin[half] = 0;
    }
    __syncthreads();
    if (u == 0 && ty < tiles_y) {
        flags[ty * tiles_x + blockIdx.x] = (uint16_t)((s_ok[half] ? first : FLAG_MIXED) | (s_bin[half] ? FLAG_BINARY : 0));
        if (LUTOUT && rec_list && s_rec[half]) rec_list[1 + atomicAdd(&rec_list[0], 1)] = ty * tiles_x + (int)blockIdx.x;
    }
}

int vsb_pack_tma_launch(const uint8_t *gray, size_t pitch, int W, int H, uint32_t *bits, int wpr, uint16_t *flags, bool lutout,
                        const uint8_t *lut, uint8_t *out, size_t opitch, const uint32_t *const *prior_bits,
                        const uint16_t *const *prior_flags, int n_priors, int *rec_list, cudaStream_t s); // vsb_pack_tma.cu

static int pack_flags_launch(const uint8_t *gray, size_t pitch, int W, int H, int thresh, uint32_t *bits, int wpr,
                             uint16_t *flags, bool lutout, const uint8_t *lut, uint8_t *out, size_t opitch,
                             const uint32_t *const *prior_bits, const uint16_t *const *prior_flags, int n_priors, int *rec_list,
                             vsb_stream_t stream)
{
    VSB_REQUIRE(gray && bits && flags, "vsb_pack_flags: null pointer");
    VSB_REQUIRE(W > 0 && H > 0 && wpr >= ((W + 127) / 128) * 4 && pitch >= (size_t)W, "vsb_pack_flags: bad geometry");
    VSB_REQUIRE(thresh >= 0 && thresh <= 255, "vsb_pack_flags: thresh out of range");
    const int tiles_x = (W + VSB_TW - 1) / VSB_TW, tiles_y = (H + VSB_TH - 1) / VSB_TH;
    dim3 grid(tiles_x, (tiles_y + 1) / 2);
    const uint32_t thr4 = (uint32_t)thresh * 0x01010101u;
    cudaStream_t s = (cudaStream_t)stream;
    const int oal = out ? (vsb_aligned16(out, opitch) ? 1 : 0) : 0;
    const bool al = vsb_aligned16(gray, pitch);
    PackPriors pr;
    memset(&pr, 0, sizeof pr);
    if (rec_list) {
        VSB_REQUIRE(n_priors >= 0 && n_priors <= VSB_MAX_PRIORS && (n_priors == 0 || prior_bits), "vsb_pack_flags_lut: priors");
        pr.n = n_priors;
        for (int i = 0; i < n_priors; ++i) pr.p[i] = prior_bits[i];
        VSB_CUDA(cudaMemsetAsync(rec_list, 0, sizeof(int), s));
    }
    // the streaming kernel (TMA-fed, one warp per tile): threshold 127, 16-byte aligned rows, width a multiple of 16
    // (the TMA unit bounds-checks the innermost dimension in 16-byte units)
    static const bool no_tma = getenv("VSB_NO_TMA") != nullptr;
    if (!no_tma && thresh == 127 && al && (W & 15) == 0 && (!lutout || oal))
        return vsb_pack_tma_launch(gray, pitch, W, H, bits, wpr, flags, lutout, lut, out, opitch, prior_bits, prior_flags, n_priors,
                                   rec_list, s);
    if (lutout) {
        if (al) k_pack_flags<true, true><<<grid, VSB_THREADS, 0, s>>>(gray, pitch, W, H, thr4, bits, wpr, flags, tiles_x, tiles_y, lut, out, opitch, oal, pr, rec_list);
        else k_pack_flags<false, true><<<grid, VSB_THREADS, 0, s>>>(gray, pitch, W, H, thr4, bits, wpr, flags, tiles_x, tiles_y, lut, out, opitch, oal, pr, rec_list);
    } else {
        if (al) k_pack_flags<true, false><<<grid, VSB_THREADS, 0, s>>>(gray, pitch, W, H, thr4, bits, wpr, flags, tiles_x, tiles_y, nullptr, nullptr, 0, 0, pr, nullptr);
        else k_pack_flags<false, false><<<grid, VSB_THREADS, 0, s>>>(gray, pitch, W, H, thr4, bits, wpr, flags, tiles_x, tiles_y, nullptr, nullptr, 0, 0, pr, nullptr);
    }
    return vsb_check_launch("k_pack_flags");
}

extern "C" int vsb_pack_flags(const uint8_t *gray, size_t pitch, int W, int H, int thresh, uint32_t *bits, int wpr,
                              uint16_t *flags, vsb_stream_t stream)
{
    return pack_flags_launch(gray, pitch, W, H, thresh, bits, wpr, flags, false, nullptr, nullptr, 0, nullptr, nullptr, 0, nullptr, stream);
}

extern "C" int vsb_pack_flags_lut(const uint8_t *gray, size_t pitch, int W, int H, int thresh, uint32_t *bits, int wpr,
                                  uint16_t *flags, const uint8_t *lut256_dev, uint8_t *out, size_t out_pitch,
                                  const uint32_t *const *prior_bits, const uint16_t *const *prior_flags, int n_priors,
                                  int *rec_tiles, vsb_stream_t stream)
{
    VSB_REQUIRE(out && out != gray && out_pitch >= (size_t)W, "vsb_pack_flags_lut: output missing / in place / pitch < width");
    return pack_flags_launch(gray, pitch, W, H, thresh, bits, wpr, flags, true, lut256_dev, out, out_pitch, prior_bits, prior_flags,
                             n_priors, rec_tiles, stream);
}

// ------------------------------------------------------------------------------------------------------
// Gaussian of one 4-pixel word: rows rr-RY..rr+RY, columns 4w-RX..4w+3+RX of S (pitch RG_W), Q8.8 taps.
template <int RX, int RY>
__device__ __forceinline__ uint32_t gauss_word(const uint8_t (*S)[RG_W], int rr, int w, const int32_t *kx, const int32_t *ky)
{
    uint32_t acc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int dy = 0; dy <= 2 * RY; ++dy) {
        const uint32_t *row = (const uint32_t *)&S[rr + dy - RY][4 * w - 4];
        const uint32_t wa = row[0], wb = row[1], wc = row[2];
        uint32_t b[12];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            b[i] = (wa >> (8 * i)) & 255u; b[4 + i] = (wb >> (8 * i)) & 255u; b[8 + i] = (wc >> (8 * i)) & 255u;
        }
        const uint32_t kyv = (uint32_t)ky[dy];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t h = 0;
#pragma unroll
            for (int dx = 0; dx <= 2 * RX; ++dx) h += (uint32_t)kx[dx] * b[4 + j + dx - RX];
            acc[j] += kyv * h;
        }
    }
    uint32_t o = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) o |= min((acc[j] + 32768u) >> 16, 255u) << (8 * j);
    return o;
}

template <int RX>
__device__ __forceinline__ uint32_t gauss_word_ry(const uint8_t (*S)[RG_W], int rr, int w, const int32_t *kx,
                                                  const int32_t *ky, int ry)
{
    switch (ry) {
    case 0: return gauss_word<RX, 0>(S, rr, w, kx, ky);
    case 1: return gauss_word<RX, 1>(S, rr, w, kx, ky);
    case 2: return gauss_word<RX, 2>(S, rr, w, kx, ky);
    default: return gauss_word<RX, 3>(S, rr, w, kx, ky);
    }
}

#define PHASE(slot)                                                                                     \
    do {                                                                                                \
        if (p.stats && t == 0) {                                                                        \
            const long long now__ = clock64();                                                          \
            atomicAdd(&p.stats[slot], (unsigned long long)(now__ - tprev));                             \
            tprev = now__;                                                                              \
        }                                                                                               \
    } while (0)

template <int ZMODE, int METRIC, bool XY>
__global__ void __launch_bounds__(VSB_THREADS, XY ? VSB_LAYER_CTAS_XY : 7) k_layer(const __grid_constant__ LayerParams p,
                                                       const BilTapsL *__restrict__ taps)
{
    // work list of the patch launch: CTAs beyond the number of listed tiles leave before any set-up
    if (!XY && p.rec_list && (int)blockIdx.x >= p.rec_list[0]) return;
    extern __shared__ __align__(16) uint8_t dsm[];
    float *s_T = (float *)(dsm + p.sm.T);
    uint8_t (*s1)[RG_W] = (uint8_t (*)[RG_W])(dsm + p.sm.s1);      // stage 1; later the Gaussian output
    uint8_t (*s2buf)[RG_W] = (uint8_t (*)[RG_W])(dsm + p.sm.s2);   // stage 2 (bilateral output)
    uint32_t (*s_bits)[LBW] = (uint32_t (*)[LBW])(dsm + p.sm.bits);
    uint16_t *s_list = (uint16_t *)(dsm + p.sm.list);             // bilateral word list, Gaussian word list at +GL_OFF
    uint16_t *s_rlist = (uint16_t *)(dsm + p.sm.s2);              // receding pixel list overlays [s2 | list]
    __shared__ uint32_t s_rowany[(LB_ROWS_MAX + 31) / 32];
    __shared__ __align__(8) uint32_t s_zpl32[RG_ROWS_MAX][2], s_wpl32[RG_ROWS_MAX][2]; // per row: words equal to LUT1[0]x4 / LUT1[255]x4
    unsigned long long *s_zpl = (unsigned long long *)s_zpl32, *s_wpl = (unsigned long long *)s_wpl32; // (32-bit halves: native shared atomics)
    __shared__ float s_cw2[512]; // colour weight by signed difference: s_cw2[255 + d] = cw[|d|]
    __shared__ __align__(4) uint8_t s_lut1[256];
    __shared__ __align__(4) uint8_t s_lut2[256];
    __shared__ int s_cnt, s_gcnt, s_quiet;

    const int t = threadIdx.x, lane = t & 31;
    const int W = p.W, H = p.H, wpr = p.wpr;
    const int np = p.z.n_priors;
    const int r = t >> 3, c = t & 7;
    const int Hh = XY ? p.halo : 0;
    const int RH = VSB_TH + 2 * Hh;
    const int R = p.z.reach;
    const int nsrc = RH * 6;
    long long tprev = p.stats ? clock64() : 0;

    // ---- once per CTA: tables ------------------------------------------------------------------------------------
    VsbLut L1, L2;
    L1.s = p.lut1 ? s_lut1 : nullptr; L1.l0 = 0u; L1.l255 = 0xFFFFFFFFu; // broadcasts filled in after the first barrier
    L2.s = p.lut2 ? s_lut2 : nullptr; L2.l0 = 0u; L2.l255 = 0xFFFFFFFFu;
    if (t < 64 && p.lut1 && !(!XY && p.sparse)) ((uint32_t *)s_lut1)[t] = ((const uint32_t *)p.lut1)[t];
    bool late_staged = false; // trailing LUT and bilateral tables: staged by the first tile that is not trivially uniform
    bool t_staged = false; // chamfer table [(R+1) x (R+1)]: staged by the first tile that needs it

    // gen_list (XY launches that follow k_layer2): a 1-D grid of CTAs strides over the listed tiles -- the tiles whose
    // region is not strictly binary; the list is empty for the input the reference asks for
    // rec_list (the patch launch of LUT-only chains) is walked the same way, by a grid of a quarter of the tiles: when up to a
    // quarter of the tiles is listed every CTA takes one tile and none starts in vain; beyond that some take a second one
    const int *glist = XY ? p.gen_list : p.rec_list;
    const int n_listed = glist ? glist[0] : 1;
    for (int li = glist ? (int)blockIdx.x : 0; li < n_listed; li += glist ? (int)gridDim.x : 1) {
        int tile = (int)(blockIdx.y * gridDim.x + blockIdx.x);
        int wl_tx = (int)blockIdx.x, wl_ty = (int)blockIdx.y;
        if (glist) {
            tile = glist[1 + li];
            wl_tx = tile % p.tiles_x; wl_ty = tile / p.tiles_x;
        }
        const int tx = wl_tx, ty = wl_ty; // 2-D grid: no division
        const int x0 = tx * VSB_TW, y0 = ty * VSB_TH;
        const int oy = y0 + r, ox = x0 + c * 16;
        const int nvalid = (oy < H) ? max(0, min(16, W - ox)) : 0;

        // ---- P1: prologue ---------------------------------------------------------------------------------------
        uint32_t rwv[2] = {0, 0};
        uint4 gv[2];
        int gstate[2]; // 0 = outside the image, 1 = full chunk, 2 = partial (byte path)
        uint32_t ref_raw;
        if (!XY) {
            // no XY filter: the region is the tile itself -- thread t owns chunk (r, c), threads < 128 one mask word
            gv[0] = gv[1] = make_uint4(0, 0, 0, 0);
            gstate[0] = gstate[1] = 0;
            if (ZMODE != ZL_NONE && t < 128) {
                const int rr = t >> 2, k = t & 3;
                const int y = y0 + rr, wx = (x0 >> 5) + k;
                if (y < H && wx < wpr) {
                    const size_t o = (size_t)y * wpr + wx;
                    uint32_t acc = 0;
                    for (int j = 0; j < np; ++j) acc |= p.prior[j][o];
                    rwv[0] = acc & ~p.cur[o];
                }
            }
            ref_raw = 0;
            if (p.sparse && ZL_SEARCHES(ZMODE)) {
                // patch launch: every listed tile has receding pixels, so the search window and the chamfer table are
                // requested together with the tile's mask words -- one memory round trip per tile instead of two
                if (METRIC == VSB_METRIC_CHAMFER5 && !t_staged) {
                    const int nT = (R + 1) * (R + 1);
                    const int nfull = ((((size_t)p.T) & 15) == 0) ? (nT >> 2) : 0;
                    for (int i = t; i < nfull; i += VSB_THREADS) ((float4 *)s_T)[i] = ((const float4 *)p.T)[i];
                    for (int i = 4 * nfull + t; i < nT; i += VSB_THREADS) s_T[i] = p.T[i];
                    t_staged = true;
                }
                stage_mask_rows<LBW>(s_bits, p.cur, wpr, H, (x0 >> 5) - 3, y0 - R, RH + 2 * R);
            }
            if (!p.sparse) {
                if (nvalid == 16 && p.aligned) { gv[0] = vsb_ld16_stream(p.in + (size_t)oy * p.ipitch + ox); gstate[0] = 1; }
                else if (nvalid > 0) gstate[0] = 2;
                ref_raw = p.in[(size_t)y0 * p.ipitch + x0];
            }
            if (t == 0) { s_cnt = 0; s_gcnt = 0; s_quiet = 0; }
            __syncthreads();
        } else {
            if (ZMODE != ZL_NONE) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int i = t + u * VSB_THREADS;
                    if (i < nsrc) {
                        const int rr = i / 6, k = i - rr * 6;
                        const int y = y0 - Hh + rr, wx = (x0 >> 5) - 1 + k;
                        if (y >= 0 && y < H && wx >= 0 && wx < wpr) {
                            const size_t o = (size_t)y * wpr + wx;
                            uint32_t acc = 0;
                            for (int j = 0; j < np; ++j) acc |= p.prior[j][o];
                            uint32_t rw = acc & ~p.cur[o];
                            if (k == 0) rw &= Hh ? (0xFFFFFFFFu << (32 - Hh)) : 0u;
                            if (k == 5) rw &= (1u << Hh) - 1u;
                            rwv[u] = rw;
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int i = t + u * VSB_THREADS;
                gv[u] = make_uint4(0, 0, 0, 0);
                gstate[u] = 0;
                if (i < RH * RG_CH) {
                    const int rr = i / RG_CH, cc = i - rr * RG_CH;
                    const int y = y0 - Hh + rr, x = x0 - 16 + cc * 16;
                    if (y >= 0 && y < H) {
                        if (x >= 0 && x + 16 <= W && p.aligned) {
                            gv[u] = vsb_ld16_stream(p.in + (size_t)y * p.ipitch + x);
                            gstate[u] = 1;
                        } else if (min(x + 16, W) > max(x, 0)) gstate[u] = 2;
                    }
                }
            }
            ref_raw = p.in[(size_t)y0 * p.ipitch + x0];
            if (t == 0) { s_cnt = 0; s_gcnt = 0; s_quiet = 0; }
            if (XY && t < RH) { s_zpl[t] = 0ULL; s_wpl[t] = 0ULL; }
            __syncthreads();
        }

        if (L1.s) { L1.l0 = (uint32_t)s_lut1[0] * 0x01010101u; L1.l255 = (uint32_t)s_lut1[255] * 0x01010101u; }
        const uint32_t ref = L1.map4(ref_raw * 0x01010101u);
        bool uni = true;
        const bool sparse = !XY && p.sparse != 0;
#pragma unroll
        for (int u = 0; u < (XY ? 2 : 1); ++u) {
            const int i = t + u * VSB_THREADS;
            if (sparse) break;
            if (!XY || i < RH * RG_CH) {
                const int rr = XY ? i / RG_CH : r, cc = XY ? i - rr * RG_CH : c + 1;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (gstate[u] == 1) {
                    v = L1.map16(gv[u]);
                    uni = uni && (v.x == ref) && (v.y == ref) && (v.z == ref) && (v.w == ref);
                    if (XY) { // value planes for the filter bypass (out-of-image and ragged chunks stay "unknown")
                        const uint32_t Z4 = L1.l0, W4 = L1.l255;
                        const uint32_t zb = (uint32_t)(v.x == Z4) | ((uint32_t)(v.y == Z4) << 1) | ((uint32_t)(v.z == Z4) << 2) | ((uint32_t)(v.w == Z4) << 3);
                        const uint32_t wb = (uint32_t)(v.x == W4) | ((uint32_t)(v.y == W4) << 1) | ((uint32_t)(v.z == W4) << 2) | ((uint32_t)(v.w == W4) << 3);
                        if (zb) atomicOr(&s_zpl32[rr][cc >> 3], zb << ((4 * cc) & 31));
                        if (wb) atomicOr(&s_wpl32[rr][cc >> 3], wb << ((4 * cc) & 31));
                    }
                } else if (gstate[u] == 2) {
                    const int y = y0 - Hh + rr, x = x0 - 16 + cc * 16;
                    const int xa = max(x, 0), xb = min(x + 16, W);
                    uint32_t w[4] = {0, 0, 0, 0};
                    const uint8_t *src = p.in + (size_t)y * p.ipitch;
                    for (int xx = xa; xx < xb; ++xx) {
                        const int k = xx - x;
                        const uint32_t bb = L1.s ? L1.s[src[xx]] : src[xx];
                        uni = uni && (bb == (ref & 255u));
                        w[k >> 2] |= bb << ((k & 3) * 8);
                    }
                    v = make_uint4(w[0], w[1], w[2], w[3]);
                }
                *(uint4 *)&s1[rr][cc * 16] = v;
            }
        }
        {   // one barrier: bit 0 = some receding pixel, bit 1 = region not uniform
            const uint32_t code = ((rwv[0] | rwv[1]) ? 1u : 0u) | (uni ? 0u : 2u);
            const uint32_t wcode = __reduce_or_sync(0xffffffffu, code);
            if (lane == 0 && wcode) atomicOr(&s_quiet, (int)wcode);
        }
        __syncthreads();
        const int tile_state = s_quiet;
        const int any_rec = tile_state & 1;
        if (p.stats && t == 0) atomicAdd(&p.stats[tile_state == 0 ? 0 : 1], 1ULL);
        PHASE(8);
        do { // one pass; `break` = tile finished
        if (sparse) {
            if (!any_rec) break; // the pre-filled output stands
            if (t < 64 && p.lut1) ((uint32_t *)s_lut1)[t] = ((const uint32_t *)p.lut1)[t]; // visible after P2's barriers
        }
        if (tile_state == 0) { // uniform region, nothing recedes: every filter of a constant is that constant
            if (nvalid > 0) {
                const uint32_t cv = p.lut2 ? (uint32_t)p.lut2[ref & 255u] * 0x01010101u : ref;
                uint8_t *dst = p.out + (size_t)oy * p.opitch + ox;
                const uint4 o = make_uint4(cv, cv, cv, cv);
                if (nvalid == 16 && p.aligned) vsb_st16_stream(dst, o);
                else vsb_st_bytes(dst, o, nvalid);
            }
            break;
        }

        if (!late_staged) { // visible to all threads after the next barrier (every path below has one before first use)
            if (t < 64 && p.lut2) ((uint32_t *)s_lut2)[t] = ((const uint32_t *)p.lut2)[t];
            if (XY && p.hb > 0) {
                const float cwt = taps->cw[t];
                s_cw2[255 + t] = cwt; s_cw2[255 - t] = cwt;
            }
            late_staged = true;
        }

        // ---- P2: distance search + gradient on the receding pixels of the region ----------------------------
        if (ZMODE != ZL_NONE && any_rec) {
            const int brows = RH + 2 * R;
            if (ZL_SEARCHES(ZMODE) && !sparse) {
                if (METRIC == VSB_METRIC_CHAMFER5 && !t_staged) {
                    const int nT = (R + 1) * (R + 1);
                    const int nfull = ((((size_t)p.T) & 15) == 0) ? (nT >> 2) : 0; // 16-byte groups when the table is aligned
                    for (int i = t; i < nfull; i += VSB_THREADS) ((float4 *)s_T)[i] = ((const float4 *)p.T)[i];
                    for (int i = 4 * nfull + t; i < nT; i += VSB_THREADS) s_T[i] = p.T[i];
                    t_staged = true;
                }
                stage_mask_rows<LBW>(s_bits, p.cur, wpr, H, (x0 >> 5) - 3, y0 - Hh - R, brows);
            }
            if (ZL_SEARCHES(ZMODE) && sparse) stage_rowany_rows<LBW>(s_rowany, s_bits, brows); // window staged before the first barrier
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int i = t + u * VSB_THREADS;
                uint32_t rw = rwv[u];
                if (rw) {
                    // XY: item i = (row i/6, word i%6 of the 6-word span starting one word left of the tile);
                    // no XY: item t = (row t>>2, tile word t&3)
                    const int rr = XY ? i / 6 : (t >> 2);
                    const int k = XY ? i - rr * 6 : (t & 3) + 1;
                    int base = atomicAdd(&s_cnt, __popc(rw));
                    while (rw) {
                        const int bq = __ffs(rw) - 1;
                        rw &= rw - 1;
                        s_rlist[base++] = (uint16_t)((rr << 8) | (32 * k + bq - 16));
                    }
                }
            }
            __syncthreads();
            if (!sparse) {
                if (ZL_SEARCHES(ZMODE)) stage_rowany_rows<LBW>(s_rowany, s_bits, brows);
                __syncthreads();
            }
            const int total = s_cnt;
            if (p.stats && t == 0) { atomicAdd(&p.stats[2], 1ULL); atomicAdd(&p.stats[3], (unsigned long long)total); }
            PHASE(9);
            const float Ff = p.z.fade;
            for (int i = t; i < total; i += VSB_THREADS) {
                const int code = s_rlist[i];
                const int rr = code >> 8, rc = code & 255;
                const int x = x0 - 16 + rc, y = y0 - Hh + rr;
                float best = 0.0f;
                if (ZL_SEARCHES(ZMODE)) best = tile_search<METRIC, LBW, false>(s_bits, s_rowany, s_T, R, rr + R, rc + 80);
                int g;
                if (ZMODE == VSB_Z_FIXED) {
                    const float n = __fdiv_rn(fminf(best, Ff), p.z.den);
                    g = (int)__fmul_rn(255.0f, __fsub_rn(1.0f, n));
                } else if (ZMODE == VSB_Z_WEIGHTED) {
                    float acc = 0.0f;
                    const size_t o = (size_t)y * wpr + (x >> 5);
                    const uint32_t ncur = ~p.cur[o];
                    for (int j = 0; j < np; ++j) {
                        if (p.z.weights[j] > 0.0f && (((p.prior[j][o] & ncur) >> (x & 31)) & 1u)) {
                            const float n = __fsub_rn(1.0f, __fdiv_rn(fminf(best, p.z.fades[j]), p.z.dens[j]));
                            acc = __fadd_rn(acc, __fmul_rn(n, p.z.weights[j]));
                        }
                    }
                    g = (int)__fmul_rn(255.0f, __fdiv_rn(acc, p.z.total_weight));
                } else if (ZMODE == VSB_Z_ENHANCED) {
                    // fade normalised by min(component maximum, F): the float32 (SciPy) or float64 (Numba) arithmetic of
                    // processing_core.py:295-302 / :329-334
                    const float d = p.enh_dist[(size_t)y * p.enh_dpitch + x];
                    const float mx = __uint_as_float(p.enh_cmax[p.enh_labels[(size_t)y * W + x]]);
                    if (p.z.enhanced_arith == 0) {
                        float den = (float)fmin((double)mx, p.z.fade_limit);
                        if (den <= 0.0f) den = 1.0f;
                        const float nrm = __fdiv_rn(fminf(fmaxf(d, 0.0f), den), den);
                        g = (int)__fmul_rn(255.0f, __fsub_rn(1.0f, nrm));
                    } else {
                        const double den = fmin((double)mx, p.z.fade_limit);
                        g = 0;
                        if (den > 0.0) g = (int)(long long)__dmul_rn(__dsub_rn(1.0, __ddiv_rn(fmin((double)d, den), den)), 255.0);
                    }
                } else {
                    g = p.z.const_val;
                }
                const int graw = (int)p.in[(size_t)y * p.ipitch + x];
                const int m = max(graw, g & 255); // merge with the raw gray value
                if (sparse) { p.out[(size_t)y * p.opitch + x] = L1.s ? L1.s[m] : (uint8_t)m; continue; }
                s1[rr][rc] = L1.s ? L1.s[m] : (uint8_t)m;
                if (XY) { // the word no longer holds a known constant
                    atomicAnd(&s_zpl32[rr][rc >> 7], ~(1u << ((rc >> 2) & 31)));
                    atomicAnd(&s_wpl32[rr][rc >> 7], ~(1u << ((rc >> 2) & 31)));
                }
            }
            if (sparse) break; // receding pixels rewritten in place
            __syncthreads();
            if (t == 0) s_cnt = 0;
            PHASE(10);
        }

        // ---- P3: REFLECT_101 fix-up of out-of-image halo positions (border tiles only) -----------------------------
        const bool border = (x0 - Hh < 0) || (x0 + VSB_TW + Hh > W) || (y0 - Hh < 0) || (y0 + VSB_TH + Hh > H);
        if (XY && border && Hh > 0) {
            for (int i = t; i < RH * RG_W; i += VSB_THREADS) {
                const int rr = i / RG_W, rc = i - rr * RG_W;
                const int y = y0 - Hh + rr, x = x0 - 16 + rc;
                if (x < -Hh || x >= W + Hh || y < -Hh || y >= H + Hh) continue;
                if (y < 0 || y >= H || x < 0 || x >= W) {
                    const int ys = vsb_reflect101(y, H), xs = vsb_reflect101(x, W);
                    s1[rr][rc] = s1[ys - (y0 - Hh)][xs - (x0 - 16)];
                }
            }
        }
        __syncthreads();

        const int hb = XY ? p.hb : 0;
        const int rx = XY ? (p.nkx >> 1) : 0, ry = XY ? (p.nky >> 1) : 0;
        const int hg = max(rx, ry);
        const bool gauss = XY && p.nkx > 0;
        uint8_t (*S2)[RG_W] = hb > 0 ? s2buf : s1;                  // Gaussian input
        uint8_t (*OUT)[RG_W] = gauss ? (hb > 0 ? s1 : s2buf) : S2;  // what the store phase reads

        if (XY && (hb > 0 || gauss)) {
            // ---- P4/P5: bypass lists from the value planes.  A row's plane has one bit per 4-byte word: "the word is
            //      LUT1[0]x4" (Z) or "LUT1[255]x4" (W); rows hold no bit for out-of-image or receding words.  One thread
            //      per stage-2 row ANDs the planes over the window rows and, with shifts, over the window words: a word
            //      whose whole window is one constant keeps its value through the bilateral (and the Gaussian); all
            //      other words are appended to the bilateral / Gaussian word lists.  Meanwhile every thread copies
            //      stage 1 to the buffer the listed words are overwritten in.
            const int r_lo = Hh - hg, r_hi = Hh + VSB_TH + hg; // region rows on which stage 2 is needed
            const int eb = hb, eg = hb + rx;
            if (t >= r_lo && t < r_hi) {
                const bool trow = t >= Hh && t < Hh + VSB_TH;
                unsigned long long Z = ~0ULL, Wp = ~0ULL;
                for (int d = -hb; d <= hb; ++d) { Z &= s_zpl[t + d]; Wp &= s_wpl[t + d]; }
                unsigned long long badb = 0ULL, badg = 0ULL;
                if (hb > 0) {
                    unsigned long long cz = Z, cw = Wp;
                    for (int sft = 1; sft <= ((eb + 3) >> 2); ++sft) { cz &= (Z >> sft) & (Z << sft); cw &= (Wp >> sft) & (Wp << sft); }
                    const int wlo = (16 - hg) >> 2, whi = (143 + hg) >> 2; // words stage 2 is needed on
                    badb = ~(cz | cw) & ((~0ULL << wlo) & ((2ULL << whi) - 1ULL));
                }
                if (gauss && trow) {
                    for (int d = hb + 1; d <= hb + ry; ++d) { Z &= s_zpl[t + d] & s_zpl[t - d]; Wp &= s_wpl[t + d] & s_wpl[t - d]; }
                    unsigned long long cz = Z, cw = Wp;
                    for (int sft = 1; sft <= ((eg + 3) >> 2); ++sft) { cz &= (Z >> sft) & (Z << sft); cw &= (Wp >> sft) & (Wp << sft); }
                    badg = ~(cz | cw) & 0xFFFFFFFF0ULL; // tile words 4 .. 35
                }
                if (badb) {
                    int base = atomicAdd(&s_cnt, __popcll(badb));
                    while (badb) { const int w = __ffsll((long long)badb) - 1; badb &= badb - 1ULL; s_list[base++] = (uint16_t)((t << 8) | w); }
                }
                if (badg) {
                    int base = GL_OFF + atomicAdd(&s_gcnt, __popcll(badg));
                    while (badg) { const int w = __ffsll((long long)badg) - 1; badg &= badg - 1ULL; s_list[base++] = (uint16_t)((t << 8) | w); }
                }
            }
            if (hb > 0) {
                for (int i = t; i < (r_hi - r_lo) * RG_CH; i += VSB_THREADS) {
                    const int rq = i / RG_CH, cc = i - rq * RG_CH;
                    *(uint4 *)&s2buf[r_lo + rq][16 * cc] = *(const uint4 *)&s1[r_lo + rq][16 * cc];
                }
            } else {
                *(uint4 *)&s2buf[r + Hh][16 + 16 * c] = *(const uint4 *)&s1[r + Hh][16 + 16 * c];
            }
            __syncthreads();
            PHASE(11);
            PHASE(12);

            if (hb > 0) {
                const int total = s_cnt; // words; four lanes per word
                if (p.stats && t == 0) atomicAdd(&p.stats[4], 4ULL * (unsigned long long)total);
                const int ntaps = p.ntaps;
                for (int i = t; i < 4 * total; i += VSB_THREADS) {
                    const int code = s_list[i >> 2];
                    const int rr = code >> 8, rc = 4 * (code & 255) + (i & 3);
                    const int y = y0 - Hh + rr, x = x0 - 16 + rc;
                    if (y < 0 || y >= H || x < 0 || x >= W) continue; // reflected below
                    const uint8_t *ctr = &s1[rr][rc];
                    const int v0 = *ctr;
                    uint32_t cwa = (uint32_t)__cvta_generic_to_shared(s_cw2) + 4u * (uint32_t)(255 - v0);
                    asm volatile("" : "+r"(cwa)); // opaque: keeps the per-tap address at one IMAD
                    float sacc = 0.0f, ws = 0.0f;
#pragma unroll 4
                    for (int k = 0; k < ntaps; ++k) {
                        const int v = ctr[p.tap_off[k]];
                        const float w = __fmul_rn(p.tap_sw[k], lds_f32(cwa + 4u * (uint32_t)v));
                        sacc = __fmaf_rn((float)v, w, sacc);
                        ws = __fadd_rn(ws, w);
                    }
                    const int o = __float2int_rn(__fdiv_rn(sacc, ws));
                    s2buf[rr][rc] = (uint8_t)min(max(o, 0), 255);
                }
                __syncthreads();
                if (border && hg > 0) {
                    const int nr = r_hi - r_lo;
                    for (int i = t; i < nr * RG_W; i += VSB_THREADS) {
                        const int rr = r_lo + i / RG_W, rc = i % RG_W;
                        const int y = y0 - Hh + rr, x = x0 - 16 + rc;
                        if (x < -hg || x >= W + hg || y < -hg || y >= H + hg) continue;
                        if (y < 0 || y >= H || x < 0 || x >= W) {
                            const int ys = vsb_reflect101(y, H), xs = vsb_reflect101(x, W);
                            s2buf[rr][rc] = s2buf[ys - (y0 - Hh)][xs - (x0 - 16)];
                        }
                    }
                    __syncthreads();
                }
            }
            PHASE(13);

            // ---- P6: Gaussian on the words that are not provably constant ----------------------------------------------
            if (gauss) {
                const int gtotal = s_gcnt;
                if (p.stats && t == 0) atomicAdd(&p.stats[5], (unsigned long long)gtotal);
                for (int i = t; i < gtotal; i += VSB_THREADS) {
                    const int code = s_list[GL_OFF + i];
                    const int rr = code >> 8, w = code & 255;
                    uint32_t o;
                    switch (rx) {
                    case 0: o = gauss_word_ry<0>(S2, rr, w, p.kx, p.ky, ry); break;
                    case 1: o = gauss_word_ry<1>(S2, rr, w, p.kx, p.ky, ry); break;
                    case 2: o = gauss_word_ry<2>(S2, rr, w, p.kx, p.ky, ry); break;
                    default: o = gauss_word_ry<3>(S2, rr, w, p.kx, p.ky, ry); break;
                    }
                    *(uint32_t *)&OUT[rr][4 * w] = o;
                }
                __syncthreads();
            }
        }
        PHASE(14);

        // ---- P7: trailing LUT, store ----------------------------------------------------------------------------------
        if (L2.s) { L2.l0 = (uint32_t)s_lut2[0] * 0x01010101u; L2.l255 = (uint32_t)s_lut2[255] * 0x01010101u; }
        if (nvalid > 0) {
            uint4 o = *(const uint4 *)&OUT[r + Hh][16 + 16 * c];
            o = L2.map16(o);
            uint8_t *dst = p.out + (size_t)oy * p.opitch + ox;
            if (nvalid == 16 && p.aligned) vsb_st16_stream(dst, o);
            else vsb_st_bytes(dst, o, nvalid);
        }
        } while (0);
        if (glist) __syncthreads(); // shared buffers are reused by the next listed tile
    }
}

#include "vsb_layer2.cuh"

template <int ZMODE, int METRIC, bool XY>
static int launch_layer_v(LayerParams &p, const BilTapsL *taps, size_t smem, cudaStream_t s)
{
    static bool seen[VSB_MAX_DEVICES] = {false}; // per instantiation and device
    if (vsb_first_on_device(seen)) {
        VSB_CUDA(cudaFuncSetAttribute(k_layer<ZMODE, METRIC, XY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    }
    dim3 grid(p.tiles_x, p.tiles_y);
    if (XY && p.gen_list) grid = dim3(p.gen_grid, 1);
    if (!XY && p.rec_list) grid = dim3(p.gen_grid, 1);
    k_layer<ZMODE, METRIC, XY><<<grid, VSB_THREADS, smem, s>>>(p, taps);
    return vsb_check_launch("k_layer");
}

// binary fast path: k_layer2 over all tiles (it lists the tiles it does not take), then k_layer over that list
template <int ZMODE, int METRIC, int SHAPE>
static int launch_layer2_s(LayerParams &p, LayerParams &pd, const BilTapsL *taps, size_t smem2, size_t smem2d, int n_sms, cudaStream_t s)
{
    static bool seen[VSB_MAX_DEVICES] = {false}; // per instantiation and device
    if (vsb_first_on_device(seen)) {
        VSB_CUDA(cudaFuncSetAttribute(k_layer2<ZMODE, METRIC, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        VSB_CUDA(cudaFuncSetAttribute(k_layer2<ZMODE, METRIC, SHAPE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        VSB_CUDA(cudaFuncSetAttribute(k_layer2_dense<ZMODE, METRIC, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    }
    dim3 grid(p.tiles_x, p.tiles_y);
    k_layer2<ZMODE, METRIC, SHAPE><<<grid, VSB_THREADS, smem2, s>>>(p, taps);
    int rc = vsb_check_launch("k_layer2");
    if (rc) return rc;
    if (ZL_SEARCHES(ZMODE)) {
        k_layer2_dense<ZMODE, METRIC, SHAPE><<<p.ntiles < 3 * n_sms ? p.ntiles : 3 * n_sms, VSB_THREADS, smem2d, s>>>(pd, taps);
        if ((rc = vsb_check_launch("k_layer2_dense"))) return rc;
    }
    return VSB_OK;
}

// binary fast path: k_layer2_classify, k_layer2 over all tiles (it lists the tiles it does not take), k_layer2_dense over the
// tiles set aside, then k_layer over the general list
template <int ZMODE, int METRIC>
static int launch_layer2(LayerParams &p, LayerParams &pd, const BilTapsL *taps, size_t smem2, size_t smem2d, size_t smem, int *gen_list,
                         cudaStream_t s)
{
    static int n_sms = 0;
    if (n_sms == 0) {
        int dev = 0, sms = 0;
        VSB_CUDA(cudaGetDevice(&dev));
        VSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        n_sms = sms > 0 ? sms : 148;
    }
    // two lists in the scratch: the general kernel's, and behind it the dense launch's; then the tile classes
    p.gen_list = pd.gen_list = gen_list;
    p.dense_list = pd.dense_list = gen_list + 1 + p.ntiles;
    VSB_CUDA(cudaMemsetAsync(p.gen_list, 0, (size_t)(2 + p.ntiles) * sizeof(int), s)); // both counters with one node (93 KB at 16K)
    p.tile_class = pd.tile_class = (uint16_t *)(gen_list + 2 * (1 + p.ntiles));
    p.route_dense = pd.route_dense = ZL_SEARCHES(ZMODE) ? 1 : 0;
    { static int df = -1; if (df < 0) { const char *e = getenv("VSB_DENSE_FROM"); df = e ? atoi(e) : L2_DENSE_MIN; } p.dense_from = pd.dense_from = df; }
    k_layer2_classify<<<(p.ntiles + 255) / 256, 256, 0, s>>>(p);
    int rc = vsb_check_launch("k_layer2_classify");
    if (rc) return rc;
    // the window sizes of the two XY chains BASELINE names are compile-time constants in their own instantiations
    // (Preset-Double-LUT: bilateral d = 7 + Gaussian 3 x 3; config 3: bilateral d = 7 alone); any other chain takes SHAPE 0
    const bool bil7 = p.hb == 3 && p.ntaps == 29;
    if (bil7 && p.nkx == 3 && p.nky == 3 && p.halo == 4) rc = launch_layer2_s<ZMODE, METRIC, 1>(p, pd, taps, smem2, smem2d, n_sms, s);
    else if (bil7 && p.nkx == 0 && p.nky == 0 && p.halo == 3) rc = launch_layer2_s<ZMODE, METRIC, 2>(p, pd, taps, smem2, smem2d, n_sms, s);
    else rc = launch_layer2_s<ZMODE, METRIC, 0>(p, pd, taps, smem2, smem2d, n_sms, s);
    if (rc) return rc;
    p.gen_grid = p.ntiles < 2 * n_sms ? p.ntiles : 2 * n_sms;
    return launch_layer_v<ZMODE, METRIC, true>(p, taps, smem, s);
}

template <int ZMODE, int METRIC>
static int launch_layer(LayerParams &p, const BilTapsL *taps, size_t smem, cudaStream_t s)
{
    const bool xy = p.hb > 0 || p.nkx > 0;
    return xy ? launch_layer_v<ZMODE, METRIC, true>(p, taps, smem, s) : launch_layer_v<ZMODE, METRIC, false>(p, taps, smem, s);
}

// Device copies of bilateral tap tables, keyed by content (a handful of distinct filters per process).
#include <algorithm>
#include <mutex>
#include <vector>
struct TapCacheEntry { unsigned long long hash; int dev; BilTapsL *d; };
static std::vector<TapCacheEntry> g_tap_cache;
static std::mutex g_tap_mutex;
static int taps_on_device(const BilTapsL &h, const BilTapsL **out)
{
    unsigned long long hash = 1469598103934665603ULL;
    const unsigned char *b = (const unsigned char *)&h;
    for (size_t i = 0; i < sizeof h; ++i) { hash ^= b[i]; hash *= 1099511628211ULL; }
    int dev = 0;
    VSB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_tap_mutex);
    for (auto &e : g_tap_cache) if (e.hash == hash && e.dev == dev) { *out = e.d; return VSB_OK; }
    BilTapsL *d = nullptr;
    VSB_CUDA(cudaMalloc((void **)&d, sizeof h));
    VSB_CUDA(cudaMemcpy(d, &h, sizeof h, cudaMemcpyHostToDevice));
    g_tap_cache.push_back({hash, dev, d});
    *out = d;
    return VSB_OK;
}

static unsigned long long *g_layer_stats = nullptr;
// debugging aid: device counters [uniform tiles, worked tiles, tiles with receding pixels, receding pixels,
// bilateral pixels, Gaussian words]; enabled by vsb_layer_stats(1), read-and-cleared by vsb_layer_stats(2)
static int g_l2_arena = 0, g_l2_rcap = 0; // vsb_layer_fast_path_limits (0 = defaults)
extern "C" int vsb_layer_fast_path_limits(int arena_entries, int receding_entries)
{
    VSB_REQUIRE(arena_entries == 0 || (arena_entries >= 1024 + 256 && arena_entries <= 7424), "vsb_layer_fast_path_limits: arena of %d entries", arena_entries);
    VSB_REQUIRE(receding_entries >= 0 && receding_entries <= L2_DENSE_MIN, "vsb_layer_fast_path_limits: list of %d entries", receding_entries);
    g_l2_arena = arena_entries;
    g_l2_rcap = receding_entries;
    return VSB_OK;
}

extern "C" int vsb_layer_stats(int op, unsigned long long *out8)
{
    if (op == 1) {
        if (!g_layer_stats) { VSB_CUDA(cudaMalloc((void **)&g_layer_stats, 128)); }
        VSB_CUDA(cudaMemset(g_layer_stats, 0, 128));
    } else if (op == 2 && g_layer_stats && out8) {
        VSB_CUDA(cudaDeviceSynchronize());
        VSB_CUDA(cudaMemcpy(out8, g_layer_stats, 128, cudaMemcpyDeviceToHost));
        VSB_CUDA(cudaMemset(g_layer_stats, 0, 128));
    } else if (op == 0 && g_layer_stats) { cudaFree(g_layer_stats); g_layer_stats = nullptr; }
    return VSB_OK;
}

struct EnhInputs { const float *dist; size_t dpitch; const int32_t *labels; const uint32_t *cmax; };

static int layer_launch(const uint8_t *in, size_t in_pitch, int W, int H, const uint32_t *cur_bits,
                        const uint32_t *const *prior_bits, const uint16_t *cur_flags,
                        const uint16_t *const *prior_flags, int wpr, const vsb_zparams *zp, const float *T_dev,
                        const uint8_t *lut1_dev, const vsb_xy_fused *xy, const uint8_t *lut2_dev, uint8_t *out,
                        size_t out_pitch, vsb_stream_t stream, int sparse, const int *rec_list = nullptr,
                        const EnhInputs *enh = nullptr, int *general_tiles = nullptr)
{
    VSB_REQUIRE(in && out && zp && in != out, "vsb_layer_fused: null pointer / in-place");
    VSB_REQUIRE(W >= 8 && H >= 8 && in_pitch >= (size_t)W && out_pitch >= (size_t)W,
                "vsb_layer_fused: images smaller than 8x8 take the unfused kernels");
    const int mode = zp->mode;
    VSB_REQUIRE(mode == VSB_Z_FIXED || mode == VSB_Z_WEIGHTED || mode == VSB_Z_CONST || mode == VSB_Z_NONE ||
                (mode == VSB_Z_ENHANCED && enh && enh->dist && enh->labels && enh->cmax && enh->dpitch >= (size_t)W),
                "vsb_layer_fused: mode %d not fused", mode);
    VSB_REQUIRE(zp->metric == VSB_METRIC_CHAMFER5 || zp->metric == VSB_METRIC_EXACT, "vsb_layer_fused: unknown metric");
    if (mode != VSB_Z_NONE) {
        VSB_REQUIRE(cur_bits && wpr >= ((W + 127) / 128) * 4, "vsb_layer_fused: mask missing or wpr not a multiple of 4 per 128 px");
        VSB_REQUIRE(zp->n_priors >= 0 && zp->n_priors <= VSB_MAX_PRIORS && (zp->n_priors == 0 || prior_bits),
                    "vsb_layer_fused: priors");
        VSB_REQUIRE(zp->reach >= 0 && zp->reach <= VSB_FAST_REACH, "vsb_layer_fused: reach %d > %d", zp->reach, VSB_FAST_REACH);
        VSB_REQUIRE(T_dev || mode == VSB_Z_CONST || mode == VSB_Z_ENHANCED || zp->metric == VSB_METRIC_EXACT,
                    "vsb_layer_fused: chamfer table missing");
    }
    LayerParams p;
    memset(&p, 0, sizeof p);
    BilTapsL taps;
    memset(&taps, 0, sizeof taps);
    p.in = in; p.ipitch = in_pitch; p.out = out; p.opitch = out_pitch;
    p.cur = cur_bits; p.cur_flags = cur_flags;
    p.z = *zp;
    if (mode == VSB_Z_NONE) p.z.n_priors = 0;
    if (mode == VSB_Z_CONST || mode == VSB_Z_ENHANCED) p.z.reach = 0;
    if (mode == VSB_Z_ENHANCED) { p.enh_dist = enh->dist; p.enh_dpitch = enh->dpitch; p.enh_labels = enh->labels; p.enh_cmax = enh->cmax; }
    for (int i = 0; i < p.z.n_priors; ++i) {
        p.prior[i] = prior_bits[i];
        p.prior_flags[i] = prior_flags ? prior_flags[i] : nullptr;
    }
    p.T = T_dev; p.lut1 = lut1_dev; p.lut2 = lut2_dev;
    p.tiles_x = (W + VSB_TW - 1) / VSB_TW; p.tiles_y = (H + VSB_TH - 1) / VSB_TH;
    p.W = W; p.H = H; p.wpr = wpr;
    p.aligned = (vsb_aligned16(in, in_pitch) && vsb_aligned16(out, out_pitch)) ? 1 : 0;
    if (xy) {
        if (xy->has_bilateral) {
            VSB_REQUIRE(xy->radius >= 1 && xy->radius <= 4 && xy->ntaps >= 1 && xy->ntaps <= 96,
                        "vsb_layer_fused: bilateral radius %d / %d taps not fused", xy->radius, xy->ntaps);
            p.hb = xy->radius; p.ntaps = xy->ntaps;
            for (int i = 0; i < xy->ntaps; ++i) { taps.off[i] = p.tap_off[i] = xy->tap_dy[i] * RG_W + xy->tap_dx[i]; taps.sw[i] = p.tap_sw[i] = xy->tap_sw[i]; }
            for (int i = 0; i < 256; ++i) taps.cw[i] = xy->cw[i];
        }
        if (xy->nkx > 0 || xy->nky > 0) {
            VSB_REQUIRE(xy->nkx >= 1 && xy->nkx <= 7 && (xy->nkx & 1) && xy->nky >= 1 && xy->nky <= 7 && (xy->nky & 1),
                        "vsb_layer_fused: Gaussian %dx%d not fused", xy->nkx, xy->nky);
            int sx = 0, sy = 0;
            for (int i = 0; i < xy->nkx; ++i) { p.kx[i] = xy->kx[i]; sx += xy->kx[i]; }
            for (int i = 0; i < xy->nky; ++i) { p.ky[i] = xy->ky[i]; sy += xy->ky[i]; }
            VSB_REQUIRE(sx == 256 && sy == 256, "vsb_layer_fused: Q8.8 taps must sum to 256");
            p.nkx = xy->nkx; p.nky = xy->nky;
        }
    }
    p.halo = p.hb + max(p.nkx >> 1, p.nky >> 1);
    p.stats = g_layer_stats;
    VSB_REQUIRE(p.halo <= LH_MAX, "vsb_layer_fused: halo %d > %d", p.halo, LH_MAX);
    p.ntiles = p.tiles_x * p.tiles_y;
    p.sparse = sparse;
    p.rec_list = sparse ? rec_list : nullptr;
    if (p.rec_list) { // CTAs of the list-walking patch launch
        static const int div_env = getenv("VSB_PATCH_GRID_DIV") ? atoi(getenv("VSB_PATCH_GRID_DIV")) : 0;
        const int dv = div_env > 0 ? div_env : 4;
        p.gen_grid = (p.ntiles + dv - 1) / dv;
    }
    const BilTapsL *taps_dev = nullptr;
    if (p.hb > 0) { const int rc = taps_on_device(taps, &taps_dev); if (rc) return rc; }
    const int RH = VSB_TH + 2 * p.halo;
    const int n1 = p.z.reach + 1;
    auto up16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const bool searching = mode != VSB_Z_NONE && mode != VSB_Z_CONST && mode != VSB_Z_ENHANCED;
    size_t off = 0;
    p.sm.T = (int)off; off += up16((searching && zp->metric == VSB_METRIC_CHAMFER5) ? (size_t)n1 * n1 * 4 : 16);
    // (the patch launch stages no pixels: it only needs the table, the window and the receding-pixel list -- the smaller its
    //  footprint, the more of its CTAs sit beside the pack pass of the next layer)
    p.sm.s1 = (int)off; off += sparse ? 16 : up16((size_t)RH * RG_W);
    p.sm.bits = (int)off; off += searching ? up16((size_t)(RH + 2 * p.z.reach) * LBW * 4) : 16;
    const size_t s2bytes = sparse ? 16 : up16((size_t)RH * RG_W);
    p.sm.s2 = (int)off; off += s2bytes;
    const size_t rlist = (mode != VSB_Z_NONE) ? (size_t)RH * (VSB_TW + 2 * p.halo) * 2 : 0; // overlays [s2 | list]
    const size_t lmin = sparse ? 16 : (size_t)(GL_OFF + 1100) * 2;
    p.sm.list = (int)off; off += up16(rlist > s2bytes + lmin ? rlist - s2bytes : lmin);
    const size_t smem = off;
    cudaStream_t s = (cudaStream_t)stream;
    const bool ex = zp->metric == VSB_METRIC_EXACT;
    // strictly binary regions (certified per tile by the pack pass) take k_layer2, which works from the masks alone
    static const bool no_layer2 = getenv("VSB_NO_LAYER2") != nullptr;
    const bool has_xy = p.hb > 0 || p.nkx > 0;
    if (!no_layer2 && has_xy && !sparse && general_tiles && cur_flags && cur_bits &&
        (mode == VSB_Z_FIXED || mode == VSB_Z_WEIGHTED || mode == VSB_Z_CONST || mode == VSB_Z_ENHANCED)) {
        p.two_valued = (xy && p.hb >= 1 && xy->two_valued_min_same >= 1 && xy->two_valued_min_same <= 3) ? 1 : 0;
        LayerParams p2 = p; // same carve for T / s1 / s2 / bits, then the three lists
        size_t o2 = 0;
        p2.sm.T = (int)o2; o2 += up16((searching && zp->metric == VSB_METRIC_CHAMFER5) ? (size_t)n1 * (n1 + 1) * 4 : 16); // (dense launch: odd row stride)
        p2.sm.s1 = (int)o2; o2 += up16((size_t)RH * RG_W);
        p2.sm.s2 = (int)o2; o2 += up16((size_t)RH * RG_W);
        p2.sm.bits = (int)o2; o2 += searching ? up16((size_t)(RH + 2 * p.z.reach) * LBW * 4) : 16;
        // the list arena (receding pixels | bilateral pixels ... | 1024 Gaussian words).  First launch: 7424 entries -- tiles
        // with more than 2560 receding pixels are set aside or handed to the general kernel, which leaves room for at least
        // 3840 bilateral pixels (all 34 x 130 of them once fewer than 1980 recede) -- four CTAs fit an SM at F = 40.  The dense
        // launch gets the largest case (46 x 142 receding, 34 x 130 bilateral), which is also the space of its table.
        p2.sm.list = (int)o2;
        LayerParams p2d = p2;
        p2.sm.glist = (g_l2_arena ? g_l2_arena : 7424) - 1024; p2.rcap = g_l2_rcap ? g_l2_rcap : L2_DENSE_MIN;
        if (p2.rcap > p2.sm.glist) p2.rcap = p2.sm.glist; // (modes without search set nothing aside: their fuller tiles take the general kernel)
        p2d.sm.glist = 6656 + 4480; p2d.rcap = 6656;
        const size_t smem2 = o2 + (size_t)(p2.sm.glist + 1024) * 2, smem2d = o2 + (size_t)(p2d.sm.glist + 1024) * 2;
        int rc2;
        switch (mode) {
        case VSB_Z_FIXED: rc2 = ex ? launch_layer2<VSB_Z_FIXED, 1>(p2, p2d, taps_dev, smem2, smem2d, smem, general_tiles, s) : launch_layer2<VSB_Z_FIXED, 0>(p2, p2d, taps_dev, smem2, smem2d, smem, general_tiles, s); break;
        case VSB_Z_WEIGHTED: rc2 = ex ? launch_layer2<VSB_Z_WEIGHTED, 1>(p2, p2d, taps_dev, smem2, smem2d, smem, general_tiles, s) : launch_layer2<VSB_Z_WEIGHTED, 0>(p2, p2d, taps_dev, smem2, smem2d, smem, general_tiles, s); break;
        case VSB_Z_CONST: rc2 = launch_layer2<VSB_Z_CONST, 0>(p2, p2d, taps_dev, smem2, smem2d, smem, general_tiles, s); break;
        default: rc2 = launch_layer2<VSB_Z_ENHANCED, 0>(p2, p2d, taps_dev, smem2, smem2d, smem, general_tiles, s); break;
        }
        return rc2;
    }
    switch (mode) {
    case VSB_Z_FIXED: return ex ? launch_layer<VSB_Z_FIXED, 1>(p, taps_dev, smem, s) : launch_layer<VSB_Z_FIXED, 0>(p, taps_dev, smem, s);
    case VSB_Z_WEIGHTED: return ex ? launch_layer<VSB_Z_WEIGHTED, 1>(p, taps_dev, smem, s) : launch_layer<VSB_Z_WEIGHTED, 0>(p, taps_dev, smem, s);
    case VSB_Z_CONST: return launch_layer<VSB_Z_CONST, 0>(p, taps_dev, smem, s);
    case VSB_Z_ENHANCED: return launch_layer<VSB_Z_ENHANCED, 0>(p, taps_dev, smem, s);
    default: return launch_layer<ZL_NONE, 0>(p, taps_dev, smem, s);
    }
}

extern "C" int vsb_bilateral_two_valued_min_same(const float *tap_sw, int ntaps, const float *cw256, int v0, int v1)
{
    if (!tap_sw || !cw256 || ntaps < 1 || ntaps > 1024 || v0 < 0 || v0 > 255 || v1 < 0 || v1 > 255) return 0;
    if (v0 == v1) return 1;
    std::vector<double> sw(tap_sw, tap_sw + ntaps);
    std::sort(sw.begin(), sw.end());
    const double d = (double)abs(v1 - v0), k = (double)cw256[abs(v1 - v0)], k0 = (double)cw256[0];
    for (int m = 1; m <= ntaps; ++m) {
        // worst case: the m taps that equal the centre carry the smallest space weights, the others the largest
        double ws = 0.0, wo = 0.0;
        for (int i = 0; i < m; ++i) ws += sw[i] * k0;
        for (int i = m; i < ntaps; ++i) wo += sw[i] * k;
        if (ws > 0.0 && d * wo / (ws + wo) < 0.49) return m;
    }
    return 0;
}

extern "C" int vsb_layer_fused(const uint8_t *in, size_t in_pitch, int W, int H, const uint32_t *cur_bits,
                               const uint32_t *const *prior_bits, const uint16_t *cur_flags,
                               const uint16_t *const *prior_flags, int wpr, const vsb_zparams *zp, const float *T_dev,
                               const uint8_t *lut1_dev, const vsb_xy_fused *xy, const uint8_t *lut2_dev, uint8_t *out,
                               size_t out_pitch, int *general_tiles, vsb_stream_t stream)
{
    return layer_launch(in, in_pitch, W, H, cur_bits, prior_bits, cur_flags, prior_flags, wpr, zp, T_dev, lut1_dev, xy, lut2_dev,
                        out, out_pitch, stream, 0, nullptr, nullptr, general_tiles);
}

extern "C" int vsb_layer_patch(const uint8_t *in, size_t in_pitch, int W, int H, const uint32_t *cur_bits,
                               const uint32_t *const *prior_bits, int wpr, const vsb_zparams *zp, const float *T_dev,
                               const uint8_t *lut1_dev, const int *rec_tiles, uint8_t *out, size_t out_pitch,
                               vsb_stream_t stream)
{
    VSB_REQUIRE(zp && (zp->mode == VSB_Z_FIXED || zp->mode == VSB_Z_WEIGHTED || zp->mode == VSB_Z_CONST),
                "vsb_layer_patch: fixed, weighted or constant fades only");
    if (zp->n_priors == 0) return VSB_OK; // nothing can recede: the pre-filled output stands
    return layer_launch(in, in_pitch, W, H, cur_bits, prior_bits, nullptr, nullptr, wpr, zp, T_dev, lut1_dev, nullptr, nullptr,
                        out, out_pitch, stream, 1, rec_tiles);
}

extern "C" int vsb_layer_fused_enhanced(const uint8_t *in, size_t in_pitch, int W, int H, const uint32_t *cur_bits,
                                        const uint32_t *const *prior_bits, int wpr, const vsb_zparams *zp, const float *dist,
                                        size_t dist_pitch_elems, const int32_t *labels, const uint32_t *compmax,
                                        const uint8_t *lut1_dev, const vsb_xy_fused *xy, const uint8_t *lut2_dev, uint8_t *out,
                                        size_t out_pitch, const uint16_t *cur_flags, int *general_tiles, vsb_stream_t stream)
{
    VSB_REQUIRE(zp, "vsb_layer_fused_enhanced: null parameters");
    vsb_zparams z = *zp;
    z.mode = VSB_Z_ENHANCED;
    const EnhInputs enh = {dist, dist_pitch_elems, labels, compmax};
    return layer_launch(in, in_pitch, W, H, cur_bits, prior_bits, cur_flags, nullptr, wpr, &z, nullptr, lut1_dev, xy, lut2_dev, out,
                        out_pitch, stream, 0, nullptr, &enh, general_tiles);
}
