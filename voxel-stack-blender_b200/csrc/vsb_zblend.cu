// vsb_zblend.cu -- K2+K3 fused: receding mask, windowed distance, fade gradient, merge, LUT.
// One CTA per 128x32-pixel tile; the float distance map never exists in HBM.
//
// Reference arithmetic replaced (processing_core.py): rec = prior & ~cur (:119), d = distanceTransform(
// ~cur, L2, 5) (:129), masked (:132), g = trunc(255*(1 - min(d,F)/den)) (:138-149), masked (:152);
// weighted accumulation (:252-282); merge = max(gray, g) (:458); lut[img] (lut_manager.py:63).
//
// Distance definition (bit-exact with oracle/c/vsb_oracle.c orc_chamfer5): D(p) = min over white
// pixels q of the current layer of T[max|p-q|][min|p-q|].  Only min(D, F) is consumed, so a tile
// needs the current mask on a halo of R = ceil(F)-1 pixels, and only rec pixels are evaluated:
//   per rec pixel: rows dy = 0, -+1, -+2 ... while |dy| < best; in each row the nearest white pixel
//   is found with two 64-bit funnel-shift/clz/ffs probes on the bit-packed row (T is monotone in
//   |dx| for fixed dy, so the nearest one is the row's minimum).
// The `exact` metric runs the same search on integer dy^2+dx^2.
// Tiles without rec pixels (the large majority) are a pure gray -> LUT stream.
#include "vsb_tile.cuh"

struct ZKParams {
    const uint8_t *gray;
    size_t gpitch;
    uint8_t *out;
    size_t opitch;
    const uint32_t *cur;
    const uint32_t *prior[VSB_MAX_PRIORS];
    const float *T;
    const uint8_t *lut;
    float *dist;
    size_t dpitch;
    uint32_t *rec_out; // VSB_Z_DIST: optional packed receding mask (what vsb_maskdiff would write)
    int W, H, wpr;
    int aligned;
    vsb_zparams z;
};

template <int MODE, int METRIC>
__global__ void __launch_bounds__(VSB_THREADS) k_zblend(const __grid_constant__ ZKParams p)
{
    extern __shared__ float s_T[];                          // (R+1)^2 chamfer table
    __shared__ uint32_t s_cur[MAX_ROWS][ROW_WORDS];         // current mask, tile + halo
    __shared__ uint32_t s_rowany[(MAX_ROWS + 31) / 32];     // staged rows that hold any white pixel
    __shared__ uint32_t s_recp[MODE == VSB_Z_WEIGHTED ? VSB_MAX_PRIORS : 1][VSB_TH][4];
    __shared__ __align__(16) uint8_t s_g[VSB_TH][VSB_TW];   // gradient of the tile
    __shared__ uint16_t s_list[VSB_TH * VSB_TW];            // compacted rec pixel list
    __shared__ int s_wsum[8];
    __shared__ __align__(4) uint8_t s_lut[256];

    const int t = threadIdx.x;
    const int x0 = blockIdx.x * VSB_TW, y0 = blockIdx.y * VSB_TH;
    const int W = p.W, H = p.H, wpr = p.wpr;
    const int R = p.z.reach, np = p.z.n_priors;

    // ---- phase 0: gray chunk of this thread, rec words of the tile -----------------------------
    const int r = t >> 3, c = t & 7;
    const int y = y0 + r, x = x0 + c * 16;
    const int nvalid = (y < H) ? min(16, W - x) : 0;
    uint4 gv = make_uint4(0, 0, 0, 0);
    if (MODE != VSB_Z_DIST) {
        if (nvalid == 16 && p.aligned) gv = vsb_ld16_stream(p.gray + (size_t)y * p.gpitch + x);
        else if (nvalid > 0) gv = vsb_ld_bytes(p.gray + (size_t)y * p.gpitch + x, nvalid);
    }
    VsbLut L = vsb_lut_stage(MODE == VSB_Z_DIST ? nullptr : p.lut, s_lut);

    uint32_t recw = 0;
    if (t < 128) {
        const int rr = t >> 2, wi = t & 3;
        const int yy = y0 + rr, wx = (x0 >> 5) + wi;
        if (yy < H && wx < wpr) {
            const size_t o = (size_t)yy * wpr + wx;
            const uint32_t ncur = ~p.cur[o];
            uint32_t acc = 0;
#pragma unroll 4
            for (int i = 0; i < np; ++i) {
                const uint32_t pw = p.prior[i][o] & ncur;
                if (MODE == VSB_Z_WEIGHTED) s_recp[i][rr][wi] = pw;
                acc |= pw;
            }
            recw = acc;
            if (MODE == VSB_Z_DIST && p.rec_out) p.rec_out[o] = acc;
        } else if (MODE == VSB_Z_WEIGHTED) {
            for (int i = 0; i < np; ++i) s_recp[i][rr][wi] = 0;
        }
    }
    const int any = __syncthreads_or(recw != 0);

    if (!any) { // ---- inactive tile: gray -> LUT stream ------------------------------------------
        if (MODE != VSB_Z_DIST && nvalid > 0) {
            uint4 o = L.map16(gv);
            uint8_t *dst = p.out + (size_t)y * p.opitch + x;
            if (nvalid == 16 && p.aligned) vsb_st16_stream(dst, o);
            else vsb_st_bytes(dst, o, nvalid);
        }
        return;
    }

    // ---- phase 1: stage table + current mask with halo, zero the gradient tile, compact rec ------
    const int nrows = VSB_TH + 2 * R;
    if (MODE != VSB_Z_CONST) {
        if (METRIC == VSB_METRIC_CHAMFER5) {
            const int n1 = R + 1;
            for (int i = t; i < n1 * n1; i += VSB_THREADS) s_T[i] = p.T[i];
        }
        stage_mask_tile(s_cur, p.cur, wpr, H, x0, y0, R);
    }
    if (MODE != VSB_Z_DIST) ((uint4 *)&s_g[0][0])[t] = make_uint4(0, 0, 0, 0);

    const int lane = t & 31, warp = t >> 5;
    const int cnt = (t < 128) ? __popc(recw) : 0;
    int incl = cnt;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    if (MODE != VSB_Z_CONST) stage_rowany(s_rowany, s_cur, nrows);
    int base = incl - cnt;
    for (int w = 0; w < warp && w < 4; ++w) base += s_wsum[w];
    const int total = s_wsum[0] + s_wsum[1] + s_wsum[2] + s_wsum[3];
    if (t < 128) {
        uint32_t m = recw;
        const int rr = t >> 2, wi = t & 3;
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            s_list[base++] = (uint16_t)((rr << 7) | (wi << 5) | b);
        }
    }
    __syncthreads();

    // ---- phase 2: windowed distance search per rec pixel, gradient ----------------------------------
    const float Ff = p.z.fade;
    for (int i = t; i < total; i += VSB_THREADS) {
        const int code = s_list[i];
        const int pr = code >> 7, pc = code & 127;
        float best = 0.0f;
        if (MODE != VSB_Z_CONST)
            best = tile_search<METRIC, ROW_WORDS>(s_cur, s_rowany, s_T, R, pr + R, pc + 32 * HALO_WORDS);
        if (MODE == VSB_Z_FIXED) {
            const float n = __fdiv_rn(fminf(best, Ff), p.z.den);
            const float v = __fmul_rn(255.0f, __fsub_rn(1.0f, n));
            s_g[pr][pc] = (uint8_t)(int)v;
        } else if (MODE == VSB_Z_WEIGHTED) {
            float acc = 0.0f;
            const int wi = pc >> 5, b = pc & 31;
            for (int j = 0; j < np; ++j) {
                if (p.z.weights[j] > 0.0f && ((s_recp[j][pr][wi] >> b) & 1u)) {
                    const float n = __fsub_rn(1.0f, __fdiv_rn(fminf(best, p.z.fades[j]), p.z.dens[j]));
                    acc = __fadd_rn(acc, __fmul_rn(n, p.z.weights[j]));
                }
            }
            const float v = __fmul_rn(255.0f, __fdiv_rn(acc, p.z.total_weight));
            s_g[pr][pc] = (uint8_t)(int)v;
        } else if (MODE == VSB_Z_CONST) {
            s_g[pr][pc] = (uint8_t)p.z.const_val;
        } else { // VSB_Z_DIST
            p.dist[(size_t)(y0 + pr) * p.dpitch + x0 + pc] = fminf(best, Ff);
        }
    }
    if (MODE == VSB_Z_DIST) return;
    __syncthreads();

    // ---- phase 3: merge with gray, LUT, store -------------------------------------------------------
    if (nvalid > 0) {
        const uint4 g = ((const uint4 *)&s_g[0][0])[t];
        uint4 m = make_uint4(__vmaxu4(gv.x, g.x), __vmaxu4(gv.y, g.y), __vmaxu4(gv.z, g.z), __vmaxu4(gv.w, g.w));
        m = L.map16(m);
        uint8_t *dst = p.out + (size_t)y * p.opitch + x;
        if (nvalid == 16 && p.aligned) vsb_st16_stream(dst, m);
        else vsb_st_bytes(dst, m, nvalid);
    }
}

template <int METRIC>
static void launch_mode(int mode, dim3 grid, size_t smem, cudaStream_t s, const ZKParams &k)
{
    switch (mode) {
    case VSB_Z_FIXED: k_zblend<VSB_Z_FIXED, METRIC><<<grid, VSB_THREADS, smem, s>>>(k); break;
    case VSB_Z_WEIGHTED: k_zblend<VSB_Z_WEIGHTED, METRIC><<<grid, VSB_THREADS, smem, s>>>(k); break;
    case VSB_Z_CONST: k_zblend<VSB_Z_CONST, METRIC><<<grid, VSB_THREADS, smem, s>>>(k); break;
    default: k_zblend<VSB_Z_DIST, METRIC><<<grid, VSB_THREADS, smem, s>>>(k); break;
    }
}

static int zblend_launch(const uint8_t *gray, size_t gray_pitch, int W, int H, const uint32_t *cur_bits,
                         const uint32_t *const *prior_bits, int wpr, const vsb_zparams *zp,
                         const float *T_dev, const uint8_t *lut256_dev, uint8_t *out, size_t out_pitch,
                         float *dist_out, size_t dist_pitch_elems, uint32_t *rec_out, vsb_stream_t stream)
{
    VSB_REQUIRE(zp && cur_bits, "vsb_zblend_fused: null pointer");
    VSB_REQUIRE(W > 0 && H > 0 && wpr >= (W + 31) / 32, "vsb_zblend_fused: bad geometry");
    VSB_REQUIRE(zp->mode >= VSB_Z_FIXED && zp->mode <= VSB_Z_DIST, "vsb_zblend_fused: unknown mode %d", zp->mode);
    VSB_REQUIRE(zp->metric == VSB_METRIC_CHAMFER5 || zp->metric == VSB_METRIC_EXACT,
                "vsb_zblend_fused: unknown metric %d", zp->metric);
    VSB_REQUIRE(zp->n_priors >= 0 && zp->n_priors <= VSB_MAX_PRIORS, "vsb_zblend_fused: n_priors %d > %d",
                zp->n_priors, VSB_MAX_PRIORS);
    VSB_REQUIRE(zp->reach >= 0 && zp->reach <= VSB_FAST_REACH,
                "vsb_zblend_fused: reach %d exceeds the tile path (%d)", zp->reach, VSB_FAST_REACH);
    VSB_REQUIRE(T_dev || zp->mode == VSB_Z_CONST || zp->metric == VSB_METRIC_EXACT,
                "vsb_zblend_fused: chamfer table missing");
    VSB_REQUIRE(zp->n_priors == 0 || prior_bits, "vsb_zblend_fused: prior pointers missing");
    if (zp->mode == VSB_Z_DIST) {
        VSB_REQUIRE(dist_out && dist_pitch_elems >= (size_t)W, "vsb_zblend_fused: dist_out missing");
    } else {
        VSB_REQUIRE(gray && out && gray_pitch >= (size_t)W && out_pitch >= (size_t)W,
                    "vsb_zblend_fused: image pointers/pitch");
    }
    ZKParams k;
    k.gray = gray; k.gpitch = gray_pitch; k.out = out; k.opitch = out_pitch;
    k.cur = cur_bits;
    for (int i = 0; i < VSB_MAX_PRIORS; ++i) k.prior[i] = i < zp->n_priors ? prior_bits[i] : nullptr;
    k.T = T_dev; k.lut = lut256_dev; k.dist = dist_out; k.dpitch = dist_pitch_elems; k.rec_out = rec_out;
    k.W = W; k.H = H; k.wpr = wpr;
    k.aligned = (zp->mode == VSB_Z_DIST) ? 1 : (vsb_aligned16(gray, gray_pitch) && vsb_aligned16(out, out_pitch));
    k.z = *zp;
    if (zp->mode == VSB_Z_CONST) k.z.reach = 0;
    dim3 grid((W + VSB_TW - 1) / VSB_TW, (H + VSB_TH - 1) / VSB_TH);
    const int n1 = k.z.reach + 1;
    const size_t smem = zp->metric == VSB_METRIC_CHAMFER5 ? (size_t)n1 * n1 * sizeof(float) : 16;
    cudaStream_t s = (cudaStream_t)stream;
    if (zp->metric == VSB_METRIC_CHAMFER5) launch_mode<VSB_METRIC_CHAMFER5>(zp->mode, grid, smem, s, k);
    else launch_mode<VSB_METRIC_EXACT>(zp->mode, grid, smem, s, k);
    return vsb_check_launch("k_zblend");
}

extern "C" int vsb_zblend_fused(const uint8_t *gray, size_t gray_pitch, int W, int H, const uint32_t *cur_bits,
                                const uint32_t *const *prior_bits, int wpr, const vsb_zparams *zp,
                                const float *T_dev, const uint8_t *lut256_dev, uint8_t *out, size_t out_pitch,
                                float *dist_out, size_t dist_pitch_elems, vsb_stream_t stream)
{
    return zblend_launch(gray, gray_pitch, W, H, cur_bits, prior_bits, wpr, zp, T_dev, lut256_dev, out, out_pitch, dist_out,
                         dist_pitch_elems, nullptr, stream);
}

extern "C" int vsb_rec_dist(const uint32_t *cur_bits, const uint32_t *const *prior_bits, int wpr, int W, int H,
                            const vsb_zparams *zp, const float *T_dev, uint32_t *rec_bits_out, float *dist_out,
                            size_t dist_pitch_elems, vsb_stream_t stream)
{
    VSB_REQUIRE(zp && rec_bits_out, "vsb_rec_dist: null pointer");
    VSB_REQUIRE(wpr >= ((W + 127) / 128) * 4, "vsb_rec_dist: wpr must be 4 words per 128 pixels (every mask word is written)");
    vsb_zparams z = *zp;
    z.mode = VSB_Z_DIST;
    return zblend_launch(nullptr, 0, W, H, cur_bits, prior_bits, wpr, &z, T_dev, nullptr, nullptr, 0, dist_out, dist_pitch_elems,
                         rec_bits_out, stream);
}
