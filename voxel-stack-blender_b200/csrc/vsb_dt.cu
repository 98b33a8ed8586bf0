// vsb_dt.cu -- K2 / K2': dense distance maps (the EDT-only path of BASELINE config 5 and the
// per-function API), both metrics.
//
//   vsb_dt_chamfer5 / vsb_dt_exact_windowed: dist = min(D, clip), exact for D < R+1 (R <= 63).  Same
//     128x32 tile + bit-mask halo as the fused Z-blend kernel, but every pixel is evaluated, so the
//     row probe is hoisted: h[row][col] = nearest set bit of the staged row (one probe per staged row
//     and column, shared by the 32 tile rows), then D(p) = min_k N(k, h[row+-k][col]).
//     replaces cv2.distanceTransform(~cur, DIST_L2, 5)      processing_core.py:129,250,379
//   (the unbounded exact transform is csrc/vsb_edt.cu)
#include "vsb_tile.cuh"

template <int METRIC>
__global__ void __launch_bounds__(VSB_THREADS) k_dt_dense(const uint32_t *__restrict__ bits, int wpr, int W, int H,
                                                          int R, const float *__restrict__ T, float clip,
                                                          float *__restrict__ dist, size_t dpitch)
{
    extern __shared__ float s_T[];
    __shared__ uint32_t s_cur[MAX_ROWS][ROW_WORDS];
    __shared__ uint8_t s_h[MAX_ROWS][VSB_TW];
    const int t = threadIdx.x;
    const int x0 = blockIdx.x * VSB_TW, y0 = blockIdx.y * VSB_TH;
    const int n1 = R + 1;
    const int nrows = VSB_TH + 2 * R;
    if (METRIC == VSB_METRIC_CHAMFER5)
        for (int i = t; i < n1 * n1; i += VSB_THREADS) s_T[i] = T[i];
    stage_mask_tile(s_cur, bits, wpr, H, x0, y0, R);
    __syncthreads();
    for (int i = t; i < nrows * VSB_TW; i += VSB_THREADS) {
        const int hr = i >> 7, pc = i & 127;
        s_h[hr][pc] = (uint8_t)nearest_dx(s_cur[hr], pc + 32 * HALO_WORDS);
    }
    __syncthreads();
    for (int i = t; i < VSB_TH * VSB_TW; i += VSB_THREADS) {
        const int pr = i >> 7, pc = i & 127;
        const int y = y0 + pr, x = x0 + pc;
        if (y >= H || x >= W) continue;
        const int hr0 = pr + R;
        float d;
        if (METRIC == VSB_METRIC_CHAMFER5) {
            float best = 3.0e38f;
            int dx = s_h[hr0][pc];
            if (dx <= R) best = s_T[dx * n1];
            for (int k = 1; k <= R && (float)k < best; ++k) {
                dx = s_h[hr0 - k][pc];
                if (dx <= R) best = fminf(best, s_T[max(k, dx) * n1 + min(k, dx)]);
                dx = s_h[hr0 + k][pc];
                if (dx <= R) best = fminf(best, s_T[max(k, dx) * n1 + min(k, dx)]);
            }
            d = best;
        } else {
            int best2 = 0x7fffffff;
            int dx = s_h[hr0][pc];
            if (dx <= R) best2 = dx * dx;
            for (int k = 1; k <= R && k * k < best2; ++k) {
                dx = s_h[hr0 - k][pc];
                if (dx <= R) best2 = min(best2, k * k + dx * dx);
                dx = s_h[hr0 + k][pc];
                if (dx <= R) best2 = min(best2, k * k + dx * dx);
            }
            d = best2 == 0x7fffffff ? 3.0e38f : __fsqrt_rn((float)best2);
        }
        dist[(size_t)y * dpitch + x] = fminf(d, clip);
    }
}

static int launch_dense(int metric, const uint32_t *bits, int wpr, int W, int H, int R, const float *T, float clip,
                        float *dist, size_t dpitch, vsb_stream_t stream)
{
    VSB_REQUIRE(bits && dist, "vsb_dt: null pointer");
    VSB_REQUIRE(W > 0 && H > 0 && wpr >= (W + 31) / 32 && dpitch >= (size_t)W, "vsb_dt: bad geometry");
    VSB_REQUIRE(R >= 0 && R <= VSB_FAST_REACH, "vsb_dt: reach %d exceeds the tile path (%d)", R, VSB_FAST_REACH);
    dim3 grid((W + VSB_TW - 1) / VSB_TW, (H + VSB_TH - 1) / VSB_TH);
    cudaStream_t s = (cudaStream_t)stream;
    if (metric == VSB_METRIC_CHAMFER5) {
        VSB_REQUIRE(T, "vsb_dt_chamfer5: table missing");
        const size_t smem = (size_t)(R + 1) * (R + 1) * sizeof(float);
        k_dt_dense<VSB_METRIC_CHAMFER5><<<grid, VSB_THREADS, smem, s>>>(bits, wpr, W, H, R, T, clip, dist, dpitch);
    } else {
        k_dt_dense<VSB_METRIC_EXACT><<<grid, VSB_THREADS, 16, s>>>(bits, wpr, W, H, R, T, clip, dist, dpitch);
    }
    return vsb_check_launch("k_dt_dense");
}

extern "C" int vsb_dt_chamfer5(const uint32_t *seed_bits, int wpr, int W, int H, int R, const float *T_dev,
                               float clip, float *dist, size_t dist_pitch_elems, vsb_stream_t stream)
{
    return launch_dense(VSB_METRIC_CHAMFER5, seed_bits, wpr, W, H, R, T_dev, clip, dist, dist_pitch_elems, stream);
}

extern "C" int vsb_dt_exact_windowed(const uint32_t *seed_bits, int wpr, int W, int H, int R, float clip,
                                     float *dist, size_t dist_pitch_elems, vsb_stream_t stream)
{
    return launch_dense(VSB_METRIC_EXACT, seed_bits, wpr, W, H, R, nullptr, clip, dist, dist_pitch_elems, stream);
}
