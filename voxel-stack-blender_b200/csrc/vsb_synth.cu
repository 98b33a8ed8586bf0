// vsb_synth.cu -- synthetic slice rasteriser (measurement utility).
// Layer z of a table of growing / shrinking disks plus an optional rounded-rectangle raft.  One thread per
// 16 pixels; each CTA first culls the primitive table against its 128x32 tile into shared memory.
#include "vsb_common.cuh"

#define SYNTH_MAX_LOCAL 1024

struct Raft { float x0, y0, x1, y1, cr, z_end; int on; };

__global__ void __launch_bounds__(VSB_THREADS) k_synth(const float *__restrict__ prims, int n_prims, Raft raft, int z,
                                                       int W, int H, uint8_t *__restrict__ out, size_t pitch,
                                                       int aligned)
{
    __shared__ float s_cx[SYNTH_MAX_LOCAL], s_cy[SYNTH_MAX_LOCAL], s_r2[SYNTH_MAX_LOCAL];
    __shared__ int s_n;
    const int x0 = blockIdx.x * VSB_TW, y0 = blockIdx.y * VSB_TH;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n_prims; i += VSB_THREADS) {
        const float *q = prims + 6 * (size_t)i;
        const float zs = q[4], ze = q[5];
        if ((float)z < zs || (float)z >= ze) continue;
        const float r = q[2] + q[3] * ((float)z - zs);
        if (r <= 0.0f) continue;
        const float cx = q[0], cy = q[1];
        if (cx + r < (float)x0 || cx - r > (float)(x0 + VSB_TW) || cy + r < (float)y0 || cy - r > (float)(y0 + VSB_TH))
            continue;
        const int slot = atomicAdd(&s_n, 1);
        if (slot < SYNTH_MAX_LOCAL) { s_cx[slot] = cx; s_cy[slot] = cy; s_r2[slot] = r * r; }
    }
    __syncthreads();
    const int n = min(s_n, SYNTH_MAX_LOCAL);
    const int r = threadIdx.x >> 3, c = threadIdx.x & 7;
    const int y = y0 + r, x = x0 + c * 16;
    if (y >= H || x >= W) return;
    uint32_t w[4] = {0, 0, 0, 0};
    for (int i = 0; i < 16; ++i) {
        const float fx = (float)(x + i), fy = (float)y;
        bool in = false;
        if (raft.on && (float)z < raft.z_end && fx >= raft.x0 && fx <= raft.x1 && fy >= raft.y0 && fy <= raft.y1) {
            const float ex = fmaxf(fmaxf(raft.x0 + raft.cr - fx, fx - (raft.x1 - raft.cr)), 0.0f);
            const float ey = fmaxf(fmaxf(raft.y0 + raft.cr - fy, fy - (raft.y1 - raft.cr)), 0.0f);
            in = ex * ex + ey * ey <= raft.cr * raft.cr;
        }
        for (int k = 0; k < n && !in; ++k) {
            const float dx = fx - s_cx[k], dy = fy - s_cy[k];
            in = dx * dx + dy * dy <= s_r2[k];
        }
        if (in) w[i >> 2] |= 0xFFu << ((i & 3) * 8);
    }
    const int nv = min(16, W - x);
    uint8_t *dst = out + (size_t)y * pitch + x;
    const uint4 v = make_uint4(w[0], w[1], w[2], w[3]);
    if (aligned && nv == 16) vsb_st16_stream(dst, v);
    else vsb_st_bytes(dst, v, nv);
}

extern "C" int vsb_synth_layer(const float *prims_dev, int n_prims, const float *raft, int z, int W, int H,
                               uint8_t *out, size_t pitch, vsb_stream_t stream)
{
    VSB_REQUIRE(out && (n_prims == 0 || prims_dev), "vsb_synth_layer: null pointer");
    VSB_REQUIRE(W > 0 && H > 0 && pitch >= (size_t)W, "vsb_synth_layer: bad geometry");
    Raft rf;
    memset(&rf, 0, sizeof rf);
    if (raft) { rf.x0 = raft[0]; rf.y0 = raft[1]; rf.x1 = raft[2]; rf.y1 = raft[3]; rf.cr = raft[4]; rf.z_end = raft[5]; rf.on = 1; }
    dim3 grid((W + VSB_TW - 1) / VSB_TW, (H + VSB_TH - 1) / VSB_TH);
    k_synth<<<grid, VSB_THREADS, 0, (cudaStream_t)stream>>>(prims_dev, n_prims, rf, z, W, H, out, pitch,
                                                             vsb_aligned16(out, pitch) ? 1 : 0);
    return vsb_check_launch("k_synth");
}
