// vsb_pack_tma.cu -- K1 as a persistent, TMA-fed streaming kernel (the bandwidth-bound pass of every layer).
//
// One WARP owns one 128x32 tile at a time and runs its own pipeline: lane 0 keeps PK_STAGES bulk tensor loads
// (cp.async.bulk.tensor.2d, 4 KB boxes) in flight into the warp's private shared-memory ring and waits on the stage's
// mbarrier; the 32 lanes then read the tile with conflict-free 128-bit shared loads (a quarter warp = one 128-byte row),
// derive the mask words, the tile flag and -- for LUT-only pipelines -- lut[gray], which is written back into the same
// stage and leaves from there through a bulk tensor store.  No block-level barrier anywhere: warps never wait for each other, the tile
// flag is a warp vote.  Out-of-image bytes arrive as zeros (TMA bounds handling) and out-of-image stores are dropped,
// so ragged right / bottom tiles need no byte loops.
//
//   replaces cv2.threshold(img, 127, 255, THRESH_BINARY)                 processing_core.py:44
//   (+ lut[max(gray, 0)] of the pixels that do not recede                 lut_manager.py:63, for LUT-only chains)
//
// Threshold is fixed at 127 here (the only value the pipeline uses): bit = most significant bit of the byte, taken for
// four pixels at once with PRMT's sign-replication mode.
#include <cstring>
#include "vsb_tma.cuh"

#define PK_WARPS 4
#define PK_TILE_BYTES (VSB_TW * VSB_TH) // 4096
#define FLAG_MIXED 0x100
#define FLAG_BINARY 0x200

// ---- host: tensor maps ----------------------------------------------------------------------------------------------
typedef CUresult (*vsb_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static vsb_encode_tiled_fn encode_tiled()
{
    static vsb_encode_tiled_fn fn = []() -> vsb_encode_tiled_fn {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (vsb_encode_tiled_fn)f;
    }();
    return fn;
}

int vsb_make_tmap_2d(CUtensorMap *out, const void *base, int elem_bytes, size_t W, size_t H, size_t pitch_bytes, int box_w, int box_h)
{
    vsb_encode_tiled_fn enc = encode_tiled();
    VSB_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from this driver");
    VSB_REQUIRE(((uintptr_t)base & 15u) == 0 && (pitch_bytes & 15u) == 0, "tensor map: base / pitch not 16-byte aligned");
    const CUtensorMapDataType dt = elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                   : elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
    const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(out, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VSB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return VSB_OK;
}

// ---- device ---------------------------------------------------------------------------------------------------------
struct PackTmaParams {
    uint32_t *bits;
    uint16_t *flags;
    const uint8_t *lut;
    const uint32_t *prior[VSB_MAX_PRIORS];
    const uint16_t *prior_flags[VSB_MAX_PRIORS];
    int *rec_list;
    int np, gate;
    int wpr, tiles_x, ntiles, W, H;
    int stages;
};

// the four bytes of w -> 0x00 / 0xFF by their most significant bit (byte > 127)
// (prmt's generic mode: a selector nibble with its top bit set replicates the sign bit of the selected byte)
__device__ __forceinline__ uint32_t msb_bytes(uint32_t w)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(0u), "r"(0xBA98u));
    return d;
}

#define PK_HELD 4 // prior masks whose words stay un-ORed in registers across one tile (the window of the usual N <= 4)

template <bool LUTOUT>
__global__ void __launch_bounds__(PK_WARPS * 32) k_pack_tma(const __grid_constant__ CUtensorMap tm_in,
                                                            const __grid_constant__ CUtensorMap tm_out,
                                                            const __grid_constant__ PackTmaParams p)
{
    extern __shared__ uint8_t dsm_raw[];
    __shared__ __align__(8) uint64_t s_bar[PK_WARPS][8];
    __shared__ __align__(4) uint8_t s_lut[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int S = p.stages;
    // 128-byte aligned ring, addressed by pointer arithmetic on the shared array (keeps the accesses LDS / STS)
    uint8_t *ring = dsm_raw + ((128u - (vsb_smem_u32(dsm_raw) & 127u)) & 127u) + (size_t)warp * S * PK_TILE_BYTES;
    uint64_t *bar = s_bar[warp];

    if (LUTOUT && p.lut && threadIdx.x < 64) ((uint32_t *)s_lut)[threadIdx.x] = ((const uint32_t *)p.lut)[threadIdx.x];
    if (lane == 0) {
        for (int s = 0; s < S; ++s) vsb_mbar_init(&bar[s], 1);
        vsb_fence_barrier_init();
    }
    __syncthreads(); // the only block barrier: LUT staged, barriers initialised

    const int gw = blockIdx.x * PK_WARPS + warp, nw = gridDim.x * PK_WARPS;
    const int count = gw < p.ntiles ? (p.ntiles - gw + nw - 1) / nw : 0;
    const uint32_t K0 = (LUTOUT && p.lut) ? (uint32_t)s_lut[0] * 0x01010101u : 0u;
    const uint32_t K1 = (LUTOUT && p.lut) ? (uint32_t)s_lut[255] * 0x01010101u : 0xFFFFFFFFu;
    // LUTOUT: lut[gray] is written back into the stage it was read from and leaves from there with a bulk tensor store;
    // the stage is refilled once that store has drained it, i.e. two tiles later -> S - 2 loads in flight per warp

    auto issue = [&](int k, int st) { // lane 0: tile k of this warp -> stage st = k % S
        const int tile = gw + k * nw;
        const int ty = (int)((unsigned)tile / (unsigned)p.tiles_x), tx = tile - ty * p.tiles_x;
        vsb_mbar_expect_tx(&bar[st], PK_TILE_BYTES);
        vsb_tma_load_2d(ring + (size_t)st * PK_TILE_BYTES, &tm_in, tx * VSB_TW, ty * VSB_TH, &bar[st]);
    };
    if (lane == 0)
        for (int k = 0; k < S && k < count; ++k) issue(k, k);

    // lane geometry: chunk c = lane & 7 of rows (lane >> 3) + 4 j, j = 0..7; after the pair exchange a lane owns the
    // mask words (row (lane >> 3) + 4 (2 m + (lane & 1)), word (lane & 7) >> 1), m = 0..3
    const int c = lane & 7, rb = lane >> 3, odd = lane & 1, wq = c >> 1;
    const bool want_rec = LUTOUT && p.rec_list != nullptr && p.np > 0;

    // receding-tile detection needs OR(priors) at this lane's four words.  Those loads are requested one tile ahead and
    // stay un-combined in registers until they are used, and whether a tile needs them at all (some prior is not
    // uniformly black there) is looked up two tiles ahead, so neither latency is on the critical path.
    uint32_t held[PK_HELD][4];
#pragma unroll
    for (int j = 0; j < PK_HELD; ++j)
#pragma unroll
        for (int m = 0; m < 4; ++m) held[j][m] = 0u;
    uint32_t gate_flag_next = 0; // this lane's prior flag of the tile after next
    auto load_gate = [&](int k) -> uint32_t {
        if (!want_rec || !p.gate || k >= count || lane >= p.np) return 0u;
        return p.prior_flags[lane][gw + k * nw];
    };
    auto needs_priors = [&](uint32_t f) -> bool { // warp-uniform: some prior's tile is mixed or not black
        if (!want_rec) return false;
        if (!p.gate) return true;
        const bool mine = lane < p.np && ((f & FLAG_MIXED) || (f & 0xFFu) > 127u);
        return __any_sync(0xffffffffu, mine);
    };
    auto load_priors = [&](int ty, int tx) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int y = ty * VSB_TH + rb + 4 * (2 * m + odd);
            if (y < p.H) {
                const size_t o = (size_t)y * p.wpr + tx * 4 + wq;
#pragma unroll
                for (int j = 0; j < PK_HELD; ++j)
                    if (j < p.np) held[j][m] = p.prior[j][o];
                for (int j = PK_HELD; j < p.np; ++j) held[0][m] |= p.prior[j][o]; // longer windows: combined right away
            }
        }
    };
    // tile coordinates advance by nw tiles per step: (dty, dtx) with a carry instead of a division
    const int dty = nw / p.tiles_x, dtx = nw - dty * p.tiles_x;
    int ty = gw / p.tiles_x, tx = gw - ty * p.tiles_x;
    if (want_rec && count > 0) {
        if (needs_priors(load_gate(0))) load_priors(ty, tx);
        gate_flag_next = load_gate(1);
    }

    int st = 0;          // k % S and (k / S) & 1, kept incrementally
    uint32_t phase = 0u;
    for (int k = 0; k < count; ++k) {
        const int tile = gw + k * nw;
        const int x0 = tx * VSB_TW, y0 = ty * VSB_TH;
        int ty1 = ty + dty, tx1 = tx + dtx; // the warp's next tile
        if (tx1 >= p.tiles_x) { tx1 -= p.tiles_x; ++ty1; }
        if (LUTOUT && lane == 0 && k >= 2 && k - 2 + S < count) { // the store of tile k-2 has drained its stage: refill it
            vsb_tma_wait_read<1>();
            issue(k - 2 + S, st >= 2 ? st - 2 : st - 2 + S);
        }
        vsb_mbar_wait(&bar[st], phase);
        uint8_t *src = ring + (size_t)st * PK_TILE_BYTES;
        uint4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = *(const uint4 *)(src + (rb + 4 * j) * VSB_TW + c * 16);
        const uint32_t first = src[0];

        // ---- mask halves, uniformity, binarity ------------------------------------------------------------------------
        // (rows / 16-byte chunks outside the image arrive as zeros: they yield 0 bits, count as binary, and are left out of
        // the uniformity test; W is a multiple of 16 on this path, so a chunk is either inside or outside)
        const uint32_t bc = first * 0x01010101u;
        const bool col_in = x0 + c * 16 < p.W;
        const int rows_in = p.H - y0;
        uint32_t diff = 0, nonbin = 0, half[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
            uint32_t m[4], dj = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                m[i] = msb_bytes(w[i]);
                dj |= w[i] ^ bc;
                nonbin |= w[i] ^ m[i];
            }
            if (col_in && rb + 4 * j < rows_in) diff |= dj;
            // four mask bits per word: bytes 0x00/0xFF -> one bit each, summed into the top byte by a multiply
            const uint32_t u01 = ((m[0] & 0x08040201u) | (m[1] & 0x80402010u)) * 0x01010101u;
            const uint32_t u23 = ((m[2] & 0x08040201u) | (m[3] & 0x80402010u)) * 0x01010101u;
            half[j] = (u01 >> 24) | ((u23 >> 16) & 0xFF00u);
        }
        const bool binary = __all_sync(0xffffffffu, nonbin == 0u);
        if (!LUTOUT) {
            // the stage has been consumed (every value above depends on the loads): refill it
            __syncwarp();
            if (lane == 0 && k + S < count) issue(k + S, st);
        } else {
            // ---- lut[gray], in place, then out through the tensor store ---------------------------------------------------
            if (binary) { // every byte is 0 or 255: select between lut[0] and lut[255] by the sign mask, no look-ups
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t m0 = msb_bytes(v[j].x), m1 = msb_bytes(v[j].y), m2 = msb_bytes(v[j].z), m3 = msb_bytes(v[j].w);
                    v[j] = make_uint4((m0 & K1) | (~m0 & K0), (m1 & K1) | (~m1 & K0), (m2 & K1) | (~m2 & K0), (m3 & K1) | (~m3 & K0));
                }
            } else if (p.lut) { // gray bytes in the tile: real table look-ups
                auto look = [&](uint32_t w) -> uint32_t {
                    return (uint32_t)s_lut[w & 255u] | ((uint32_t)s_lut[(w >> 8) & 255u] << 8) |
                           ((uint32_t)s_lut[(w >> 16) & 255u] << 16) | ((uint32_t)s_lut[w >> 24] << 24);
                };
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = make_uint4(look(v[j].x), look(v[j].y), look(v[j].z), look(v[j].w));
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) *(uint4 *)(src + (rb + 4 * j) * VSB_TW + c * 16) = v[j];
            vsb_fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                vsb_tma_store_2d(&tm_out, x0, y0, src);
                vsb_tma_commit();
            }
        }

        // ---- pair exchange -> four full mask words per lane -------------------------------------------------------------
        const uint32_t s0 = odd ? (half[0] | (half[2] << 16)) : (half[1] | (half[3] << 16));
        const uint32_t s1 = odd ? (half[4] | (half[6] << 16)) : (half[5] | (half[7] << 16));
        const uint32_t r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
        uint32_t word[4];
        if (!odd) { // rows j = 0, 2, 4, 6: my chunk is the low half
            word[0] = half[0] | (r0 << 16);
            word[1] = half[2] | (r0 & 0xFFFF0000u);
            word[2] = half[4] | (r1 << 16);
            word[3] = half[6] | (r1 & 0xFFFF0000u);
        } else { // rows j = 1, 3, 5, 7: my chunk is the high half
            word[0] = (r0 & 0xFFFFu) | (half[1] << 16);
            word[1] = (r0 >> 16) | (half[3] << 16);
            word[2] = (r1 & 0xFFFFu) | (half[5] << 16);
            word[3] = (r1 >> 16) | (half[7] << 16);
        }
        bool rec = false;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int y = y0 + rb + 4 * (2 * m + odd);
            if (y < p.H) p.bits[(size_t)y * p.wpr + tx * 4 + wq] = word[m];
            if (want_rec) { // first use of the words requested one tile ago
                uint32_t pa = held[0][m];
#pragma unroll
                for (int j = 1; j < PK_HELD; ++j) pa |= held[j][m];
                rec = rec || (pa & ~word[m]) != 0u;
            }
        }

        // ---- tile flag: uniform value | MIXED, BINARY --------------------------------------------------------------------
        const bool uni = __all_sync(0xffffffffu, diff == 0u);
        const bool any_rec = want_rec && __any_sync(0xffffffffu, rec);
        if (lane == 0) {
            p.flags[tile] = (uint16_t)((uni ? first : FLAG_MIXED) | (binary ? FLAG_BINARY : 0));
            if (any_rec) p.rec_list[1 + atomicAdd(&p.rec_list[0], 1)] = tile;
        }
        if (want_rec) { // requests for the tiles ahead (consumed at the end of the next iteration)
#pragma unroll
            for (int j = 0; j < PK_HELD; ++j)
#pragma unroll
                for (int m = 0; m < 4; ++m) held[j][m] = 0u;
            if (k + 1 < count && needs_priors(gate_flag_next)) load_priors(ty1, tx1);
            gate_flag_next = load_gate(k + 2);
        }
        ty = ty1; tx = tx1;
        if (++st == S) { st = 0; phase ^= 1u; }
    }
    if (LUTOUT && lane == 0) vsb_tma_wait_all<0>();
}

// ---- launcher (called from vsb_layer.cu's pack entry points) -----------------------------------------------------------
int vsb_pack_tma_launch(const uint8_t *gray, size_t pitch, int W, int H, uint32_t *bits, int wpr, uint16_t *flags, bool lutout,
                        const uint8_t *lut, uint8_t *out, size_t opitch, const uint32_t *const *prior_bits,
                        const uint16_t *const *prior_flags, int n_priors, int *rec_list, cudaStream_t s)
{
    static int n_sms = 0, stages_env = -1, ctas_env = -1;
    if (n_sms == 0) {
        int dev = 0, sms = 0;
        VSB_CUDA(cudaGetDevice(&dev));
        VSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        n_sms = sms > 0 ? sms : 148;
        const char *e = getenv("VSB_PACK_STAGES");
        stages_env = e ? atoi(e) : 0;
        e = getenv("VSB_PACK_CTAS");
        ctas_env = e ? atoi(e) : 0;
    }
    static bool seen[VSB_MAX_DEVICES] = {false};
    if (vsb_first_on_device(seen)) {
        VSB_CUDA(cudaFuncSetAttribute(k_pack_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        VSB_CUDA(cudaFuncSetAttribute(k_pack_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    PackTmaParams p;
    memset(&p, 0, sizeof p);
    p.bits = bits; p.flags = flags; p.lut = lut; p.rec_list = rec_list;
    p.np = rec_list ? n_priors : 0;
    p.gate = (rec_list && prior_flags) ? 1 : 0;
    for (int i = 0; i < p.np; ++i) {
        p.prior[i] = prior_bits[i];
        p.prior_flags[i] = prior_flags ? prior_flags[i] : nullptr;
        if (prior_flags && !prior_flags[i]) p.gate = 0;
    }
    p.wpr = wpr; p.tiles_x = (W + VSB_TW - 1) / VSB_TW;
    const int tiles_y = (H + VSB_TH - 1) / VSB_TH;
    p.ntiles = p.tiles_x * tiles_y; p.W = W; p.H = H;
    p.stages = stages_env >= 3 && stages_env <= 8 ? stages_env : (lutout ? 5 : 3);
    CUtensorMap tm_in, tm_out;
    int rc = vsb_make_tmap_2d(&tm_in, gray, 1, (size_t)W, (size_t)H, pitch, VSB_TW, VSB_TH);
    if (rc) return rc;
    if (lutout) {
        rc = vsb_make_tmap_2d(&tm_out, out, 1, (size_t)W, (size_t)H, opitch, VSB_TW, VSB_TH);
        if (rc) return rc;
    } else {
        tm_out = tm_in;
    }
    const size_t smem = (size_t)PK_WARPS * p.stages * PK_TILE_BYTES + 128;
    int per_sm = ctas_env > 0 ? ctas_env : (int)((220 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int grid = n_sms * per_sm;
    const int need = (p.ntiles + PK_WARPS - 1) / PK_WARPS;
    if (grid > need) grid = need;
    if (lutout) k_pack_tma<true><<<grid, PK_WARPS * 32, smem, s>>>(tm_in, tm_out, p);
    else k_pack_tma<false><<<grid, PK_WARPS * 32, smem, s>>>(tm_in, tm_out, p);
    return vsb_check_launch("k_pack_tma");
}
