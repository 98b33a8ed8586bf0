// vsb_mask.cu -- K1: threshold + bit-pack, mask OR / difference on packed words.
//
// Replaces cv2.threshold (processing_core.py:44), the bitwise_or loop of
// find_prior_combined_white_mask (:61-65) and the bitwise_not/and/countNonZero prologue of every
// blend mode (:119-122, :259-260, :350-351).  HBM-bound byte work: one 2x128-bit load per thread
// (32 pixels -> one mask word), coalesced over the row, grid sized in whole rows.
#include "vsb_common.cuh"

// One thread = 32 pixels = one mask word.  blockDim.x threads along x, blockDim.y rows per CTA.
template <bool ALIGNED>
__global__ void __launch_bounds__(256) k_threshold_pack(const uint8_t *__restrict__ gray, size_t pitch, int W,
                                                        int H, uint32_t thr4, uint32_t *__restrict__ bits,
                                                        int wpr)
{
    const int wx = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (y >= H || wx >= wpr) return;
    const int x = wx << 5;
    uint32_t word = 0;
    if (x < W) {
        const uint8_t *p = gray + (size_t)y * pitch + x;
        const int n = W - x;
        uint4 a, b;
        if (ALIGNED && n >= 32) {
            a = vsb_ld16_stream(p);
            b = vsb_ld16_stream(p + 16);
        } else {
            a = vsb_ld_bytes(p, n);
            b = n > 16 ? vsb_ld_bytes(p + 16, n - 16) : make_uint4(0, 0, 0, 0);
        }
        word = vsb_gt_bits16(a, thr4) | (vsb_gt_bits16(b, thr4) << 16);
        if (n < 32) word &= (1u << n) - 1u; // zero-filled bytes never exceed thr, mask anyway
    }
    bits[(size_t)y * wpr + wx] = word;
}

template <bool ALIGNED>
__global__ void __launch_bounds__(256) k_bits_unpack(const uint32_t *__restrict__ bits, int wpr, int W, int H,
                                                     uint8_t *__restrict__ out, size_t pitch)
{
    // one thread = 16 pixels (half a word)
    const int hx = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x = hx << 4;
    if (y >= H || x >= W) return;
    uint32_t w = bits[(size_t)y * wpr + (x >> 5)];
    uint32_t h = (x & 16) ? (w >> 16) : (w & 0xFFFFu);
    uint4 v = make_uint4(vsb_nibble_to_bytes(h & 15u), vsb_nibble_to_bytes((h >> 4) & 15u),
                         vsb_nibble_to_bytes((h >> 8) & 15u), vsb_nibble_to_bytes(h >> 12));
    uint8_t *p = out + (size_t)y * pitch + x;
    const int n = W - x;
    if (ALIGNED && n >= 16) vsb_st16_stream(p, v);
    else vsb_st_bytes(p, v, n);
}

struct PriorPtrs {
    const uint32_t *p[VSB_MAX_PRIORS];
};

__global__ void __launch_bounds__(256) k_bits_or(PriorPtrs src, int n, size_t nwords, uint32_t *__restrict__ dst)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < nwords; i += stride) {
        uint32_t v = 0;
        for (int k = 0; k < n; ++k) v |= src.p[k][i];
        dst[i] = v;
    }
}

__global__ void __launch_bounds__(256) k_maskdiff(const uint32_t *__restrict__ cur, PriorPtrs pri, int n,
                                                  size_t nwords, uint32_t *__restrict__ rec,
                                                  unsigned long long *__restrict__ count)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned int local = 0;
    for (; i < nwords; i += stride) {
        uint32_t v = 0;
        for (int k = 0; k < n; ++k) v |= pri.p[k][i];
        v &= ~cur[i];
        if (rec) rec[i] = v;
        local += __popc(v);
    }
    if (count) {
        for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, (unsigned long long)local);
    }
}

extern "C" int vsb_threshold_pack(const uint8_t *gray, size_t pitch, int W, int H, int thresh, uint32_t *bits,
                                  int wpr, vsb_stream_t stream)
{
    VSB_REQUIRE(gray && bits, "vsb_threshold_pack: null pointer");
    VSB_REQUIRE(W > 0 && H > 0 && wpr >= (W + 31) / 32 && pitch >= (size_t)W, "vsb_threshold_pack: bad geometry");
    VSB_REQUIRE(thresh >= 0 && thresh <= 255, "vsb_threshold_pack: thresh out of range");
    uint32_t thr4 = (uint32_t)thresh * 0x01010101u;
    dim3 block(64, 4);
    dim3 grid((wpr + block.x - 1) / block.x, (H + block.y - 1) / block.y);
    cudaStream_t s = (cudaStream_t)stream;
    if (vsb_aligned16(gray, pitch))
        k_threshold_pack<true><<<grid, block, 0, s>>>(gray, pitch, W, H, thr4, bits, wpr);
    else
        k_threshold_pack<false><<<grid, block, 0, s>>>(gray, pitch, W, H, thr4, bits, wpr);
    return vsb_check_launch("k_threshold_pack");
}

extern "C" int vsb_bits_unpack(const uint32_t *bits, int wpr, int W, int H, uint8_t *out, size_t pitch,
                               vsb_stream_t stream)
{
    VSB_REQUIRE(bits && out, "vsb_bits_unpack: null pointer");
    VSB_REQUIRE(W > 0 && H > 0 && wpr >= (W + 31) / 32 && pitch >= (size_t)W, "vsb_bits_unpack: bad geometry");
    dim3 block(64, 4);
    dim3 grid(((W + 15) / 16 + block.x - 1) / block.x, (H + block.y - 1) / block.y);
    cudaStream_t s = (cudaStream_t)stream;
    if (vsb_aligned16(out, pitch)) k_bits_unpack<true><<<grid, block, 0, s>>>(bits, wpr, W, H, out, pitch);
    else k_bits_unpack<false><<<grid, block, 0, s>>>(bits, wpr, W, H, out, pitch);
    return vsb_check_launch("k_bits_unpack");
}

static int grid_for(size_t n)
{
    size_t g = (n + 255) / 256;
    size_t cap = 148 * 8;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

extern "C" int vsb_bits_or(const uint32_t *const *srcs, int n, int wpr, int H, uint32_t *dst, vsb_stream_t stream)
{
    VSB_REQUIRE(srcs && dst && n >= 1 && n <= VSB_MAX_PRIORS, "vsb_bits_or: need 1..%d sources", VSB_MAX_PRIORS);
    PriorPtrs pp;
    for (int i = 0; i < VSB_MAX_PRIORS; ++i) pp.p[i] = i < n ? srcs[i] : nullptr;
    size_t nwords = (size_t)wpr * H;
    k_bits_or<<<grid_for(nwords), 256, 0, (cudaStream_t)stream>>>(pp, n, nwords, dst);
    return vsb_check_launch("k_bits_or");
}

extern "C" int vsb_maskdiff(const uint32_t *cur, const uint32_t *const *priors, int n, int wpr, int H,
                            uint32_t *rec_out, unsigned long long *count_out, vsb_stream_t stream)
{
    VSB_REQUIRE(cur && (n == 0 || priors) && n >= 0 && n <= VSB_MAX_PRIORS, "vsb_maskdiff: bad arguments");
    PriorPtrs pp;
    for (int i = 0; i < VSB_MAX_PRIORS; ++i) pp.p[i] = i < n ? priors[i] : nullptr;
    size_t nwords = (size_t)wpr * H;
    k_maskdiff<<<grid_for(nwords), 256, 0, (cudaStream_t)stream>>>(cur, pp, n, nwords, rec_out, count_out);
    return vsb_check_launch("k_maskdiff");
}
