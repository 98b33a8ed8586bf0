// vsb_common.cuh -- shared device/host helpers of libvsb_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#pragma GCC visibility push(default)
#include "../../include/vsb.h"
#pragma GCC visibility pop

#define VSB_TW 128 // tile width  (pixels) = 4 mask words = one 128-byte line of gray
#define VSB_TH 32  // tile height (rows)
#define VSB_THREADS 256

// ---- error plumbing (vsb_api.cu) ---------------------------------------------------------------
void vsb_set_error(const char *fmt, ...);
void vsb_count_launch(int n = 1);
int vsb_check_launch(const char *what);

#define VSB_REQUIRE(cond, ...)                                                                     \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            vsb_set_error(__VA_ARGS__);                                                            \
            return VSB_ERR_ARG;                                                                    \
        }                                                                                          \
    } while (0)

#define VSB_CUDA(call)                                                                             \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            vsb_set_error("%s failed: %s", #call, cudaGetErrorString(e__));                        \
            return VSB_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

// Function attributes (opt-in shared memory, carve-out) belong to the device the call runs on: a process that drives several
// devices (config.devices: one engine thread per GPU) has to set them once on each.  `seen` is a per-call-site flag array.
#define VSB_MAX_DEVICES 32
static inline bool vsb_first_on_device(bool (&seen)[VSB_MAX_DEVICES])
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= VSB_MAX_DEVICES) return true;
    if (seen[dev]) return false;
    seen[dev] = true;
    return true;
}

static inline bool vsb_aligned16(const void *p, size_t pitch)
{
    return (((uintptr_t)p) & 15u) == 0 && (pitch & 15u) == 0;
}

// ---- device helpers ----------------------------------------------------------------------------
#ifdef __CUDACC__
// streaming 128-bit accesses: every gray/out byte is touched once per launch, keep it out of L1
__device__ __forceinline__ uint4 vsb_ld16_stream(const void *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void vsb_st16_stream(void *p, uint4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// 16 bytes of one row with a ragged right edge / unaligned base: byte loads, zero fill
__device__ __forceinline__ uint4 vsb_ld_bytes(const uint8_t *p, int n)
{
    uint32_t w[4] = {0, 0, 0, 0};
    for (int i = 0; i < n && i < 16; ++i) w[i >> 2] |= (uint32_t)p[i] << ((i & 3) * 8);
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void vsb_st_bytes(uint8_t *p, uint4 v, int n)
{
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    for (int i = 0; i < n && i < 16; ++i) p[i] = (uint8_t)(w[i >> 2] >> ((i & 3) * 8));
}

// 4 packed bytes -> 4 mask bits (bit i = byte i > thr), thr4 = thr replicated
__device__ __forceinline__ uint32_t vsb_gt_nibble(uint32_t v, uint32_t thr4)
{
    uint32_t m = __vcmpgtu4(v, thr4) & 0x01010101u;
    return (m * 0x01020408u) >> 24;
}
__device__ __forceinline__ uint32_t vsb_gt_bits16(uint4 v, uint32_t thr4)
{
    return vsb_gt_nibble(v.x, thr4) | (vsb_gt_nibble(v.y, thr4) << 4) |
           (vsb_gt_nibble(v.z, thr4) << 8) | (vsb_gt_nibble(v.w, thr4) << 12);
}

// 4 mask bits -> 4 bytes of 0x00 / 0xFF
__device__ __forceinline__ uint32_t vsb_nibble_to_bytes(uint32_t nib)
{
    uint32_t t = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
    return t * 0xFFu;
}

// 256-entry LUT in shared memory applied to 4 packed bytes; 0x00000000 / 0xFFFFFFFF words (the
// bulk of a slice) take the precomputed broadcast
struct VsbLut {
    const uint8_t *s; // shared memory table (nullptr = identity)
    uint32_t l0, l255;
    __device__ __forceinline__ uint32_t map4(uint32_t v) const
    {
        if (s == nullptr) return v;
        if (v == 0u) return l0;
        if (v == 0xFFFFFFFFu) return l255;
        return (uint32_t)s[v & 255u] | ((uint32_t)s[(v >> 8) & 255u] << 8) |
               ((uint32_t)s[(v >> 16) & 255u] << 16) | ((uint32_t)s[v >> 24] << 24);
    }
    __device__ __forceinline__ uint4 map16(uint4 v) const
    {
        return make_uint4(map4(v.x), map4(v.y), map4(v.z), map4(v.w));
    }
};

// cooperative load of a device LUT into shared memory; returns the accessor (call before a
// __syncthreads()).  lut == nullptr -> identity.
__device__ __forceinline__ VsbLut vsb_lut_stage(const uint8_t *lut, uint8_t *s_lut)
{
    VsbLut L;
    if (lut == nullptr) {
        L.s = nullptr; L.l0 = 0; L.l255 = 0xFFFFFFFFu;
        return L;
    }
    if (threadIdx.x < 64) ((uint32_t *)s_lut)[threadIdx.x] = ((const uint32_t *)lut)[threadIdx.x];
    L.s = s_lut;
    L.l0 = (uint32_t)lut[0] * 0x01010101u;
    L.l255 = (uint32_t)lut[255] * 0x01010101u;
    return L;
}

__device__ __forceinline__ int vsb_reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        if (i >= n) i = 2 * (n - 1) - i;
    }
    return i;
}
__device__ __forceinline__ int vsb_clampi(int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }
#endif // __CUDACC__
