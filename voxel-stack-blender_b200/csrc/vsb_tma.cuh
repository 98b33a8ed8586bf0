// vsb_tma.cuh -- the sm_100a bulk-copy plumbing shared by the TMA-staged kernels: mbarrier phases,
// cp.async.bulk.tensor loads / stores (SASS: UTMALDG / UTMASTG), proxy fences, and the host-side tensor-map
// encoder (cuTensorMapEncodeTiled through cudaGetDriverEntryPoint: no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include "vsb_common.cuh"

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t vsb_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void vsb_mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(vsb_smem_u32(bar)), "r"(count) : "memory");
}
// makes the barrier initialisation visible to the async proxy (the TMA unit) before the first copy names it
__device__ __forceinline__ void vsb_fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to a following bulk store
__device__ __forceinline__ void vsb_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void vsb_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(vsb_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void vsb_mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "VSB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra VSB_DONE;\n"
        "bra VSB_WAIT;\n"
        "VSB_DONE:\n"
        "}" ::"r"(vsb_smem_u32(bar)), "r"(parity) : "memory");
}

// 2-D tiled load: box (tensor-map box size) at element coordinates (x, y) -> shared memory, completion on `bar`.
// Out-of-bounds elements arrive as zeros.
__device__ __forceinline__ void vsb_tma_load_2d(void *smem_dst, const CUtensorMap *tm, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     vsb_smem_u32(smem_dst)),
                 "l"(tm), "r"(x), "r"(y), "r"(vsb_smem_u32(bar))
                 : "memory");
}
// 2-D tiled store: shared memory -> box at (x, y); elements outside the tensor are dropped
__device__ __forceinline__ void vsb_tma_store_2d(const CUtensorMap *tm, int x, int y, const void *smem_src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(vsb_smem_u32(smem_src)),
                 "r"(x), "r"(y)
                 : "memory");
}
__device__ __forceinline__ void vsb_tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most N committed store groups still READING their shared-memory source
template <int N>
__device__ __forceinline__ void vsb_tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void vsb_tma_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
#endif

// Tensor map over a row-major 2-D array of `elem_bytes`-sized elements: W x H elements, `pitch_bytes` between rows
// (multiple of 16, base 16-byte aligned), box_w x box_h elements per copy (box_w * elem_bytes a multiple of 16, both
// <= 256).  Returns VSB_OK or sets the error string.
int vsb_make_tmap_2d(CUtensorMap *out, const void *base, int elem_bytes, size_t W, size_t H, size_t pitch_bytes, int box_w,
                     int box_h);
