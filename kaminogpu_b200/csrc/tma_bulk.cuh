// 1-D bulk asynchronous copies global -> shared memory through the TMA unit
// (cp.async.bulk, completion signalled on an mbarrier), sm_90+/sm_100a.
//
// Usage pattern in the kernels (one producer thread, everybody consumes):
//     __shared__ __align__(8) uint64_t bar;
//     if (threadIdx.x == 0) { mbarInit(&bar, 1); fenceBarrierInit(); }
//     __syncthreads();
//     if (threadIdx.x == 0) { mbarExpectTx(&bar, bytes); bulkLoad(dst, src, bytes, &bar); ... }
//     ... independent work ...
//     mbarWait(&bar, 0);          // every consumer thread; phase parity 0 for the first use
// Sizes and both addresses must be multiples of 16 bytes.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace kb {
namespace tma {

__device__ __forceinline__ uint32_t smemAddr(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbarInit(uint64_t* bar, int arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(arrivals) : "memory");
}

// make the initialised barrier visible to the async (TMA) proxy
__device__ __forceinline__ void fenceBarrierInit()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulkLoad(void* dstShared, const void* srcGlobal, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smemAddr(dstShared)), "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar))
                 : "memory");
}

__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "KB_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra KB_WAIT_DONE;\n"
        "bra KB_WAIT_LOOP;\n"
        "KB_WAIT_DONE:\n"
        "}\n" ::"r"(smemAddr(bar)), "r"(parity)
        : "memory");
}

} // namespace tma
} // namespace kb
