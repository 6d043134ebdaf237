// tma.cuh — the few raw PTX pieces of Blackwell's asynchronous copy machinery this library uses: 1-D bulk copies
// global -> shared memory issued by ONE thread (cp.async.bulk, the TMA unit without a tensor map) and the mbarrier
// whose transaction count tells the consumers that the bytes have landed. sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace slpr {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
// make the initialised barriers visible to the async proxy (the copy unit) before the first copy is issued
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// one arrival + "expect `bytes` more bytes of copy transactions" (the phase completes when both are satisfied)
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// plain arrival (release at CTA scope: what the arriving thread wrote before is visible to a thread that has seen
// the phase complete)
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

// `bytes` (a multiple of 16) from 16-byte aligned global memory to 16-byte aligned shared memory; completion is
// counted on `bar`. Without a cluster launch the CTA is its own cluster, so its shared::cta addresses are valid
// shared::cluster addresses.
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// barrier among a subset of the block's warps (the producer / look-back warp does not take part)
__device__ __forceinline__ void named_barrier_sync(int id, int n_threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

}  // namespace slpr
