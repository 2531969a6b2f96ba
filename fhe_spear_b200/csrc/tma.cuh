// tma.cuh -- mbarrier / TMA / L2 cache-policy helpers shared by the kernels that stream rotation keys and diagonals
// (ops.cu: key inner product, diagonal MAC; ntt.cu: forward pass B fused with the key inner product).
#pragma once
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"

// ---- TMA / mbarrier / cache-policy helpers ------------------------------------------------------
static __device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((u32)__cvta_generic_to_shared(bar)), "r"(count));
}
static __device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((u32)__cvta_generic_to_shared(bar)),
                 "r"(bytes));
}
static __device__ __forceinline__ void mbar_wait(uint64_t* bar, u32 parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"((u32)__cvta_generic_to_shared(bar)),
        "r"(parity));
}
static __device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            (u32)__cvta_generic_to_shared(dst)),
        "l"(map), "r"((u32)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

static __device__ __forceinline__ void tma_load_3d_hint(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                                 uint64_t* bar, u64 policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, "
        "%5}], [%2], %6;" ::"r"((u32)__cvta_generic_to_shared(dst)),
        "l"(map), "r"((u32)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}
static __device__ __forceinline__ u64 evict_last_policy() {
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
static __device__ __forceinline__ u64 ld_keep(const u64* p, u64 pol) {
    u64 v;
    asm volatile("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    return v;
}
