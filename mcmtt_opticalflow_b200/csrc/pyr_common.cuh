// Device helpers shared by the per-level (pyramid.cu) and the fused multi-level (pyramid_fused.cu) pyramid kernels:
// TMA / mbarrier wrappers and the dp4a coefficient words of the separable integer filters.
#pragma once
#include "common.cuh"

namespace psnw {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(phase)
            : "memory");
    } while (!ok);
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

// dp4a coefficient words (byte 0 = lowest address)
constexpr int kDiffLo = 0x000100FF;     // (-1, 0, 1, 0)
constexpr int kDiffHi = 0x0100FF00;     // ( 0,-1, 0, 1)
constexpr int kSmoothLo = 0x00030A03;   // ( 3,10, 3, 0)
constexpr int kSmoothHi = 0x030A0300;   // ( 0, 3,10, 3)
constexpr int kGauss = 0x04060401;      // ( 1, 4, 6, 4)

}  // namespace
}  // namespace psnw
