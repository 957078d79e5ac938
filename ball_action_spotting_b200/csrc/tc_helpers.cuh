// Helpers shared by the implicit-GEMM tcgen05 kernels (conv_tc.cuh, conv_tc_ws.cuh, stem_tc.cuh) and the TMA depthwise kernels:
// the no-swizzle K-major UMMA operand descriptor and the 5-D TMA tensor load.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace mds {

// no-swizzle K-major operand: rows 16 B apart, 8-row groups SBO = 128 B apart, the two 8-element K halves LBO apart
__device__ __forceinline__ uint64_t tc_desc_nosw(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) |
           ((uint64_t)1 << 46);
}

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

}  // namespace mds
