// Micro-benchmark (measurement tool): time per tcgen05.mma for the conv_tc operand pattern.
//   real   : A start = tile + (m*128 + r*PW + s) * 16 B (16-byte aligned only), LBO = PLANE, descriptors change per MMA
//   align  : same loop, pixel offsets rounded down to multiples of 8 pixels (128-byte aligned core matrices)
//   const  : one descriptor pair re-issued (issue-rate floor)
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../ball_action_spotting_b200/csrc/common.cuh"
#include "../ball_action_spotting_b200/csrc/gemm_tc.cuh"
#include "../ball_action_spotting_b200/csrc/tc_helpers.cuh"

using namespace mds;

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

template <int CIN, int CMID>
__global__ void __launch_bounds__(128, 1) probe(int mode, int mtiles, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint32_t s_tmem;
    __shared__ uint64_t bar;
    constexpr int PW = 34, PIX = 34 * 17, PLANE = PIX * 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = s_tmem;
    if (warp == 1 && mode == 6) {          // whole warp runs the loop, one elected lane issues
        const uint32_t idesc = tc_idesc(128, CMID);
        const uint32_t ta = smem_u32(smem), w1a = smem_u32(smem) + 48 * 1024;
        long long t0 = clock64();
        for (int t = 0; t < mtiles; ++t) {
            const int m = t & 3;
            const uint32_t d1 = tb + (t & 1) * CMID;
#pragma unroll
            for (int rs = 0; rs < 9; ++rs) {
                const int r = rs / 3, s = rs - r * 3;
                const int off = m * 128 + r * PW + s;
                const uint32_t a_addr = ta + (uint32_t)off * 16u;
#pragma unroll
                for (int kc = 0; kc < CIN / 16; ++kc) {
                    const uint64_t ad = tc_desc_nosw(a_addr + kc * 2 * PLANE, PLANE);
                    const uint64_t bd = tc_desc_nosw(w1a + (rs * (CIN / 8) + kc * 2) * (CMID * 16), CMID * 16);
                    if (elect_one()) tc_mma_f16(d1, ad, bd, idesc, 1);
                }
            }
            __syncwarp();
        }
        if (elect_one()) tc_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (lane == 0) out[blockIdx.x] = t1 - t0;
    } else if (warp == 1 && lane == 0) {
        const uint32_t idesc = tc_idesc(128, CMID);
        const uint32_t ta = smem_u32(smem), w1a = smem_u32(smem) + 48 * 1024;
        long long t0 = clock64();
        for (int t = 0; t < (mode < 3 ? mtiles : 0); ++t) {
            const int m = t & 3;
            const uint32_t d1 = tb + (t & 1) * CMID;
#pragma unroll
            for (int rs = 0; rs < 9; ++rs) {
                const int r = rs / 3, s = rs - r * 3;
                int off = m * 128 + r * PW + s;
                if (mode == 1) off &= ~7;
                if (mode == 2) off = 0;
                const uint32_t a_addr = ta + (uint32_t)off * 16u;
#pragma unroll
                for (int kc = 0; kc < CIN / 16; ++kc) {
                    const int kcc = mode == 2 ? 0 : kc;
                    const int rss = mode == 2 ? 0 : rs;
                    tc_mma_f16(d1, tc_desc_nosw(a_addr + kcc * 2 * PLANE, PLANE),
                               tc_desc_nosw(w1a + (rss * (CIN / 8) + kcc * 2) * (CMID * 16), CMID * 16), idesc, 1);
                }
            }
        }
        if (mode == 3 || mode == 4 || mode == 5) {        // TS form: A = 8 TMEM columns (packed fp16), B from smem; N = CMID
            t0 = clock64();
            const int nn = mode == 5 ? 16 : CMID;
            const uint32_t id2 = tc_idesc(128, nn);
            for (int t = 0; t < mtiles; ++t) {
#pragma unroll
                for (int kk = 0; kk < 18; ++kk) {
                    const uint64_t bdesc = tc_desc_nosw(w1a + kk * 2 * (nn * 16), nn * 16);
                    if (mode == 4) tc_mma_f16(tb + 256, tc_desc_nosw(ta + kk * 4096, 2048), bdesc, id2, 1);     // SS for comparison
                    else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                      ::"r"(tb + 256), "r"(tb + (uint32_t)(kk * 8)), "l"(bdesc), "r"(id2), "r"(1u) : "memory");
                }
            }
        }
        tc_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
    }
}

template <int CIN, int CMID>
static void run() {
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    cudaFuncSetAttribute(probe<CIN, CMID>, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
    const int mtiles = 512;
    const char* names[7] = {"real", "align", "const", "TS", "SS-k18", "TS-N16", "elect"};
    for (int mode = 0; mode < 7; ++mode) {
        probe<CIN, CMID><<<148, 128, 202 * 1024>>>(mode, mtiles, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); exit(1); }
        long long h[148];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        const int nmma = (mode < 3 || mode == 6) ? mtiles * 9 * (CIN / 16) : mtiles * 18;
        printf("CIN %2d N %3d %-5s: %8lld clk = %6.1f clk/MMA (math floor %d)\n", CIN, CMID, names[mode], h[0], (double)h[0] / nmma, CMID / 2);
    }
    cudaFree(d);
}

int main() {
    run<32, 16>();
    run<16, 64>();
    run<32, 128>();
    return 0;
}
