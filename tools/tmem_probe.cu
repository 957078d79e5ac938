// Micro-benchmark of the tensor-memory read path on sm_100a (measurement tool, not part of the library):
// how many bytes per clock tcgen05.ld delivers for different shapes / warp counts / waits, and how much a concurrent
// stream of tcgen05.mma slows it down (and vice versa).  Build + run: tools/tmem_probe.sh (on the GPU box).
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../ball_action_spotting_b200/csrc/common.cuh"
#include "../ball_action_spotting_b200/csrc/gemm_tc.cuh"
#include "../ball_action_spotting_b200/csrc/tc_helpers.cuh"

using namespace mds;

__device__ __forceinline__ void ld_32x32b_x8(uint32_t a, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld_32x32b_x32(uint32_t a, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
        "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(a) : "memory");
}
// 16x256b.x2: 16 lanes x 512 bits = 16 columns of 16 lanes, 8 registers per thread
__device__ __forceinline__ void ld_16x256b_x2(uint32_t a, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(a) : "memory");
}
// 16x128b.x4: 16 lanes x 512 bits
__device__ __forceinline__ void ld_16x128b_x4(uint32_t a, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(a) : "memory");
}

struct Res { long long clk_ld, clk_mma; unsigned sink; };

// mode: 0 = 32x32b.x8, 1 = 32x32b.x16, 2 = 32x32b.x32, 3 = 16x256b.x2 (two per 32 lanes), 4 = 16x128b.x4 (two per 32 lanes)
// depth: loads issued per tcgen05.wait::ld; nwarps: reader warps (warps 2..); mma_n: 0 = no MMA stream, else N of a
// back-to-back M=128,K=16 MMA stream issued by warp 1 while the readers run; iters: loads per reader warp
template <int MODE>
__global__ void __launch_bounds__(64 + 32 * 16, 1) probe(int nwarps, int depth, int iters, int mma_n, int mma_count, Res* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint32_t s_tmem;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 65536 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
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
    long long t0 = 0, t1 = 0;
    unsigned sink = 0;
    if (warp == 1) {
        if (lane == 0 && mma_n > 0) {
            const uint32_t idesc = tc_idesc(128, mma_n);
            const uint64_t ad = tc_desc_nosw(smem_u32(smem), 128 * 16), bd = tc_desc_nosw(smem_u32(smem) + 8192, 256 * 16);
            t0 = clock64();
            for (int i = 0; i < mma_count; ++i) tc_mma_f16(tb + 256, ad, bd, idesc, 1);       // accumulator in columns 256..
            tc_commit(&bar);
            mbar_wait(&bar, 0);
            t1 = clock64();
            out[blockIdx.x].clk_mma = t1 - t0;
        }
    } else if (warp >= 2 && warp < 2 + nwarps) {
        const uint32_t t_row = tb + ((uint32_t)((warp & 3) * 32) << 16);
        const int wsel = (warp - 2) >> 2;                 // warps sharing a quadrant read different columns
        t0 = clock64();
        for (int i = 0; i < iters; i += depth) {
            for (int d = 0; d < depth; ++d) {
                const uint32_t col = (uint32_t)(((i + d) * 32 + wsel * 64) & 255);
                if (MODE == 0) { uint32_t v[8]; ld_32x32b_x8(t_row + col, v); sink ^= v[0] ^ v[7]; }
                if (MODE == 1) { uint32_t v[16]; tc_ld16(t_row + col, v); sink ^= v[0] ^ v[15]; }
                if (MODE == 2) { uint32_t v[32]; ld_32x32b_x32(t_row + col, v); sink ^= v[0] ^ v[31]; }
                if (MODE == 3) { uint32_t v[8], w[8]; ld_16x256b_x2(t_row + col, v); ld_16x256b_x2(t_row + (16u << 16) + col, w); sink ^= v[0] ^ w[7]; }
                if (MODE == 4) { uint32_t v[8], w[8]; ld_16x128b_x4(t_row + col, v); ld_16x128b_x4(t_row + (16u << 16) + col, w); sink ^= v[0] ^ w[7]; }
            }
            tc_wait_ld();
        }
        t1 = clock64();
        if (warp == 2 && lane == 0) out[blockIdx.x].clk_ld = t1 - t0;
    }
    if (sink == 0x12345678u) out[blockIdx.x].sink = sink;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512u) : "memory");
    }
}

template <int MODE>
static void run(const char* name, int bytes_per_ld) {
    Res* d;
    cudaMalloc(&d, 148 * sizeof(Res));
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    const int iters = 4096;
    for (int nw : {4, 8, 16})
        for (int depth : {1, 2, 4})
            for (int mma_n : {0, 128, 256}) {
                const int mma_count = 4096;
                cudaMemset(d, 0, 148 * sizeof(Res));
                probe<MODE><<<148, 64 + 32 * 16, 80 * 1024>>>(nw, depth, iters, mma_n, mma_count, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
                Res h[148];
                cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                const double bpc = (double)nw * iters * bytes_per_ld / (double)h[0].clk_ld;
                printf("%-12s warps %2d depth %d mma_n %3d: ld %8lld clk = %6.1f B/clk/SM", name, nw, depth, mma_n, h[0].clk_ld, bpc);
                if (mma_n) printf("   mma %8lld clk = %6.1f clk/MMA (floor %d)", h[0].clk_mma, (double)h[0].clk_mma / mma_count, mma_n / 2);
                printf("\n");
            }
    cudaFree(d);
}

int main() {
    run<0>("32x32b.x8", 32 * 8 * 4);
    run<1>("32x32b.x16", 32 * 16 * 4);
    run<2>("32x32b.x32", 32 * 32 * 4);
    run<3>("16x256b.x2x2", 32 * 16 * 4);
    run<4>("16x128b.x4x2", 32 * 16 * 4);
    return 0;
}
