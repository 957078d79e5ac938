// Micro-benchmark (measurement tool): global-store throughput of the GEMM epilogue's pattern (each lane owns a row and stores
// 32 bytes, rows `pitch` bytes apart) against coalesced alternatives.  One persistent CTA of 512 threads per SM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void st_v8(void* p, uint32_t v) {
    asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_v4(void* p, uint32_t v) {
    asm volatile("st.global.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}
// mode 0: lane = row, 32 B per lane (st.v8), columns advance by 32 B per iteration          (the GEMM epilogue today)
// mode 1: same with two st.v4
// mode 2: lane = 16 B column chunk of ONE row (512 B contiguous per warp instruction), rows advance per iteration
// mode 3: 8 lanes cover 128 B of a row, 4 rows per instruction (st.v4)
__global__ void __launch_bounds__(512, 1) probe(unsigned char* out, int pitch, int cols_bytes, int tiles, int mode, long long* clk) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long t0 = clock64();
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        unsigned char* base = out + (size_t)t * 128 * pitch;          // a 128-row tile
        if (mode == 0 || mode == 1) {
            const int row = (warp & 3) * 32 + lane;                    // 4 quadrants x 4 column parts
            const int part = warp >> 2, per = cols_bytes / 4;
            for (int c = 0; c < per; c += 32) {
                unsigned char* p = base + (size_t)row * pitch + part * per + c;
                if (mode == 0) st_v8(p, t + c); else { st_v4(p, t + c); st_v4(p + 16, t + c); }
            }
        } else if (mode == 2) {
            for (int r = warp; r < 128; r += 16)
                for (int c = lane * 16; c < cols_bytes; c += 512) st_v4(base + (size_t)r * pitch + c, t + c);
        } else {
            for (int r = warp * 4 + (lane >> 3); r < 128; r += 64)
                for (int c = (lane & 7) * 16; c < cols_bytes; c += 128) st_v4(base + (size_t)r * pitch + c, t + c);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}
int main() {
    const int pitch = 1344, cols = 1280, tiles = 148 * 40;
    unsigned char* d; long long* c;
    cudaMalloc(&d, (size_t)tiles * 128 * pitch);
    cudaMalloc(&c, 148 * sizeof(long long));
    for (int rep = 0; rep < 2; ++rep)
        for (int mode = 0; mode < 4; ++mode) {
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a);
            probe<<<148, 512>>>(d, pitch, cols, tiles, mode, c);
            cudaEventRecord(b);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
            float ms; cudaEventElapsedTime(&ms, a, b);
            long long h[148]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
            const double bytes = (double)tiles * 128 * cols;
            printf("mode %d: %.1f us  %.0f GB/s  %.1f B/clk/SM\n", mode, ms * 1e3, bytes / ms * 1e-6, bytes / 148 / (double)h[0]);
        }
    return 0;
}
