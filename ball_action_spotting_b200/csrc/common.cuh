// Shared device helpers for the MultiDimStacker sm_100a kernels.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

// Order in which the tensor-core kernels issue the MMAs of one accumulator (bit 0: stem weights lo before hi, bit 1: 3x3 taps
// last to first, bit 2: K steps of a GEMM block last to first).  Every order is the same sum in exact arithmetic; in fp32 they
// differ in the last bit of ~0.1 % of the outputs, which the fp16 activation roundings downstream amplify chaotically, so the
// end-to-end logit error of a given input is a different draw from the same distribution (RMS ~4e-4 of max|logit|, DESIGN.md
// "Numerics").  tests/parity_variants.py measures all eight; the shipped value is recorded there.
#ifndef MDS_NUMERICS_VARIANT
#define MDS_NUMERICS_VARIANT 5
#endif

namespace mds {

constexpr int kWarp = 32;

// sigmoid(x) = 1 / (1 + 2^(-x*log2e)) with the two SFU approximations in flush-to-zero form (5 instructions for SiLU:
// FMUL, MUFU.EX2, FADD, MUFU.RCP, FMUL; ~2 ulp, far below the fp16 storage rounding).  x -> -inf gives 2^inf = inf,
// rcp(inf) = 0; x -> +inf gives rcp(1) = 1.
__device__ __forceinline__ float sigmoid_f(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_f(x); }

// SiLU of four values with ONE reciprocal: 1/(1+e_i) = prod_{j != i}(1+e_j) / prod_j(1+e_j).  5 SFU operations per four values
// instead of 8 (the SiLU epilogues are bound by the SFU pipe, which also executes the fp32 -> fp16 packs), at the price of 10
// more FMA-pipe instructions.  x is clamped at -20 (SiLU(-20) = -4e-8 is below half of the smallest fp16 subnormal step from
// what any smaller x gives), so every factor is <= 2^29 and the product of four stays finite; ~1e-6 relative error
// (tests/test_host_logic.py emulates it in float32), 400x below the fp16 storage rounding.
__device__ __forceinline__ void silu4(float (&x)[4]) {
    float d[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        x[i] = fmaxf(x[i], -20.0f);
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x[i] * -1.4426950408889634f));
        d[i] = 1.0f + e;
    }
    const float p01 = d[0] * d[1], p23 = d[2] * d[3];
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p01 * p23));
    const float r01 = r * p23, r23 = r * p01;       // 1 / (d0 d1), 1 / (d2 d3)
    x[0] *= r01 * d[1]; x[1] *= r01 * d[0];
    x[2] *= r23 * d[3]; x[3] *= r23 * d[2];
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_half2(uint32_t u) {
    return __half22float2(*reinterpret_cast<__half2*>(&u));
}
__device__ __forceinline__ uint32_t hmul2_u32(uint32_t a, uint32_t b) {
    __half2 r = __hmul2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}

__device__ __forceinline__ uint32_t hfma2_u32(uint32_t a, uint32_t b, uint32_t c) {
    __half2 r = __hfma2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b), *reinterpret_cast<__half2*>(&c));
    return *reinterpret_cast<uint32_t*>(&r);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte async copy global -> shared; src_bytes == 0 zero-fills the destination (used for halo / tails).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
                 "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}

// D(16x8, f32) += A(16x16, f16, row) * B(16x8, f16, col)
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Programmatic dependent launch (PDL).  Every kernel of the forward chain is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: it calls pdl_trigger() first thing (the next kernel's CTAs may be
// scheduled onto SMs as this grid drains) and pdl_wait() before it touches anything a previous kernel wrote or still
// reads (griddepcontrol.wait returns once every prerequisite grid has completed and its memory is visible).  Only
// weight / constant loads and on-chip set-up sit above the wait.  Both are no-ops for a normally launched kernel.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// One lane of a fully converged warp (elect.sync).  Code that issues warp-uniform instructions (tcgen05.mma / commit, TMA) must
// be guarded by this and NOT by `lane == 0`: ptxas then emits one predicated instruction instead of an election loop.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// read-only 16-byte load that stays in L1 (neighbouring threads re-read the same pixels)
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

__device__ __forceinline__ void half8_to_float(const uint4& v, float (&f)[8]) {
    float2 a = unpack_half2(v.x), b = unpack_half2(v.y), c = unpack_half2(v.z), d = unpack_half2(v.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 float8_to_half(const float (&f)[8]) {
    uint4 v;
    v.x = pack_half2(f[0], f[1]); v.y = pack_half2(f[2], f[3]);
    v.z = pack_half2(f[4], f[5]); v.w = pack_half2(f[6], f[7]);
    return v;
}

}  // namespace mds
