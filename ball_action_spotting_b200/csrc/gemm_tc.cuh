// K3 on the 5th-gen tensor cores: persistent, warp-specialised 1x1-conv GEMM for sm_100a.
//   C[m][n] = act( sum_k A[m][k] * W[n][k] + bias[n] )            A = NHWC fp16 activations, W = [cout][cin] fp16
// These GEMMs are HBM-bound (AI < 200 FLOP/B) and, for the expand convs, output-dominated, so the design goal is
// to stream A once, write C once and keep the bias+SiLU epilogue off the critical path:
//   * a CTA owns a 128-row A tile (K <= 192 -> at most three 128x64 blocks, resident in smem, double-buffered
//     across row tiles) and sweeps ALL N tiles of it, streaming only W blocks (L2-resident) through a TMA ring;
//   * warp 0           : TMA producer (cp.async.bulk.tensor 2D, 128-byte swizzle, OOB rows / K tail zero-filled by the TMA unit,
//     mbarrier expect_tx); the whole warp runs the loops converged and one lane elected by elect.sync issues (under `lane == 0`
//     ptxas wraps every UTMALDG / UTCHMMA in an election loop of 60-90 clocks);
//   * warp 1           : TMEM allocator and MMA issuer, same scheme: tcgen05.mma (M=128, N=BN, K=16, fp16 -> fp32 in TMEM) and
//     tcgen05.commit to release smem and to publish finished accumulators;
//   * warps 2-17       : epilogue, two groups of 8 (one per accumulator): a warp streams 16-column chunks - the tcgen05.ld (and
//     residual load) of the next chunk in flight, SiLU with one reciprocal per four values, one rounding to fp16, one 32-byte
//     (full sector) st.global.v8 per chunk; no smem staging; templated on ACT / RES so that only two chunk buffers are live;
//   * the bias is added by the tensor core: one extra K=16 MMA per tile multiplies a constant "ones" A tile with a
//     [N][64] fp16 matrix holding bias as a hi/lo pair (exact to ~2^-22), so the epilogue has no loads at all;
//   * two accumulators (2*BN TMEM columns): the epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace mds {

struct TcGemmParams {
    __half* C;            // [M][N]
    const __half* res;    // [M][N] residual added in fp32 before the rounding, or nullptr
    long long M;          // total rows
    int rows_per_img;     // streamed mode: M tiles never straddle images (B differs per image)
    int tiles_per_img;
    int N, K, BN;
    int act;
    int m_tiles, n_tiles;
    int stages;           // ring depth
    int tmem_cols;        // power of two >= 2*BN
    int streamed;         // 0: A row tile resident in smem, swept over all N tiles (K <= 192)
                          // 1: A streamed with B through the ring (any K), one N tile
    long long* trace;     // -DMDS_GEMM_TRACE builds only: clock64 stamps of CTA 0 (tools/conv_trace.sh)
    const float* gate;    // streamed mode: nullptr -> B = per-image pre-gated weights (3-D map);
                          // else [n_img][K] fp32 SE gates: B = the shared weights (2-D map) and warps 10-17 multiply every A
                          // block by the image's gate in shared memory (fp32 product, one rounding) before the MMAs read it
};

#ifdef MDS_GEMM_TRACE
#define GT_TRACE(T_, ev_) do { if (p.trace && blockIdx.x == 0 && (T_) < 64 && (threadIdx.x & 31) == 0) p.trace[(T_) * 16 + (ev_)] = clock64(); } while (0)
#else
#define GT_TRACE(T_, ev_) do { } while (0)
#endif

constexpr int kTcBM = 128, kTcBK = 64, kTcMaxKB = 3;
constexpr int kTcEpiWarps = 16;
constexpr int kTcThreads = 64 + 32 * kTcEpiWarps;
constexpr int kTcABytes = kTcBM * kTcBK * 2;          // 16 KB per K block
constexpr int kTcMaxGateK = 1152;                     // largest gated K (blocks.5.x conv_pwl)

__host__ __device__ inline int tc_stages(int BN, int streamed) { return streamed ? 4 : (BN > 192 ? 3 : 4); }
__host__ __device__ inline size_t tc_smem_bytes(int BN, int streamed) {
    const size_t ring = streamed ? (size_t)tc_stages(BN, 1) * (kTcABytes + (size_t)BN * 128)
                                 : (size_t)2 * kTcMaxKB * kTcABytes + (size_t)tc_stages(BN, 0) * BN * 128;
    return 1024 /*align slack*/ + kTcABytes /*ones tile*/ + ring + 256 /*barriers*/ + (streamed ? kTcMaxGateK * 4 : 0) /*gate vector*/;
}

// ---- PTX wrappers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded spin: a protocol bug must surface as a trap (-> CUDA error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 22)) __trap();
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void ld_global_v8(const void* ptr, uint32_t (&v)[8]) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(ptr));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 (fp16 inputs, fp32 accumulate)
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzled operand tile: rows are 128 B, 8-row groups are 1024 B apart (cute::UMMA::SmemDescriptor:
// start>>4 [0,14), LBO>>4 [16,30) (unused for swizzled K-major, 1), SBO>>4 [32,46) = 64, version [46,48) = 1,
// layout [61,64) = 2 (SWIZZLE_128B)).
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6), a/b_format F16 (0), K-major A and B, N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t tc_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void st_global_v8(void* ptr, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

template <bool ACT, bool RES>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmBias, TcGemmParams p) {
    extern __shared__ unsigned char tc_smem_raw[];
    // 1024-byte alignment is required by the 128-byte swizzle atoms
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_bytes = p.BN * 128;
    unsigned char* s_ones = smem;                                        // 128x64 tile: columns 0,1 = 1.0, rest 0
    unsigned char* s_a = s_ones + kTcABytes;                             // resident mode: [2][kTcMaxKB][16 KB]
    unsigned char* s_ring = p.streamed ? s_ones + kTcABytes : s_a + (size_t)2 * kTcMaxKB * kTcABytes;
    const int stage_bytes = p.streamed ? kTcABytes + b_bytes : b_bytes; // streamed: [A block | W block], resident: [W block]
    const int b_off = p.streamed ? kTcABytes : 0;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + (size_t)p.stages * stage_bytes);
    uint64_t* b_full = bars;                  // [4]
    uint64_t* b_empty = bars + 4;             // [4]
    uint64_t* a_full = bars + 8;              // [2]
    uint64_t* a_empty = bars + 10;            // [2]
    uint64_t* acc_full = bars + 12;           // [2]
    uint64_t* acc_empty = bars + 14;          // [2]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 16);
    uint64_t* a_gated = bars + 17;            // [4]  gated mode: A block of the stage has been multiplied by the gate
    uint32_t* s_gate_hi = reinterpret_cast<uint32_t*>(bars + 32);      // [kTcMaxGateK / 2] half2 (streamed mode)
    uint32_t* s_gate_lo = s_gate_hi + kTcMaxGateK / 2;
    const bool gated = p.streamed && p.gate != nullptr;

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (p.K + kTcBK - 1) / kTcBK;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); mbar_init(&a_gated[i], kTcEpiWarps / 2); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1);
            mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], gated ? kTcEpiWarps / 4 : kTcEpiWarps / 2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBias) : "memory");
    }
    // ones tile in the 128-byte-swizzled K-major layout: element (r, 0..7) lives in 16-byte chunk (0 ^ (r & 7)) of row r
    for (int i = threadIdx.x; i < kTcABytes / 16; i += kTcThreads) {
        const int r = i >> 3, ch = i & 7;
        reinterpret_cast<uint4*>(s_ones)[i] = (ch == (r & 7)) ? make_uint4(0x3C003C00u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    }
    if (gated)
        for (int i = threadIdx.x; i < kTcMaxGateK; i += kTcThreads) s_gate_hi[i] = 0u;      // (hi and lo) entries beyond K stay zero
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to tcgen05.mma
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    pdl_wait();       // barriers, ones tile and TMEM are set up; every global access below depends on earlier kernels

    if (warp == 0) {
        // ================= TMA producer =================
        // The whole warp runs the loops converged and one elected lane (elect.sync) issues: under `if (lane == 0)` ptxas wraps
        // every warp-uniform instruction (UTMALDG, UTCHMMA, UTCBAR) in an election loop, 60-90 clocks each (tools/mma_probe.cu).
        {
            int stage = 0;
            uint32_t phase = 0;
            if (p.streamed) {
                for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
                    const int img = mt / p.tiles_per_img, ti = mt - img * p.tiles_per_img;
                    const int row0 = img * p.rows_per_img + ti * kTcBM;
                    for (int kb = 0; kb <= num_kb; ++kb) {              // block num_kb = the bias block
                        mbar_wait(&b_empty[stage], phase ^ 1);
                        unsigned char* st = s_ring + (size_t)stage * stage_bytes;
                        if (elect_one()) {
                            if (kb < num_kb) {
                                mbar_expect_tx(&b_full[stage], (uint32_t)(kTcABytes + b_bytes));
                                tma_load_2d(st, &tmA, &b_full[stage], kb * kTcBK, row0);
                                if (gated) tma_load_2d(st + b_off, &tmB, &b_full[stage], kb * kTcBK, 0);
                                else tma_load_3d(st + b_off, &tmB, &b_full[stage], kb * kTcBK, 0, img);
                            } else {
                                mbar_expect_tx(&b_full[stage], (uint32_t)b_bytes);
                                tma_load_2d(st + b_off, &tmBias, &b_full[stage], 0, 0);
                            }
                        }
                        __syncwarp();
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            } else {
                auto load_a = [&](int i, int mt) {      // A row tile i of this CTA -> buffer i & 1
                    const int ab = i & 1;
                    mbar_wait(&a_empty[ab], (((uint32_t)i >> 1) & 1) ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&a_full[ab], (uint32_t)(num_kb * kTcABytes));
                        for (int kb = 0; kb < num_kb; ++kb)
                            tma_load_2d(s_a + (size_t)(ab * kTcMaxKB + kb) * kTcABytes, &tmA, &a_full[ab], kb * kTcBK, mt * kTcBM);
                    }
                    __syncwarp();
                };
                int i = 0;
                if ((int)blockIdx.x < p.m_tiles) load_a(0, blockIdx.x);
                for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x, ++i) {
                    // the next row tile's A is requested a whole tile ahead, so its HBM latency hides behind this tile
                    if (mt + (int)gridDim.x < p.m_tiles) load_a(i + 1, mt + gridDim.x);
                    for (int nt = 0; nt < p.n_tiles; ++nt)
                        for (int kb = 0; kb <= num_kb; ++kb) {              // block num_kb = the bias block
                            mbar_wait(&b_empty[stage], phase ^ 1);
                            if (elect_one()) {
                                mbar_expect_tx(&b_full[stage], (uint32_t)b_bytes);
                                if (kb < num_kb) tma_load_2d(s_ring + (size_t)stage * stage_bytes, &tmB, &b_full[stage], kb * kTcBK, nt * p.BN);
                                else tma_load_2d(s_ring + (size_t)stage * stage_bytes, &tmBias, &b_full[stage], 0, nt * p.BN);
                            }
                            __syncwarp();
                            if (++stage == p.stages) { stage = 0; phase ^= 1; }
                        }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (whole warp converged, one elected lane issues) =================
        {
            const uint32_t idesc = tc_idesc(kTcBM, p.BN);
            int stage = 0;
            uint32_t phase = 0;
            int i = 0, t = 0;
            for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x, ++i) {
                const int ab = i & 1;
                if (!p.streamed) mbar_wait(&a_full[ab], ((uint32_t)i >> 1) & 1);
                for (int nt = 0; nt < p.n_tiles; ++nt, ++t) {
                    const int acc = t & 1;
                    GT_TRACE(t, 0);
                    mbar_wait(&acc_empty[acc], (((uint32_t)t >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
                    tc_fence_after();
                    GT_TRACE(t, 1);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
                    for (int kb = 0; kb <= num_kb; ++kb) {
                        mbar_wait(gated ? &a_gated[stage] : &b_full[stage], phase);
                        tc_fence_after();
                        if (kb < 4) GT_TRACE(t, 2 + kb);
                        const bool bias_blk = kb == num_kb;
                        unsigned char* st = s_ring + (size_t)stage * stage_bytes;
                        const unsigned char* a_src = bias_blk ? s_ones : (p.streamed ? st : s_a + (size_t)(ab * kTcMaxKB + kb) * kTcABytes);
                        const uint64_t adesc = tc_smem_desc(smem_u32(a_src));
                        const uint64_t bdesc = tc_smem_desc(smem_u32(st + b_off));
                        if (elect_one()) {
                            if (bias_blk) {                  // the bias block only has K = 16 worth of data
                                tc_mma_f16(d_tmem, adesc, bdesc, idesc, kb != 0);
                            } else {
#pragma unroll
                                for (int ki = 0; ki < kTcBK / 16; ++ki) {     // advance 16 K-elements = 32 B inside the swizzle atom (>>4 = 2)
                                    const int k = (MDS_NUMERICS_VARIANT & 4) ? kTcBK / 16 - 1 - ki : ki;
                                    tc_mma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | ki) != 0);
                                }
                            }
                            tc_commit(&b_empty[stage]);                  // stage reusable once these MMAs have read it
                            if (bias_blk) tc_commit(&acc_full[acc]);     // accumulator complete
                        }
                        __syncwarp();
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                    GT_TRACE(t, 6);
                }
                if (!p.streamed) {
                    if (elect_one()) tc_commit(&a_empty[ab]);            // every MMA that reads this A tile has been issued
                    __syncwarp();
                }
            }
        }
    } else if (gated && warp >= 2 + kTcEpiWarps / 2) {
        // ================= gaters (warps 10..17, gated mode): A block *= gate[img], in place =================
        // 128 rows x 8 chunks of 16 B, 128-byte swizzle: thread -> physical chunk pc = gt & 7 of rows (gt >> 3) + 32 i; the
        // logical chunk j = pc ^ (row & 7) is the same for all four rows (32 i keeps row & 7)
        const int gt = threadIdx.x - 32 * (2 + kTcEpiWarps / 2);       // 0..255
        int stage = 0;
        uint32_t phase = 0;
        int gate_img = -1;
        for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
            const int img = mt / p.tiles_per_img;
            asm volatile("bar.sync 1, 256;" ::: "memory");             // all gaters are done with the previous tile's gate vector
            if (img != gate_img) {
                // the gate is split into an fp16 (hi, lo) pair so that a * g = a * hi + a * lo runs on the packed fp16 pipe
                // (HMUL2 + HFMA2 per two channels) and still carries the gate to ~22 bits
                gate_img = img;
                const float* gsrc = p.gate + (size_t)img * p.K;
                for (int i = gt; i < (p.K >> 1); i += 256) {
                    const float2 g = __ldg(reinterpret_cast<const float2*>(gsrc) + i);
                    const __half2 hi = __floats2half2_rn(g.x, g.y);
                    const float2 hf = __half22float2(hi);
                    const __half2 lo = __floats2half2_rn(g.x - hf.x, g.y - hf.y);
                    s_gate_hi[i] = *reinterpret_cast<const uint32_t*>(&hi);
                    s_gate_lo[i] = *reinterpret_cast<const uint32_t*>(&lo);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            for (int kb = 0; kb <= num_kb; ++kb) {       // block num_kb = bias block: nothing to gate, but every use of a stage
                                                         // completes one phase of each of its barriers
                mbar_wait(&b_full[stage], phase);
                if (kb < num_kb) {
                    unsigned char* a_blk = s_ring + (size_t)stage * stage_bytes;
                    const int pc = gt & 7, r0 = gt >> 3;
                    const int j = pc ^ (r0 & 7);
                    const uint4 gh = *reinterpret_cast<const uint4*>(s_gate_hi + kb * (kTcBK / 2) + j * 4);
                    const uint4 gl = *reinterpret_cast<const uint4*>(s_gate_lo + kb * (kTcBK / 2) + j * 4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4* q = reinterpret_cast<uint4*>(a_blk + (size_t)(r0 + 32 * i) * 128 + pc * 16);
                        uint4 v = *q;
                        v.x = hfma2_u32(v.x, gl.x, hmul2_u32(v.x, gh.x));
                        v.y = hfma2_u32(v.y, gl.y, hmul2_u32(v.y, gh.y));
                        v.z = hfma2_u32(v.z, gl.z, hmul2_u32(v.z, gh.z));
                        v.w = hfma2_u32(v.w, gl.w, hmul2_u32(v.w, gh.w));
                        *q = v;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_gated[stage]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ================= epilogue (warps 2..17; gated mode: warps 2..9) =================
        // Two groups of 8 warps; group e drains accumulator e (tiles t with t % 2 == e).  TMEM lane quadrant = warp % 4; the two
        // warps of a group that share a quadrant take alternate 16-column chunks.  A warp streams its chunks: the tcgen05.ld (and
        // the residual load) of chunk k+1 is in flight while chunk k goes through bias-free SiLU (four values per reciprocal),
        // one rounding to fp16 and one 32-byte store.  Only two 16-register buffers are live, so the compiler can interleave the
        // 16 independent SiLU chains of a chunk (with three chunks in registers and the 96-register cap of a 576-thread CTA it
        // serialised them: ~100 clocks per element, measured with clock64 stamps, tools/conv_trace.sh).
        const int q = warp & 3;
        const int e = ((warp - 2) >> 2) & 1;
        const int half = (warp - 2) >> 3;              // 0 in gated mode (8 epilogue warps: one per quadrant and accumulator)
        const int nhalf = gated ? 1 : 2;
        const int ngroups = p.BN >> 4;
        const int nk = (ngroups - half + nhalf - 1) / nhalf;      // chunks of this warp per tile
        int t = 0;
        for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
            long long row;
            bool row_ok;
            if (p.streamed) {
                const int img = mt / p.tiles_per_img, ti = mt - img * p.tiles_per_img;
                const int r_in = ti * kTcBM + q * 32 + lane;
                row_ok = r_in < p.rows_per_img;
                row = (long long)img * p.rows_per_img + r_in;
            } else {
                row = (long long)mt * kTcBM + q * 32 + lane;
                row_ok = row < p.M;
            }
            for (int nt = 0; nt < p.n_tiles; ++nt, ++t) {
                if ((t & 1) != e) continue;
                __half* c_ptr = p.C + row * p.N + nt * p.BN + half * 16;
                const __half* r_ptr = p.res + row * p.N + nt * p.BN + half * 16;
                const uint32_t t_col = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(e * p.BN + half * 16);
                uint32_t va[16], vb[16], ra[8], rb[8];
                // chunk k: columns [16 * (half + nhalf * k), + 16) of the tile
                auto request = [&](int k, uint32_t (&v)[16], uint32_t (&r)[8]) {
                    if constexpr (RES) {
                        if (row_ok) ld_global_v8(r_ptr + k * nhalf * 16, r);
                    }
                    tc_ld16(t_col + (uint32_t)(k * nhalf * 16), v);
                };
                auto consume = [&](int k, uint32_t (&v)[16], uint32_t (&r)[8]) {
                    uint32_t pk[8];
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        float x[4] = {__uint_as_float(v[4 * h]), __uint_as_float(v[4 * h + 1]), __uint_as_float(v[4 * h + 2]),
                                      __uint_as_float(v[4 * h + 3])};
                        if constexpr (ACT) silu4(x);
                        if constexpr (RES) {
                            const float2 r01 = unpack_half2(r[2 * h]), r23 = unpack_half2(r[2 * h + 1]);
                            x[0] += r01.x; x[1] += r01.y; x[2] += r23.x; x[3] += r23.y;
                        }
                        pk[2 * h] = pack_half2(x[0], x[1]);
                        pk[2 * h + 1] = pack_half2(x[2], x[3]);
                    }
                    if (row_ok) st_global_v8(c_ptr + k * nhalf * 16, pk);
                };
                auto release = [&]() {       // last TMEM read of this tile is in registers: hand the accumulator back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[e]);
                };
                if (warp == 2 || warp == 6) GT_TRACE(t, 8);
                mbar_wait(&acc_full[e], ((uint32_t)t >> 1) & 1);
                tc_fence_after();
                if (warp == 2 || warp == 6) GT_TRACE(t, 9);
                request(0, va, ra);
                tc_wait_ld();
                if (warp == 2 || warp == 6) GT_TRACE(t, 10);
                if (nk == 1) release();
                for (int k = 0; k < nk; k += 2) {
                    if (k + 1 < nk) request(k + 1, vb, rb);
                    consume(k, va, ra);
                    if (k + 1 < nk) {
                        tc_wait_ld();
                        if (k + 2 == nk) release();
                        if (k + 2 < nk) request(k + 2, va, ra);
                        consume(k + 1, vb, rb);
                        if (k + 2 < nk) {
                            tc_wait_ld();
                            if (k + 3 == nk) release();
                        }
                    }
                }
                if (warp == 2 || warp == 6) GT_TRACE(t, 12);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

}  // namespace mds
