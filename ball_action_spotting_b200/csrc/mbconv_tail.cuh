// MBConv tail in ONE launch: depthwise 3x3 (2D, stride 1/2, TF-SAME) or 3x3x3 (3D) + folded BN + SiLU, SE squeeze,
// SE excitation MLP, SE gating and the 1x1 projection GEMM (+ folded BN, + shortcut)
// (timm InvertedResidual conv_dw/bn2/se/conv_pwl/bn3 built at multidim_stacker.py:166-176; InvertedResidual3d
// multidim_stacker.py:110-134; SqueezeExcite :72-90).
//
// The SE squeeze is a reduction over a whole image (2D) / stack (3D), so the projection of an image can only start when
// its depthwise pass is complete.  Instead of three kernels separated by grid-wide barriers this is one persistent
// kernel (one CTA per SM) in which every CTA runs TWO concurrent streams of work items on disjoint warps and disjoint
// shared-memory rings, coupled only by per-image flags in global memory:
//   * depthwise stream (FMA / SFU pipes): item = (image, [plane t], row chunk, 40 output columns, 128-channel slab).
//     Input rows are staged by TMA (cp.async.bulk.tensor, 4-D / 5-D map over the NHWC tensor, hardware zero fill = the
//     conv padding) into an mbarrier ring; 16 compute warps (two 64-channel groups x 8 warps x 5 columns, lane = 2
//     channels as one f32x2, packed FFMA2 / FMUL2 / FADD2) write the SiLU'd fp16 output and leave one squeeze partial per
//     (image, part, channel).  The CTA that completes the LAST item of an image (device-scope counter) evaluates the SE
//     MLP for that image from the partials in a fixed order (deterministic, independent of the batch), writes the fp32
//     gate vector and releases the image's flag;
//   * projection stream (TMA / tensor pipe): item = (image, 128-row tile).  The TMA producer acquires the image's flag,
//     then streams [A | W] K-blocks; 4 "gater" warps multiply the A block by the gate in shared memory (fp32 product, one
//     rounding), the tcgen05 issuer accumulates in TMEM (bias through a ones-tile MMA), 4 epilogue warps add the shortcut
//     and store.
// Both streams walk their items image-major (item blockIdx.x, blockIdx.x + grid, ...), so the projection of an image
// follows its depthwise pass closely enough for the depthwise output to still be in L2.  The depthwise stream never
// waits for anything but its own ring and all CTAs are co-resident, so the flags are always released: no deadlock.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "tc_helpers.cuh"
#include "gemm_tc.cuh"

namespace mds {

struct TailParams {
    // ---- depthwise ----
    __half* m2;              // [n][T][Ho][Wo][C] depthwise output (fp16, before gating) = A operand of the projection
    const float* dw_w;       // [taps][C]  tap = (dt*3 + r)*3 + s, BN scale folded
    const float* dw_b;       // [C]
    float* partials;         // [n][nparts][C] squeeze partial sums (fp32)
    // ---- SE ----
    const float* se_w1;      // [rd][C]
    const float* se_b1;      // [rd]
    const float* se_w2t;     // [rd][C]
    const float* se_b2;      // [C]
    float* gate;             // [n][C] fp32
    int* sync;               // [3][sync_stride]: dw_done, flag, gemm_done; all zero between launches
    int sync_stride;
    // ---- projection ----
    __half* out;             // [n * rows_per_img][N]
    const __half* res;       // same shape or nullptr
    int n, T, H, W, C, Ho, Wo;
    int rd;
    float inv_count;         // 1 / (T * Ho * Wo)
    int rows_per_chunk, chunks, xtiles, slab_pairs, nparts;
    int dw_per_img;          // dw items per image = T * chunks * xtiles * slab_pairs
    int rows_per_img, tiles_per_img, N, num_kb;      // tiles_per_img = 0: depthwise + SE only
    int tmem_cols;
    int g_stages;            // projection ring depth (2..4)
};

// warps 0-3: service warpgroup (0 = depthwise TMA producer, 1 = tcgen05 issuer, 2 = projection TMA producer, 3 idle) that
// hands most of its registers to the others through setmaxnreg; warps 4-19: depthwise; warps 20-23: epilogue; 24-27: gaters
constexpr int kTlDwWarps = 16, kTlGemmWarps = 8;
constexpr int kTlThreads = 128 + 32 * (kTlDwWarps + kTlGemmWarps);
constexpr int kTlServiceRegs = 24, kTlDwRegs = 80, kTlGemmRegs = 80;
constexpr int kTlPXW = 5, kTlTWX = 40, kTlCS = 128;      // columns per warp / per item, channels per item
constexpr int kTlMaxC = 1152, kTlMaxRd = 64;
constexpr int kTlGP = 2;               // 16-column accumulator groups an epilogue warp holds in registers at once
constexpr int kTlOnesBytes = 2 * 128 * 16;      // ones tile, no-swizzle K-major: [2 planes][128 rows][8 halves]

template <int KT, int STRIDE>
struct TailCfg {
    static constexpr int IW = (STRIDE == 1) ? kTlTWX + 2 : 2 * kTlTWX + 1;     // input columns per row tile
    static constexpr int NV = (STRIDE == 1) ? kTlPXW + 2 : 2 * kTlPXW + 1;
    static constexpr int RPS = (KT == 1 && STRIDE == 1) ? 2 : 1;               // input rows per stage
    static constexpr int ROWB = IW * kTlCS * 2;                                // bytes of one input row tile (one plane)
    static constexpr int STAGEB = KT * RPS * ROWB;
    static constexpr int NST = (KT == 3) ? 3 : 4;
    static constexpr int DW_RING = NST * STAGEB;
};

__host__ __device__ inline int tl_gemm_stage_bytes(int N) { return kTcABytes + N * 128; }
// shared memory map (after 1024-byte alignment):
//   projection ring [g_stages][A 16 KB | W N*128] | depthwise ring | ones tile | s_gate[kTlMaxC] | s_mean[kTlMaxC] |
//   s_hid[kTlMaxRd] | s_part[16][64] | s_w3[27][128] (3D only) | barriers
__host__ __device__ inline size_t tl_smem_bytes(int dw_ring, int N, int g_stages, int kt) {
    return 1024 + (size_t)g_stages * tl_gemm_stage_bytes(N) + dw_ring + kTlOnesBytes +
           (size_t)(2 * kTlMaxC + kTlMaxRd + 16 * 64) * 4 + (kt == 3 ? 27 * kTlCS * 4 : 0) + 512;
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_cg_f32(const float* p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t saddr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr));
    return v;
}
__device__ __forceinline__ float2 tl_ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 tl_lds_half2(uint32_t saddr) {
    uint32_t u;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(saddr));
    return unpack_half2(u);
}
// SiLU of two channels with the packed fp32 pipe: 1 FMUL2 + 2 MUFU.EX2 + 1 FADD2 + 2 MUFU.RCP + 1 FMUL2
__device__ __forceinline__ float2 silu2(float2 x) {
    const float2 t = __fmul2_rn(x, make_float2(-1.4426950408889634f, -1.4426950408889634f));
    float2 e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(t.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(t.y));
    const float2 d = __fadd2_rn(e, make_float2(1.0f, 1.0f));
    float2 r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(d.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(d.y));
    return __fmul2_rn(x, r);
}

template <int KT, int STRIDE>
__global__ void __launch_bounds__(kTlThreads, 1)
mbconv_tail_kernel(const __grid_constant__ CUtensorMap tmDw, const __grid_constant__ CUtensorMap tmA,
                   const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBias, TailParams p) {
    using Cfg = TailCfg<KT, STRIDE>;
    extern __shared__ unsigned char tl_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tl_smem_raw) + 1023) & ~uintptr_t(1023));
    const int g_stage_bytes = tl_gemm_stage_bytes(p.N);
    unsigned char* s_gring = smem;                                                     // 1024-byte aligned stages (128-byte swizzle)
    unsigned char* s_dring = s_gring + (size_t)p.g_stages * g_stage_bytes;
    unsigned char* s_ones = s_dring + Cfg::DW_RING;
    float* s_gate = reinterpret_cast<float*>(s_ones + kTlOnesBytes);                   // [kTlMaxC]
    float* s_mean = s_gate + kTlMaxC;
    float* s_hid = s_mean + kTlMaxC;
    float* s_part = s_hid + kTlMaxRd;                                                  // [16][64]
    float* s_w3 = s_part + 16 * 64;                                                    // [27][128] (KT == 3)
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_w3 + (KT == 3 ? 27 * kTlCS : 0));
    uint64_t* d_full = bars;              // [4]
    uint64_t* d_empty = bars + 4;         // [4]
    uint64_t* g_full = bars + 8;          // [4]
    uint64_t* g_gated = bars + 12;        // [4]
    uint64_t* g_empty = bars + 16;        // [4]
    uint64_t* acc_full = bars + 20;       // [2]
    uint64_t* acc_empty = bars + 22;      // [2]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 24);
    int* s_last = reinterpret_cast<int*>(bars + 25);

    pdl_trigger();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int* dw_done = p.sync;
    int* flag = p.sync + p.sync_stride;
    int* gemm_done = p.sync + 2 * p.sync_stride;
    const int n_dw = p.n * p.dw_per_img, n_tiles = p.n * p.tiles_per_img;

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) {
            mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], kTlDwWarps);
            mbar_init(&g_full[i], 1); mbar_init(&g_gated[i], 4); mbar_init(&g_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDw) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBias) : "memory");
    }
    for (int i = tid; i < 2 * 128; i += kTlThreads)                   // ones tile: (row, k = 0, 1) = 1.0 (bias hi + lo columns)
        reinterpret_cast<uint4*>(s_ones)[i] = (i < 128) ? make_uint4(0x3C003C00u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < kTlMaxC; i += kTlThreads) s_gate[i] = 0.f;  // entries beyond C stay zero (K tail of the last block)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1 && n_tiles > 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    pdl_wait();

    // geometry of a depthwise item
    auto dw_geom = [&](int idx, int& n, int& t, int& chunk, int& xt, int& sp) {
        n = idx / p.dw_per_img;
        int loc = idx - n * p.dw_per_img;
        sp = loc % p.slab_pairs; loc /= p.slab_pairs;
        xt = loc % p.xtiles; loc /= p.xtiles;
        chunk = loc % p.chunks; t = loc / p.chunks;
    };

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kTlServiceRegs));
    if (warp == 0) {
        // =========================================== depthwise TMA producer ===========================================
        if (lane == 0) {
            uint32_t dcnt = 0;
            for (int idx = blockIdx.x; idx < n_dw; idx += gridDim.x) {
                int n, t, chunk, xt, sp;
                dw_geom(idx, n, t, chunk, xt, sp);
                const int yo0 = chunk * p.rows_per_chunk, yo1 = min(p.Ho, yo0 + p.rows_per_chunk);
                const int yi0 = (STRIDE == 1) ? yo0 - 1 : 2 * yo0;
                const int NR = (STRIDE == 1) ? (yo1 - yo0) + 2 : 2 * (yo1 - yo0) + 1;
                const int xi0 = (STRIDE == 1) ? xt * kTlTWX - 1 : 2 * xt * kTlTWX;
                const int nstg = (NR + Cfg::RPS - 1) / Cfg::RPS;
                for (int js = 0; js < nstg; ++js, ++dcnt) {
                    const int st = dcnt % Cfg::NST;
                    mbar_wait(&d_empty[st], ((dcnt / Cfg::NST) & 1) ^ 1);
                    mbar_expect_tx(&d_full[st], (uint32_t)Cfg::STAGEB);
                    if constexpr (KT == 3)
                        tma_load_5d(s_dring + (size_t)st * Cfg::STAGEB, &tmDw, &d_full[st], sp * kTlCS, xi0, yi0 + js, t - 1, n);
                    else
                        tma_load_4d(s_dring + (size_t)st * Cfg::STAGEB, &tmDw, &d_full[st], sp * kTlCS, xi0, yi0 + js * Cfg::RPS, n);
                }
            }
        }
    } else if (warp == 2) {
        // =========================================== projection TMA producer ===========================================
        if (lane == 0) {
            uint32_t gcnt = 0;
            int seen_img = -1;
            for (int idx = blockIdx.x; idx < n_tiles; idx += gridDim.x) {
                const int img = idx / p.tiles_per_img, ti = idx - img * p.tiles_per_img;
                if (img != seen_img) {     // the image's depthwise output and gate are complete (flag released by the SE CTA)
                    uint32_t spins = 0;
                    while (ld_acquire_gpu(&flag[img]) == 0) {
                        __nanosleep(64);
                        if (++spins > (1u << 24)) __trap();
                    }
                    seen_img = img;
                    asm volatile("fence.proxy.async;" ::: "memory");      // generic-proxy global writes -> async-proxy (TMA) reads
                }
                const int row0 = img * p.rows_per_img + ti * kTcBM;
                for (int kb = 0; kb <= p.num_kb; ++kb, ++gcnt) {
                    const int st = gcnt % p.g_stages;
                    mbar_wait(&g_empty[st], ((gcnt / p.g_stages) & 1) ^ 1);
                    unsigned char* dst = s_gring + (size_t)st * g_stage_bytes;
                    if (kb < p.num_kb) {
                        mbar_expect_tx(&g_full[st], (uint32_t)g_stage_bytes);
                        tma_load_2d(dst, &tmA, &g_full[st], kb * kTcBK, row0);
                        tma_load_2d(dst + kTcABytes, &tmB, &g_full[st], kb * kTcBK, 0);
                    } else {               // bias block: W part only
                        mbar_expect_tx(&g_full[st], (uint32_t)(p.N * 128));
                        tma_load_2d(dst + kTcABytes, &tmBias, &g_full[st], 0, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =========================================== MMA issuer ===========================================
        if (lane == 0) {
            const uint32_t idesc = tc_idesc(kTcBM, p.N);
            const uint64_t ones_desc = tc_desc_nosw(smem_u32(s_ones), 128 * 16);
            uint32_t gcnt = 0;
            int t = 0;
            for (int idx = blockIdx.x; idx < n_tiles; idx += gridDim.x, ++t) {
                const int acc = t & 1;
                mbar_wait(&acc_empty[acc], (((uint32_t)t >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.N);
                for (int kb = 0; kb <= p.num_kb; ++kb, ++gcnt) {
                    const int st = gcnt % p.g_stages;
                    mbar_wait(&g_gated[st], (gcnt / p.g_stages) & 1);
                    tc_fence_after();
                    unsigned char* sp = s_gring + (size_t)st * g_stage_bytes;
                    const uint64_t bdesc = tc_smem_desc(smem_u32(sp + kTcABytes));
                    if (kb < p.num_kb) {
                        const uint64_t adesc = tc_smem_desc(smem_u32(sp));
#pragma unroll
                        for (int k = 0; k < kTcBK / 16; ++k)
                            tc_mma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                    } else {
                        tc_mma_f16(d_tmem, ones_desc, bdesc, idesc, 1);       // + bias (hi + lo columns of the bias matrix)
                    }
                    tc_commit(&g_empty[st]);
                }
                tc_commit(&acc_full[acc]);
            }
        }
    }
    } else if (warp < 4 + kTlDwWarps) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTlDwRegs));
        // ================================== depthwise warps: conv + SiLU + squeeze (+ SE MLP) ==================================
        const int cw = warp - 4;                    // 0..15
        const int ctid = tid - 128;                 // 0..511
        const int grp = cw >> 3, wi = cw & 7;       // channel group (64 ch) and column strip
        uint32_t dcnt = 0;
        const uint32_t ring_u32 = smem_u32(s_dring);
        const uint32_t lds_lane = (uint32_t)(((STRIDE == 1) ? wi * kTlPXW : 2 * wi * kTlPXW) * kTlCS + grp * 64 + 2 * lane) * 2u;
        const uint32_t w3_u32 = smem_u32(s_w3) + (uint32_t)(grp * 64 + 2 * lane) * 4u;

        for (int idx = blockIdx.x; idx < n_dw; idx += gridDim.x) {
            int n, t, chunk, xt, sp;
            dw_geom(idx, n, t, chunk, xt, sp);
            const int yo0 = chunk * p.rows_per_chunk, yo1 = min(p.Ho, yo0 + p.rows_per_chunk);
            const int c = sp * kTlCS + grp * 64 + 2 * lane;
            const bool c_ok = c < p.C;
            const int c_ld = c_ok ? c : 0;
            const int NR = (STRIDE == 1) ? (yo1 - yo0) + 2 : 2 * (yo1 - yo0) + 1;
            const int nstg = (NR + Cfg::RPS - 1) / Cfg::RPS;

            float2 w[9];
            if constexpr (KT == 1) {
#pragma unroll
                for (int i = 0; i < 9; ++i) w[i] = __ldg(reinterpret_cast<const float2*>(p.dw_w + (size_t)i * p.C + c_ld));
            } else {
                // stage this item's 27 x 128 weights in shared memory (3 x 27 float2 do not fit the register budget)
                named_bar_sync(1, 512);            // the previous item's readers are done
                for (int i = ctid; i < 27 * (kTlCS / 2); i += 512) {
                    const int tap = i / (kTlCS / 2), cc = (i - tap * (kTlCS / 2)) * 2;
                    const int cg = sp * kTlCS + cc;
                    float2 v = make_float2(0.f, 0.f);
                    if (cg < p.C) v = __ldg(reinterpret_cast<const float2*>(p.dw_w + (size_t)tap * p.C + cg));
                    *reinterpret_cast<float2*>(s_w3 + tap * kTlCS + cc) = v;
                }
                named_bar_sync(1, 512);
            }
            const float2 bias = __ldg(reinterpret_cast<const float2*>(p.dw_b + c_ld));

            float2 lsum = make_float2(0.f, 0.f);
            float2 a0[kTlPXW], a1[kTlPXW], a2[kTlPXW];      // accumulators start from the bias
#pragma unroll
            for (int j = 0; j < kTlPXW; ++j) a0[j] = a1[j] = a2[j] = bias;
            const int xw = xt * kTlTWX + wi * kTlPXW;
            const long long out_pitch = (long long)p.Wo * p.C;
            __half* o_ptr = p.m2 + (((size_t)n * p.T + t) * p.Ho + yo0) * (size_t)out_pitch + (long long)xw * p.C + c;
            const int npx = c_ok ? min(kTlPXW, p.Wo - xw) : 0;       // valid output columns of this lane (may be <= 0)

            auto emit = [&](const float2 (&acc)[kTlPXW]) {
#pragma unroll
                for (int j = 0; j < kTlPXW; ++j) {
                    if (j < npx) {
                        const float2 o = silu2(acc[j]);
                        lsum = __fadd2_rn(lsum, o);
                        *reinterpret_cast<uint32_t*>(o_ptr + j * p.C) = pack_half2(o.x, o.y);
                    }
                }
                o_ptr += out_pitch;
            };

            int k = 0;                              // input row counter of this item
            for (int js = 0; js < nstg; ++js, ++dcnt) {
                const int st = dcnt % Cfg::NST;
                mbar_wait(&d_full[st], (dcnt / Cfg::NST) & 1);
                const uint32_t sbase = ring_u32 + (uint32_t)st * Cfg::STAGEB + lds_lane;
#pragma unroll
                for (int rr = 0; rr < Cfg::RPS; ++rr, ++k) {
                    if (k >= NR) break;
                    const uint32_t srow = sbase + rr * Cfg::ROWB;
                    if constexpr (STRIDE == 1) {
                        // input row yi = yo0 - 1 + k feeds out rows yi + 1 (kernel row 0), yi (row 1), yi - 1 (row 2)
#pragma unroll
                        for (int dt = 0; dt < KT; ++dt) {
                            if constexpr (KT == 3) {
#pragma unroll
                                for (int i = 0; i < 9; ++i) w[i] = lds_f32x2(w3_u32 + (uint32_t)((dt * 9 + i) * kTlCS) * 4u);
                            }
                            float2 v[Cfg::NV];
#pragma unroll
                            for (int i = 0; i < Cfg::NV; ++i) v[i] = tl_lds_half2(srow + (uint32_t)(dt * Cfg::ROWB + i * kTlCS * 2));
#pragma unroll
                            for (int j = 0; j < kTlPXW; ++j)
#pragma unroll
                                for (int s = 0; s < 3; ++s) {
                                    a2[j] = tl_ffma2(w[0 + s], v[j + s], a2[j]);
                                    a1[j] = tl_ffma2(w[3 + s], v[j + s], a1[j]);
                                    a0[j] = tl_ffma2(w[6 + s], v[j + s], a0[j]);
                                }
                        }
                        if (k >= 2) emit(a0);       // out row yo0 + k - 2
#pragma unroll
                        for (int j = 0; j < kTlPXW; ++j) { a0[j] = a1[j]; a1[j] = a2[j]; a2[j] = bias; }
                    } else {
                        // stride 2: even input row 2yo is kernel row 0 of out yo and kernel row 2 of out yo-1; odd row 2yo+1 is row 1
                        float2 v[Cfg::NV];
#pragma unroll
                        for (int i = 0; i < Cfg::NV; ++i) v[i] = tl_lds_half2(srow + (uint32_t)(i * kTlCS * 2));
                        if ((k & 1) == 0) {
#pragma unroll
                            for (int j = 0; j < kTlPXW; ++j) {
                                a2[j] = bias;
#pragma unroll
                                for (int s = 0; s < 3; ++s) {
                                    a1[j] = tl_ffma2(w[6 + s], v[2 * j + s], a1[j]);     // closes out row k/2 - 1
                                    a2[j] = tl_ffma2(w[0 + s], v[2 * j + s], a2[j]);     // opens out row k/2
                                }
                            }
                            if (k > 0) emit(a1);
#pragma unroll
                            for (int j = 0; j < kTlPXW; ++j) a1[j] = a2[j];
                        } else {
#pragma unroll
                            for (int j = 0; j < kTlPXW; ++j)
#pragma unroll
                                for (int s = 0; s < 3; ++s) a1[j] = tl_ffma2(w[3 + s], v[2 * j + s], a1[j]);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[st]);
            }

            // ---- squeeze partial of this item: fixed-order sum of the 8 column strips, one store per channel ----
            s_part[cw * 64 + 2 * lane] = lsum.x;
            s_part[cw * 64 + 2 * lane + 1] = lsum.y;
            named_bar_sync(1, 512);
            if (ctid < kTlCS) {
                const int g2 = ctid >> 6, cc = ctid & 63;
                const int cg = sp * kTlCS + ctid;
                if (cg < p.C) {
                    float s = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) s += s_part[(g2 * 8 + i) * 64 + cc];
                    const int part = (t * p.chunks + chunk) * p.xtiles + xt;
                    p.partials[((size_t)n * p.nparts + part) * p.C + cg] = s;
                }
            }
            __threadfence();                 // this thread's m2 / partial stores are visible device-wide before the count below
            named_bar_sync(1, 512);
            if (ctid == 0) {
                const int old = atomicAdd(&dw_done[n], 1);
                *s_last = (old == p.dw_per_img - 1) ? 1 : 0;
            }
            named_bar_sync(1, 512);
            if (*s_last) {
                // ---- this CTA completed the image: SE excitation MLP (multidim_stacker.py:86-90 / timm SqueezeExcite) ----
                __threadfence();
                const float* part = p.partials + (size_t)n * p.nparts * p.C;
                for (int cc = ctid; cc < p.C; cc += 512) {
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                    int q = 0;
                    for (; q + 3 < p.nparts; q += 4) {
                        s0 += ld_cg_f32(part + (size_t)q * p.C + cc);       s1 += ld_cg_f32(part + (size_t)(q + 1) * p.C + cc);
                        s2 += ld_cg_f32(part + (size_t)(q + 2) * p.C + cc); s3 += ld_cg_f32(part + (size_t)(q + 3) * p.C + cc);
                    }
                    for (; q < p.nparts; ++q) s0 += ld_cg_f32(part + (size_t)q * p.C + cc);
                    s_mean[cc] = ((s0 + s1) + (s2 + s3)) * p.inv_count;
                }
                named_bar_sync(1, 512);
                {
                    const float4* m = reinterpret_cast<const float4*>(s_mean);
                    const int c4n = p.C >> 2;
                    for (int j0 = cw; j0 < p.rd; j0 += 16 * 3) {
                        float acc[3];
                        const float4* wrow[3];
#pragma unroll
                        for (int u = 0; u < 3; ++u) {
                            acc[u] = 0.f;
                            const int j = j0 + u * 16;
                            wrow[u] = reinterpret_cast<const float4*>(p.se_w1 + (size_t)(j < p.rd ? j : j0) * p.C);
                        }
#pragma unroll 3
                        for (int c4 = lane; c4 < c4n; c4 += 32) {
                            const float4 mv = m[c4];
#pragma unroll
                            for (int u = 0; u < 3; ++u) {
                                const float4 wv = __ldg(wrow[u] + c4);
                                acc[u] = fmaf(wv.x, mv.x, fmaf(wv.y, mv.y, fmaf(wv.z, mv.z, fmaf(wv.w, mv.w, acc[u]))));
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 3; ++u) {
                            const int j = j0 + u * 16;
                            const float a = warp_sum(acc[u]);
                            if (lane == 0 && j < p.rd) s_hid[j] = silu_f(a + __ldg(p.se_b1 + j));
                        }
                    }
                }
                named_bar_sync(1, 512);
                for (int cc = ctid; cc < p.C; cc += 512) {
                    float a0s = 0.f, a1s = 0.f;
                    int j = 0;
                    for (; j + 1 < p.rd; j += 2) {
                        a0s = fmaf(__ldg(p.se_w2t + (size_t)j * p.C + cc), s_hid[j], a0s);
                        a1s = fmaf(__ldg(p.se_w2t + (size_t)(j + 1) * p.C + cc), s_hid[j + 1], a1s);
                    }
                    if (j < p.rd) a0s = fmaf(__ldg(p.se_w2t + (size_t)j * p.C + cc), s_hid[j], a0s);
                    p.gate[(size_t)n * p.C + cc] = sigmoid_f(a0s + a1s + __ldg(p.se_b2 + cc));
                }
                __threadfence();
                named_bar_sync(1, 512);
                if (ctid == 0) {
                    if (p.tiles_per_img > 0) st_release_gpu(&flag[n], 1);
                    else dw_done[n] = 0;          // depthwise + SE only: nobody consumes the flag, leave the words clean
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTlGemmRegs));
        // ================================== projection warps: 4 epilogue + 4 gaters ==================================
        const int gw = warp - (4 + kTlDwWarps);     // 0..7
        if (gw >= 4) {
            // ---- gaters: A block *= gate, in place (128 rows x 8 chunks of 16 B; 128-byte swizzle) ----
            const int gt = tid - 32 * (4 + kTlDwWarps + 4);     // 0..127
            uint32_t gcnt = 0;
            int gate_img = -1;
            for (int idx = blockIdx.x; idx < n_tiles; idx += gridDim.x) {
                const int img = idx / p.tiles_per_img;
                named_bar_sync(2, 128);             // every gater warp has finished the previous tile (s_gate is single-buffered)
                for (int kb = 0; kb <= p.num_kb; ++kb, ++gcnt) {   // block num_kb = bias block: nothing to gate, but every use of
                                                                    // a stage completes one phase of each of its barriers
                    const int st = gcnt % p.g_stages;
                    mbar_wait(&g_full[st], (gcnt / p.g_stages) & 1);
                    if (kb == 0 && gate_img != img) {
                        // g_full of the first block completes only after the producer acquired the image's flag, so the gate
                        // vector written by the SE CTA is visible (read through L2: the buffer is re-used by every layer)
                        gate_img = img;
                        const float* gsrc = p.gate + (size_t)img * p.C;
                        for (int i = gt; i < p.C; i += 128) s_gate[i] = ld_cg_f32(gsrc + i);
                        named_bar_sync(2, 128);
                    }
                    if (kb < p.num_kb) {
                        // thread -> physical chunk pc = gt & 7 of rows (gt >> 3) + 16 i; logical chunk j = pc ^ (row & 7) is
                        // the same for all eight rows (16 i keeps row & 7)
                        unsigned char* a_blk = s_gring + (size_t)st * g_stage_bytes;
                        const int pc = gt & 7, r0 = gt >> 3;
                        const int j = pc ^ (r0 & 7);
                        const float4 g0 = *reinterpret_cast<const float4*>(s_gate + kb * kTcBK + j * 8);
                        const float4 g1 = *reinterpret_cast<const float4*>(s_gate + kb * kTcBK + j * 8 + 4);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            uint4* q = reinterpret_cast<uint4*>(a_blk + (size_t)(r0 + 16 * i) * 128 + pc * 16);
                            uint4 v = *q;
                            float2 f;
                            f = unpack_half2(v.x); v.x = pack_half2(f.x * g0.x, f.y * g0.y);
                            f = unpack_half2(v.y); v.y = pack_half2(f.x * g0.z, f.y * g0.w);
                            f = unpack_half2(v.z); v.z = pack_half2(f.x * g1.x, f.y * g1.y);
                            f = unpack_half2(v.w); v.w = pack_half2(f.x * g1.z, f.y * g1.w);
                            *q = v;
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&g_gated[st]);
                }
            }
        } else {
            // ---- epilogue: TMEM -> (+ shortcut) -> fp16 -> global; warp q owns TMEM lanes [32q, 32q + 32) ----
            const int q = warp & 3;                 // warps 20..23 -> 0..3
            int t_acc = 0;
            const bool has_res = p.res != nullptr;
            const int ngroups = p.N >> 4;
            for (int idx = blockIdx.x; idx < n_tiles; idx += gridDim.x, ++t_acc) {
                const int img = idx / p.tiles_per_img, ti = idx - img * p.tiles_per_img;
                const int e = t_acc & 1;
                const int r_in = ti * kTcBM + q * 32 + lane;
                const bool row_ok = r_in < p.rows_per_img;
                const long long row = (long long)img * p.rows_per_img + r_in;
                __half* c_row = p.out + row * p.N;
                const __half* r_row = p.res + row * p.N;
                for (int g0 = 0; g0 < ngroups; g0 += kTlGP) {
                    uint32_t rv[kTlGP][8];
                    if (has_res) {
#pragma unroll
                        for (int j = 0; j < kTlGP; ++j)
                            if (g0 + j < ngroups && row_ok) ld_global_v8(r_row + (g0 + j) * 16, rv[j]);
                    }
                    if (g0 == 0) {
                        mbar_wait(&acc_full[e], ((uint32_t)t_acc >> 1) & 1);
                        tc_fence_after();
                    }
                    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(e * p.N);
                    uint32_t v[kTlGP][16];
#pragma unroll
                    for (int j = 0; j < kTlGP; ++j)
                        if (g0 + j < ngroups) tc_ld16(t_row + (uint32_t)((g0 + j) * 16), v[j]);
                    tc_wait_ld();
                    if (g0 + kTlGP >= ngroups) {     // last TMEM read of this tile: hand the accumulator back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[e]);
                    }
#pragma unroll
                    for (int j = 0; j < kTlGP; ++j) {
                        if (g0 + j < ngroups) {
                            uint32_t pk[8];
#pragma unroll
                            for (int h = 0; h < 4; ++h) {
                                float x0 = __uint_as_float(v[j][4 * h]), x1 = __uint_as_float(v[j][4 * h + 1]);
                                float x2 = __uint_as_float(v[j][4 * h + 2]), x3 = __uint_as_float(v[j][4 * h + 3]);
                                if (has_res && row_ok) {
                                    const float2 r01 = unpack_half2(rv[j][2 * h]), r23 = unpack_half2(rv[j][2 * h + 1]);
                                    x0 += r01.x; x1 += r01.y; x2 += r23.x; x3 += r23.y;
                                }
                                pk[2 * h] = pack_half2(x0, x1);
                                pk[2 * h + 1] = pack_half2(x2, x3);
                            }
                            if (row_ok) st_global_v8(c_row + (g0 + j) * 16, pk);
                        }
                    }
                }
                // last tile of the image: leave the sync words zero for the next launch
                __syncwarp();
                if (lane == 0) {
                    const int old = atomicAdd(&gemm_done[img], 1);
                    if (old == 4 * p.tiles_per_img - 1) { dw_done[img] = 0; flag[img] = 0; gemm_done[img] = 0; }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1 && n_tiles > 0) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

}  // namespace mds
