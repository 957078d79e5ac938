// MBConv tail in ONE launch: depthwise 3x3 (2D, stride 1/2, TF-SAME) or 3x3x3 (3D) + folded BN + SiLU, SE squeeze,
// SE excitation MLP, SE gating and the 1x1 projection GEMM (+ folded BN, + shortcut)
// (timm InvertedResidual conv_dw/bn2/se/conv_pwl/bn3 built at multidim_stacker.py:166-176; InvertedResidual3d
// multidim_stacker.py:110-134; SqueezeExcite :72-90).
//
// The SE squeeze is a reduction over a whole image (2D) / stack (3D), so the projection of an image can only start when
// its depthwise pass is complete.  Instead of three kernels separated by grid-wide barriers this is one persistent
// kernel (one CTA per SM) that walks a statically ordered list of work items with per-image dependencies:
//   * dw item   = (image, [plane t], row chunk, 40 output columns, 128-channel slab): input rows are staged by TMA
//     (cp.async.bulk.tensor, 4-D / 5-D map over the NHWC tensor, hardware zero fill = the conv padding) into an
//     mbarrier ring; 16 compute warps (two 64-channel groups x 8 warps x 5 columns, lane = 2 channels as one f32x2,
//     packed FFMA2) write the SiLU'd fp16 output and leave one squeeze partial per (image, part, channel);
//     the CTA that completes the LAST dw item of an image (device-scope counter) evaluates the SE MLP for that image
//     from the partials in a fixed order (deterministic, independent of the batch), writes the fp32 gate vector and
//     releases the image's flag;
//   * gemm item = (image, 128-row tile): the TMA producer acquires the image's flag, then streams [A | W] K-blocks;
//     8 "gater" warps multiply the A block by the gate in shared memory (fp32 product, one rounding), the tcgen05
//     issuer accumulates in TMEM (bias through a ones-tile MMA), 8 epilogue warps add the shortcut and store.
// Items are ordered in blocks of images, dw(B0) | dw(B1) gemm(B0) | dw(B2) gemm(B1) | ..., so that an image's depthwise
// output is still in L2 when its projection reads it; every CTA takes items blockIdx.x, blockIdx.x + grid, ... in order.
// The order is topological and all CTAs are co-resident, so the earliest unfinished item can always run: no deadlock.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "conv3x3_tc.cuh"
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
    int rows_per_img, tiles_per_img, N, num_kb;
    int lag;                 // images of dw work issued ahead of the first projection
    int n_items;             // n * (dw_per_img + tiles_per_img)  (tiles_per_img = 0: depthwise + SE only)
    int tmem_cols;
    int g_stages;            // projection ring depth
};

// warps 0-3: service warpgroup (0 = TMA producer, 1 = tcgen05 issuer, 2-3 idle) that hands most of its registers to the
// four compute warpgroups (warps 4-19) through setmaxnreg
constexpr int kTlThreads = 128 + 32 * 16;
constexpr int kTlServiceRegs = 24, kTlComputeRegs = 112;
constexpr int kTlPXW = 5, kTlTWX = 40, kTlCS = 128;      // columns per warp / per item, channels per item
constexpr int kTlMaxC = 1152, kTlMaxRd = 64;
constexpr int kTlGP = 2;               // 16-column accumulator groups an epilogue warp holds in registers at once

template <int KT, int STRIDE>
struct TailCfg {
    static constexpr int IW = (STRIDE == 1) ? kTlTWX + 2 : 2 * kTlTWX + 1;     // input columns per row tile
    static constexpr int NV = (STRIDE == 1) ? kTlPXW + 2 : 2 * kTlPXW + 1;
    static constexpr int RPS = (KT == 3) ? 1 : 2;                              // input rows per stage
    static constexpr int ROWB = IW * kTlCS * 2;                                // bytes of one input row tile (one plane)
    static constexpr int STAGEB = KT * RPS * ROWB;
    static constexpr int NST = (STRIDE == 2) ? 3 : 4;
    static constexpr int DW_RING = NST * STAGEB;
};

__host__ __device__ inline int tl_gemm_stage_bytes(int N) { return kTcABytes + N * 128; }
__host__ __device__ inline int tl_ring_bytes(int dw_ring, int N, int g_stages) {
    const int g = g_stages * tl_gemm_stage_bytes(N);
    return g > dw_ring ? g : dw_ring;
}
// shared memory map (after 1024-byte alignment): ones tile | ring | s_gate[kTlMaxC] | s_mean[kTlMaxC] | s_hid[kTlMaxRd]
//                                               | s_part[16][64] | s_w3[27][128] (3D only) | barriers
__host__ __device__ inline size_t tl_smem_bytes(int dw_ring, int N, int g_stages, int kt) {
    return 1024 + kTcABytes + (size_t)tl_ring_bytes(dw_ring, N, g_stages) + (size_t)(2 * kTlMaxC + kTlMaxRd + 16 * 64) * 4 +
           (kt == 3 ? 27 * kTlCS * 4 : 0) + 512;
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

// item index -> (is_gemm, image, local index).  Images are taken in blocks of L = p.lag; order:
//   dw(B0) | dw(B1) gemm(B0) | dw(B2) gemm(B1) | ... | gemm(B_last)
// i.e. an image's projection tiles come one block of depthwise work after its own depthwise items: their dependency is
// (almost always) already satisfied when a CTA reaches them, the depthwise output is still in L2 (2-3 blocks live), and a
// CTA changes between the two kinds of work only twice per block (the shared-memory ring is drained at every change).
struct TailItem { int gemm, img, loc; };
__device__ __forceinline__ TailItem tail_item(const TailParams& p, int idx) {
    TailItem it;
    const int D = p.dw_per_img, G = p.tiles_per_img;
    const int L = p.lag < p.n ? p.lag : p.n;
    const int nb = (p.n + L - 1) / L;
    auto bsize = [&](int k) { return (k + 1) * L <= p.n ? L : p.n - k * L; };
    int seg = bsize(0) * D;                      // dw(B0)
    if (idx < seg) { it.gemm = 0; it.img = idx / D; it.loc = idx - it.img * D; return it; }
    idx -= seg;
    for (int k = 0; k < nb; ++k) {
        if (k + 1 < nb) {                        // dw(B_{k+1})
            seg = bsize(k + 1) * D;
            if (idx < seg) { const int i = idx / D; it.gemm = 0; it.img = (k + 1) * L + i; it.loc = idx - i * D; return it; }
            idx -= seg;
        }
        seg = bsize(k) * G;                      // gemm(B_k)
        if (idx < seg) { const int i = idx / G; it.gemm = 1; it.img = k * L + i; it.loc = idx - i * G; return it; }
        idx -= seg;
    }
    it.gemm = 0; it.img = 0; it.loc = 0;         // unreachable for idx < n_items
    return it;
}

template <int KT, int STRIDE>
__global__ void __launch_bounds__(kTlThreads, 1)
mbconv_tail_kernel(const __grid_constant__ CUtensorMap tmDw, const __grid_constant__ CUtensorMap tmA,
                   const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBias, TailParams p) {
    using Cfg = TailCfg<KT, STRIDE>;
    extern __shared__ unsigned char tl_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tl_smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* s_ones = smem;
    unsigned char* s_ring = s_ones + kTcABytes;
    const int g_stage_bytes = tl_gemm_stage_bytes(p.N);
    float* s_gate = reinterpret_cast<float*>(s_ring + tl_ring_bytes(Cfg::DW_RING, p.N, p.g_stages));     // [kTlMaxC]
    float* s_mean = s_gate + kTlMaxC;
    float* s_hid = s_mean + kTlMaxC;
    float* s_part = s_hid + kTlMaxRd;                                                                  // [16][64]
    float* s_w3 = s_part + 16 * 64;                                                                    // [27][128] (KT == 3)
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

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) {
            mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 16);
            mbar_init(&g_full[i], 1); mbar_init(&g_gated[i], 8); mbar_init(&g_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDw) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBias) : "memory");
    }
    for (int i = tid; i < kTcABytes / 16; i += kTlThreads) {        // ones tile (128-byte swizzled K-major), see gemm_tc.cuh
        const int r = i >> 3, ch = i & 7;
        reinterpret_cast<uint4*>(s_ones)[i] = (ch == (r & 7)) ? make_uint4(0x3C003C00u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    }
    for (int i = tid; i < kTlMaxC; i += kTlThreads) s_gate[i] = 0.f;      // entries beyond C stay zero (K tail of the last block)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1 && p.tiles_per_img > 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    pdl_wait();

    // geometry of a dw item
    auto dw_geom = [&](int loc, int& t, int& chunk, int& xt, int& sp) {
        sp = loc % p.slab_pairs; loc /= p.slab_pairs;
        xt = loc % p.xtiles; loc /= p.xtiles;
        chunk = loc % p.chunks; t = loc / p.chunks;
    };

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kTlServiceRegs));
    if (warp == 0) {
        // =========================================== TMA producer ===========================================
        if (lane == 0) {
            uint32_t dcnt = 0, gcnt = 0;          // stages issued so far in each ring
            int seen_img = -1;
            bool ring_is_dw = true;               // which layout the shared ring currently holds
            for (int idx = blockIdx.x; idx < p.n_items; idx += gridDim.x) {
                const TailItem it = tail_item(p, idx);
                if (!it.gemm) {
                    if (!ring_is_dw) {            // every projection stage must have been consumed by the MMAs
                        for (int s = 0; s < p.g_stages; ++s) {
                            const uint32_t c = gcnt + s;            // the next p.g_stages uses: wait for their "empty" phase
                            mbar_wait(&g_empty[c % p.g_stages], ((c / p.g_stages) & 1) ^ 1);
                        }
                        ring_is_dw = true;
                    }
                    int t, chunk, xt, sp;
                    dw_geom(it.loc, t, chunk, xt, sp);
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
                            tma_load_5d(s_ring + (size_t)st * Cfg::STAGEB, &tmDw, &d_full[st], sp * kTlCS, xi0, yi0 + js, t - 1, it.img);
                        else
                            tma_load_4d(s_ring + (size_t)st * Cfg::STAGEB, &tmDw, &d_full[st], sp * kTlCS, xi0, yi0 + js * Cfg::RPS, it.img);
                    }
                } else {
                    if (ring_is_dw) {             // every depthwise stage must have been released by the compute warps
                        for (int s = 0; s < Cfg::NST; ++s) {
                            const uint32_t c = dcnt + s;
                            mbar_wait(&d_empty[c % Cfg::NST], ((c / Cfg::NST) & 1) ^ 1);
                        }
                        ring_is_dw = false;
                    }
                    if (it.img != seen_img) {     // the image's depthwise output and gate are complete (flag released by the SE CTA)
                        uint32_t spins = 0;
                        while (ld_acquire_gpu(&flag[it.img]) == 0) {
                            __nanosleep(64);
                            if (++spins > (1u << 24)) __trap();
                        }
                        seen_img = it.img;
                        asm volatile("fence.proxy.async;" ::: "memory");      // generic-proxy global writes -> async-proxy (TMA) reads
                    }
                    const int row0 = it.img * p.rows_per_img + it.loc * kTcBM;
                    for (int kb = 0; kb <= p.num_kb; ++kb, ++gcnt) {
                        const int st = gcnt % p.g_stages;
                        mbar_wait(&g_empty[st], ((gcnt / p.g_stages) & 1) ^ 1);
                        unsigned char* dst = s_ring + (size_t)st * g_stage_bytes;
                        if (kb < p.num_kb) {
                            mbar_expect_tx(&g_full[st], (uint32_t)g_stage_bytes);
                            tma_load_2d(dst, &tmA, &g_full[st], kb * kTcBK, row0);
                            tma_load_2d(dst + kTcABytes, &tmB, &g_full[st], kb * kTcBK, 0);
                        } else {
                            mbar_expect_tx(&g_full[st], (uint32_t)(p.N * 128));
                            tma_load_2d(dst + kTcABytes, &tmBias, &g_full[st], 0, 0);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =========================================== MMA issuer ===========================================
        if (lane == 0) {
            const uint32_t idesc = tc_idesc(kTcBM, p.N);
            uint32_t gcnt = 0;
            int t = 0;
            for (int idx = blockIdx.x; idx < p.n_items; idx += gridDim.x) {
                const TailItem it = tail_item(p, idx);
                if (!it.gemm) continue;
                const int acc = t & 1;
                mbar_wait(&acc_empty[acc], (((uint32_t)t >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.N);
                for (int kb = 0; kb <= p.num_kb; ++kb, ++gcnt) {
                    const int st = gcnt % p.g_stages;
                    const uint32_t ph = (gcnt / p.g_stages) & 1;
                    const bool bias_blk = kb == p.num_kb;
                    mbar_wait(&g_gated[st], ph);
                    tc_fence_after();
                    unsigned char* sp = s_ring + (size_t)st * g_stage_bytes;
                    const uint64_t adesc = tc_smem_desc(smem_u32(bias_blk ? s_ones : sp));
                    const uint64_t bdesc = tc_smem_desc(smem_u32(sp + kTcABytes));
                    const int nk = bias_blk ? 1 : kTcBK / 16;
                    for (int k = 0; k < nk; ++k)
                        tc_mma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                    tc_commit(&g_empty[st]);
                }
                tc_commit(&acc_full[acc]);
                ++t;
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTlComputeRegs));
        // ============================ compute warps: depthwise + SE / gating + epilogue ============================
        const int cw = warp - 4;                    // 0..15
        const int ctid = tid - 128;                 // 0..511
        const int grp = cw >> 3, wi = cw & 7;       // depthwise: channel group (64 ch) and column strip
        const bool is_gater = cw >= 8;              // projection: warps 12..19 gate A, warps 4..11 run the epilogue
        uint32_t dcnt = 0, gcnt = 0;
        int t_acc = 0;
        int gate_img = -1;
        const uint32_t ring_u32 = smem_u32(s_ring);

        for (int idx = blockIdx.x; idx < p.n_items; idx += gridDim.x) {
            const TailItem it = tail_item(p, idx);
            if (!it.gemm) {
                // ------------------------------------------------ depthwise item ------------------------------------------------
                int t, chunk, xt, sp;
                dw_geom(it.loc, t, chunk, xt, sp);
                const int n = it.img;
                const int yo0 = chunk * p.rows_per_chunk, yo1 = min(p.Ho, yo0 + p.rows_per_chunk);
                const int xo0 = xt * kTlTWX;
                const int c = sp * kTlCS + grp * 64 + 2 * lane;
                const bool c_ok = c < p.C;
                const int c_ld = c_ok ? c : 0;
                const int yi0 = (STRIDE == 1) ? yo0 - 1 : 2 * yo0;
                const int NR = (STRIDE == 1) ? (yo1 - yo0) + 2 : 2 * (yo1 - yo0) + 1;
                const int nstg = (NR + Cfg::RPS - 1) / Cfg::RPS;

                float2 w[9];
                if constexpr (KT == 1) {
#pragma unroll
                    for (int i = 0; i < 9; ++i) w[i] = __ldg(reinterpret_cast<const float2*>(p.dw_w + (size_t)i * p.C + c_ld));
                } else {
                    // stage this item's 27 x 128 weights in shared memory (the previous item's readers are past the barrier below)
                    named_bar_sync(1, 512);
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
                const uint32_t w3_u32 = smem_u32(s_w3) + (uint32_t)(grp * 64 + 2 * lane) * 4u;

                float2 lsum = make_float2(0.f, 0.f);
                float2 a0[kTlPXW], a1[kTlPXW], a2[kTlPXW];
#pragma unroll
                for (int j = 0; j < kTlPXW; ++j) a0[j] = a1[j] = a2[j] = make_float2(0.f, 0.f);
                const int xw = xo0 + wi * kTlPXW;
                const int px_base = (STRIDE == 1) ? wi * kTlPXW : 2 * wi * kTlPXW;
                const long long out_pitch = (long long)p.Wo * p.C;
                __half* o_ptr = p.m2 + ((size_t)n * p.T + t) * (size_t)p.Ho * out_pitch + (long long)yo0 * out_pitch + (long long)xw * p.C + c;
                bool px_ok[kTlPXW];
#pragma unroll
                for (int j = 0; j < kTlPXW; ++j) px_ok[j] = c_ok && (xw + j < p.Wo);
                const int pxC = p.C;

                auto emit = [&](const float2 (&acc)[kTlPXW]) {
#pragma unroll
                    for (int j = 0; j < kTlPXW; ++j) {
                        if (px_ok[j]) {
                            const float ox = silu_f(acc[j].x + bias.x), oy = silu_f(acc[j].y + bias.y);
                            lsum.x += ox; lsum.y += oy;
                            *reinterpret_cast<uint32_t*>(o_ptr + j * pxC) = pack_half2(ox, oy);
                        }
                    }
                    o_ptr += out_pitch;
                };

                const uint32_t lds_lane = (uint32_t)(px_base * kTlCS + grp * 64 + 2 * lane) * 2u;
                for (int js = 0; js < nstg; ++js, ++dcnt) {
                    const int st = dcnt % Cfg::NST;
                    mbar_wait(&d_full[st], (dcnt / Cfg::NST) & 1);
                    const uint32_t sbase = ring_u32 + (uint32_t)st * Cfg::STAGEB + lds_lane;
#pragma unroll
                    for (int rr = 0; rr < Cfg::RPS; ++rr) {
                        const int k = js * Cfg::RPS + rr;
                        if (k >= NR) break;
                        const uint32_t srow = sbase + rr * Cfg::ROWB;
                        const int yi = yi0 + k;
                        if constexpr (STRIDE == 1) {
#pragma unroll
                            for (int j = 0; j < kTlPXW; ++j) a2[j] = make_float2(0.f, 0.f);
#pragma unroll
                            for (int dt = 0; dt < KT; ++dt) {
                                if constexpr (KT == 3) {
#pragma unroll
                                    for (int i = 0; i < 9; ++i) w[i] = lds_f32x2(w3_u32 + (uint32_t)((dt * 9 + i) * kTlCS) * 4u);
                                }
                                float2 v[Cfg::NV];
#pragma unroll
                                for (int i = 0; i < Cfg::NV; ++i) v[i] = lds_half2(srow + (uint32_t)(dt * Cfg::ROWB + i * kTlCS * 2));
#pragma unroll
                                for (int j = 0; j < kTlPXW; ++j)
#pragma unroll
                                    for (int s = 0; s < 3; ++s) {
                                        a2[j] = ffma2(w[0 + s], v[j + s], a2[j]);
                                        a1[j] = ffma2(w[3 + s], v[j + s], a1[j]);
                                        a0[j] = ffma2(w[6 + s], v[j + s], a0[j]);
                                    }
                            }
                            if (yi - 1 >= yo0) emit(a0);
#pragma unroll
                            for (int j = 0; j < kTlPXW; ++j) { a0[j] = a1[j]; a1[j] = a2[j]; }
                        } else {
                            float2 v[Cfg::NV];
#pragma unroll
                            for (int i = 0; i < Cfg::NV; ++i) v[i] = lds_half2(srow + (uint32_t)(i * kTlCS * 2));
                            if ((k & 1) == 0) {
#pragma unroll
                                for (int j = 0; j < kTlPXW; ++j) {
                                    a2[j] = make_float2(0.f, 0.f);
#pragma unroll
                                    for (int s = 0; s < 3; ++s) {
                                        a1[j] = ffma2(w[6 + s], v[2 * j + s], a1[j]);
                                        a2[j] = ffma2(w[0 + s], v[2 * j + s], a2[j]);
                                    }
                                }
                                if (k > 0) emit(a1);
#pragma unroll
                                for (int j = 0; j < kTlPXW; ++j) a1[j] = a2[j];
                            } else {
#pragma unroll
                                for (int j = 0; j < kTlPXW; ++j)
#pragma unroll
                                    for (int s = 0; s < 3; ++s) a1[j] = ffma2(w[3 + s], v[2 * j + s], a1[j]);
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
                    if (ctid == 0) st_release_gpu(&flag[n], 1);
                }
            } else {
                // ------------------------------------------------ projection item ------------------------------------------------
                const int img = it.img;
                if (is_gater) {
                    const int gt = ctid - 256;                 // 0..255
                    named_bar_sync(2, 256);                    // every gater warp has finished the previous tile (s_gate is single-buffered)
                    for (int kb = 0; kb <= p.num_kb; ++kb, ++gcnt) {      // block num_kb = bias block: nothing to gate, but every
                                                                              // use of a stage completes one phase of each barrier
                        const int st = gcnt % p.g_stages;
                        mbar_wait(&g_full[st], (gcnt / p.g_stages) & 1);
                        if (kb == 0 && gate_img != img) {
                            // g_full of the first block completes only after the producer acquired the image's flag, so the
                            // gate vector written by the SE CTA is visible (read through L2: the buffer is re-used by every layer)
                            gate_img = img;
                            const float* gsrc = p.gate + (size_t)img * p.C;
                            for (int i = gt; i < p.C; i += 256) s_gate[i] = ld_cg_f32(gsrc + i);
                            named_bar_sync(2, 256);
                        }
                        if (kb < p.num_kb) {
                        // 128 rows x 8 chunks of 16 B; thread -> physical chunk pc = gt & 7 of rows (gt >> 3) + 32 i
                        unsigned char* a_blk = s_ring + (size_t)st * g_stage_bytes;
                        const int pc = gt & 7, r0 = gt >> 3;
                        const int j = pc ^ (r0 & 7);          // logical 16-byte chunk (the same for all four rows: 32 i keeps r & 7)
                        const float4 g0 = *reinterpret_cast<const float4*>(s_gate + kb * kTcBK + j * 8);
                        const float4 g1 = *reinterpret_cast<const float4*>(s_gate + kb * kTcBK + j * 8 + 4);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint4* q = reinterpret_cast<uint4*>(a_blk + (size_t)(r0 + 32 * i) * 128 + pc * 16);
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
                } else {
                    // epilogue: two groups of 4 warps; group e drains accumulator e
                    const int q = warp & 3;
                    const int e = cw >> 2;                     // cw in 0..7
                    if ((t_acc & 1) == e) {
                        const int r_in = it.loc * kTcBM + q * 32 + lane;
                        const bool row_ok = r_in < p.rows_per_img;
                        const long long row = (long long)img * p.rows_per_img + r_in;
                        __half* c_row = p.out + row * p.N;
                        const __half* r_row = p.res + row * p.N;
                        const bool has_res = p.res != nullptr;
                        const int ngroups = p.N >> 4;
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
                            if (g0 + kTlGP >= ngroups) {
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
                    ++t_acc;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1 && p.tiles_per_img > 0) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

}  // namespace mds
