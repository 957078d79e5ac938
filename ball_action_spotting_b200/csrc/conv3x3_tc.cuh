// K2+K3 on the 5th-gen tensor cores: stride-1 FusedMBConv (timm EdgeResidual: blocks.1.1 / blocks.2.1)
//     y = conv1x1( SiLU( conv3x3(x) * s1 + b1 ) ) * s2 + b2 + x
// as ONE persistent, warp-specialised tcgen05 kernel.  The expanded tensor never leaves the SM.
//
// Implicit GEMM without im2col: the input halo tile is loaded by TMA (5-D tensor map over NHWC seen as
// [n][C/8][H][W][8], hardware zero fill = the conv's zero padding) into shared memory as 8-channel PLANES,
// plane[c8][pixel][8 ch] with pixels in row-major halo order.  In the no-swizzle K-major UMMA layout a row is 16 bytes
// and 8-row groups are 128 bytes apart, so "row m of the A operand" is simply "pixel p0 + m of the linearised tile":
// tap (r, s) of the 3x3 stencil is nothing but a start-address offset of (r * (TW+2) + s) pixels.  The two halo
// columns of every tile row become garbage rows of M (6 %), which the epilogue discards.
//   warp 0 / lane 0 : TMA producer (double-buffered halo tiles)
//   warp 1 / lane 0 : MMA issuer: per 128-pixel M tile, 9 taps x CIN/16 tcgen05.mma (N = CMID) into TMEM, one more
//                     against a "ones" tile that adds the bias; later CMID/16 MMAs (N = COUT) for the projection
//   warps 2-17      : epilogue, two ping-pong groups: TMEM -> SiLU -> fp16 -> smem (A operand of the projection MMA);
//                     then TMEM -> + residual -> fp16 -> 32-byte global stores
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace mds {

struct Conv3TcParams {
    const __half* in;     // [n][H][W][CIN]  (also the residual)
    __half* out;          // [n][H][W][COUT]
    const __half* w1;     // [CMID][9*CIN]   k = (r*3+s)*CIN + ci, BN folded
    const float* b1;      // [CMID]
    const __half* w2;     // [COUT][CMID]
    const float* b2;      // [COUT]
    int n, H, W;
    int tiles_x, tiles_y;
};

template <int CIN, int CMID, int COUT>
struct Conv3TcCfg {
    static constexpr int TW = 32, TH = 15;
    static constexpr int SW = TW + 2, SH = TH + 2;            // halo tile 34 x 17
    static constexpr int PIX = SW * SH;                        // 578 pixels per plane
    static constexpr int PLANE = PIX * 16;                     // bytes per 8-channel plane
    static constexpr int TILE_BYTES = (CIN / 8) * PLANE;
    static constexpr int TILE_ALLOC = ((TILE_BYTES + 64 + 127) / 128) * 128;   // + 64 B: the last M tile over-reads 3 pixels
    static constexpr int MT_MAX = (TH * SW + 127) / 128;       // 4 M tiles of 128 linear pixels
    static constexpr int W1_BYTES = 9 * CIN * CMID * 2;        // [tap][c8][n][8]
    static constexpr int W2_BYTES = CMID * COUT * 2;           // [c8][n][8]
    static constexpr int P_BYTES = 128 * CMID * 2;             // [c8][row][8], one per accumulator
    static constexpr int ONES_BYTES = 2 * 128 * 16;            // [2 planes][128 rows][8]: (row, k=0,1) = 1
    static constexpr int BM1_BYTES = 2 * CMID * 16;            // [2 planes][CMID][8]: (n, k=0,1) = bias hi/lo
    static constexpr int BM2_BYTES = 2 * COUT * 16;
    static constexpr int D2_STRIDE = 64;                       // TMEM columns reserved per projection accumulator
    static constexpr size_t SMEM = 128 + 2 * (size_t)TILE_ALLOC + W1_BYTES + W2_BYTES + 2 * P_BYTES + ONES_BYTES + BM1_BYTES +
                                   BM2_BYTES + 256;
    static_assert(CIN % 16 == 0 && CMID % 16 == 0 && COUT % 16 == 0 && CIN == COUT, "shape");
    static_assert(2 * CMID + 2 * D2_STRIDE <= 512, "TMEM columns");
};

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

template <int CIN, int CMID, int COUT>
__global__ void __launch_bounds__(kTcThreads, 1) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmIn, Conv3TcParams p) {
    using Cfg = Conv3TcCfg<CIN, CMID, COUT>;
    extern __shared__ unsigned char c3tc_smem_raw[];
    pdl_trigger();
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(c3tc_smem_raw) + 127) & ~uintptr_t(127));
    unsigned char* s_tile = smem;                                   // [2][TILE_ALLOC]
    unsigned char* s_w1 = s_tile + 2 * Cfg::TILE_ALLOC;             // [9][CIN/8][CMID][16 B]
    unsigned char* s_w2 = s_w1 + Cfg::W1_BYTES;                     // [CMID/8][COUT][16 B]
    unsigned char* s_p = s_w2 + Cfg::W2_BYTES;                      // [2][CMID/8][128][16 B]
    unsigned char* s_ones = s_p + 2 * Cfg::P_BYTES;
    unsigned char* s_bm1 = s_ones + Cfg::ONES_BYTES;
    unsigned char* s_bm2 = s_bm1 + Cfg::BM1_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bm2 + Cfg::BM2_BYTES);
    uint64_t* tile_full = bars;          // [2]
    uint64_t* tile_empty = bars + 2;     // [2]
    uint64_t* d1_full = bars + 4;        // [2]
    uint64_t* d1_empty = bars + 6;       // [2]
    uint64_t* p_full = bars + 8;         // [2]
    uint64_t* d2_full = bars + 10;       // [2]
    uint64_t* d2_empty = bars + 12;      // [2]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 14);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int ntiles = tiles_per_img * p.n;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tile_full[i], 1); mbar_init(&tile_empty[i], 1);
            mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], kTcEpiWarps / 2);
            mbar_init(&p_full[i], kTcEpiWarps / 2);
            mbar_init(&d2_full[i], 1); mbar_init(&d2_empty[i], kTcEpiWarps / 2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmIn) : "memory");
    }
    // ---- weights, ones tile and bias tiles -> smem in the no-swizzle K-major plane layout (once per CTA) ----
    for (int i = tid; i < 9 * (CIN / 8) * CMID; i += kTcThreads) {          // dest chunk (tap, c8, n)
        const int nrow = i % CMID, pl = i / CMID;                            // pl = tap * (CIN/8) + c8
        reinterpret_cast<uint4*>(s_w1)[i] = __ldg(reinterpret_cast<const uint4*>(p.w1 + (size_t)nrow * 9 * CIN + pl * 8));
    }
    for (int i = tid; i < (CMID / 8) * COUT; i += kTcThreads) {
        const int nrow = i % COUT, c8 = i / COUT;
        reinterpret_cast<uint4*>(s_w2)[i] = __ldg(reinterpret_cast<const uint4*>(p.w2 + (size_t)nrow * CMID + c8 * 8));
    }
    for (int i = tid; i < 2 * 128; i += kTcThreads)
        reinterpret_cast<uint4*>(s_ones)[i] = (i < 128) ? make_uint4(0x3C003C00u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < 2 * CMID; i += kTcThreads) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (i < CMID) {
            const float b = __ldg(p.b1 + i);
            const __half hi = __float2half_rn(b), lo = __float2half_rn(b - __half2float(hi));
            v.x = (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
        }
        reinterpret_cast<uint4*>(s_bm1)[i] = v;
    }
    for (int i = tid; i < 2 * COUT; i += kTcThreads) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (i < COUT) {
            const float b = __ldg(p.b2 + i);
            const __half hi = __float2half_rn(b), lo = __float2half_rn(b - __half2float(hi));
            v.x = (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
        }
        reinterpret_cast<uint4*>(s_bm2)[i] = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    pdl_wait();       // on-chip set-up done; activations (and the output buffer) belong to earlier kernels until now

    auto tile_geom = [&](int tile, int& n, int& y0, int& x0, int& nm) {
        n = tile / tiles_per_img;
        const int rem = tile - n * tiles_per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        y0 = ty * Cfg::TH; x0 = tx * Cfg::TW;
        const int rows = min(Cfg::TH, p.H - y0);
        nm = (rows * Cfg::SW + 127) / 128;                       // M tiles that contain valid output pixels
    };

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int i = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
                int n, y0, x0, nm;
                tile_geom(tile, n, y0, x0, nm);
                const int buf = i & 1;
                mbar_wait(&tile_empty[buf], (((uint32_t)i >> 1) & 1) ^ 1);
                mbar_expect_tx(&tile_full[buf], (uint32_t)Cfg::TILE_BYTES);
                tma_load_5d(s_tile + (size_t)buf * Cfg::TILE_ALLOC, &tmIn, &tile_full[buf], 0, x0 - 1, y0 - 1, 0, n);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc1 = tc_idesc(128, CMID), idesc2 = tc_idesc(128, COUT);
            const uint32_t w1a = smem_u32(s_w1), w2a = smem_u32(s_w2), pa = smem_u32(s_p), onesa = smem_u32(s_ones);
            const uint64_t ones_desc = tc_desc_nosw(onesa, 128 * 16);
            const uint64_t bm1_desc = tc_desc_nosw(smem_u32(s_bm1), CMID * 16), bm2_desc = tc_desc_nosw(smem_u32(s_bm2), COUT * 16);
            auto mma2 = [&](int u) {             // projection of M tile u: D2 = P . W2^T + b2
                const int a = u & 1;
                const uint32_t ph = ((uint32_t)u >> 1) & 1;
                mbar_wait(&p_full[a], ph);
                mbar_wait(&d2_empty[a], ph ^ 1);
                tc_fence_after();
                const uint32_t d2 = tmem_base + 2 * CMID + a * Cfg::D2_STRIDE;
                tc_mma_f16(d2, ones_desc, bm2_desc, idesc2, 0);
#pragma unroll
                for (int kk = 0; kk < CMID / 16; ++kk)
                    tc_mma_f16(d2, tc_desc_nosw(pa + a * Cfg::P_BYTES + kk * 2 * (128 * 16), 128 * 16),
                               tc_desc_nosw(w2a + kk * 2 * (COUT * 16), COUT * 16), idesc2, 1);
                tc_commit(&d2_full[a]);
            };
            int i = 0, t = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
                int n, y0, x0, nm;
                tile_geom(tile, n, y0, x0, nm);
                const int buf = i & 1;
                mbar_wait(&tile_full[buf], ((uint32_t)i >> 1) & 1);
                tc_fence_after();
                const uint32_t ta = smem_u32(s_tile + (size_t)buf * Cfg::TILE_ALLOC);
                for (int m = 0; m < nm; ++m, ++t) {
                    const int a = t & 1;
                    mbar_wait(&d1_empty[a], (((uint32_t)t >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d1 = tmem_base + a * CMID;
                    tc_mma_f16(d1, ones_desc, bm1_desc, idesc1, 0);                     // D1 = bias
#pragma unroll
                    for (int rs = 0; rs < 9; ++rs) {
                        const int r = rs / 3, s = rs - r * 3;
                        const uint32_t a_pix = (uint32_t)(m * 128 + r * Cfg::SW + s) * 16u;
#pragma unroll
                        for (int kc = 0; kc < CIN / 16; ++kc)
                            tc_mma_f16(d1, tc_desc_nosw(ta + kc * 2 * Cfg::PLANE + a_pix, Cfg::PLANE),
                                       tc_desc_nosw(w1a + (rs * (CIN / 8) + kc * 2) * (CMID * 16), CMID * 16), idesc1, 1);
                    }
                    tc_commit(&d1_full[a]);
                    if (t >= 1) mma2(t - 1);
                }
                tc_commit(&tile_empty[buf]);            // every MMA that reads this halo tile has been issued
            }
            if (t >= 1) mma2(t - 1);
        }
    } else {
        // ================= epilogue (warps 2..17), two groups of 8 warps =================
        const int q = warp & 3;                         // TMEM lane quadrant
        const int e = ((warp - 2) >> 2) & 1;            // group = accumulator index
        const int half = (warp - 2) >> 3;               // which half of the columns
        const int row = q * 32 + lane;                  // row of the M tile = linear halo pixel
        int i = 0, t = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
            int n, y0, x0, nm;
            tile_geom(tile, n, y0, x0, nm);
            for (int m = 0; m < nm; ++m, ++t) {
                if ((t & 1) != e) continue;
                const uint32_t ph = ((uint32_t)t >> 1) & 1;
                // ---- epilogue 1: D1 -> SiLU -> fp16 -> P (A operand of the projection) ----
                mbar_wait(&d1_full[e], ph);
                tc_fence_after();
                const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
                unsigned char* pbuf = s_p + e * Cfg::P_BYTES;
                constexpr int G1 = CMID / 32;            // 16-column groups per warp (this half)
#pragma unroll
                for (int g0 = 0; g0 < G1; g0 += 2) {
                    uint32_t v[2][16];
#pragma unroll
                    for (int j = 0; j < 2; ++j) tc_ld16(t_row + (uint32_t)(e * CMID + (half * G1 + g0 + j) * 16), v[j]);
                    tc_wait_ld();
                    if (g0 + 2 >= G1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&d1_empty[e]);
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int g = half * G1 + g0 + j;       // columns [16g, 16g+16) = planes 2g, 2g+1
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            uint4 pk;
                            pk.x = pack_half2(silu_f(__uint_as_float(v[j][8 * h + 0])), silu_f(__uint_as_float(v[j][8 * h + 1])));
                            pk.y = pack_half2(silu_f(__uint_as_float(v[j][8 * h + 2])), silu_f(__uint_as_float(v[j][8 * h + 3])));
                            pk.z = pack_half2(silu_f(__uint_as_float(v[j][8 * h + 4])), silu_f(__uint_as_float(v[j][8 * h + 5])));
                            pk.w = pack_half2(silu_f(__uint_as_float(v[j][8 * h + 6])), silu_f(__uint_as_float(v[j][8 * h + 7])));
                            *reinterpret_cast<uint4*>(pbuf + (size_t)(2 * g + h) * (128 * 16) + row * 16) = pk;
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // P written by the generic proxy, read by tcgen05.mma
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[e]);

                // ---- epilogue 2: D2 (+ residual) -> fp16 -> global ----
                const int lp = m * 128 + row;                                    // linear pixel inside the halo-row-major tile
                const int ry = lp / Cfg::SW, cx = lp - ry * Cfg::SW;
                const int oy = y0 + ry, ox = x0 + cx;
                const bool ok = (cx < Cfg::TW) && (ry < Cfg::TH) && (oy < p.H) && (ox < p.W);
                const size_t pix = ((size_t)n * p.H + (ok ? oy : 0)) * p.W + (ok ? ox : 0);
                constexpr int G2 = COUT / 32;            // 16-column groups per half... COUT = 32 -> 1, 48 -> handled below
                uint32_t rv[2][8];
                constexpr int NG2 = COUT / 16;           // total 16-col groups: 2 (COUT 32) or 3 (COUT 48)
                // half 0 takes groups 0 and 2, half 1 takes group 1
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int g = half + 2 * j;
                        if (g < NG2) ld_global_v8(p.in + pix * CIN + g * 16, rv[j]);
                    }
                }
                mbar_wait(&d2_full[e], ph);
                tc_fence_after();
                uint32_t v2[2][16];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int g = half + 2 * j;
                    if (g < NG2) tc_ld16(t_row + (uint32_t)(2 * CMID + e * Cfg::D2_STRIDE + g * 16), v2[j]);
                }
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d2_empty[e]);
                (void)G2;
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int g = half + 2 * j;
                        if (g < NG2) {
                            uint32_t pk[8];
#pragma unroll
                            for (int h = 0; h < 8; ++h) {
                                const float2 r2 = unpack_half2(rv[j][h]);
                                pk[h] = pack_half2(__uint_as_float(v2[j][2 * h]) + r2.x, __uint_as_float(v2[j][2 * h + 1]) + r2.y);
                            }
                            st_global_v8(p.out + pix * COUT + g * 16, pk);
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace mds
