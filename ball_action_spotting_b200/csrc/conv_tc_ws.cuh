// blocks.2.1 (timm EdgeResidual 48 -> 192 -> 48, stride 1, skip; multidim_stacker.py:166-176) on the 5th-gen tensor cores.
// Same implicit GEMM as conv_tc.cuh (halo tiles by TMA as 8-channel planes, taps = start-address offsets, expanded tensor
// kept in tensor memory as the projection's A operand), with two differences forced by its size:
//   * the 3x3 weights are 166 KB of fp16 and do not fit next to the tiles, so they are STREAMED: a dedicated producer warp
//     brings one tap slice (all K of that tap, 18 KB, contiguous in the tap-major copy made by conv_w1_tapmajor_kernel) per
//     ring slot with one 1-D bulk copy from L2, the MMA warp consumes a slot with CIN/16 MMAs and frees it by tcgen05.commit;
//   * two 192-column conv accumulators + two projection accumulators fill the 512 TMEM columns, so the packed fp16 expanded
//     tensor P (96 columns) is written over the low half of its own accumulator: every epilogue warp first loads ALL its
//     accumulator columns, the 16 warps meet at a named barrier, and only then P is stored.  The tensor pipe executes in order
//     (... MMA2(t) reading P, then MMA1(t+2) overwriting that accumulator), so no further hand-shake is needed.
//   warp 0: halo-tile TMA producer, warp 1: MMA issuer, warp 2: weight-slice producer, warps 4-19: epilogue
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "conv_tc.cuh"

namespace mds {

// w1 [CMID][9*CIN] (k = tap*CIN + ci)  ->  w1t [tap][CIN/8][CMID][8]
__global__ void conv_w1_tapmajor_kernel(const __half* __restrict__ w1, __half* __restrict__ w1t, int cin, int cmid) {
    const int chunks = 9 * (cin / 8) * cmid;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += gridDim.x * blockDim.x) {
        const int n = i % cmid, pl = i / cmid;                   // pl = tap * (cin/8) + c8
        reinterpret_cast<uint4*>(w1t)[i] = __ldg(reinterpret_cast<const uint4*>(w1 + (size_t)n * 9 * cin + pl * 8));
    }
}

template <int CIN, int CMID, int CPROJ>
struct ConvWsCfg {
    static constexpr int TW = 32, TH = 15, PW = TW + 2, PH = TH + 2;
    static constexpr int PIX = PW * PH, PLANE = PIX * 16;
    static constexpr int TILE_BYTES = (CIN / 8) * PLANE;
    static constexpr int MT_MAX = (TH * PW + 127) / 128;
    static constexpr int OVER = MT_MAX * 128 + 2 * PW + 2 > PIX ? MT_MAX * 128 + 2 * PW + 2 - PIX : 0;
    static constexpr int TILE_ALLOC = ((TILE_BYTES + OVER * 16 + 127) / 128) * 128;
    static constexpr int NW = 4;                                 // weight ring slots (one tap each)
    static constexpr int WSLICE = CIN * CMID * 2;                // [c8][CMID][16 B]
    static constexpr int W2_BYTES = CMID * CPROJ * 2;            // [c8][CPROJ][16 B]
    static constexpr int ONES_BYTES = 2 * 128 * 16;
    static constexpr int BM1_BYTES = 2 * CMID * 16, BM2_BYTES = 2 * CPROJ * 16;
    static constexpr int EPI_WARPS = 16, THREADS = 128 + 32 * EPI_WARPS;
    static constexpr int PARTS = 4, COLS = CMID / PARTS;         // accumulator columns per epilogue warp
    static constexpr int D2_COL = 2 * CMID, D2_STRIDE = 64;
    static constexpr size_t SMEM = 128 + 2 * (size_t)TILE_ALLOC + (size_t)NW * WSLICE + W2_BYTES + ONES_BYTES + BM1_BYTES + BM2_BYTES + 256;
    static_assert(CIN % 16 == 0 && CMID % 64 == 0 && CPROJ % 16 == 0 && CPROJ == CIN && CPROJ <= 64, "shape");
    static_assert(COLS % 16 == 0 && CPROJ / 16 <= PARTS, "column split");
    static_assert(2 * CMID + 2 * D2_STRIDE <= 512, "TMEM columns");
    static_assert(SMEM <= 232448, "shared memory");
    static_assert(WSLICE % 16 == 0, "bulk copy size");
};

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct ConvWsParams {
    const __half* in;     // [n][H][W][CIN]  (also the residual)
    __half* out;          // [n][H][W][CPROJ]
    const __half* w1t;    // [9][CIN/8][CMID][8]  tap-major, BN folded
    const float* b1;      // [CMID]
    const __half* w2;     // [CPROJ][CMID]
    const float* b2;      // [CPROJ]
    int n, H, W;
    int tiles_x, tiles_y;
};

template <int CIN, int CMID, int CPROJ>
__global__ void __launch_bounds__((ConvWsCfg<CIN, CMID, CPROJ>::THREADS), 1)
conv_tc_ws_kernel(const __grid_constant__ CUtensorMap tmIn, ConvWsParams p) {
    using Cfg = ConvWsCfg<CIN, CMID, CPROJ>;
    constexpr int NT = Cfg::THREADS;
    extern __shared__ unsigned char cws_smem_raw[];
    pdl_trigger();
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(cws_smem_raw) + 127) & ~uintptr_t(127));
    unsigned char* s_tile = smem;                                   // [2][TILE_ALLOC]
    unsigned char* s_wr = s_tile + 2 * Cfg::TILE_ALLOC;             // [NW][CIN/8][CMID][16 B]
    unsigned char* s_w2 = s_wr + Cfg::NW * Cfg::WSLICE;             // [CMID/8][CPROJ][16 B]
    unsigned char* s_ones = s_w2 + Cfg::W2_BYTES;
    unsigned char* s_bm1 = s_ones + Cfg::ONES_BYTES;
    unsigned char* s_bm2 = s_bm1 + Cfg::BM1_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bm2 + Cfg::BM2_BYTES);
    uint64_t* tile_full = bars;          // [2]
    uint64_t* tile_empty = bars + 2;     // [2]
    uint64_t* w_full = bars + 4;         // [NW]
    uint64_t* w_empty = bars + 8;        // [NW]
    uint64_t* d1_full = bars + 12;       // [2]
    uint64_t* d1_empty = bars + 14;      // [2]
    uint64_t* p_full = bars + 16;        // [2]
    uint64_t* d2_full = bars + 18;       // [2]
    uint64_t* d2_empty = bars + 20;      // [2]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 22);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int ntiles = tiles_per_img * p.n;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tile_full[i], 1); mbar_init(&tile_empty[i], 1);
            mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], Cfg::EPI_WARPS);
            mbar_init(&p_full[i], Cfg::EPI_WARPS);
            mbar_init(&d2_full[i], 1); mbar_init(&d2_empty[i], Cfg::EPI_WARPS);
        }
        for (int i = 0; i < Cfg::NW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmIn) : "memory");
    }
    for (int i = tid; i < (CMID / 8) * CPROJ; i += NT) {
        const int nrow = i % CPROJ, c8 = i / CPROJ;
        reinterpret_cast<uint4*>(s_w2)[i] = __ldg(reinterpret_cast<const uint4*>(p.w2 + (size_t)nrow * CMID + c8 * 8));
    }
    for (int i = tid; i < 2 * 128; i += NT)
        reinterpret_cast<uint4*>(s_ones)[i] = (i < 128) ? make_uint4(0x3C003C00u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < 2 * CMID; i += NT) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (i < CMID) {
            const float b = __ldg(p.b1 + i);
            const __half hi = __float2half_rn(b), lo = __float2half_rn(b - __half2float(hi));
            v.x = (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
        }
        reinterpret_cast<uint4*>(s_bm1)[i] = v;
    }
    for (int i = tid; i < 2 * CPROJ; i += NT) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (i < CPROJ) {
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
        nm = (rows * Cfg::PW + 127) / 128;
    };

    if (warp == 0) {
        // ================= halo-tile producer (whole warp converged, one elected lane issues) =================
        int i = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
            int n, y0, x0, nm;
            tile_geom(tile, n, y0, x0, nm);
            const int buf = i & 1;
            mbar_wait(&tile_empty[buf], (((uint32_t)i >> 1) & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&tile_full[buf], (uint32_t)Cfg::TILE_BYTES);
                tma_load_5d(s_tile + (size_t)buf * Cfg::TILE_ALLOC, &tmIn, &tile_full[buf], 0, x0 - 1, y0 - 1, 0, n);
            }
            __syncwarp();
        }
    } else if (warp == 2) {
        // ================= weight-slice producer: one tap (all K of it) per ring slot, 9 slots per M tile =================
        uint32_t wi = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int n, y0, x0, nm;
            tile_geom(tile, n, y0, x0, nm);
            for (int m = 0; m < nm; ++m)
                for (int rs = 0; rs < 9; ++rs, ++wi) {
                    const uint32_t slot = wi % Cfg::NW;
                    mbar_wait(&w_empty[slot], ((wi / Cfg::NW) & 1) ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&w_full[slot], (uint32_t)Cfg::WSLICE);
                        bulk_load_1d(s_wr + (size_t)slot * Cfg::WSLICE, reinterpret_cast<const unsigned char*>(p.w1t) + (size_t)rs * Cfg::WSLICE,
                                     (uint32_t)Cfg::WSLICE, &w_full[slot]);
                    }
                    __syncwarp();
                }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (whole warp converged, one elected lane issues) =================
        const uint32_t idesc1 = tc_idesc(128, CMID), idesc2 = tc_idesc(128, CPROJ);
        const uint32_t wra = smem_u32(s_wr), w2a = smem_u32(s_w2);
        const uint64_t ones_desc = tc_desc_nosw(smem_u32(s_ones), 128 * 16);
        const uint64_t bm1_desc = tc_desc_nosw(smem_u32(s_bm1), CMID * 16), bm2_desc = tc_desc_nosw(smem_u32(s_bm2), CPROJ * 16);
        auto mma2 = [&](int u) {             // projection of M tile u: D2 = P . W2^T + b2, P = packed fp16 in the low half of D1[u & 1]
            const int a = u & 1;
            const uint32_t ph = ((uint32_t)u >> 1) & 1;
            mbar_wait(&p_full[a], ph);
            mbar_wait(&d2_empty[a], ph ^ 1);
            tc_fence_after();
            const uint32_t d2 = tmem_base + Cfg::D2_COL + a * Cfg::D2_STRIDE;
            if (elect_one()) {
                tc_mma_f16(d2, ones_desc, bm2_desc, idesc2, 0);
#pragma unroll
                for (int kk = 0; kk < CMID / 16; ++kk)
                    tc_mma_f16_ts(d2, tmem_base + a * CMID + kk * 8, tc_desc_nosw(w2a + kk * 2 * (CPROJ * 16), CPROJ * 16), idesc2, 1);
                tc_commit(&d2_full[a]);
            }
            __syncwarp();
        };
        int i = 0, t = 0;
        uint32_t wi = 0;
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
                if (elect_one()) tc_mma_f16(d1, ones_desc, bm1_desc, idesc1, 0);           // D1 = bias
                __syncwarp();
                for (int rs = 0; rs < 9; ++rs, ++wi) {
                    const uint32_t slot = wi % Cfg::NW;
                    mbar_wait(&w_full[slot], (wi / Cfg::NW) & 1);
                    tc_fence_after();
                    const int r = rs / 3, s = rs - r * 3;
                    const uint32_t a_addr = ta + (uint32_t)(m * 128 + r * Cfg::PW + s) * 16u;
                    const uint32_t b_addr = wra + slot * Cfg::WSLICE;
                    if (elect_one()) {
#pragma unroll
                        for (int kc = 0; kc < CIN / 16; ++kc)
                            tc_mma_f16(d1, tc_desc_nosw(a_addr + kc * 2 * Cfg::PLANE, Cfg::PLANE),
                                       tc_desc_nosw(b_addr + kc * 2 * (CMID * 16), CMID * 16), idesc1, 1);
                        tc_commit(&w_empty[slot]);                   // slot reusable once these MMAs have read it
                        if (rs == 8) tc_commit(&d1_full[a]);
                    }
                    __syncwarp();
                }
                if (t >= 1) mma2(t - 1);
            }
            if (elect_one()) tc_commit(&tile_empty[buf]);            // every MMA that reads this halo tile has been issued
            __syncwarp();
        }
        if (t >= 1) mma2(t - 1);
    } else if (warp >= 4) {
        // ================= epilogue: 4 lane quadrants x 4 column parts =================
        const int q = warp & 3;
        const int part = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        constexpr int NG2 = CPROJ / 16;
        constexpr int NCH = Cfg::COLS / 16;             // 16-column chunks per warp (3)
        uint32_t rv[8], rv_next[8];
        auto epi2 = [&](int u, bool okp, size_t pixp) {
            const int a = u & 1;
            mbar_wait(&d2_full[a], ((uint32_t)u >> 1) & 1);
            tc_fence_after();
            uint32_t v2[16];
            if (part < NG2) tc_ld16(t_row + (uint32_t)(Cfg::D2_COL + a * Cfg::D2_STRIDE + part * 16), v2);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d2_empty[a]);
            if (okp && part < NG2) {
                uint32_t pk[8];
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                    const float2 r2 = unpack_half2(rv[h]);
                    pk[h] = pack_half2(__uint_as_float(v2[2 * h]) + r2.x, __uint_as_float(v2[2 * h + 1]) + r2.y);
                }
                st_global_v8(p.out + pixp * CPROJ + part * 16, pk);
            }
        };
        int t = 0;
        bool ok_prev = false;
        size_t pix_prev = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int n, y0, x0, nm;
            tile_geom(tile, n, y0, x0, nm);
            for (int m = 0; m < nm; ++m, ++t) {
                const int a = t & 1;
                const int lp = m * 128 + row;
                const int ry = lp / Cfg::PW, cx = lp - ry * Cfg::PW;
                const int oy = y0 + ry, ox = x0 + cx;
                const bool ok = (cx < Cfg::TW) && (ry < Cfg::TH) && (oy < p.H) && (ox < p.W);
                const size_t pix = ((size_t)n * p.H + (ok ? oy : 0)) * p.W + (ok ? ox : 0);
                if (ok && part < NG2) ld_global_v8(p.in + pix * CIN + part * 16, rv_next);      // residual, a whole E1 phase ahead

                // ---- E1: every warp loads ALL its accumulator columns, the 16 warps meet, then P overwrites the low half ----
                mbar_wait(&d1_full[a], ((uint32_t)t >> 1) & 1);
                tc_fence_after();
                uint32_t v[NCH][16];
#pragma unroll
                for (int j = 0; j < NCH; ++j) tc_ld16(t_row + (uint32_t)(a * CMID + part * Cfg::COLS + j * 16), v[j]);
                tc_wait_ld();
                tc_fence_before();
                asm volatile("bar.sync 1, %0;" ::"n"(32 * Cfg::EPI_WARPS) : "memory");
                tc_fence_after();
                if (lane == 0) mbar_arrive(&d1_empty[a]);
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    uint32_t pk[8];
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        float x4[4] = {__uint_as_float(v[j][4 * h]), __uint_as_float(v[j][4 * h + 1]), __uint_as_float(v[j][4 * h + 2]),
                                       __uint_as_float(v[j][4 * h + 3])};
                        silu4(x4);
                        pk[2 * h] = pack_half2(x4[0], x4[1]);
                        pk[2 * h + 1] = pack_half2(x4[2], x4[3]);
                    }
                    tc_st8(t_row + (uint32_t)(a * CMID + (part * Cfg::COLS + j * 16) / 2), pk);
                }
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[a]);
                // ---- E2 of the previous M tile: its projection ran on the tensor core during this tile's E1 ----
                if (t >= 1) epi2(t - 1, ok_prev, pix_prev);
                ok_prev = ok; pix_prev = pix;
#pragma unroll
                for (int h = 0; h < 8; ++h) rv[h] = rv_next[h];
            }
        }
        if (t >= 1) epi2(t - 1, ok_prev, pix_prev);
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
