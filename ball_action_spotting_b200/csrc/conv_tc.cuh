// K2 (+K3) on the 5th-gen tensor cores, every dense 3x3 block of the encoder:
//   ConvBnAct    (timm blocks.0.0):                 y = SiLU(conv3x3(x) * s + b)
//   EdgeResidual (timm blocks.1.0 / 1.1 / 2.0):     y = conv1x1( SiLU( conv3x3_stride(x) * s1 + b1 ) ) * s2 + b2 (+ x)
// as ONE persistent, warp-specialised tcgen05 kernel (reference call site: multidim_stacker.py:166-176, the timm blocks
// it builds).  The expanded tensor never leaves the SM, and with PT = true it never leaves TENSOR MEMORY: the SiLU
// epilogue writes it back to TMEM as packed fp16 (tcgen05.st) and the projection MMA takes its A operand from there.
//
// Implicit GEMM without im2col.  The halo tile is loaded by TMA as 8-channel PLANES, plane[c8][pixel][8 ch], pixels in
// row-major order of the tile (5-D tensor map over NHWC seen as [n][C/8][H][W][8]; hardware zero fill = the conv padding).
// In the no-swizzle K-major UMMA layout a row is 16 bytes and 8-row groups are 128 bytes apart, so "row m of the A
// operand" is "pixel p0 + m of the linearised tile" and a tap of the stencil is only a start-address offset.
//   stride 1: one plane set of (TW+2) x (TH+2) pixels; tap (r, s) = offset r * (TW+2) + s; 2 garbage rows of M per tile row
//   stride 2: the input is read as its four (row parity, column parity) PHASES, each a dense (TW+1) x (TH+1) image
//             (four tensor maps with element strides doubled and the base moved by one row / pixel; TF-SAME on even sizes
//             pads bottom / right only = TMA zero fill); tap (r, s) = phase (r&1, s&1) at offset (r>>1)*(TW+1) + (s>>1)
//   warp 0 / lane 0 : TMA producer (double-buffered halo tiles)
//   warp 1 / lane 0 : MMA issuer: per 128-pixel M tile 9 x CIN/16 tcgen05.mma (N = CMID) + one against a "ones" tile that
//                     adds the bias; later CMID/16 MMAs (N = CPROJ) for the projection
//   other warps     : epilogue (4 TMEM lane quadrants x up to 4 column parts): E1 = TMEM -> SiLU -> fp16 -> TMEM / smem
//                     (A operand of the projection), E2 = TMEM -> + residual -> fp16 -> 32-byte global stores.  Every warp
//                     streams 8-column chunks with the next tcgen05.ld in flight, E2 of M tile t-1 comes after E1 of tile t
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "tc_helpers.cuh"
#include "gemm_tc.cuh"

// Timeline instrumentation, compiled out of the product build (tools/conv_trace.sh builds a separate library with it)
#ifdef MDS_CONV_TRACE
#define CTC_TRACE(T_, ev_) do { if (p.trace && blockIdx.x == 0 && (T_) < 64 && (threadIdx.x & 31) == 0) p.trace[(T_) * 16 + (ev_)] = clock64(); } while (0)
#else
#define CTC_TRACE(T_, ev_) do { } while (0)
#endif

#ifndef MDS_CONV_SILU4
#define MDS_CONV_SILU4 1
#endif

namespace mds {

struct ConvTcMaps { CUtensorMap m[4]; };     // stride 1: m[0]; stride 2: phase (py, px) = m[py * 2 + px]

struct ConvTcParams {
    const __half* in;     // [n][H][W][CIN]  (also the residual)
    __half* out;          // [n][Ho][Wo][COUT],  COUT = CPROJ ? CPROJ : CMID
    const __half* w1;     // [CMID][9*CIN]   k = (r*3+s)*CIN + ci, BN folded
    const float* b1;      // [CMID]
    const __half* w2;     // [CPROJ][CMID]
    const float* b2;      // [CPROJ]
    int n, H, W, Ho, Wo;
    int tiles_x, tiles_y;
    long long* trace;     // -DMDS_CONV_TRACE builds only: clock64 stamps of CTA 0, [M tile < 64][16 events] (tools/conv_trace.sh)
};

template <int CIN, int CMID, int CPROJ, int STRIDE, bool RES, int TH_, bool PT, bool FOLD = false>
struct ConvTcCfg {
    // FOLD (stride 1, no projection, small CMID): the three COLUMN taps are folded into N.  One MMA per (row tap, K step) computes
    // D[pixel][(s, co)] for s = 0..2 (N = 3 * CMID) from the UNSHIFTED pixel rows, and the epilogue adds the three column groups
    // of lanes l, l+1, l+2 (warp shuffles).  tcgen05.mma costs >= ~32-39 clocks per instruction whatever N is (the 128 x 16 A block
    // is fetched from shared memory every time), so for CMID = 16 this is 7 MMAs per M tile instead of 19.  Tile rows are exactly
    // 32 pixels (30 outputs + 2 halo) = one TMEM lane quadrant = one warp, so lane l + s never leaves the warp.
    static constexpr int TW = FOLD ? 30 : 32, TH = TH_;
    static constexpr int NPH = STRIDE == 1 ? 1 : 4;
    static constexpr int PW = STRIDE == 1 ? TW + 2 : TW + 1;
    static constexpr int PH = STRIDE == 1 ? TH + 2 : TH + 1;
    static constexpr int PIX = PW * PH;                          // pixels per plane
    static constexpr int PLANE = PIX * 16;                       // bytes per 8-channel plane
    static constexpr int PHASE_BYTES = (CIN / 8) * PLANE;
    static constexpr int TILE_BYTES = NPH * PHASE_BYTES;
    static constexpr int MT_MAX = (TH * PW + 127) / 128;         // M tiles of 128 linear pixels
    static constexpr int MAX_OFF = STRIDE == 1 ? 2 * PW + 2 : PW + 1;
    static constexpr int OVER = MT_MAX * 128 + MAX_OFF > PIX ? MT_MAX * 128 + MAX_OFF - PIX : 0;   // pixels the last M tile over-reads
    static constexpr int TILE_ALLOC = ((TILE_BYTES + OVER * 16 + 127) / 128) * 128;
    static constexpr int GROUPS = 2;                             // epilogue groups; group e takes the M tiles t with t % 2 == e, so one
                                                                 // group's SFU-bound E1 phase covers the other's latency-bound phases
    static constexpr int PARTS = CMID >= 32 ? 2 : 1;             // warps of a group that share a TMEM lane quadrant (column parts)
    static constexpr int ND1 = CPROJ > 0 ? 2 : 4;                // conv accumulators in flight (buffer t % ND1; P and D2: t % 2 = group)
    static constexpr int EPI_WARPS = 4 * PARTS * GROUPS;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
    static constexpr int N1 = FOLD ? 3 * CMID : CMID;            // N of the conv MMAs
    static constexpr int W1_BYTES = 9 * CIN * CMID * 2;          // [tap][c8][n][8]  (FOLD: [row tap][c8][s * CMID + co][8])
    static constexpr int W2_BYTES = CMID * CPROJ * 2;            // [c8][n][8]
    static constexpr int P_BYTES = (CPROJ > 0 && !PT) ? 128 * CMID * 2 : 0;   // [c8][row][8], one per accumulator
    static constexpr int ONES_BYTES = 2 * 128 * 16;              // [2 planes][128 rows][8]: (row, k=0,1) = 1
    static constexpr int BM1_BYTES = 2 * N1 * 16;                // [2 planes][N1][8]: (n, k=0,1) = bias hi/lo
    static constexpr int BM2_BYTES = 2 * CPROJ * 16;
    // tensor memory columns: D1[ND1] | P[2] (packed fp16, PT only) | D2[2]
    static constexpr int D1_STRIDE = N1 <= 32 ? 32 : N1 <= 64 ? 64 : N1;
    static constexpr int P_COL = ND1 * D1_STRIDE;
    static constexpr int D2_COL = P_COL + (CPROJ > 0 ? CMID : 0);
    static constexpr int D2_STRIDE = 64;
    static constexpr int TMEM_USED = D2_COL + (CPROJ > 0 ? 2 * D2_STRIDE : 0);
    static constexpr int TMEM_COLS = TMEM_USED <= 32 ? 32 : TMEM_USED <= 64 ? 64 : TMEM_USED <= 128 ? 128 : TMEM_USED <= 256 ? 256 : 512;
    static constexpr size_t SMEM = 128 + 2 * (size_t)TILE_ALLOC + W1_BYTES + W2_BYTES + 2 * P_BYTES + ONES_BYTES + BM1_BYTES +
                                   BM2_BYTES + 256;
    static_assert(CIN % 16 == 0 && CMID % 16 == 0 && CPROJ % 16 == 0, "channel counts must be multiples of 16");
    static_assert(!RES || (STRIDE == 1 && CPROJ == CIN), "residual needs same shape");
    static_assert(PHASE_BYTES % 128 == 0, "TMA destinations are 128-byte aligned");
    static_assert(TMEM_USED <= 512, "TMEM columns");
    static_assert(CPROJ <= 64, "projection accumulator stride");
    static_assert(CPROJ == 0 || CPROJ / 16 <= 2 * PARTS, "at most two projection column groups per epilogue part");
    static_assert((CMID / 16) % PARTS == 0, "column groups split evenly over the parts");
    static_assert(SMEM <= 232448, "shared memory");
    static_assert(!FOLD || (STRIDE == 1 && CPROJ == 0 && PW == 32 && N1 % 16 == 0 && N1 <= 256), "FOLD: stride 1, no projection");
};

__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int CIN, int CMID, int CPROJ, int STRIDE, bool RES, int TH_, bool PT, int MINB, bool FOLD = false>
__global__ void __launch_bounds__((ConvTcCfg<CIN, CMID, CPROJ, STRIDE, RES, TH_, PT, FOLD>::THREADS), MINB)
conv_tc_kernel(const __grid_constant__ ConvTcMaps maps, ConvTcParams p) {
    using Cfg = ConvTcCfg<CIN, CMID, CPROJ, STRIDE, RES, TH_, PT, FOLD>;
    constexpr int NT = Cfg::THREADS;
    constexpr int COUT = CPROJ ? CPROJ : CMID;
    constexpr bool PROJ = CPROJ > 0;
    extern __shared__ unsigned char ctc_smem_raw[];
    pdl_trigger();
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ctc_smem_raw) + 127) & ~uintptr_t(127));
    unsigned char* s_tile = smem;                                   // [2][TILE_ALLOC]
    unsigned char* s_w1 = s_tile + 2 * Cfg::TILE_ALLOC;             // [9][CIN/8][CMID][16 B]
    unsigned char* s_w2 = s_w1 + Cfg::W1_BYTES;                     // [CMID/8][CPROJ][16 B]
    unsigned char* s_p = s_w2 + Cfg::W2_BYTES;                      // [2][CMID/8][128][16 B]   (PT = false)
    unsigned char* s_ones = s_p + 2 * Cfg::P_BYTES;
    unsigned char* s_bm1 = s_ones + Cfg::ONES_BYTES;
    unsigned char* s_bm2 = s_bm1 + Cfg::BM1_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bm2 + Cfg::BM2_BYTES);
    uint64_t* tile_full = bars;          // [2]
    uint64_t* tile_empty = bars + 2;     // [2]
    uint64_t* d1_full = bars + 4;        // [ND1 <= 4]
    uint64_t* d1_empty = bars + 8;       // [ND1 <= 4]
    uint64_t* p_full = bars + 12;        // [2]
    uint64_t* d2_full = bars + 14;       // [2]
    uint64_t* d2_empty = bars + 16;      // [2]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 18);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int ntiles = tiles_per_img * p.n;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tile_full[i], 1); mbar_init(&tile_empty[i], 1);
            mbar_init(&p_full[i], Cfg::EPI_WARPS / Cfg::GROUPS);
            mbar_init(&d2_full[i], 1); mbar_init(&d2_empty[i], Cfg::EPI_WARPS / Cfg::GROUPS);
        }
        for (int i = 0; i < Cfg::ND1; ++i) { mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], Cfg::EPI_WARPS / Cfg::GROUPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (int i = 0; i < Cfg::NPH; ++i) asm volatile("prefetch.tensormap [%0];" ::"l"(&maps.m[i]) : "memory");
    }
    // ---- weights, ones tile and bias tiles -> smem in the no-swizzle K-major plane layout (once per CTA) ----
    for (int i = tid; i < 9 * (CIN / 8) * CMID; i += NT) {                  // dest chunk (tap, c8, n)
        if constexpr (FOLD) {                                                // dest chunk (r, c8, n = s * CMID + co)
            const int nn = i % Cfg::N1, pl = i / Cfg::N1;
            const int sx = nn / CMID, co = nn - sx * CMID, r = pl / (CIN / 8), c8 = pl - r * (CIN / 8);
            reinterpret_cast<uint4*>(s_w1)[i] = __ldg(reinterpret_cast<const uint4*>(p.w1 + (size_t)co * 9 * CIN + (r * 3 + sx) * CIN + c8 * 8));
        } else {
            const int nrow = i % CMID, pl = i / CMID;                        // pl = tap * (CIN/8) + c8
            reinterpret_cast<uint4*>(s_w1)[i] = __ldg(reinterpret_cast<const uint4*>(p.w1 + (size_t)nrow * 9 * CIN + pl * 8));
        }
    }
    if constexpr (PROJ) {
        for (int i = tid; i < (CMID / 8) * CPROJ; i += NT) {
            const int nrow = i % CPROJ, c8 = i / CPROJ;
            reinterpret_cast<uint4*>(s_w2)[i] = __ldg(reinterpret_cast<const uint4*>(p.w2 + (size_t)nrow * CMID + c8 * 8));
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
    }
    for (int i = tid; i < 2 * 128; i += NT)
        reinterpret_cast<uint4*>(s_ones)[i] = (i < 128) ? make_uint4(0x3C003C00u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < 2 * Cfg::N1; i += NT) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (i < CMID) {                                  // FOLD: the bias goes to the s = 0 column group only
            const float b = __ldg(p.b1 + i);
            const __half hi = __float2half_rn(b), lo = __float2half_rn(b - __half2float(hi));
            v.x = (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
        }
        reinterpret_cast<uint4*>(s_bm1)[i] = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
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
        y0 = ty * Cfg::TH; x0 = tx * Cfg::TW;                    // output coordinates
        const int rows = min(Cfg::TH, p.Ho - y0);
        nm = (rows * Cfg::PW + 127) / 128;                       // M tiles that contain valid output pixels
    };

    if (warp == 0) {
        // ================= TMA producer (whole warp converged, one elected lane issues) =================
        {
            int i = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
                int n, y0, x0, nm;
                tile_geom(tile, n, y0, x0, nm);
                const int buf = i & 1;
                mbar_wait(&tile_empty[buf], (((uint32_t)i >> 1) & 1) ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&tile_full[buf], (uint32_t)Cfg::TILE_BYTES);
                    unsigned char* dst = s_tile + (size_t)buf * Cfg::TILE_ALLOC;
                    if constexpr (STRIDE == 1) {
                        tma_load_5d(dst, &maps.m[0], &tile_full[buf], 0, x0 - 1, y0 - 1, 0, n);
                    } else {
#pragma unroll
                        for (int ph = 0; ph < 4; ++ph)
                            tma_load_5d(dst + (size_t)ph * Cfg::PHASE_BYTES, &maps.m[ph], &tile_full[buf], 0, x0, y0, 0, n);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The WHOLE warp runs the control flow (waits, loops) converged and one elected lane issues: a tcgen05.mma under
        // `if (lane == 0)` makes ptxas wrap every UTCHMMA in an election loop (ELECT / BRA.U.ANY), measured at 62-94 clocks per
        // MMA whatever its shape; under elect.sync it is a single predicated instruction (tools/mma_probe.cu: 39 clocks at N = 16).
        {
            const uint32_t idesc1 = tc_idesc(128, Cfg::N1);
            const uint32_t w1a = smem_u32(s_w1), onesa = smem_u32(s_ones);
            const uint64_t ones_desc = tc_desc_nosw(onesa, 128 * 16);
            const uint64_t bm1_desc = tc_desc_nosw(smem_u32(s_bm1), Cfg::N1 * 16);
            auto mma2 = [&](int u) {             // projection of M tile u: D2 = P . W2^T + b2
                if constexpr (PROJ) {
                    const uint32_t idesc2 = tc_idesc(128, CPROJ);
                    const uint32_t w2a = smem_u32(s_w2);
                    const uint64_t bm2_desc = tc_desc_nosw(smem_u32(s_bm2), CPROJ * 16);
                    const int a = u & 1;
                    const uint32_t ph = ((uint32_t)u >> 1) & 1;
                    CTC_TRACE(u, 4);
                    mbar_wait(&p_full[a], ph);
                    CTC_TRACE(u, 5);
                    mbar_wait(&d2_empty[a], ph ^ 1);
                    tc_fence_after();
                    CTC_TRACE(u, 6);
                    const uint32_t d2 = tmem_base + Cfg::D2_COL + a * Cfg::D2_STRIDE;
                    if (elect_one()) {
                        tc_mma_f16(d2, ones_desc, bm2_desc, idesc2, 0);
#pragma unroll
                        for (int kk = 0; kk < CMID / 16; ++kk) {
                            const uint64_t bdesc = tc_desc_nosw(w2a + kk * 2 * (CPROJ * 16), CPROJ * 16);
                            if constexpr (PT) {
                                tc_mma_f16_ts(d2, tmem_base + Cfg::P_COL + a * (CMID / 2) + kk * 8, bdesc, idesc2, 1);
                            } else {
                                tc_mma_f16(d2, tc_desc_nosw(smem_u32(s_p) + a * Cfg::P_BYTES + kk * 2 * (128 * 16), 128 * 16), bdesc, idesc2, 1);
                            }
                        }
                        tc_commit(&d2_full[a]);
                    }
                    __syncwarp();
                    CTC_TRACE(u, 7);
                }
            };
            int i = 0, t = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
                int n, y0, x0, nm;
                tile_geom(tile, n, y0, x0, nm);
                const int buf = i & 1;
                CTC_TRACE(t, 3);
                mbar_wait(&tile_full[buf], ((uint32_t)i >> 1) & 1);
                tc_fence_after();
                const uint32_t ta = smem_u32(s_tile + (size_t)buf * Cfg::TILE_ALLOC);
                for (int m = 0; m < nm; ++m, ++t) {
                    const int a = t % Cfg::ND1;
                    CTC_TRACE(t, 0);
                    mbar_wait(&d1_empty[a], (((uint32_t)t / Cfg::ND1) & 1) ^ 1);
                    tc_fence_after();
                    CTC_TRACE(t, 1);
                    const uint32_t d1 = tmem_base + a * Cfg::D1_STRIDE;
                    if (elect_one()) {
                        tc_mma_f16(d1, ones_desc, bm1_desc, idesc1, 0);                     // D1 = bias
                        if constexpr (FOLD) {
#pragma unroll
                            for (int r = 0; r < 3; ++r) {
                                const uint32_t a_addr = ta + (uint32_t)(m * 128 + r * Cfg::PW) * 16u;
#pragma unroll
                                for (int kc = 0; kc < CIN / 16; ++kc)
                                    tc_mma_f16(d1, tc_desc_nosw(a_addr + kc * 2 * Cfg::PLANE, Cfg::PLANE),
                                               tc_desc_nosw(w1a + (r * (CIN / 8) + kc * 2) * (Cfg::N1 * 16), Cfg::N1 * 16), idesc1, 1);
                            }
                        } else {
#pragma unroll
                        for (int rsi = 0; rsi < 9; ++rsi) {
                            const int rs = (MDS_NUMERICS_VARIANT & 2) ? 8 - rsi : rsi;
                            const int r = rs / 3, s = rs - r * 3;
                            const int phase = STRIDE == 1 ? 0 : (r & 1) * 2 + (s & 1);
                            const int off = STRIDE == 1 ? r * Cfg::PW + s : (r >> 1) * Cfg::PW + (s >> 1);
                            const uint32_t a_addr = ta + (uint32_t)(phase * Cfg::PHASE_BYTES) + (uint32_t)(m * 128 + off) * 16u;
#pragma unroll
                            for (int kc = 0; kc < CIN / 16; ++kc)
                                tc_mma_f16(d1, tc_desc_nosw(a_addr + kc * 2 * Cfg::PLANE, Cfg::PLANE),
                                           tc_desc_nosw(w1a + (rs * (CIN / 8) + kc * 2) * (CMID * 16), CMID * 16), idesc1, 1);
                        }
                        }
                        tc_commit(&d1_full[a]);
                    }
                    __syncwarp();
                    CTC_TRACE(t, 2);
                    if (t >= 1) mma2(t - 1);
                }
                if (elect_one()) tc_commit(&tile_empty[buf]);            // every MMA that reads this halo tile has been issued
                __syncwarp();
            }
            if (t >= 1) mma2(t - 1);
        }
    } else {
        // ================= epilogue: GROUPS x 4 lane quadrants x PARTS column parts =================
        // Group e takes the M tiles t with t % 2 == e, so one group's SFU-bound E1 phase covers the other's waits.  Each warp
        // streams its columns of an M tile in 8-column chunks: the tcgen05.ld of chunk c+1 is in flight while chunk c goes through
        // SiLU (a tcgen05.ld + wait::ld round trip is ~180 clocks whatever its width, tools/tmem_probe.cu).  Order per M tile:
        // E1(t) [D1 -> SiLU -> P], E2 of the group's previous tile [D2 + residual -> global; its projection ran on the tensor core
        // meanwhile, and the projection of THIS tile waits for that accumulator], then the first chunk of the group's next tile.
        const int q = warp & 3;                         // TMEM lane quadrant
        const int ew = (warp - 2) >> 2;                 // 0 .. GROUPS * PARTS - 1
        const int grp = ew / Cfg::PARTS, part = ew - grp * Cfg::PARTS;
        const int row = q * 32 + lane;                  // row of the M tile = linear tile pixel
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        constexpr int NG2 = CPROJ / 16;                 // 16-column groups of the projection
        constexpr int NJ = NG2 > Cfg::PARTS ? 2 : 1;    // ... per part: part g takes the groups g and g + PARTS
        constexpr int CW = 8;                           // columns per chunk
        constexpr int COLS = CMID / Cfg::PARTS;         // columns per warp
        constexpr int NCH = COLS / CW;                  // chunks per warp and M tile (even)
        static_assert(NCH % 2 == 0, "chunks are consumed in pairs (16 channels = one 32-byte store)");
        uint32_t rv[NJ][8], rv_next[NJ][8];    // residual of the M tile in E2 / of the tile in E1 (requested a whole E1 phase ahead)
        auto epi2 = [&](int u, bool okp, size_t pixp) {
            if constexpr (PROJ) {
                const int a = u & 1;
                mbar_wait(&d2_full[a], ((uint32_t)u >> 1) & 1);
                tc_fence_after();
                if (warp == 2) CTC_TRACE(u, 14);
                uint32_t v2[NJ][16];
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int g = part + Cfg::PARTS * j;
                    if (g < NG2) tc_ld16(t_row + (uint32_t)(Cfg::D2_COL + a * Cfg::D2_STRIDE + g * 16), v2[j]);
                }
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d2_empty[a]);
                if (okp) {
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const int g = part + Cfg::PARTS * j;
                        if (g < NG2) {
                            uint32_t pk[8];
#pragma unroll
                            for (int h = 0; h < 8; ++h) {
                                float x0f = __uint_as_float(v2[j][2 * h]), x1f = __uint_as_float(v2[j][2 * h + 1]);
                                if constexpr (RES) {
                                    const float2 r2 = unpack_half2(rv[j][h]);
                                    x0f += r2.x; x1f += r2.y;
                                }
                                pk[h] = pack_half2(x0f, x1f);
                            }
                            st_global_v8(p.out + pixp * COUT + g * 16, pk);
                        }
                    }
                }
            }
        };
        // iterate over the M tiles (tile, m) of this CTA; T = running index over all of them, this group owns T % GROUPS == grp
        int tile = blockIdx.x, m = 0, T = 0, n = 0, y0 = 0, x0 = 0, nm = 0;
        bool have = tile < ntiles;
        if (have) tile_geom(tile, n, y0, x0, nm);
        auto step = [&]() {          // advance (tile, m, T) to the next M tile of the CTA
            ++T;
            if (++m == nm) {
                tile += gridDim.x; m = 0;
                have = tile < ntiles;
                if (have) tile_geom(tile, n, y0, x0, nm);
            }
        };
        while (have && T % Cfg::GROUPS != grp) step();
        if constexpr (FOLD) {
            // D[pixel][(s, co)]: out(lane l) = D[l][(0, .)] + D[l+1][(1, .)] + D[l+2][(2, .)], lanes = the 32 pixels of one tile row
            while (have) {
                const int a = T % Cfg::ND1;
                mbar_wait(&d1_full[a], ((uint32_t)T / Cfg::ND1) & 1);
                tc_fence_after();
                uint32_t v[3][CMID];
#pragma unroll
                for (int sx = 0; sx < 3; ++sx)
#pragma unroll
                    for (int g = 0; g < CMID / 16; ++g)
                        tc_ld16(t_row + (uint32_t)(a * Cfg::D1_STRIDE + sx * CMID + g * 16), *reinterpret_cast<uint32_t(*)[16]>(&v[sx][g * 16]));
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d1_empty[a]);
                const int lp = m * 128 + row;
                const int ry = lp / Cfg::PW, cx = lp - ry * Cfg::PW;
                const int oy = y0 + ry, ox = x0 + cx;
                const bool ok = (cx < Cfg::TW) && (ry < Cfg::TH) && (oy < p.Ho) && (ox < p.Wo);
                const size_t pix = ((size_t)n * p.Ho + (ok ? oy : 0)) * p.Wo + (ok ? ox : 0);
#pragma unroll
                for (int g = 0; g < CMID / 16; ++g) {
                    uint32_t pk[8];
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        float x4[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int co = g * 16 + h * 4 + e;
                            x4[e] = __uint_as_float(v[0][co]) + __shfl_down_sync(0xffffffffu, __uint_as_float(v[1][co]), 1) +
                                    __shfl_down_sync(0xffffffffu, __uint_as_float(v[2][co]), 2);
                        }
                        silu4(x4);
                        pk[2 * h] = pack_half2(x4[0], x4[1]);
                        pk[2 * h + 1] = pack_half2(x4[2], x4[3]);
                    }
                    if (ok) st_global_v8(p.out + pix * COUT + g * 16, pk);
                }
                do step(); while (have && T % Cfg::GROUPS != grp);
            }
        }
        uint32_t buf[2][CW];
        auto ld_first = [&]() {      // wait for the accumulator of M tile T and request its first chunk
            const int a = T % Cfg::ND1;
            if (warp == 2) CTC_TRACE(T, 8);
            mbar_wait(&d1_full[a], ((uint32_t)T / Cfg::ND1) & 1);
            tc_fence_after();
            if (warp == 2) CTC_TRACE(T, 9);
            tc_ld8(t_row + (uint32_t)(a * Cfg::D1_STRIDE + part * COLS), buf[0]);
        };
        if (have) { ld_first(); tc_wait_ld(); }
        int u_prev = -1;
        bool ok_prev = false;
        size_t pix_prev = 0;
        while (have) {
            const int a = T % Cfg::ND1;
            const int lp = m * 128 + row;                                    // linear pixel inside the row-major tile
            const int ry = lp / Cfg::PW, cx = lp - ry * Cfg::PW;
            const int oy = y0 + ry, ox = x0 + cx;
            const bool ok = (cx < Cfg::TW) && (ry < Cfg::TH) && (oy < p.Ho) && (ox < p.Wo);
            const size_t pix = ((size_t)n * p.Ho + (ok ? oy : 0)) * p.Wo + (ok ? ox : 0);
            const int u_cur = T;

            if constexpr (RES) {
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int g = part + Cfg::PARTS * j;
                    if (ok && g < NG2) ld_global_v8(p.in + pix * CIN + g * 16, rv_next[j]);
                }
            }
            // ---- E1: D1 -> SiLU -> fp16 -> P (A operand of the projection) or, without projection, global ----
            if (warp == 2) CTC_TRACE(u_cur, 10);
            uint32_t pk[8];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if (c + 1 < NCH) tc_ld8(t_row + (uint32_t)(a * Cfg::D1_STRIDE + part * COLS + (c + 1) * CW), buf[(c + 1) & 1]);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float x4[4] = {__uint_as_float(buf[c & 1][4 * h]), __uint_as_float(buf[c & 1][4 * h + 1]),
                                   __uint_as_float(buf[c & 1][4 * h + 2]), __uint_as_float(buf[c & 1][4 * h + 3])};
#if MDS_CONV_SILU4
                    silu4(x4);
#else
                    x4[0] = silu_f(x4[0]); x4[1] = silu_f(x4[1]); x4[2] = silu_f(x4[2]); x4[3] = silu_f(x4[3]);
#endif
                    pk[(c & 1) * 4 + 2 * h] = pack_half2(x4[0], x4[1]);
                    pk[(c & 1) * 4 + 2 * h + 1] = pack_half2(x4[2], x4[3]);
                }
                if (c & 1) {                                 // 16 channels ready: columns [col, col + 16)
                    const int col = part * COLS + (c - 1) * CW;
                    if constexpr (!PROJ) {
                        if (ok) st_global_v8(p.out + pix * COUT + col, pk);
                    } else if constexpr (PT) {
                        tc_st8(t_row + (uint32_t)(Cfg::P_COL + a * (CMID / 2) + col / 2), pk);
                    } else {
                        unsigned char* pbuf = s_p + a * Cfg::P_BYTES;
                        *reinterpret_cast<uint4*>(pbuf + (size_t)(col / 8) * (128 * 16) + row * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4*>(pbuf + (size_t)(col / 8 + 1) * (128 * 16) + row * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                }
                if (c + 1 < NCH) tc_wait_ld();
                if (c + 2 == NCH) {                          // the last chunk of this accumulator is in registers
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&d1_empty[a]);
                }
            }
            if constexpr (PROJ) {
                if constexpr (PT) {
                    tc_wait_st();
                    tc_fence_before();
                } else {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // P written by the generic proxy, read by tcgen05.mma
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[a]);
            }
            if (warp == 2) CTC_TRACE(u_cur, 11);
                // E2 of the group's previous M tile first: the projection of THIS tile waits for that accumulator (d2_empty), and the
            // conv MMAs of the group's next tile are issued behind it
            if (PROJ && u_prev >= 0) epi2(u_prev, ok_prev, pix_prev);
            if (warp == 2) CTC_TRACE(u_cur, 12);
            do step(); while (have && T % Cfg::GROUPS != grp);
            if (have) { ld_first(); tc_wait_ld(); }
            if (warp == 2) CTC_TRACE(u_cur, 13);
            u_prev = u_cur; ok_prev = ok; pix_prev = pix;
            if constexpr (RES) {
#pragma unroll
                for (int h = 0; h < 8; ++h)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) rv[j][h] = rv_next[j][h];
            }
        }
        if (PROJ && u_prev >= 0) epi2(u_prev, ok_prev, pix_prev);
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    }
}

}  // namespace mds
