// K2 (+K3 for FusedMBConv): dense 3x3 conv as implicit GEMM on tensor cores, NHWC fp16, fp32 accumulate.
//   ConvBnAct   (timm blocks.0.0):        y = SiLU(conv3x3(x) * s + b)
//   EdgeResidual (timm blocks.1.*, 2.*):  y = conv1x1(SiLU(conv3x3(x)*s1 + b1))*s2 + b2 (+ x)
// The expanded tensor of an EdgeResidual block (up to 15 MB / image) never leaves the SM: the conv3x3
// accumulator fragments are re-used in registers as the A operand of the 1x1 projection GEMM.
// Persistent CTAs: weights live in shared memory for the whole kernel, halo tiles are double-buffered with
// cp.async (zero-fill provides both the symmetric pad-1 of stride-1 convs and TF-SAME (0,1) of stride-2 convs).
#pragma once
#include "common.cuh"

namespace mds {

struct Conv3Params {
    const __half* in;    // [n][H][W][CIN]
    __half* out;         // [n][Ho][Wo][COUT]   COUT = CPROJ ? CPROJ : CMID
    const __half* w1;    // [CMID][9*CIN]  k = (r*3+s)*CIN + ci, BN scale folded
    const float* b1;     // [CMID]
    const __half* w2;    // [CPROJ][CMID]
    const float* b2;     // [CPROJ]
    int n, H, W, Ho, Wo;
};

template <int CIN, int CMID, int STRIDE, int CPROJ, bool RES>
struct Conv3Cfg {
    static constexpr int TH = 8, TW = 16;
    static constexpr int PAD = (STRIDE == 1) ? 1 : 0;
    static constexpr int IH = (TH - 1) * STRIDE + 3, IW = (TW - 1) * STRIDE + 3;
    static constexpr int PIXP = CIN + 8;            // halves per pixel in smem (odd multiple of 16 B)
    static constexpr int TILE = IH * IW * PIXP;     // halves
    static constexpr int W1P = 9 * CIN + 8;
    static constexpr int W2P = CMID + 8;
    static constexpr int W1_HALVES = CMID * W1P;
    static constexpr int W2_HALVES = CPROJ * W2P;
    static constexpr size_t SMEM = (size_t)(W1_HALVES + W2_HALVES + 2 * TILE) * 2 + (size_t)(CMID + CPROJ) * 4;
    static_assert(CIN % 16 == 0 && CMID % 16 == 0 && CPROJ % 16 == 0, "channel counts must be multiples of 16");
    static_assert(!RES || (STRIDE == 1 && CPROJ == CIN), "residual needs same shape");
};

// MT = m16 pixel-row tiles per warp.  MT = 2: a warp owns two output rows, so every weight (B) fragment it loads from
// shared memory feeds twice the MMAs (the kernel is bound by ldmatrix traffic, not by the tensor pipe); the CTA then
// has 4 warps (128 threads) for the same 8x16 pixel tile.
template <int CIN, int CMID, int STRIDE, int CPROJ, bool RES, int MINB, int MT>
__global__ void __launch_bounds__(256 / MT, MINB) conv3x3_kernel(Conv3Params p) {
    using Cfg = Conv3Cfg<CIN, CMID, STRIDE, CPROJ, RES>;
    constexpr int NT = 256 / MT;           // threads per CTA
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* s_w1 = reinterpret_cast<__half*>(smem_raw);
    __half* s_w2 = s_w1 + Cfg::W1_HALVES;
    __half* s_tile = s_w2 + Cfg::W2_HALVES;
    float* s_b1 = reinterpret_cast<float*>(s_tile + 2 * Cfg::TILE);
    float* s_b2 = s_b1 + CMID;

    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_x = (p.Wo + Cfg::TW - 1) / Cfg::TW, tiles_y = (p.Ho + Cfg::TH - 1) / Cfg::TH;
    const int tiles_per_img = tiles_x * tiles_y;
    const int ntiles = tiles_per_img * p.n;

    // ---- weights -> smem (once per CTA) ----
    {
        constexpr int CH1 = 9 * CIN / 8;   // 16 B chunks per W1 row
        for (int i = tid; i < CMID * CH1; i += NT) {
            int row = i / CH1, c = i - row * CH1;
            cp_async16(s_w1 + row * Cfg::W1P + c * 8, p.w1 + (size_t)row * 9 * CIN + c * 8, 16);
        }
        if constexpr (CPROJ > 0) {
            constexpr int CH2 = CMID / 8;
            for (int i = tid; i < CPROJ * CH2; i += NT) {
                int row = i / CH2, c = i - row * CH2;
                cp_async16(s_w2 + row * Cfg::W2P + c * 8, p.w2 + (size_t)row * CMID + c * 8, 16);
            }
            for (int i = tid; i < CPROJ; i += NT) s_b2[i] = p.b2[i];
        }
        for (int i = tid; i < CMID; i += NT) s_b1[i] = p.b1[i];
    }

    auto load_tile = [&](int t, __half* dst) {
        int n = t / tiles_per_img;
        int rem = t - n * tiles_per_img;
        int ty = rem / tiles_x, tx = rem - ty * tiles_x;
        int gy0 = ty * Cfg::TH * STRIDE - Cfg::PAD, gx0 = tx * Cfg::TW * STRIDE - Cfg::PAD;
        const __half* base = p.in + (size_t)n * p.H * p.W * CIN;
        constexpr int CPP = CIN / 8;
        for (int i = tid; i < Cfg::IH * Cfg::IW * CPP; i += NT) {
            int pix = i / CPP, c8 = i - pix * CPP;
            int iy = pix / Cfg::IW, ix = pix - iy * Cfg::IW;
            int gy = gy0 + iy, gx = gx0 + ix;
            bool ok = (gy >= 0) && (gy < p.H) && (gx >= 0) && (gx < p.W);
            const __half* src = ok ? base + ((size_t)gy * p.W + gx) * CIN + c8 * 8 : p.in;
            cp_async16(dst + pix * Cfg::PIXP + c8 * 8, src, ok ? 16 : 0);
        }
    };

    pdl_wait();       // weights above are constants; activations below come from the previous kernel
    int t = blockIdx.x;
    if (t < ntiles) load_tile(t, s_tile);
    cp_async_commit();

    // lane-dependent ldmatrix offsets
    const int lm = lane >> 3, lr = lane & 7;
    const int a_pix = lr + (lm & 1) * 8;       // pixel (row of A) this lane addresses
    const int a_kof = (lm >> 1) * 8;
    const int b_nof = (lm >> 1) * 8 + lr;      // row of W this lane addresses inside an n16 chunk
    const int b_kof = (lm & 1) * 8;
    const int g = lane >> 2, tq = lane & 3;

    uint32_t breg[CMID == 16 ? 9 * CIN / 16 : 1][4];      // register-resident weights (CMID == 16 only)
    for (int it = 0; t < ntiles; t += gridDim.x, ++it) {
        __half* cur = s_tile + (it & 1) * Cfg::TILE;
        int nxt = t + gridDim.x;
        if (nxt < ntiles) load_tile(nxt, s_tile + ((it + 1) & 1) * Cfg::TILE);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        float acc[MT][CMID / 8][4];
#pragma unroll
        for (int mi = 0; mi < MT; ++mi)
#pragma unroll
            for (int j = 0; j < CMID / 8; ++j) { acc[mi][j][0] = acc[mi][j][1] = acc[mi][j][2] = acc[mi][j][3] = 0.f; }

        const uint32_t tile_addr = smem_u32(cur);
        const uint32_t w1_addr = smem_u32(s_w1);
        if constexpr (CMID == 16) {
            // 16 output channels: the whole weight matrix (9*CIN x 16) is 9*CIN/4 registers per thread; keeping it in
            // registers halves the shared-memory traffic of this bandwidth-bound layer (each A fragment feeds only 2 MMAs)
            if (it == 0) {
#pragma unroll
                for (int q = 0; q < 9 * CIN / 16; ++q) ldmatrix_x4(breg[q], w1_addr + 2u * (uint32_t)(b_nof * Cfg::W1P + q * 16 + b_kof));
            }
#pragma unroll
            for (int rs = 0; rs < 9; ++rs) {
                const int r = rs / 3, s = rs - r * 3;
#pragma unroll
                for (int mi = 0; mi < MT; ++mi) {
                    const uint32_t a_base = tile_addr +
                        2u * (uint32_t)((((warp * MT + mi) * STRIDE + r) * Cfg::IW + a_pix * STRIDE + s) * Cfg::PIXP + a_kof);
#pragma unroll
                    for (int kc = 0; kc < CIN / 16; ++kc) {
                        uint32_t a[4];
                        ldmatrix_x4(a, a_base + kc * 32);
                        const int q = rs * (CIN / 16) + kc;
                        mma16816(acc[mi][0], a, breg[q][0], breg[q][1]);
                        mma16816(acc[mi][1], a, breg[q][2], breg[q][3]);
                    }
                }
            }
        } else {
#pragma unroll 1
        for (int rs = 0; rs < 9; ++rs) {
            const int r = rs / 3, s = rs - r * 3;
            const uint32_t b_base = w1_addr + 2u * (uint32_t)(b_nof * Cfg::W1P + rs * CIN + b_kof);
#pragma unroll
            for (int kc = 0; kc < CIN / 16; ++kc) {
                uint32_t a[MT][4];
#pragma unroll
                for (int mi = 0; mi < MT; ++mi)
                    ldmatrix_x4(a[mi], tile_addr + 2u * (uint32_t)((((warp * MT + mi) * STRIDE + r) * Cfg::IW + a_pix * STRIDE + s) * Cfg::PIXP + a_kof) + kc * 32);
#pragma unroll
                for (int nc = 0; nc < CMID / 16; ++nc) {
                    uint32_t b[4];
                    ldmatrix_x4(b, b_base + 2u * (uint32_t)(nc * 16 * Cfg::W1P) + kc * 32);
#pragma unroll
                    for (int mi = 0; mi < MT; ++mi) {
                        mma16816(acc[mi][2 * nc], a[mi], b[0], b[1]);
                        mma16816(acc[mi][2 * nc + 1], a[mi], b[2], b[3]);
                    }
                }
            }
        }
        }

        // ---- epilogue ----
        const int n = t / tiles_per_img;
        const int rem = t - n * tiles_per_img;
        const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
#pragma unroll
        for (int mi = 0; mi < MT; ++mi) {
        const int oy = ty * Cfg::TH + warp * MT + mi;
        const int ox_lo = tx * Cfg::TW + g, ox_hi = ox_lo + 8;
        const bool ok_lo = (oy < p.Ho) && (ox_lo < p.Wo), ok_hi = (oy < p.Ho) && (ox_hi < p.Wo);
        constexpr int COUT = CPROJ ? CPROJ : CMID;
        __half* out_lo = p.out + (((size_t)n * p.Ho + oy) * p.Wo + ox_lo) * COUT;
        __half* out_hi = out_lo + 8 * COUT;

        if constexpr (CPROJ == 0) {
#pragma unroll
            for (int j = 0; j < CMID / 8; ++j) {
                const int c = j * 8 + tq * 2;
                const float bx = s_b1[c], by = s_b1[c + 1];
                if (ok_lo) *reinterpret_cast<uint32_t*>(out_lo + c) = pack_half2(silu_f(acc[mi][j][0] + bx), silu_f(acc[mi][j][1] + by));
                if (ok_hi) *reinterpret_cast<uint32_t*>(out_hi + c) = pack_half2(silu_f(acc[mi][j][2] + bx), silu_f(acc[mi][j][3] + by));
            }
        } else {
            float acc2[CPROJ / 8][4];
#pragma unroll
            for (int j = 0; j < CPROJ / 8; ++j) { acc2[j][0] = acc2[j][1] = acc2[j][2] = acc2[j][3] = 0.f; }
            const uint32_t w2_base = smem_u32(s_w2) + 2u * (uint32_t)(b_nof * Cfg::W2P + b_kof);
#pragma unroll
            for (int kk = 0; kk < CMID / 16; ++kk) {
                uint32_t a[4];
                {
                    const int c0 = kk * 16 + tq * 2, c1 = c0 + 8;
                    const float b00 = s_b1[c0], b01 = s_b1[c0 + 1], b10 = s_b1[c1], b11 = s_b1[c1 + 1];
                    a[0] = pack_half2(silu_f(acc[mi][2 * kk][0] + b00), silu_f(acc[mi][2 * kk][1] + b01));
                    a[1] = pack_half2(silu_f(acc[mi][2 * kk][2] + b00), silu_f(acc[mi][2 * kk][3] + b01));
                    a[2] = pack_half2(silu_f(acc[mi][2 * kk + 1][0] + b10), silu_f(acc[mi][2 * kk + 1][1] + b11));
                    a[3] = pack_half2(silu_f(acc[mi][2 * kk + 1][2] + b10), silu_f(acc[mi][2 * kk + 1][3] + b11));
                }
#pragma unroll
                for (int nc = 0; nc < CPROJ / 16; ++nc) {
                    uint32_t b[4];
                    ldmatrix_x4(b, w2_base + 2u * (uint32_t)(nc * 16 * Cfg::W2P) + kk * 32);
                    mma16816(acc2[2 * nc], a, b[0], b[1]);
                    mma16816(acc2[2 * nc + 1], a, b[2], b[3]);
                }
            }
#pragma unroll
            for (int j = 0; j < CPROJ / 8; ++j) {
                const int c = j * 8 + tq * 2;
                const float bx = s_b2[c], by = s_b2[c + 1];
                float v0 = acc2[j][0] + bx, v1 = acc2[j][1] + by, v2 = acc2[j][2] + bx, v3 = acc2[j][3] + by;
                if constexpr (RES) {   // shortcut = block input = centre tap of the halo tile
                    const __half* rl = cur + ((warp * MT + mi + 1) * Cfg::IW + (g + 1)) * Cfg::PIXP + c;
                    const __half* rh = rl + 8 * Cfg::PIXP;
                    float2 fl = __half22float2(*reinterpret_cast<const __half2*>(rl));
                    float2 fh = __half22float2(*reinterpret_cast<const __half2*>(rh));
                    v0 += fl.x; v1 += fl.y; v2 += fh.x; v3 += fh.y;
                }
                if (ok_lo) *reinterpret_cast<uint32_t*>(out_lo + c) = pack_half2(v0, v1);
                if (ok_hi) *reinterpret_cast<uint32_t*>(out_hi + c) = pack_half2(v2, v3);
            }
        }
        }   // mi
        __syncthreads();   // everyone is done with `cur` before the next-next prefetch overwrites it
    }
    cp_async_wait<0>();
}

}  // namespace mds
