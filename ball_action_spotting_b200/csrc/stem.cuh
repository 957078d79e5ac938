// K0+K1: frame pre-processing (frames.py:7-31: zero-pad H, /255) fused into the encoder stem
// (timm conv_stem 3x3 s2 TF-SAME + bn1 + SiLU).  Reads planar uint8 (or already-normalised float) frames,
// treats `stack_size`=3 frames as the input channels (multidim_stacker.py:214), writes NHWC fp16 [n][H/2][W/2][32].
//
// The 3->32 conv runs on tensor cores as an implicit GEMM (M = 16 output pixels of a row, K = 27 taps padded to 32,
// N = 32): A fragments are gathered straight from the fp16 input tile in shared memory.  uint8 pixels are exact in
// fp16 and the folded weights are split into an fp16 hi/lo pair (two MMAs), so the products are exact and the
// result carries fp32-level accuracy; float input is split hi/lo as well (three MMAs).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace mds {

struct StemParams {
    const void* in;        // planar frames
    long long img_stride;  // elements between consecutive images (3-frame groups)
    long long plane_stride;  // elements between the 3 channel planes of one image
    int stored_h;          // rows physically present per plane (720 for raw frames, H for padded input)
    int pad_top;           // logical row y maps to stored row y - pad_top (frames.py:19)
    int H, W;              // logical (padded) size; H even, W a multiple of 4
    int hflip;             // TTA: read columns mirrored (predictors.py:63)
    float scale;           // 1/255 for uint8 frames (frames.py:8), 1 for float input; applied to the fp32 accumulator
    const __half* wh;      // [2][32][32]  (hi, lo) x cout x k,  k = (ci*3 + r)*3 + s (27..31 zero), BN scale folded
    const float* bias;     // [32]
    __half* out;           // [n][H/2][W/2][32]
};

constexpr int kStemTW = 32, kStemTH = 8, kStemC = 32;
constexpr int kStemIW = 2 * kStemTW + 1, kStemIH = 2 * kStemTH + 1;
constexpr int kStemWPR = (kStemIW + 3) / 4, kStemIWP = kStemWPR * 4;   // tile rows are loaded as 4-pixel words
constexpr int kStemPlane = kStemIH * kStemIWP;                          // halves per channel plane
constexpr int kStemStgPitch = 40;                                       // halves per pixel in the output staging tile

template <typename IN_T>
__global__ void __launch_bounds__(256, 3) stem_kernel(StemParams p) {   // 3 CTAs / SM: measured 7 % faster than 2 (106 -> 80 registers), 4 spills
    constexpr bool kFloat = sizeof(IN_T) == 4;
    constexpr int kParts = kFloat ? 2 : 1;                // float input is split into fp16 hi + lo
    __shared__ __align__(16) __half s_in[kParts][3 * kStemPlane + 8];
    __shared__ __align__(16) __half s_stg[8][16 * kStemStgPitch];

    pdl_trigger();
    pdl_wait();       // the output buffer may still be read by the previous forward's last kernels
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int Ho = p.H >> 1, Wo = p.W >> 1;
    const int ox0 = blockIdx.x * kStemTW, oy0 = blockIdx.y * kStemTH, n = blockIdx.z;

    // ---- input tile -> smem (fp16).  All of a thread's loads are issued before any is consumed. ----
    const IN_T* img = reinterpret_cast<const IN_T*>(p.in) + (long long)n * p.img_stride;
    constexpr int kItems = 3 * kStemIH * kStemWPR, kIters = (kItems + 255) / 256;
    using Word = typename std::conditional<kFloat, float4, uint32_t>::type;
    Word vals[kIters];
    bool okv[kIters];
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
        const int i = tid + it * 256;
        const int ci = i / (kStemIH * kStemWPR);
        const int rem = i - ci * (kStemIH * kStemWPR);
        const int iy = rem / kStemWPR, wd = rem - iy * kStemWPR;
        const int y = 2 * oy0 + iy, x = 2 * ox0 + 4 * wd;   // TF-SAME on even input: pad bottom/right only
        const int ys = y - p.pad_top;
        okv[it] = (i < kItems) && (y < p.H) && (x < p.W) && (ys >= 0) && (ys < p.stored_h);
        if (okv[it]) {
            const int xs = p.hflip ? (p.W - 4 - x) : x;
            vals[it] = __ldg(reinterpret_cast<const Word*>(img + ci * p.plane_stride + (long long)ys * p.W + xs));
        }
    }
    if (tid < 8) {   // the zero element that the padded K columns (27..31) point at
        s_in[0][3 * kStemPlane + tid] = __float2half(0.f);
        if constexpr (kFloat) s_in[1][3 * kStemPlane + tid] = __float2half(0.f);
    }
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
        const int i = tid + it * 256;
        if (i >= kItems) continue;
        const int ci = i / (kStemIH * kStemWPR);
        const int rem = i - ci * (kStemIH * kStemWPR);
        const int iy = rem / kStemWPR, wd = rem - iy * kStemWPR;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (okv[it]) {
            if constexpr (kFloat) {
                v = vals[it];
            } else {
                const uint32_t u = vals[it];
                v = make_float4((float)(u & 0xffu), (float)((u >> 8) & 0xffu), (float)((u >> 16) & 0xffu), (float)(u >> 24));
            }
            if (p.hflip) v = make_float4(v.w, v.z, v.y, v.x);
        }
        const int o = ci * kStemPlane + iy * kStemIWP + 4 * wd;
        const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);   // exact for uint8
        *reinterpret_cast<__half2*>(&s_in[0][o]) = h01;
        *reinterpret_cast<__half2*>(&s_in[0][o + 2]) = h23;
        if constexpr (kFloat) {
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            *reinterpret_cast<__half2*>(&s_in[1][o]) = __floats2half2_rn(v.x - f01.x, v.y - f01.y);
            *reinterpret_cast<__half2*>(&s_in[1][o + 2]) = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
        }
    }

    // ---- per-thread constants: the 8 K indices of this lane's A fragments and its B (weight) fragments ----
    int koff[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = (i >> 2) * 16 + ((i >> 1) & 1) * 8 + 2 * t + (i & 1);
        const int ci = k / 9, r9 = k - ci * 9, r = r9 / 3, s = r9 - r * 3;
        koff[i] = (k < 27) ? (ci * kStemPlane + r * kStemIWP + s) : -1;
    }
    uint32_t bw[2][2][4][2];      // [hi/lo][k16 step][n8 tile][b0,b1]
#pragma unroll
    for (int part = 0; part < 2; ++part)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __half* wp = p.wh + ((part * 32 + j * 8 + g) * 32 + kk * 16 + 2 * t);
                bw[part][kk][j][0] = __ldg(reinterpret_cast<const uint32_t*>(wp));
                bw[part][kk][j][1] = __ldg(reinterpret_cast<const uint32_t*>(wp + 8));
            }
    float bias_r[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) { bias_r[j][0] = __ldg(p.bias + j * 8 + 2 * t); bias_r[j][1] = __ldg(p.bias + j * 8 + 2 * t + 1); }
    __syncthreads();

    // ---- warp = one output row of the tile, two m16 pixel groups ----
    const int ty = warp, yo = oy0 + ty;
    __half* stg = s_stg[warp];
    constexpr int kZero = 3 * kStemPlane;
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
        const int base = (2 * ty) * kStemIWP + 2 * (mt * 16 + g);     // input element of (row g, tap 0)
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
        for (int part = 0; part < kParts; ++part) {
            uint32_t a[2][4];
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i0 = kk * 4 + h * 2;
                    const int o0 = koff[i0] >= 0 ? base + koff[i0] : kZero, o1 = koff[i0 + 1] >= 0 ? base + koff[i0 + 1] : kZero;
                    const int o0b = koff[i0] >= 0 ? o0 + 16 : kZero, o1b = koff[i0 + 1] >= 0 ? o1 + 16 : kZero;   // row g + 8
                    const __half2 lo = __halves2half2(s_in[part][o0], s_in[part][o1]);
                    const __half2 hi = __halves2half2(s_in[part][o0b], s_in[part][o1b]);
                    a[kk][h * 2 + 0] = *reinterpret_cast<const uint32_t*>(&lo);
                    a[kk][h * 2 + 1] = *reinterpret_cast<const uint32_t*>(&hi);
                }
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    mma16816(acc[j], a[kk], bw[0][kk][j][0], bw[0][kk][j][1]);
                    if (part == 0) mma16816(acc[j], a[kk], bw[1][kk][j][0], bw[1][kk][j][1]);   // x_hi * w_lo
                }
        }
        // bias + SiLU, stage the 16 px x 32 ch tile, then 16-byte coalesced stores (64 B per pixel, 512 B per request)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float v0 = silu_f(fmaf(acc[j][0], p.scale, bias_r[j][0])), v1 = silu_f(fmaf(acc[j][1], p.scale, bias_r[j][1]));
            const float v2 = silu_f(fmaf(acc[j][2], p.scale, bias_r[j][0])), v3 = silu_f(fmaf(acc[j][3], p.scale, bias_r[j][1]));
            *reinterpret_cast<uint32_t*>(stg + g * kStemStgPitch + j * 8 + 2 * t) = pack_half2(v0, v1);
            *reinterpret_cast<uint32_t*>(stg + (g + 8) * kStemStgPitch + j * 8 + 2 * t) = pack_half2(v2, v3);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = lane + 32 * i, px = idx >> 2, c16 = idx & 3;
            const int xo = ox0 + mt * 16 + px;
            if (yo < Ho && xo < Wo)
                *reinterpret_cast<uint4*>(p.out + (((long long)n * Ho + yo) * Wo + xo) * kStemC + c16 * 8) =
                    *reinterpret_cast<const uint4*>(stg + px * kStemStgPitch + c16 * 8);
        }
        __syncwarp();
    }
}

}  // namespace mds
