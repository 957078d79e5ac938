// K0+K1: frame pre-processing (frames.py:7-31: zero-pad H, /255) fused into the encoder stem
// (timm conv_stem 3x3 s2 TF-SAME + bn1 + SiLU).  Reads planar uint8 (or already-normalised float) frames,
// treats `stack_size`=3 frames as the input channels (multidim_stacker.py:214), writes NHWC fp16 [n][H/2][W/2][32].
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
    int H, W;              // logical (padded) size, both even
    int hflip;             // TTA: read columns mirrored (predictors.py:63)
    float divisor;         // 255 for uint8 frames (frames.py:8), 1 for float input
    const float* w;        // [27][32]  k = (ci*3 + r)*3 + s, BN scale folded
    const float* bias;     // [32]
    __half* out;           // [n][H/2][W/2][32]
};

constexpr int kStemTW = 32, kStemTH = 8, kStemC = 32;
constexpr int kStemIW = 2 * kStemTW + 1, kStemIH = 2 * kStemTH + 1;
constexpr int kStemWPR = (kStemIW + 3) / 4, kStemIWP = kStemWPR * 4;   // tile rows are loaded as 4-element words

template <typename IN_T>
__global__ void __launch_bounds__(256, 2) stem_kernel(StemParams p) {
    __shared__ __align__(16) float s_in[3][kStemIH][kStemIWP];
    __shared__ __align__(16) float s_w[27 * kStemC];
    __shared__ float s_b[kStemC];

    const int tid = threadIdx.x;
    const int Ho = p.H >> 1, Wo = p.W >> 1;
    const int ox0 = blockIdx.x * kStemTW, oy0 = blockIdx.y * kStemTH, n = blockIdx.z;

    for (int i = tid; i < 27 * kStemC; i += 256) s_w[i] = p.w[i];
    if (tid < kStemC) s_b[tid] = p.bias[tid];

    const IN_T* img = reinterpret_cast<const IN_T*>(p.in) + (long long)n * p.img_stride;
    // Tile = 3 planes x 17 rows x 17 words of 4 pixels.  All of a thread's loads are issued before any is consumed.
    constexpr int kItems = 3 * kStemIH * kStemWPR, kIters = (kItems + 255) / 256;
    using Word = typename std::conditional<sizeof(IN_T) == 1, uint32_t, float4>::type;
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
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
        const int i = tid + it * 256;
        if (i >= kItems) continue;
        const int ci = i / (kStemIH * kStemWPR);
        const int rem = i - ci * (kStemIH * kStemWPR);
        const int iy = rem / kStemWPR, wd = rem - iy * kStemWPR;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (okv[it]) {
            if constexpr (sizeof(IN_T) == 1) {
                const uint32_t u = vals[it];
                v = make_float4((float)(u & 0xffu), (float)((u >> 8) & 0xffu), (float)((u >> 16) & 0xffu), (float)(u >> 24));
            } else {
                v = vals[it];
            }
            if (p.hflip) v = make_float4(v.w, v.z, v.y, v.x);
            v.x /= p.divisor; v.y /= p.divisor; v.z /= p.divisor; v.w /= p.divisor;
        }
        *reinterpret_cast<float4*>(&s_in[ci][iy][4 * wd]) = v;
    }
    __syncthreads();

    const int cg = tid & 3, px = (tid >> 2) & 31, py2 = tid >> 7;
#pragma unroll 1
    for (int it = 0; it < kStemTH / 2; ++it) {
        const int ty = it * 2 + py2;
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = s_b[cg * 8 + c];
#pragma unroll 1
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    float v = s_in[ci][2 * ty + r][2 * px + s];
                    const float4* wp = reinterpret_cast<const float4*>(&s_w[((ci * 3 + r) * 3 + s) * kStemC + cg * 8]);
                    float4 w0 = wp[0], w1 = wp[1];
                    acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]);
                    acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
                    acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]);
                    acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
                }
        const int yo = oy0 + ty, xo = ox0 + px;
        if (yo < Ho && xo < Wo) {
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = silu_f(acc[c]);
            uint4 v = float8_to_half(acc);
            *reinterpret_cast<uint4*>(p.out + (((long long)n * Ho + yo) * Wo + xo) * kStemC + cg * 8) = v;
        }
    }
}

}  // namespace mds
