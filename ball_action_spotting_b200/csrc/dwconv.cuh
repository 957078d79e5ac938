// K4/K6: depthwise 3x3 (2D, stride 1/2, TF-SAME) and 3x3x3 (3D, pad 1) convolution + folded BN + SiLU, NHWC fp16,
// with the SE squeeze (per-image channel sums over the output, fp32) produced in the same pass
// (timm InvertedResidual conv_dw/bn2/se; multidim_stacker.py:110-114,86).
// HBM-bound: each thread owns 4 channels (8-byte loads, a warp covers 256 contiguous bytes of one pixel) and one
// output column, and slides down the rows; every loaded input row is scattered into the three output rows it
// contributes to, so the only live state is 3 accumulators x 4 channels.  The next row is prefetched before the
// current one is consumed.
#pragma once
#include "common.cuh"

namespace mds {

struct DwParams {
    const __half* in;    // [n][T][H][W][C]
    __half* out;         // [n][T][Ho][Wo][C]
    const float* w;      // [taps][C]  tap = (dt*3 + r)*3 + s, BN scale folded
    const float* bias;   // [C]
    float* sums;         // [n][C]  += sum over (T,Ho,Wo) of the fp32 SiLU output
    int n, T, H, W, C, Ho, Wo;
    int rows_per_chunk, chunks;
};

__device__ __forceinline__ void half4_to_float(const uint2& v, float (&f)[4]) {
    float2 a = unpack_half2(v.x), b = unpack_half2(v.y);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}

__device__ __forceinline__ void dw_block_sums(float (&lsum)[4], int c0, bool active, float* s_sum, float* gsum, int C) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_sum[i] = 0.f;
    __syncthreads();
    if (active) {
#pragma unroll
        for (int i = 0; i < 4; ++i) atomicAdd(&s_sum[c0 + i], lsum[i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        float v = s_sum[i];
        if (v != 0.f) atomicAdd(&gsum[i], v);
    }
}

// ---- 2D: weights in registers -------------------------------------------------------------------------------
template <int STRIDE>
__global__ void __launch_bounds__(256) dwconv2d_kernel(DwParams p) {
    extern __shared__ float s_sum[];
    const int C4 = p.C >> 2;
    const int f = blockIdx.x * 256 + threadIdx.x;
    const bool active = f < p.Wo * C4;
    const int xo = active ? f / C4 : 0;
    const int c0 = active ? (f - xo * C4) * 4 : 0;
    const int n = blockIdx.z;
    const int yo0 = blockIdx.y * p.rows_per_chunk;
    const int yo1 = min(p.Ho, yo0 + p.rows_per_chunk);
    float lsum[4] = {0.f, 0.f, 0.f, 0.f};

    if (active && yo0 < yo1) {
        float w[9][4], b[4];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            float4 v = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)t * p.C + c0));
            w[t][0] = v.x; w[t][1] = v.y; w[t][2] = v.z; w[t][3] = v.w;
        }
        {
            float4 v = __ldg(reinterpret_cast<const float4*>(p.bias + c0));
            b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
        }
        const __half* in = p.in + (size_t)n * p.H * p.W * p.C + c0;
        __half* out = p.out + (size_t)n * p.Ho * p.Wo * p.C + (size_t)xo * p.C + c0;
        // input columns of the three taps; TF-SAME: stride 1 -> pad 1 both sides, stride 2 (even W) -> pad right only
        const int xi0 = xo * STRIDE - (STRIDE == 1 ? 1 : 0);
        bool xok[3];
#pragma unroll
        for (int s = 0; s < 3; ++s) xok[s] = (xi0 + s >= 0) && (xi0 + s < p.W);

        auto load_row = [&](int yi, uint2 (&v)[3]) {
            const bool yok = (yi >= 0) && (yi < p.H);
            const __half* row = in + ((size_t)yi * p.W + xi0) * p.C;
#pragma unroll
            for (int s = 0; s < 3; ++s)
                v[s] = (yok && xok[s]) ? __ldg(reinterpret_cast<const uint2*>(row + (size_t)s * p.C)) : make_uint2(0u, 0u);
        };
        auto emit = [&](int yo, float (&a)[4]) {
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { o[i] = silu_f(a[i] + b[i]); lsum[i] += o[i]; }
            uint2 v; v.x = pack_half2(o[0], o[1]); v.y = pack_half2(o[2], o[3]);
            *reinterpret_cast<uint2*>(out + (size_t)yo * p.Wo * p.C) = v;
        };

        if constexpr (STRIDE == 1) {
            // input row yi feeds out rows yi+1 (kernel row 0), yi (row 1), yi-1 (row 2)
            float a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0}, a2[4];   // a0: out yi-1, a1: out yi, a2: out yi+1
            uint2 cur[3], nxt[3];
            load_row(yo0 - 1, cur);
            for (int yi = yo0 - 1; yi <= yo1; ++yi) {
                if (yi < yo1) load_row(yi + 1, nxt);
                float v[3][4];
#pragma unroll
                for (int s = 0; s < 3; ++s) half4_to_float(cur[s], v[s]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    a0[i] = fmaf(w[6][i], v[0][i], fmaf(w[7][i], v[1][i], fmaf(w[8][i], v[2][i], a0[i])));
                    a1[i] = fmaf(w[3][i], v[0][i], fmaf(w[4][i], v[1][i], fmaf(w[5][i], v[2][i], a1[i])));
                    a2[i] = fmaf(w[0][i], v[0][i], fmaf(w[1][i], v[1][i], w[2][i] * v[2][i]));
                }
                if (yi - 1 >= yo0) emit(yi - 1, a0);
#pragma unroll
                for (int i = 0; i < 4; ++i) { a0[i] = a1[i]; a1[i] = a2[i]; }
#pragma unroll
                for (int s = 0; s < 3; ++s) cur[s] = nxt[s];
            }
        } else {
            // out row yo reads input rows 2yo (kernel row 0), 2yo+1, 2yo+2; row 2yo+2 is also kernel row 0 of yo+1
            float acc[4];
            uint2 r0[3], ra[3], rb[3];
            load_row(2 * yo0, r0);
            {
                float v[3][4];
#pragma unroll
                for (int s = 0; s < 3; ++s) half4_to_float(r0[s], v[s]);
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = fmaf(w[0][i], v[0][i], fmaf(w[1][i], v[1][i], w[2][i] * v[2][i]));
            }
            load_row(2 * yo0 + 1, ra);
            load_row(2 * yo0 + 2, rb);
            for (int yo = yo0; yo < yo1; ++yo) {
                uint2 na[3], nb[3];
                if (yo + 1 < yo1) { load_row(2 * yo + 3, na); load_row(2 * yo + 4, nb); }
                float va[3][4], vb[3][4], nacc[4];
#pragma unroll
                for (int s = 0; s < 3; ++s) { half4_to_float(ra[s], va[s]); half4_to_float(rb[s], vb[s]); }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    acc[i] = fmaf(w[3][i], va[0][i], fmaf(w[4][i], va[1][i], fmaf(w[5][i], va[2][i], acc[i])));
                    acc[i] = fmaf(w[6][i], vb[0][i], fmaf(w[7][i], vb[1][i], fmaf(w[8][i], vb[2][i], acc[i])));
                    nacc[i] = fmaf(w[0][i], vb[0][i], fmaf(w[1][i], vb[1][i], w[2][i] * vb[2][i]));
                }
                emit(yo, acc);
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = nacc[i];
#pragma unroll
                for (int s = 0; s < 3; ++s) { ra[s] = na[s]; rb[s] = nb[s]; }
            }
        }
    }
    dw_block_sums(lsum, c0, active, s_sum, p.sums + (size_t)n * p.C, p.C);
}

// ---- 3D (3x3x3, stride 1, pad 1): weights streamed through L1 -------------------------------------------------
__global__ void __launch_bounds__(256) dwconv3d_kernel(DwParams p) {
    extern __shared__ float s_sum[];
    const int C4 = p.C >> 2;
    const int f = blockIdx.x * 256 + threadIdx.x;
    const bool active = f < p.Wo * C4;
    const int xo = active ? f / C4 : 0;
    const int c0 = active ? (f - xo * C4) * 4 : 0;
    const int n = blockIdx.z;
    const int t = blockIdx.y / p.chunks;
    const int chunk = blockIdx.y - t * p.chunks;
    const int yo0 = chunk * p.rows_per_chunk;
    const int yo1 = min(p.Ho, yo0 + p.rows_per_chunk);
    float lsum[4] = {0.f, 0.f, 0.f, 0.f};

    if (active && yo0 < yo1) {
        float b[4];
        {
            float4 v = __ldg(reinterpret_cast<const float4*>(p.bias + c0));
            b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
        }
        const size_t plane = (size_t)p.H * p.W * p.C;
        const __half* in = p.in + (size_t)n * p.T * plane + c0;
        __half* out = p.out + ((size_t)n * p.T + t) * plane + (size_t)xo * p.C + c0;
        const float* wp = p.w + c0;
        bool xok[3];
#pragma unroll
        for (int s = 0; s < 3; ++s) xok[s] = (xo - 1 + s >= 0) && (xo - 1 + s < p.W);

        float a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
        for (int yi = yo0 - 1; yi <= yo1; ++yi) {
#pragma unroll
            for (int i = 0; i < 4; ++i) a2[i] = 0.f;
            if (yi >= 0 && yi < p.H) {
#pragma unroll
                for (int dt = 0; dt < 3; ++dt) {
                    const int ti = t + dt - 1;
                    if (ti < 0 || ti >= p.T) continue;
                    const __half* row = in + (size_t)ti * plane + ((size_t)yi * p.W + (xo - 1)) * p.C;
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        if (!xok[s]) continue;
                        float v[4];
                        half4_to_float(__ldg(reinterpret_cast<const uint2*>(row + (size_t)s * p.C)), v);
                        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)((dt * 3 + 0) * 3 + s) * p.C));
                        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)((dt * 3 + 1) * 3 + s) * p.C));
                        const float4 w2 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)((dt * 3 + 2) * 3 + s) * p.C));
                        a2[0] = fmaf(w0.x, v[0], a2[0]); a2[1] = fmaf(w0.y, v[1], a2[1]);
                        a2[2] = fmaf(w0.z, v[2], a2[2]); a2[3] = fmaf(w0.w, v[3], a2[3]);
                        a1[0] = fmaf(w1.x, v[0], a1[0]); a1[1] = fmaf(w1.y, v[1], a1[1]);
                        a1[2] = fmaf(w1.z, v[2], a1[2]); a1[3] = fmaf(w1.w, v[3], a1[3]);
                        a0[0] = fmaf(w2.x, v[0], a0[0]); a0[1] = fmaf(w2.y, v[1], a0[1]);
                        a0[2] = fmaf(w2.z, v[2], a0[2]); a0[3] = fmaf(w2.w, v[3], a0[3]);
                    }
                }
            }
            if (yi - 1 >= yo0) {
                float o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { o[i] = silu_f(a0[i] + b[i]); lsum[i] += o[i]; }
                uint2 v; v.x = pack_half2(o[0], o[1]); v.y = pack_half2(o[2], o[3]);
                *reinterpret_cast<uint2*>(out + (size_t)(yi - 1) * p.Wo * p.C) = v;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { a0[i] = a1[i]; a1[i] = a2[i]; }
        }
    }
    dw_block_sums(lsum, c0, active, s_sum, p.sums + (size_t)n * p.C, p.C);
}

}  // namespace mds
