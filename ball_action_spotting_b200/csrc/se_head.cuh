// K5 (SE excitation MLP), K7 (GeM pooling + classifier) and the NCHW<->NHWC boundary converters.
#pragma once
#include "common.cuh"

namespace mds {

// gate[n][c] = sigmoid(W2 . SiLU(W1 . mean + b1) + b2), mean = sums / count  (timm SqueezeExcite; multidim_stacker.py:86-90).
// grid = (images, kSeSlices): every CTA recomputes the tiny squeeze FC, then owns one slice of the channels: it writes
// their gates and — when wg != nullptr — the slice's columns of this image's gated projection weights
//     wg[n][o][c] = fp16( w32[o][c] * gate[c] )          (fp32 product, ONE rounding)
// which the tcgen05 projection GEMM consumes through a 3-D tensor map.  The squeeze is deterministic: the depthwise
// kernel leaves one partial sum per CTA and they are added here in a fixed order (no float atomics anywhere on the
// path, so the whole forward is bit-reproducible).
struct SeParams {
    const float* partials;   // [n][nparts][C]
    int nparts;
    const float* w1;      // [rd][C]
    const float* b1;      // [rd]
    const float* w2t;     // [rd][C]  (conv_expand weight transposed)
    const float* b2;      // [C]
    __half* gate;         // [n][C]
    const float* w32;     // [N][C] folded projection weights (fp32) or nullptr
    __half* wg;           // [n][N][C] or nullptr
    int C, rd, N;
    float inv_count;
};

constexpr int kSeThreads = 512;
constexpr int kSeSlices = 8;
constexpr int kSeUnits = 3;          // hidden units a warp accumulates concurrently (rd <= 48 -> one pass)
__global__ void __launch_bounds__(kSeThreads) se_fc_kernel(SeParams p) {
    extern __shared__ float s_se[];
    float* s_mean = s_se;            // [C]
    float* s_hid = s_se + p.C;       // [rd]
    float* s_gate = s_hid + ((p.rd + 3) & ~3);   // [slice width]
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.x, slice = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* part = p.partials + (size_t)n * p.nparts * p.C;
    for (int c = tid; c < p.C; c += kSeThreads) {      // fixed summation order; 4 loads in flight per step
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int q = 0;
        for (; q + 3 < p.nparts; q += 4) {
            a0 += __ldg(part + (size_t)q * p.C + c);       a1 += __ldg(part + (size_t)(q + 1) * p.C + c);
            a2 += __ldg(part + (size_t)(q + 2) * p.C + c); a3 += __ldg(part + (size_t)(q + 3) * p.C + c);
        }
        for (; q < p.nparts; ++q) a0 += __ldg(part + (size_t)q * p.C + c);
        s_mean[c] = ((a0 + a1) + (a2 + a3)) * p.inv_count;
    }
    __syncthreads();
    // squeeze FC: each warp owns up to kSeUnits hidden units and walks the channels ONCE for all of them, so their
    // weight loads are in flight together (the kernel is a chain of L2 latencies, not bandwidth)
    {
        constexpr int kWarps = kSeThreads / 32;
        const float4* m = reinterpret_cast<const float4*>(s_mean);
        const int c4n = p.C >> 2;
        for (int j0 = warp; j0 < p.rd; j0 += kWarps * kSeUnits) {
            float acc[kSeUnits];
            const float4* wrow[kSeUnits];
#pragma unroll
            for (int u = 0; u < kSeUnits; ++u) {
                acc[u] = 0.f;
                const int j = j0 + u * kWarps;
                wrow[u] = reinterpret_cast<const float4*>(p.w1 + (size_t)(j < p.rd ? j : j0) * p.C);
            }
#pragma unroll 3
            for (int c = lane; c < c4n; c += 32) {
                const float4 mv = m[c];
#pragma unroll
                for (int u = 0; u < kSeUnits; ++u) {
                    const float4 wv = __ldg(wrow[u] + c);
                    acc[u] = fmaf(wv.x, mv.x, fmaf(wv.y, mv.y, fmaf(wv.z, mv.z, fmaf(wv.w, mv.w, acc[u]))));
                }
            }
#pragma unroll
            for (int u = 0; u < kSeUnits; ++u) {
                const int j = j0 + u * kWarps;
                const float a = warp_sum(acc[u]);
                if (lane == 0 && j < p.rd) s_hid[j] = silu_f(a + __ldg(p.b1 + j));
            }
        }
    }
    __syncthreads();
    // this CTA's channel slice, in units of 8 channels (16 bytes of fp16)
    const int groups = p.C >> 3, gps = (groups + kSeSlices - 1) / kSeSlices;
    const int c_lo = min(p.C, slice * gps * 8), c_hi = min(p.C, c_lo + gps * 8);
    // expand FC + sigmoid: 4 lanes per channel split the hidden units, so each lane has rd/4 independent loads
    for (int c0 = c_lo; c0 < c_hi; c0 += kSeThreads / 4) {
        const int c = c0 + (tid >> 2), sub = tid & 3;
        float a = 0.f;
        if (c < c_hi) {
#pragma unroll 4
            for (int j = sub; j < p.rd; j += 4) a = fmaf(__ldg(p.w2t + (size_t)j * p.C + c), s_hid[j], a);
        }
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        if (c < c_hi && sub == 0) {
            const float gv = sigmoid_f(a + __ldg(p.b2 + c));
            s_gate[c - c_lo] = gv;
            p.gate[(size_t)n * p.C + c] = __float2half_rn(gv);
        }
    }
    if (p.wg == nullptr) return;
    __syncthreads();
    const int wgrp = (c_hi - c_lo) >> 3;                 // 8-channel groups in this slice
    __half* wg = p.wg + (size_t)n * p.N * p.C;
    const int total = p.N * wgrp;
    auto gate8 = [&](int i, const float4& x0, const float4& x1) {
        const int o = i / wgrp, gq = i - o * wgrp;
        const float* gp = s_gate + gq * 8;
        uint4 v;
        v.x = pack_half2(x0.x * gp[0], x0.y * gp[1]); v.y = pack_half2(x0.z * gp[2], x0.w * gp[3]);
        v.z = pack_half2(x1.x * gp[4], x1.y * gp[5]); v.w = pack_half2(x1.z * gp[6], x1.w * gp[7]);
        *reinterpret_cast<uint4*>(wg + (size_t)o * p.C + c_lo + gq * 8) = v;
    };
    auto src = [&](int i) {
        const int o = i / wgrp, gq = i - o * wgrp;
        return reinterpret_cast<const float4*>(p.w32 + (size_t)o * p.C + c_lo + gq * 8);
    };
    int i = tid;
    for (; i + 3 * kSeThreads < total; i += 4 * kSeThreads) {  // four items (eight 16-byte loads) in flight per thread: this loop is a
        const float4* s0 = src(i);                             // chain of L2 round trips, so its time is the number of trips
        const float4* s1 = src(i + kSeThreads);
        const float4* s2 = src(i + 2 * kSeThreads);
        const float4* s3 = src(i + 3 * kSeThreads);
        const float4 a0 = __ldg(s0), a1 = __ldg(s0 + 1), b0 = __ldg(s1), b1 = __ldg(s1 + 1);
        const float4 c0 = __ldg(s2), c1 = __ldg(s2 + 1), d0 = __ldg(s3), d1 = __ldg(s3 + 1);
        gate8(i, a0, a1);
        gate8(i + kSeThreads, b0, b1);
        gate8(i + 2 * kSeThreads, c0, c1);
        gate8(i + 3 * kSeThreads, d0, d1);
    }
    for (; i < total; i += kSeThreads) {
        const float4* s0 = src(i);
        const float4 a0 = __ldg(s0), a1 = __ldg(s0 + 1);
        gate8(i, a0, a1);
    }
}

// GeM (multidim_stacker.py:42-45) over the P = h*w positions of x[b][t][P][C]: feat[b][t*C + c].
// Two launches: partial sums of clamp(x, eps)^p over kGemSplit slices of the positions (grid (T, b, kGemSplit): enough
// CTAs to hide the latency at small batches), then the fixed-order sum, mean and ^(1/p).
constexpr int kGemSplit = 8;
struct GemParams {
    const __half* x;   // [b][T][P][C]
    float* part;       // [b][T][kGemSplit][C]
    float* feat;       // [b][T*C]
    int T, P, C;
    float p, eps;
};

__global__ void __launch_bounds__(256) gem_kernel(GemParams g) {
    __shared__ float s_part[8][260];
    pdl_trigger();
    pdl_wait();
    const int t = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x;
    const int C8 = g.C >> 3;                 // threads along channels (8 ch each)
    const int lanes_p = min(256 / C8, 8);    // position lanes (s_part holds 8)
    const int cg = tid % C8, pl = tid / C8;
    const bool cube = (g.p == 3.0f);
    const int per = (g.P + kGemSplit - 1) / kGemSplit;
    const int p0 = blockIdx.z * per, p1 = min(g.P, p0 + per);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const __half* base = g.x + (((size_t)b * g.T + t) * g.P) * g.C + cg * 8;
    if (pl < lanes_p) {
        for (int pos = p0 + pl; pos < p1; pos += lanes_p) {
            float v[8];
            half8_to_float(ldg16(base + (size_t)pos * g.C), v);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float x = fmaxf(v[i], g.eps);
                acc[i] += cube ? x * x * x : powf(x, g.p);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) s_part[pl][cg * 8 + i] = acc[i];
    }
    __syncthreads();
    if (tid < g.C) {
        float s = 0.f;
        for (int l = 0; l < lanes_p; ++l) s += s_part[l][tid];
        g.part[(((size_t)b * g.T + t) * kGemSplit + blockIdx.z) * g.C + tid] = s;
    }
}
// grid ceil(b * T * C / 256): feat = (sum of the slices / P)^(1/p)
__global__ void __launch_bounds__(256) gem_finish_kernel(GemParams g, int total) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * 256 + threadIdx.x;         // i = (b * T + t) * C + c
    if (i >= total) return;
    const int bt = i / g.C, c = i - bt * g.C;
    float s = 0.f;
#pragma unroll
    for (int z = 0; z < kGemSplit; ++z) s += g.part[((size_t)bt * kGemSplit + z) * g.C + c];
    const float m = s / (float)g.P;
    g.feat[i] = (g.p == 3.0f) ? cbrtf(m) : powf(m, 1.0f / g.p);
}

// logits[b][k] = feat[b] . W[k] + bias[k]  (nn.Linear, multidim_stacker.py:236); optional sigmoid (argus_models.py:26)
__global__ void __launch_bounds__(256) linear_head_kernel(const float* feat, const float* w, const float* bias, float* out,
                                                          int F, int num_classes, int apply_sigmoid) {
    __shared__ float s_red[8];
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = 0; k < num_classes; ++k) {
        float acc = 0.f;
        for (int i = tid; i < F; i += 256) acc = fmaf(feat[(size_t)b * F + i], __ldg(w + (size_t)k * F + i), acc);
        acc = warp_sum(acc);
        if (lane == 0) s_red[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            float s = 0.f;
            for (int i = 0; i < 8; ++i) s += s_red[i];
            s += bias[k];
            out[(size_t)b * num_classes + k] = apply_sigmoid ? 1.0f / (1.0f + expf(-s)) : s;
        }
        __syncthreads();
    }
}

// ---- sliding-window assembly (src/predictors.py:66-68: torch.cat of the cached per-triple features) ------------------
// out[p][t] = feats[first + p + hop * t]: window p stacks the cached encoder features of T triples `hop` frames apart.
// grid (T, n_pred); plane_vec = 16-byte vectors per (h, w, C) feature plane
__global__ void __launch_bounds__(256) gather_stacks_kernel(const uint4* feats, uint4* out, long long first, int hop, int T,
                                                            long long plane_vec) {
    const int t = blockIdx.x, pidx = blockIdx.y;
    const uint4* src = feats + (first + pidx + (long long)hop * t) * plane_vec;
    uint4* dst = out + ((long long)pidx * T + t) * plane_vec;
    for (long long i = threadIdx.x; i < plane_vec; i += 256) dst[i] = __ldg(src + i);
}
// y = a * y + b * x (mean of the TTA branches, predictors.py:72)
__global__ void __launch_bounds__(256) axpby_kernel(float* y, const float* x, float a, float b, long long n) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i < n) y[i] = a * y[i] + b * x[i];
}

// ---- boundary layout converters (the reference API is NCHW float32; the engine is NHWC fp16) ----------------
// src [n][C][P] f32 -> dst [n][P][C] f16
__global__ void __launch_bounds__(256) nchw32_to_nhwc16_kernel(const float* src, __half* dst, int C, int P) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        int c = c0 + i, pp = p0 + tx;
        tile[i][tx] = (c < C && pp < P) ? src[((size_t)n * C + c) * P + pp] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        int pp = p0 + i, c = c0 + tx;
        if (pp < P && c < C) dst[((size_t)n * P + pp) * C + c] = __float2half_rn(tile[tx][i]);
    }
}
// src [n][P][C] f16 -> dst [n][C][P] f32
__global__ void __launch_bounds__(256) nhwc16_to_nchw32_kernel(const __half* src, float* dst, int C, int P) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        int pp = p0 + i, c = c0 + tx;
        tile[i][tx] = (pp < P && c < C) ? __half2float(src[((size_t)n * P + pp) * C + c]) : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        int c = c0 + i, pp = p0 + tx;
        if (c < C && pp < P) dst[((size_t)n * C + c) * P + pp] = tile[tx][i];
    }
}

}  // namespace mds
