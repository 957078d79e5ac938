// K5 (SE excitation MLP), K7 (GeM pooling + classifier) and the NCHW<->NHWC boundary converters.
#pragma once
#include "common.cuh"

namespace mds {

// gate[n][c] = sigmoid(W2 . SiLU(W1 . mean + b1) + b2), mean = sums / count  (timm SqueezeExcite; multidim_stacker.py:86-90).
// The kernel is a chain of latencies, not bandwidth, so it is built to make the chain short:
//   * one thread-block CLUSTER of 8 CTAs per image; CTA r owns hidden units [r U, r U + U) of the squeeze FC and one
//     eighth of the channels of the excitation FC; the 8 x U hidden values are exchanged through distributed shared memory
//     (one DSMEM store per peer + one cluster barrier), so no CTA recomputes the squeeze FC;
//   * everything that does not depend on the depthwise kernel — this CTA's rows of W1 and its slice of W2 — is staged in
//     shared memory by cp.async BEFORE griddepcontrol.wait, i.e. while the depthwise kernel is still draining; after the
//     wait only the partial sums (one round trip) and the fp32 projection weights are read from L2.
// When wg != nullptr the CTA also writes its slice of this image's gated projection weights
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
constexpr int kSeSlices = 8;          // = cluster size
__host__ __device__ inline int se_units(int rd) { return (rd + kSeSlices - 1) / kSeSlices; }        // hidden units per CTA
__host__ __device__ inline int se_slice(int C) { return (((C >> 3) + kSeSlices - 1) / kSeSlices) * 8; }   // channels per CTA
__host__ __device__ inline size_t se_smem_bytes(int C, int rd) {
    return (size_t)(se_units(rd) * C + rd * se_slice(C) + C + ((rd + 3) & ~3) + se_slice(C)) * sizeof(float);
}

__device__ __forceinline__ uint32_t se_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void se_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store a float into the shared memory of CTA `rank` of this cluster at the same offset as local address `saddr`
__device__ __forceinline__ void se_st_cluster(uint32_t saddr, uint32_t rank, float v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(saddr), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}

__global__ void __launch_bounds__(kSeThreads) se_fc_kernel(SeParams p) {
    extern __shared__ __align__(16) float s_se[];
    const int U = se_units(p.rd), SW = se_slice(p.C);
    float* s_w1 = s_se;                          // [U][C]   rows rank*U .. of W1
    float* s_w2 = s_w1 + U * p.C;                // [rd][SW] this CTA's channel slice of W2^T
    float* s_mean = s_w2 + p.rd * SW;            // [C]
    float* s_hid = s_mean + p.C;                 // [rd] (all hidden units, filled by the 8 CTAs of the cluster)
    float* s_gate = s_hid + ((p.rd + 3) & ~3);   // [SW]
    pdl_trigger();
    const int n = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)se_cluster_rank();     // == blockIdx.x
    const int c_lo = min(p.C, rank * SW), c_hi = min(p.C, c_lo + SW);
    const int sw = c_hi - c_lo;                  // may be 0 for the last ranks of narrow layers
    const int u_lo = rank * U;

    // ---- constants -> shared memory, before the dependency on the depthwise kernel ----
    {
        const int c4n = p.C >> 2;
        for (int i = tid; i < U * c4n; i += kSeThreads) {
            const int u = i / c4n, c4 = i - u * c4n;
            if (u_lo + u < p.rd) cp_async16(s_w1 + u * p.C + c4 * 4, p.w1 + (size_t)(u_lo + u) * p.C + c4 * 4, 16);
        }
        const int s4n = sw >> 2;
        for (int i = tid; i < p.rd * s4n; i += kSeThreads) {
            const int j = i / s4n, c4 = i - j * s4n;
            cp_async16(s_w2 + j * SW + c4 * 4, p.w2t + (size_t)j * p.C + c_lo + c4 * 4, 16);
        }
        cp_async_commit();
    }
    pdl_wait();
    // ---- squeeze: fixed-order sum of the depthwise kernel's partials (all loads of a thread in flight together) ----
    const float* part = p.partials + (size_t)n * p.nparts * p.C;
    for (int c = tid; c < p.C; c += kSeThreads) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int q = 0;
        for (; q + 3 < p.nparts; q += 4) {
            a0 += __ldg(part + (size_t)q * p.C + c);       a1 += __ldg(part + (size_t)(q + 1) * p.C + c);
            a2 += __ldg(part + (size_t)(q + 2) * p.C + c); a3 += __ldg(part + (size_t)(q + 3) * p.C + c);
        }
        for (; q < p.nparts; ++q) a0 += __ldg(part + (size_t)q * p.C + c);
        s_mean[c] = ((a0 + a1) + (a2 + a3)) * p.inv_count;
    }
    cp_async_wait<0>();
    __syncthreads();
    // ---- squeeze FC: warp u -> hidden unit u_lo + u (U <= 8 warps busy), result broadcast to every CTA of the cluster ----
    if (warp < U && u_lo + warp < p.rd) {
        const float4* m = reinterpret_cast<const float4*>(s_mean);
        const float4* wr = reinterpret_cast<const float4*>(s_w1 + warp * p.C);
        float acc = 0.f;
        for (int c4 = lane; c4 < (p.C >> 2); c4 += 32) {
            const float4 mv = m[c4], wv = wr[c4];
            acc = fmaf(wv.x, mv.x, fmaf(wv.y, mv.y, fmaf(wv.z, mv.z, fmaf(wv.w, mv.w, acc))));
        }
        acc = warp_sum(acc);
        const int j = u_lo + warp;
        const float hv = silu_f(acc + __ldg(p.b1 + j));
        if (lane < kSeSlices) se_st_cluster(smem_u32(s_hid + j), (uint32_t)lane, hv);
    }
    se_cluster_sync();            // all hidden units of the image are in every CTA's s_hid
    // ---- excitation FC + sigmoid for this CTA's channel slice ----
    for (int cc = tid; cc < sw; cc += kSeThreads) {
        float a0 = 0.f, a1 = 0.f;
        int j = 0;
        for (; j + 1 < p.rd; j += 2) {
            a0 = fmaf(s_w2[j * SW + cc], s_hid[j], a0);
            a1 = fmaf(s_w2[(j + 1) * SW + cc], s_hid[j + 1], a1);
        }
        if (j < p.rd) a0 = fmaf(s_w2[j * SW + cc], s_hid[j], a0);
        const float gv = sigmoid_f(a0 + a1 + __ldg(p.b2 + c_lo + cc));
        s_gate[cc] = gv;
        p.gate[(size_t)n * p.C + c_lo + cc] = __float2half_rn(gv);
    }
    if (p.wg == nullptr || sw == 0) return;
    __syncthreads();
    const int wgrp = sw >> 3;                            // 8-channel groups in this slice
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
    // four items (eight 16-byte loads) in flight per thread, also in the last, partial round: the loop is a chain of L2 round
    // trips, so its time is the number of rounds
    for (int i = tid; i < total; i += 4 * kSeThreads) {
        float4 x[4][2];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int it = i + q * kSeThreads;
            if (it < total) { const float4* s0 = src(it); x[q][0] = __ldg(s0); x[q][1] = __ldg(s0 + 1); }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int it = i + q * kSeThreads;
            if (it < total) gate8(it, x[q][0], x[q][1]);
        }
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
