// Kernels of the frozen-encoder training step (BASELINE.json configs[4]; reference: BallActionModel.train_step,
// src/argus_models.py:41-74, over conv2d_projection / InvertedResidual3d / conv3d_projection / GeM / classifier,
// src/models/multidim_stacker.py:53-134,198-237) -- forward in train mode, backward, SGD.
//
// Layout: every activation / activation-gradient is a row-major [M = b*T*h*w][C] fp16 matrix (NDHWC), rows of one
// sample contiguous.  Reductions over rows are two-level and fixed-order (per-CTA partials, then a finalize kernel):
// no floating-point atomics, the step is bit-reproducible.
#pragma once
#include "common.cuh"
#include "dwconv.cuh"

namespace mds {

constexpr int kEwThreads = 256;
constexpr int kEwMaxRed = 4608;      // floats of shared memory for the cross-row-lane reduction (2 sums x RL x C)

__device__ __forceinline__ float dsilu_f(float z) {        // d/dz [z * sigmoid(z)]
    const float s = sigmoid_f(z);
    return s * (1.0f + z * (1.0f - s));
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}

// ------------------------------------------------------------------------------------------------------------
// Column kernels over [M][C]: grid (chunks, samples); thread = 8 channels x a strided set of the chunk's rows.
// ------------------------------------------------------------------------------------------------------------
struct EwParams {
    const __half* y;        // [M][C] pre-BatchNorm tensor (conv output)
    const __half* g;        // [M][C] second operand: incoming gradient, or the residual (forward mode 3)
    __half* out;            // [M][C]
    float* partials;        // [samples * chunks][NACC][C]
    const float* scale;     // [C] gamma * rstd
    const float* shift;     // [C] beta - mean * scale
    const float* mean;      // [C]
    const float* rstd;      // [C]
    const float* smul;      // [samples][C] or nullptr (SE gate)
    const float* sadd;      // [samples][C] or nullptr (SE squeeze gradient, already divided by the row count)
    const float* bmul;      // [samples]    or nullptr (DropPath mask)
    const float* c1;        // [C] A   (backward pass 2: dy = gr * dz + A * y + B)
    const float* c2;        // [C] B
    const float* gr;        // [C] gamma * rstd
    int C, rows_per_sample, rows_per_chunk;
};

struct EwCtx {
    int C8, RL, cg, rl, smp, r0, r1;
    bool active;
    size_t base;            // first row of this sample
};
__device__ __forceinline__ EwCtx ew_ctx(const EwParams& p) {
    EwCtx c;
    c.C8 = p.C >> 3;
    c.RL = kEwThreads / c.C8;
    c.cg = threadIdx.x % c.C8;
    c.rl = threadIdx.x / c.C8;
    c.active = c.rl < c.RL;
    c.smp = blockIdx.y;
    c.r0 = blockIdx.x * p.rows_per_chunk;
    c.r1 = min(c.r0 + p.rows_per_chunk, p.rows_per_sample);
    c.base = (size_t)c.smp * p.rows_per_sample;
    return c;
}
__device__ __forceinline__ void load8(const float* src, int cg, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + cg * 2);
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + cg * 2 + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load_row8(const __half* t, size_t row, int C, int cg, float (&v)[8]) {
    half8_to_float(__ldg(reinterpret_cast<const uint4*>(t + row * C) + cg), v);
}
__device__ __forceinline__ uint4 ldg_row(const __half* t, size_t row, int C, int cg) {
    return __ldg(reinterpret_cast<const uint4*>(t + row * C) + cg);
}
__device__ __forceinline__ void store_row8(__half* t, size_t row, int C, int cg, const float (&v)[8]) {
    *(reinterpret_cast<uint4*>(t + row * C) + cg) = float8_to_half(v);
}

// cross-row-lane reduction of NACC x 8 per-thread sums, written to partials[(smp * gridDim.x + chunk)][NACC][C]
template <int NACC>
__device__ __forceinline__ void ew_reduce_store(const EwParams& p, const EwCtx& c, const float (&acc)[NACC][8]) {
    __shared__ float s_red[kEwMaxRed];
    const int C = p.C;
    if (c.active) {
#pragma unroll
        for (int a = 0; a < NACC; ++a)
#pragma unroll
            for (int i = 0; i < 8; ++i) s_red[(a * c.RL + c.rl) * C + c.cg * 8 + i] = acc[a][i];
    }
    __syncthreads();
    float* out = p.partials + ((size_t)c.smp * gridDim.x + blockIdx.x) * NACC * C;
    for (int i = threadIdx.x; i < NACC * C; i += kEwThreads) {
        const int a = i / C, ch = i - a * C;
        float s = 0.f;
        for (int l = 0; l < c.RL; ++l) s += s_red[(a * c.RL + l) * C + ch];
        out[i] = s;
    }
}

// BatchNorm batch statistics: partial sums of (y - ref) and (y - ref)^2, ref = p.mean = the layer's running mean (a
// data-independent guess of the batch mean: the shift keeps the one-pass variance well conditioned).
__global__ void __launch_bounds__(kEwThreads) bn_stats_kernel(EwParams p) {
    pdl_trigger();
    pdl_wait();
    const EwCtx c = ew_ctx(p);
    float acc[2][8] = {};
    if (c.active) {
        float ref[8];
        load8(p.mean, c.cg, ref);
        for (int r = c.r0 + c.rl; r < c.r1; r += 2 * c.RL) {        // two rows (four 16-byte loads per pair) in flight
            const bool two = r + c.RL < c.r1;
            const uint4 ra = ldg_row(p.y, c.base + r, p.C, c.cg);
            const uint4 rb = two ? ldg_row(p.y, c.base + r + c.RL, p.C, c.cg) : make_uint4(0, 0, 0, 0);
            float v[8];
            half8_to_float(ra, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float d = v[i] - ref[i];
                acc[0][i] += d;
                acc[1][i] = fmaf(d, d, acc[1][i]);
            }
            if (two) {
                half8_to_float(rb, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float d = v[i] - ref[i];
                    acc[0][i] += d;
                    acc[1][i] = fmaf(d, d, acc[1][i]);
                }
            }
        }
    }
    ew_reduce_store<2>(p, c, acc);
}

// Forward application of BatchNorm (+ SiLU).  MODE 0: out = silu(z); 1: per-sample column sums of silu(z) only (SE
// squeeze, multidim_stacker.py:86); 2: out = silu(z) * gate[sample] (:90); 3: out = z * mask[sample] + residual (:133).
template <int MODE>
__global__ void __launch_bounds__(kEwThreads) bn_fwd_kernel(EwParams p) {
    pdl_trigger();
    pdl_wait();
    const EwCtx c = ew_ctx(p);
    float acc[1][8] = {};
    if (c.active) {
        float sc[8], sh[8], gt[8];
        load8(p.scale, c.cg, sc);
        load8(p.shift, c.cg, sh);
        if (MODE == 2) load8(p.smul + (size_t)c.smp * p.C, c.cg, gt);
        const float bm = (MODE == 3 && p.bmul) ? __ldg(p.bmul + c.smp) : 1.0f;
        auto row = [&](const uint4& ry, const uint4& rr, int r) {
            float v[8], o[8];
            half8_to_float(ry, v);
            if (MODE == 3) {
                float res[8];
                half8_to_float(rr, res);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = fmaf(fmaf(v[i], sc[i], sh[i]), bm, res[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    o[i] = silu_f(fmaf(v[i], sc[i], sh[i]));
                    if (MODE == 1) acc[0][i] += o[i];
                    if (MODE == 2) o[i] *= gt[i];
                }
            }
            if (MODE != 1) store_row8(p.out, c.base + r, p.C, c.cg, o);
        };
        const uint4 zero = make_uint4(0, 0, 0, 0);
        for (int r = c.r0 + c.rl; r < c.r1; r += 2 * c.RL) {
            const bool two = r + c.RL < c.r1;
            const uint4 ya = ldg_row(p.y, c.base + r, p.C, c.cg);
            const uint4 ga = MODE == 3 ? ldg_row(p.g, c.base + r, p.C, c.cg) : zero;
            const uint4 yb = two ? ldg_row(p.y, c.base + r + c.RL, p.C, c.cg) : zero;
            const uint4 gb = (two && MODE == 3) ? ldg_row(p.g, c.base + r + c.RL, p.C, c.cg) : zero;
            row(ya, ga, r);
            if (two) row(yb, gb, r + c.RL);
        }
    }
    if (MODE == 1) ew_reduce_store<1>(p, c, acc);
}

// incoming gradient of the BatchNorm(+SiLU) output for one row: da = g * smul + sadd, times the DropPath mask.
// The per-channel constants live in registers, so the kernels keep as few of them as the algebra allows:
//   pass 1 accumulates sum(dz) and sum(dz * (y - mean))            -> needs scale, shift (for z) and mean
//   pass 2 writes dy = gr * dz + A * y + B, with A = -rstd * gr * c2 and B = -gr * c1 - mean * A folded by the finalize
//   kernel (gr = gamma * rstd, c1 = sum(dz) / M, c2 = sum(dz * yhat) / M)   -> needs scale, shift, gr, A, B
struct BwdConst {
    float sc[8], sh[8], sm[8], sa[8];
    float bm;
    bool has_s;
};
__device__ __forceinline__ void bwd_load_const(const EwParams& p, const EwCtx& c, BwdConst& k) {
    load8(p.scale, c.cg, k.sc);
    load8(p.shift, c.cg, k.sh);
    k.has_s = p.smul != nullptr;
    if (k.has_s) {
        load8(p.smul + (size_t)c.smp * p.C, c.cg, k.sm);
        load8(p.sadd + (size_t)c.smp * p.C, c.cg, k.sa);
    }
    k.bm = p.bmul ? __ldg(p.bmul + c.smp) : 1.0f;
}
template <bool ACT>
__device__ __forceinline__ void bwd_row(const BwdConst& k, const uint4& ry, const uint4& rg, float (&dz)[8], float (&v)[8]) {
    float g[8];
    half8_to_float(ry, v);
    half8_to_float(rg, g);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float da = g[i];
        if (k.has_s) da = fmaf(da, k.sm[i], k.sa[i]);
        da *= k.bm;
        dz[i] = ACT ? da * dsilu_f(fmaf(v[i], k.sc[i], k.sh[i])) : da;
    }
}
constexpr int kBwdRows = 4;      // rows (pairs of 16-byte loads) in flight per thread

// BatchNorm backward, pass 1: partial sums of dz and dz * (y - mean)   (d beta, d gamma / rstd)
template <bool ACT>
__global__ void __launch_bounds__(kEwThreads) bn_bwd_reduce_kernel(EwParams p) {
    pdl_trigger();
    pdl_wait();
    const EwCtx c = ew_ctx(p);
    float acc[2][8] = {};
    if (c.active) {
        BwdConst k;
        bwd_load_const(p, c, k);
        float mu[8];
        load8(p.mean, c.cg, mu);
        for (int r = c.r0 + c.rl; r < c.r1; r += kBwdRows * c.RL) {
            uint4 ry[kBwdRows], rg[kBwdRows];
#pragma unroll
            for (int u = 0; u < kBwdRows; ++u) {
                const int rr = r + u * c.RL;
                if (rr < c.r1) { ry[u] = ldg_row(p.y, c.base + rr, p.C, c.cg); rg[u] = ldg_row(p.g, c.base + rr, p.C, c.cg); }
            }
#pragma unroll
            for (int u = 0; u < kBwdRows; ++u) {
                if (r + u * c.RL >= c.r1) break;
                float dz[8], v[8];
                bwd_row<ACT>(k, ry[u], rg[u], dz, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    acc[0][i] += dz[i];
                    acc[1][i] = fmaf(dz[i], v[i] - mu[i], acc[1][i]);
                }
            }
        }
    }
    ew_reduce_store<2>(p, c, acc);
}

// BatchNorm backward, pass 2: dy = gamma * rstd * (dz - mean(dz) - yhat * mean(dz * yhat)) = gr * dz + A * y + B
template <bool ACT>
__global__ void __launch_bounds__(kEwThreads) bn_bwd_apply_kernel(EwParams p) {
    pdl_trigger();
    pdl_wait();
    const EwCtx c = ew_ctx(p);
    if (!c.active) return;
    BwdConst k;
    bwd_load_const(p, c, k);
    float ca[8], cb[8], gr[8];
    load8(p.c1, c.cg, ca);       // A (see bn_bwd_finalize_kernel)
    load8(p.c2, c.cg, cb);       // B
    load8(p.gr, c.cg, gr);
    for (int r = c.r0 + c.rl; r < c.r1; r += kBwdRows * c.RL) {
        uint4 ry[kBwdRows], rg[kBwdRows];
#pragma unroll
        for (int u = 0; u < kBwdRows; ++u) {
            const int rr = r + u * c.RL;
            if (rr < c.r1) { ry[u] = ldg_row(p.y, c.base + rr, p.C, c.cg); rg[u] = ldg_row(p.g, c.base + rr, p.C, c.cg); }
        }
#pragma unroll
        for (int u = 0; u < kBwdRows; ++u) {
            const int rr = r + u * c.RL;
            if (rr >= c.r1) break;
            float dz[8], v[8], o[8];
            bwd_row<ACT>(k, ry[u], rg[u], dz, v);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = fmaf(gr[i], dz[i], fmaf(ca[i], v[i], cb[i]));
            store_row8(p.out, c.base + rr, p.C, c.cg, o);
        }
    }
}

// d gate[sample][c] = sum over the sample's rows of g * silu(bn(y))   (backward of x * gate, multidim_stacker.py:90)
__global__ void __launch_bounds__(kEwThreads) dgate_kernel(EwParams p) {
    pdl_trigger();
    pdl_wait();
    const EwCtx c = ew_ctx(p);
    float acc[1][8] = {};
    if (c.active) {
        float sc[8], sh[8];
        load8(p.scale, c.cg, sc);
        load8(p.shift, c.cg, sh);
        for (int r = c.r0 + c.rl; r < c.r1; r += c.RL) {
            float v[8], g[8];
            load_row8(p.y, c.base + r, p.C, c.cg, v);
            load_row8(p.g, c.base + r, p.C, c.cg, g);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[0][i] = fmaf(g[i], silu_f(fmaf(v[i], sc[i], sh[i])), acc[0][i]);
        }
    }
    ew_reduce_store<1>(p, c, acc);
}

// ---- finalize kernels: fixed-order reduction of the per-CTA partials, 8 channels per CTA, 128 part lanes ----------
constexpr int kFinCh = 8, kFinLanes = 128, kFinThreads = kFinCh * kFinLanes;
template <int NACC>
__device__ __forceinline__ bool fin_reduce(const float* partials, int nparts, int C, double (&tot)[NACC]) {
    __shared__ float s_p[NACC][kFinCh][kFinLanes + 1];
    const int cl = threadIdx.x & (kFinCh - 1), pl = threadIdx.x / kFinCh;
    const int ch = blockIdx.x * kFinCh + cl;
    float a[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) a[i] = 0.f;
    if (ch < C)
        for (int q = pl; q < nparts; q += kFinLanes)
#pragma unroll
            for (int i = 0; i < NACC; ++i) a[i] += partials[((size_t)q * NACC + i) * C + ch];
#pragma unroll
    for (int i = 0; i < NACC; ++i) s_p[i][cl][pl] = a[i];
    __syncthreads();
    if (threadIdx.x >= kFinCh || ch >= C) return false;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        double s = 0.0;
        for (int l = 0; l < kFinLanes; ++l) s += (double)s_p[i][cl][l];
        tot[i] = s;
    }
    return true;
}

struct BnFwdFin {
    const float* partials; int nparts;
    const float* ref;             // [C] the shift the partial sums were taken against (read before running_mean is updated)
    const float *gamma, *beta;
    float *running_mean, *running_var;
    float *scale, *shift, *mean, *rstd;
    int C; float count, eps, momentum;
};
__global__ void __launch_bounds__(kFinThreads) bn_fwd_finalize_kernel(BnFwdFin f) {
    pdl_trigger();
    pdl_wait();
    double tot[2];
    if (!fin_reduce<2>(f.partials, f.nparts, f.C, tot)) return;
    const int ch = blockIdx.x * kFinCh + (threadIdx.x & (kFinCh - 1));
    const double n = (double)f.count;
    const double d1 = tot[0] / n;
    const double mean = (double)f.ref[ch] + d1;
    double var = tot[1] / n - d1 * d1;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)f.eps));
    const float sc = f.gamma[ch] * rstd;
    f.scale[ch] = sc;
    f.shift[ch] = f.beta[ch] - (float)mean * sc;
    f.mean[ch] = (float)mean;
    f.rstd[ch] = rstd;
    // nn.BatchNorm train mode: running stats move towards (mean, unbiased variance) by `momentum`
    const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
    f.running_mean[ch] = (1.0f - f.momentum) * f.running_mean[ch] + f.momentum * (float)mean;
    f.running_var[ch] = (1.0f - f.momentum) * f.running_var[ch] + f.momentum * (float)unbiased;
}

struct BnBwdFin {
    const float* partials; int nparts;
    const float *gamma, *rstd, *mean;
    float *dgamma, *dbeta;        // gradient slices (still multiplied by the loss scale)
    float *c1, *c2, *gr;
    int C; float count;
};
// per-sample column sums: out[b][c] = scale * sum_q partials[b][q][c]   (SE squeeze, d gate); grid (C / kFinCh, b)
__global__ void __launch_bounds__(kFinThreads) colsum_finalize_kernel(const float* partials, int nparts, int C, float scale, float* out) {
    pdl_trigger();
    pdl_wait();
    double tot[1];
    if (!fin_reduce<1>(partials + (size_t)blockIdx.y * nparts * C, nparts, C, tot)) return;
    out[(size_t)blockIdx.y * C + blockIdx.x * kFinCh + (threadIdx.x & (kFinCh - 1))] = (float)tot[0] * scale;
}
__global__ void __launch_bounds__(kFinThreads) bn_bwd_finalize_kernel(BnBwdFin f) {
    pdl_trigger();
    pdl_wait();
    double tot[2];
    if (!fin_reduce<2>(f.partials, f.nparts, f.C, tot)) return;
    const int ch = blockIdx.x * kFinCh + (threadIdx.x & (kFinCh - 1));
    const double rstd = (double)f.rstd[ch], gr = (double)f.gamma[ch] * rstd;
    const double dgamma = tot[1] * rstd;                    // pass 1 accumulated dz * (y - mean)
    f.dbeta[ch] = (float)tot[0];
    f.dgamma[ch] = (float)dgamma;
    const double c1 = tot[0] / (double)f.count, c2 = dgamma / (double)f.count;
    const double A = -rstd * gr * c2;
    f.c1[ch] = (float)A;                                    // dy = gr * dz + A * y + B
    f.c2[ch] = (float)(-gr * c1 - (double)f.mean[ch] * A);
    f.gr[ch] = (float)gr;
}

// ------------------------------------------------------------------------------------------------------------
// Depthwise 3x3x3 convolution (conv_dw, multidim_stacker.py:110-113).  Forward and data gradient run on the streaming
// inference kernel (dwconv_kernel<3, 1, LIN>, dwconv.cuh) with tap-major weights; the data gradient correlates with
// the mirrored kernel.  The weight gradient below uses the same CTA shape and cp.async ring.
// ------------------------------------------------------------------------------------------------------------
// fp32 master weight [C][27] -> tap-major [27][C] and its mirror (tap 26 - k)
__global__ void __launch_bounds__(256) dw3_weights_kernel(const float* src, float* w27, float* w27_flip, int C) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * 256 + threadIdx.x;      // i = k * C + c
    if (i >= 27 * C) return;
    const int k = i / C, c = i - k * C;
    const float v = src[(size_t)c * 27 + k];
    w27[i] = v;
    w27_flip[(size_t)(26 - k) * C + c] = v;
}

// dW[c][tap] = sum_pos dy[pos] * in[pos + offset(tap)].  CTA = (sample, plane t, 64-channel slab, row chunk); warp = 5
// output columns, lane = one channel pair (packed f32x2).  A ring stage holds input row yi of the three planes
// t-1, t, t+1 plus dy row yi + 1; the three dy rows an input row pairs with (yi+1, yi, yi-1 <-> kernel rows 0, 1, 2)
// live in registers and shift by one per stage.  27 x 2 accumulators per thread, reduced over the 8 warps in a fixed
// order into one partial per CTA.
struct Dw3WgParams {
    const __half* in;       // [n][T][H][W][C]  conv_dw input (silu(bn1(.)))
    const __half* dy;       // [n][T][H][W][C]  gradient of the conv_dw output
    float* partials;        // [n * gridDim.y * xtiles][27][C]
    int n, T, H, W, C, rows_per_chunk, chunks, xtiles;
};
struct Dw3WgCfg {
    static constexpr int IW = kDwTWX + 2, NV = kDwPXW + 2;
    static constexpr int ROW_HALVES = IW * kDwCS;                       // one plane of one input row
    static constexpr int DY_HALVES = kDwTWX * kDwCS;
    static constexpr int STAGE_HALVES = 3 * ROW_HALVES + DY_HALVES;
    static constexpr int NST = 4;
    static constexpr int IN_CHUNKS = 3 * IW * 8, CHUNKS = IN_CHUNKS + kDwTWX * 8;
    static constexpr int SLOTS = (CHUNKS + 255) / 256;
    static constexpr size_t RING = (size_t)NST * STAGE_HALVES * 2;
    static constexpr size_t RED = (size_t)8 * 27 * kDwCS * sizeof(float);
    static constexpr size_t SMEM = RING > RED ? RING : RED;
};

__global__ void __launch_bounds__(256) dw3_wgrad_kernel(Dw3WgParams p) {
    pdl_trigger();
    pdl_wait();
    using Cfg = Dw3WgCfg;
    extern __shared__ __align__(16) unsigned char dwg_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int xt = blockIdx.x % p.xtiles, slab = blockIdx.x / p.xtiles;
    const int t = blockIdx.y / p.chunks, chunk = blockIdx.y - t * p.chunks;
    const int n = blockIdx.z;
    const int yo0 = chunk * p.rows_per_chunk, yo1 = min(p.H, yo0 + p.rows_per_chunk);
    const int xo0 = xt * kDwTWX;
    const int c_slab = slab * kDwCS;
    const int yi0 = yo0 - 1, NR = (yo1 - yo0) + 2;
    const size_t plane = (size_t)p.H * p.W * p.C;
    const long long row_pitch = (long long)p.W * p.C;
    const __half* in_n = p.in + (size_t)n * p.T * plane;
    const __half* dy_nt = p.dy + ((size_t)n * p.T + t) * plane;

    uint32_t s_off[Cfg::SLOTS];
    const __half* g_ptr[Cfg::SLOTS];
    bool s_ok[Cfg::SLOTS], s_dy[Cfg::SLOTS];
#pragma unroll
    for (int sl = 0; sl < Cfg::SLOTS; ++sl) {
        const int idx = tid + sl * 256;
        const bool is_dy = idx >= Cfg::IN_CHUNKS;
        s_dy[sl] = is_dy;
        if (!is_dy) {
            const int pl = idx / (Cfg::IW * 8);
            const int rem = idx - pl * (Cfg::IW * 8);
            const int px = rem >> 3, c16 = rem & 7;
            const int xi = xo0 + px - 1, ti = t + pl - 1, cc = c_slab + c16 * 8;
            s_ok[sl] = (xi >= 0) && (xi < p.W) && (ti >= 0) && (ti < p.T) && (cc < p.C);
            s_off[sl] = (uint32_t)(pl * Cfg::ROW_HALVES + px * kDwCS + c16 * 8) * 2u;
            g_ptr[sl] = s_ok[sl] ? in_n + (long long)ti * (long long)plane + (long long)yi0 * row_pitch + (long long)xi * p.C + cc : p.in;
        } else {
            const int rem = idx - Cfg::IN_CHUNKS;
            const int px = rem >> 3, c16 = rem & 7;
            const int cc = c_slab + c16 * 8;
            s_ok[sl] = (idx < Cfg::CHUNKS) && (xo0 + px < p.W) && (cc < p.C);
            s_off[sl] = (uint32_t)(3 * Cfg::ROW_HALVES + px * kDwCS + c16 * 8) * 2u;
            g_ptr[sl] = s_ok[sl] ? dy_nt + (long long)(yi0 + 1) * row_pitch + (long long)(xo0 + px) * p.C + cc : p.in;   // dy row yi + 1
        }
        if (idx >= Cfg::CHUNKS) s_off[sl] = 0xffffffffu;
    }
    const uint32_t ring_u32 = smem_u32(dwg_smem);
    uint32_t iss_off = 0;
    int iss_y = yi0;
    auto issue = [&]() {
        const bool in_ok = (unsigned)iss_y < (unsigned)p.H;
        const bool dy_ok = (iss_y + 1 >= yo0) && (iss_y + 1 < yo1);      // only this chunk's output rows contribute here
#pragma unroll
        for (int sl = 0; sl < Cfg::SLOTS; ++sl) {
            if (s_off[sl] != 0xffffffffu) {
                const bool ok = (s_dy[sl] ? dy_ok : in_ok) && s_ok[sl];
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(ring_u32 + iss_off + s_off[sl]),
                             "l"(ok ? g_ptr[sl] : p.in), "r"(ok ? 16 : 0));
                if (s_ok[sl]) g_ptr[sl] += row_pitch;
            }
        }
        ++iss_y;
        iss_off += Cfg::STAGE_HALVES * 2;
        if (iss_off == Cfg::NST * Cfg::STAGE_HALVES * 2) iss_off = 0;
    };
#pragma unroll
    for (int j = 0; j < Cfg::NST - 1; ++j) {
        if (j < NR) issue();
        cp_async_commit();
    }
    float2 acc[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[k] = make_float2(0.f, 0.f);
    float2 d0[kDwPXW], d1[kDwPXW], d2[kDwPXW];
#pragma unroll
    for (int j = 0; j < kDwPXW; ++j) d0[j] = d1[j] = d2[j] = make_float2(0.f, 0.f);

    uint32_t cons_off = 0;
    const uint32_t in_base = ring_u32 + (uint32_t)(warp * kDwPXW * kDwCS + 2 * lane) * 2u;
    const uint32_t dy_base = in_base + (uint32_t)(3 * Cfg::ROW_HALVES) * 2u;
    for (int k = 0; k < NR; ++k) {
        cp_async_wait<Cfg::NST - 2>();
        __syncthreads();
        if (k + Cfg::NST - 1 < NR) issue();
        cp_async_commit();
#pragma unroll
        for (int j = 0; j < kDwPXW; ++j) {
            d0[j] = d1[j];
            d1[j] = d2[j];
            d2[j] = lds_half2(dy_base + cons_off + j * kDwCS * 2);
        }
#pragma unroll
        for (int dt = 0; dt < 3; ++dt) {
            float2 v[Cfg::NV];
#pragma unroll
            for (int i = 0; i < Cfg::NV; ++i) v[i] = lds_half2(in_base + cons_off + (dt * Cfg::ROW_HALVES + i * kDwCS) * 2);
#pragma unroll
            for (int j = 0; j < kDwPXW; ++j)
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    acc[dt * 9 + 0 + s] = ffma2(d2[j], v[j + s], acc[dt * 9 + 0 + s]);
                    acc[dt * 9 + 3 + s] = ffma2(d1[j], v[j + s], acc[dt * 9 + 3 + s]);
                    acc[dt * 9 + 6 + s] = ffma2(d0[j], v[j + s], acc[dt * 9 + 6 + s]);
                }
        }
        cons_off += Cfg::STAGE_HALVES * 2;
        if (cons_off == Cfg::NST * Cfg::STAGE_HALVES * 2) cons_off = 0;
    }
    cp_async_wait<0>();
    __syncthreads();                                   // the ring is dead: reuse it for the cross-warp reduction
    float* s_acc = reinterpret_cast<float*>(dwg_smem);  // [8][27][64]
#pragma unroll
    for (int k = 0; k < 27; ++k)
        *reinterpret_cast<float2*>(s_acc + ((size_t)warp * 27 + k) * kDwCS + 2 * lane) = acc[k];
    __syncthreads();
    float* out = p.partials + (((size_t)n * gridDim.y + blockIdx.y) * p.xtiles + xt) * 27 * p.C;
    for (int i = tid; i < 27 * kDwCS; i += 256) {
        const int k = i >> 6, cl = i & 63;
        if (c_slab + cl >= p.C) continue;
        float sum = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) sum += s_acc[((size_t)wv * 27 + k) * kDwCS + cl];
        out[(size_t)k * p.C + c_slab + cl] = sum;
    }
}
// grad[c][tap] = sum over partials of partials[q][tap][c]
__global__ void __launch_bounds__(256) dw3_wgrad_reduce_kernel(const float* partials, int nparts, int C, float* grad) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * 256 + threadIdx.x;      // i = tap * C + c
    if (i >= 27 * C) return;
    float s = 0.f;
    for (int q = 0; q < nparts; ++q) s += partials[(size_t)q * 27 * C + i];
    const int k = i / C, c = i - k * C;
    grad[(size_t)c * 27 + k] = s;
}

// ------------------------------------------------------------------------------------------------------------
// Weight-gradient GEMM (reduction over rows): dW[n][k] = sum_m dY[m][n] * X[m][k], fp16 operands, fp32 accumulate.
// CTA = 64 (n) x 64 (k) output tile over one slice of M, 4 warps of 32 x 32; both operands are stored with the
// reduction index slowest, so both fragments come from ldmatrix.trans.  Split-M partials are reduced in a fixed order.
// ------------------------------------------------------------------------------------------------------------
struct WgradParams {
    const __half* dY;     // [M][N]
    const __half* X;      // [M][K]
    float* partials;      // [splits][N][K]
    long long M;
    int N, K, rows_per_split;
};
constexpr int kWgBM = 32, kWgPitch = 72, kWgStages = 3;

__global__ void __launch_bounds__(128) wgrad_gemm_kernel(WgradParams p) {
    pdl_trigger();
    pdl_wait();
    __shared__ __align__(16) __half s_a[kWgStages][kWgBM * kWgPitch];
    __shared__ __align__(16) __half s_b[kWgStages][kWgBM * kWgPitch];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wn = warp & 1, wk = warp >> 1;
    const int k0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const long long m_begin = (long long)blockIdx.z * p.rows_per_split;
    long long m_end = m_begin + p.rows_per_split;
    if (m_end > p.M) m_end = p.M;
    const int steps = m_end > m_begin ? (int)((m_end - m_begin + kWgBM - 1) / kWgBM) : 0;

    auto load_stage = [&](int step, int st) {
        const long long m0 = m_begin + (long long)step * kWgBM;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int q = tid + i * 128;          // 256 chunks of 16 bytes per operand tile
            const int r = q >> 3, cc = q & 7;
            const bool ok = m0 + r < m_end;
            const __half* sa = ok ? p.dY + (size_t)(m0 + r) * p.N + n0 + cc * 8 : p.dY;
            const __half* sb = ok ? p.X + (size_t)(m0 + r) * p.K + k0 + cc * 8 : p.X;
            cp_async16(&s_a[st][r * kWgPitch + cc * 8], sa, ok ? 16 : 0);
            cp_async16(&s_b[st][r * kWgPitch + cc * 8], sb, ok ? 16 : 0);
        }
    };
#pragma unroll
    for (int s = 0; s < kWgStages - 1; ++s) {
        if (s < steps) load_stage(s, s);
        cp_async_commit();
    }
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

    const int mi = lane >> 3, lr = lane & 7;
    // A (rows = n, cols = m): matrices (m 0-7, n 0-7), (m 0-7, n 8-15), (m 8-15, n 0-7), (m 8-15, n 8-15)
    const int a_m = (mi >> 1) * 8 + lr, a_n = wn * 32 + (mi & 1) * 8;
    // B (rows = m, cols = k): matrices (m 0-7, k 0-7), (m 8-15, k 0-7), (m 0-7, k 8-15), (m 8-15, k 8-15)
    const int b_m = (mi & 1) * 8 + lr, b_k = wk * 32 + (mi >> 1) * 8;

    for (int step = 0; step < steps; ++step) {
        cp_async_wait<kWgStages - 2>();
        __syncthreads();
        {
            const int nx = step + kWgStages - 1;
            if (nx < steps) load_stage(nx, nx % kWgStages);
            cp_async_commit();
        }
        const __half* sa = s_a[step % kWgStages];
        const __half* sb = s_b[step % kWgStages];
#pragma unroll
        for (int ms = 0; ms < kWgBM / 16; ++ms) {
            uint32_t a[2][4], bf[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
                ldmatrix_x4_trans(a[i], smem_u32(sa + (ms * 16 + a_m) * kWgPitch + a_n + i * 16));
#pragma unroll
            for (int j = 0; j < 2; ++j)
                ldmatrix_x4_trans(bf[j], smem_u32(sb + (ms * 16 + b_m) * kWgPitch + b_k + j * 16));
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    mma16816(acc[i][2 * j], a[i], bf[j][0], bf[j][1]);
                    mma16816(acc[i][2 * j + 1], a[i], bf[j][2], bf[j][3]);
                }
        }
    }
    cp_async_wait<0>();
    float* out = p.partials + (size_t)blockIdx.z * p.N * p.K;
    const int g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + wn * 32 + i * 16 + g;
            const int k = k0 + wk * 32 + j * 8 + tq * 2;
            *reinterpret_cast<float2*>(out + (size_t)n * p.K + k) = make_float2(acc[i][j][0], acc[i][j][1]);
            *reinterpret_cast<float2*>(out + (size_t)(n + 8) * p.K + k) = make_float2(acc[i][j][2], acc[i][j][3]);
        }
}
__global__ void __launch_bounds__(256) sum_partials_kernel(const float* partials, int nparts, size_t count, float* out) {
    pdl_trigger();
    pdl_wait();
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= count) return;
    float s = 0.f;
    for (int q = 0; q < nparts; ++q) s += partials[(size_t)q * count + i];
    out[i] = s;
}

// ------------------------------------------------------------------------------------------------------------
// Squeeze-and-excitation, train mode (multidim_stacker.py:85-90): forward saves what the backward needs
// ------------------------------------------------------------------------------------------------------------
struct SeTrainParams {
    const float* sums;                      // fwd: squeeze s[b][C] (mean of silu(bn2)); bwd: d gate[b][C] (colsum_finalize_kernel)
    const float *w1, *b1, *w2, *b2;         // conv_reduce [rd][C], [rd]; conv_expand [C][rd], [C]
    float *hpre, *gate;                     // saved by the forward: [b][rd], [b][C]
    float *dgpre, *dhpre, *sadd;            // backward outputs: [b][C], [b][rd], [b][C]
    int C, rd; float inv_count;
};
// grid b, 256 threads, dynamic smem (2 * C + 2 * rd) floats
__global__ void __launch_bounds__(256) se_train_fwd_kernel(SeTrainParams p) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float s_se[];
    float* s_s = s_se;
    float* s_h = s_se + p.C;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int c = tid; c < p.C; c += 256) s_s[c] = p.sums[(size_t)b * p.C + c];
    __syncthreads();
    for (int j = warp; j < p.rd; j += 8) {
        float a = 0.f;
        for (int c = lane; c < p.C; c += 32) a = fmaf(__ldg(p.w1 + (size_t)j * p.C + c), s_s[c], a);
        a = warp_sum(a);
        if (lane == 0) {
            a += p.b1[j];
            p.hpre[(size_t)b * p.rd + j] = a;
            s_h[j] = silu_f(a);
        }
    }
    __syncthreads();
    for (int c = tid; c < p.C; c += 256) {
        float a = p.b2[c];
        for (int j = 0; j < p.rd; ++j) a = fmaf(__ldg(p.w2 + (size_t)c * p.rd + j), s_h[j], a);
        p.gate[(size_t)b * p.C + c] = sigmoid_f(a);
    }
}
// grid b: d gate -> d(pre-sigmoid), d(pre-SiLU hidden), d squeeze / row count
__global__ void __launch_bounds__(256) se_train_bwd_kernel(SeTrainParams p) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float s_se[];
    float* s_dg = s_se;             // [C] d gpre
    float* s_dh = s_se + p.C;       // [rd] d hpre
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int c = tid; c < p.C; c += 256) {
        const float g = p.gate[(size_t)b * p.C + c];
        const float d = p.sums[(size_t)b * p.C + c] * g * (1.0f - g);
        s_dg[c] = d;
        p.dgpre[(size_t)b * p.C + c] = d;
    }
    __syncthreads();
    for (int j = warp; j < p.rd; j += 8) {
        float a = 0.f;
        for (int c = lane; c < p.C; c += 32) a = fmaf(__ldg(p.w2 + (size_t)c * p.rd + j), s_dg[c], a);
        a = warp_sum(a);
        if (lane == 0) {
            const float d = a * dsilu_f(p.hpre[(size_t)b * p.rd + j]);
            s_dh[j] = d;
            p.dhpre[(size_t)b * p.rd + j] = d;
        }
    }
    __syncthreads();
    for (int c = tid; c < p.C; c += 256) {
        float a = 0.f;
        for (int j = 0; j < p.rd; ++j) a = fmaf(__ldg(p.w1 + (size_t)j * p.C + c), s_dh[j], a);
        p.sadd[(size_t)b * p.C + c] = a * p.inv_count;
    }
}
// parameter gradients of the two SE convolutions, summed over samples in a fixed order
struct SeGradParams {
    const float *s, *hpre, *dgpre, *dhpre;
    float *dw1, *db1, *dw2, *db2;
    int b, C, rd;
};
__global__ void __launch_bounds__(256) se_train_wgrad_kernel(SeGradParams p) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int nW = p.C * p.rd;
    if (i < nW) {
        {   // dw2[c][j] = sum_b dgpre[b][c] * silu(hpre[b][j])
            const int c = i / p.rd, j = i - c * p.rd;
            float a = 0.f;
            for (int b = 0; b < p.b; ++b) a = fmaf(p.dgpre[(size_t)b * p.C + c], silu_f(p.hpre[(size_t)b * p.rd + j]), a);
            p.dw2[i] = a;
        }
        {   // dw1[j][c] = sum_b dhpre[b][j] * s[b][c]
            const int j = i / p.C, c = i - j * p.C;
            float a = 0.f;
            for (int b = 0; b < p.b; ++b) a = fmaf(p.dhpre[(size_t)b * p.rd + j], p.s[(size_t)b * p.C + c], a);
            p.dw1[i] = a;
        }
    }
    if (i < p.C) {
        float a = 0.f;
        for (int b = 0; b < p.b; ++b) a += p.dgpre[(size_t)b * p.C + i];
        p.db2[i] = a;
    }
    if (i < p.rd) {
        float a = 0.f;
        for (int b = 0; b < p.b; ++b) a += p.dhpre[(size_t)b * p.rd + i];
        p.db1[i] = a;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Head: GeM with learnable p (multidim_stacker.py:20-45), dropout, classifier, focal loss (src/losses.py:31-48)
// ------------------------------------------------------------------------------------------------------------
constexpr int kGemChunks = 8;       // CTAs per (sample, t) plane in the GeM kernels
struct GemTrainParams {
    const __half* x;        // [b][T][P][C] = silu(bn(conv3d_projection))
    const float* p;         // device scalar (global_pool.p)
    float* partials;        // [b][T][kGemChunks][2][C]: sums of c^p and c^p * ln c   (c = max(x, eps))
    int T, P, C; float eps;
};
// grid (T, b, kGemChunks)
__global__ void __launch_bounds__(256) gem_train_fwd_kernel(GemTrainParams g) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_part[2][8][260];
    const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int C8 = g.C >> 3, lanes_p = min(256 / C8, 8);
    const int cg = tid % C8, pl = tid / C8;
    const float pw = __ldg(g.p);
    const int per = (g.P + kGemChunks - 1) / kGemChunks;
    const int p0 = blockIdx.z * per, p1 = min(g.P, p0 + per);
    float acc[8] = {}, accl[8] = {};
    const __half* base = g.x + (((size_t)b * g.T + t) * g.P) * g.C + cg * 8;
    if (pl < lanes_p) {
        for (int pos = p0 + pl; pos < p1; pos += lanes_p) {
            float v[8];
            half8_to_float(ldg16(base + (size_t)pos * g.C), v);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float c = fmaxf(v[i], g.eps);
                const float lc = __logf(c);
                const float cp = __expf(pw * lc);
                acc[i] += cp;
                accl[i] = fmaf(cp, lc, accl[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { s_part[0][pl][cg * 8 + i] = acc[i]; s_part[1][pl][cg * 8 + i] = accl[i]; }
    }
    __syncthreads();
    if (tid < g.C) {
        float s = 0.f, sl = 0.f;
        for (int l = 0; l < lanes_p; ++l) { s += s_part[0][l][tid]; sl += s_part[1][l][tid]; }
        float* o = g.partials + ((((size_t)b * g.T + t) * kGemChunks + blockIdx.z) * 2) * g.C;
        o[tid] = s;
        o[g.C + tid] = sl;
    }
}

struct HeadTrainParams {
    const float* gem_partials;             // gem_train_fwd_kernel output
    float *feat, *pooled, *mlog;           // [b][F]: mean^(1/p), mean of c^p, mean of c^p * ln c
    const float* dmask;                    // [b][F] dropout mask (0 or 1/(1-p)) or nullptr
    const float *w, *bias;                 // classifier [K][F], [K]
    const float* targets;                  // [b][K]
    const float* gem_p;                    // device scalar
    const float* scaler;                   // [0] = loss scale
    float *logits, *loss;                  // [b][K], [1] (unscaled)
    float *dw, *dbias, *dgem_p;            // gradient slices (scaled)
    float* logit_partials;                 // [b][ceil(F / 256)][K]
    float* dp_partials;                    // [gridDim.x] partial sums of d p
    float* coef;                           // [b][F]: d feat * mean^(1/p - 1) / (p P), consumed by gem_bwd_kernel
    int b, F, K, P, C; float alpha, gamma;
};
// grid (ceil(F / 256), b), thread = one pooled feature: finish GeM (pooled, mlog, feat) and reduce this CTA's share of
// the sample's logits into logit_partials[b][gridDim.x][K]
__global__ void __launch_bounds__(256) head_logits_kernel(HeadTrainParams h) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_red[8];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = blockIdx.x * 256 + tid;
    const float pw = __ldg(h.gem_p);
    float f = 0.f;
    if (i < h.F) {
        const size_t o = (size_t)b * h.F + i;
        const int tt = i / h.C, c = i - tt * h.C;
        const float* q = h.gem_partials + (((size_t)b * (h.F / h.C) + tt) * kGemChunks * 2) * h.C + c;
        float s = 0.f, sl = 0.f;
#pragma unroll
        for (int z = 0; z < kGemChunks; ++z) { s += q[(size_t)z * 2 * h.C]; sl += q[(size_t)(z * 2 + 1) * h.C]; }
        const float m = s / (float)h.P;
        f = powf(m, 1.0f / pw);
        h.pooled[o] = m;
        h.mlog[o] = sl / (float)h.P;
        h.feat[o] = f;
        if (h.dmask) f *= h.dmask[o];
    }
    for (int k = 0; k < h.K; ++k) {
        const float a = warp_sum(i < h.F ? f * __ldg(h.w + (size_t)k * h.F + i) : 0.f);
        if (lane == 0) s_red[warp] = a;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int w = 0; w < 8; ++w) t += s_red[w];
            h.logit_partials[((size_t)b * gridDim.x + blockIdx.x) * h.K + k] = t;
        }
        __syncthreads();
    }
}
// grid ceil(F / 256), thread = one pooled feature: focal loss gradient (recomputed per CTA from the b * K logits),
// d classifier, d feat -> GeM backward coefficients, partial d p.  dynamic smem: 2 * b * K floats
__global__ void __launch_bounds__(256) head_grad_kernel(HeadTrainParams h) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float s_dl[];          // [b][K] d loss / d logit (scaled), then [b][K] logits
    __shared__ float s_red[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float pw = __ldg(h.gem_p);
    const int BK = h.b * h.K;
    const float scale = h.scaler[0] / (float)BK;
    float* s_x = s_dl + BK;
    for (int o = tid; o < BK; o += 256) {
        const int bb = o / h.K, k = o - bb * h.K;
        float x = h.bias[k];
        for (int q = 0; q < (int)gridDim.x; ++q) x += h.logit_partials[((size_t)bb * gridDim.x + q) * h.K + k];
        if (blockIdx.x == 0) h.logits[o] = x;
        s_x[o] = x;
        const float t = h.targets[o];
        const float pr = 1.0f / (1.0f + expf(-x));
        const float ce = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
        const float q = fminf(fmaxf(pr + t - 2.0f * pr * t, 0.f), 1.0f);          // 1 - p_t
        const float at = h.alpha >= 0.f ? h.alpha * t + (1.0f - h.alpha) * (1.0f - t) : 1.0f;
        const float qg = powf(q, h.gamma);
        const float dq = q > 0.f ? h.gamma * powf(q, h.gamma - 1.0f) * (1.0f - 2.0f * t) * pr * (1.0f - pr) : 0.f;
        s_dl[o] = at * ((pr - t) * qg + ce * dq) * scale;
    }
    __syncthreads();
    if (blockIdx.x == 0) {
        if (tid == 0) {                                   // loss value, fixed summation order
            float total = 0.f;
            for (int o = 0; o < BK; ++o) {
                const float x = s_x[o], t = h.targets[o];
                const float pr = 1.0f / (1.0f + expf(-x));
                const float ce = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
                const float q = fminf(fmaxf(pr + t - 2.0f * pr * t, 0.f), 1.0f);
                const float at = h.alpha >= 0.f ? h.alpha * t + (1.0f - h.alpha) * (1.0f - t) : 1.0f;
                total += at * ce * powf(q, h.gamma);
            }
            h.loss[0] = total / (float)BK;
        }
        if (tid < h.K) {
            float a = 0.f;
            for (int b = 0; b < h.b; ++b) a += s_dl[b * h.K + tid];
            h.dbias[tid] = a;
        }
    }
    float dp_acc = 0.f;
    const int i = blockIdx.x * 256 + tid;
    if (i < h.F) {
        for (int k = 0; k < h.K; ++k) {                   // d classifier.weight
            float a = 0.f;
            for (int b = 0; b < h.b; ++b) {
                const float m = h.dmask ? h.dmask[(size_t)b * h.F + i] : 1.0f;
                a = fmaf(s_dl[b * h.K + k], h.feat[(size_t)b * h.F + i] * m, a);
            }
            h.dw[(size_t)k * h.F + i] = a;
        }
        for (int b = 0; b < h.b; ++b) {                   // d feat -> GeM backward coefficients and d p
            float df = 0.f;
            for (int k = 0; k < h.K; ++k) df = fmaf(__ldg(h.w + (size_t)k * h.F + i), s_dl[b * h.K + k], df);
            if (h.dmask) df *= h.dmask[(size_t)b * h.F + i];
            const size_t o = (size_t)b * h.F + i;
            const float m = h.pooled[o], f = h.feat[o];
            // f = m^(1/p):  df/dm = f / (p m);  df/dp = f * (-ln(m) / p^2 + mlog / (p m))
            h.coef[o] = df * f / (pw * m * (float)h.P);
            dp_acc += df * f * (-__logf(m) / (pw * pw) + h.mlog[o] / (pw * m));
        }
    }
    dp_acc = warp_sum(dp_acc);
    if (lane == 0) s_red[warp] = dp_acc;
    __syncthreads();
    if (tid == 0) {
        float a = 0.f;
        for (int w = 0; w < 8; ++w) a += s_red[w];
        h.dp_partials[blockIdx.x] = a;
    }
}
__global__ void head_dp_reduce_kernel(const float* partials, int n, float* out) {
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float a = 0.f;
    for (int i = 0; i < n; ++i) a += partials[i];
    out[0] = a;
}

// sigmoid focal loss (src/losses.py:31-48, mean reduction) and the sigmoid probabilities of a small logit matrix: the
// validation step (BallActionModel.val_step, src/argus_models.py:76-91).  One CTA; the sum runs in a fixed order.
__global__ void __launch_bounds__(256) focal_loss_kernel(const float* logits, const float* targets, int n, float alpha, float gamma,
                                                         float* loss, float* probs) {
    pdl_trigger();
    pdl_wait();
    for (int o = threadIdx.x; o < n; o += 256) probs[o] = 1.0f / (1.0f + expf(-logits[o]));
    if (threadIdx.x != 0) return;
    float total = 0.f;
    for (int o = 0; o < n; ++o) {
        const float x = logits[o], t = targets[o];
        const float pr = 1.0f / (1.0f + expf(-x));
        const float ce = fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
        const float q = fminf(fmaxf(pr + t - 2.0f * pr * t, 0.f), 1.0f);          // 1 - p_t
        const float at = alpha >= 0.f ? alpha * t + (1.0f - alpha) * (1.0f - t) : 1.0f;
        total += at * ce * powf(q, gamma);
    }
    loss[0] = total / (float)n;
}

// d x[b][t][pos][c] = coef[b][t*C + c] * p * c^(p-1) for x >= eps (clamp passes no gradient below eps); grid (T, b, kGemChunks)
struct GemBwdParams {
    const __half* x; const float* coef; const float* p; __half* dx;
    int T, P, C; float eps;
};
__global__ void __launch_bounds__(256) gem_bwd_kernel(GemBwdParams g) {
    pdl_trigger();
    pdl_wait();
    const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int C8 = g.C >> 3, lanes_p = 256 / C8;
    const int cg = tid % C8, pl = tid / C8;
    if (pl >= lanes_p) return;
    const float pw = __ldg(g.p);
    const int per = (g.P + kGemChunks - 1) / kGemChunks;
    const int p0 = blockIdx.z * per, p1 = min(g.P, p0 + per);
    float cf[8];
    load8(g.coef + (size_t)b * g.T * g.C + (size_t)t * g.C, cg, cf);
    const size_t base = (((size_t)b * g.T + t) * g.P) * g.C + cg * 8;
    for (int pos = p0 + pl; pos < p1; pos += lanes_p) {
        float v[8], o[8];
        half8_to_float(ldg16(g.x + base + (size_t)pos * g.C), v);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = v[i] >= g.eps ? cf[i] * pw * __expf((pw - 1.0f) * __logf(v[i])) : 0.f;
        *reinterpret_cast<uint4*>(g.dx + base + (size_t)pos * g.C) = float8_to_half(o);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Optimizer: GradScaler inf check + unscale, SGD with Nesterov momentum (torch.optim.SGD), derived fp16 weights
// scaler[0] = loss scale, [1] = growth tracker, [2] = found_inf of this step, [3] = optimizer steps performed
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grad_check_kernel(const float* grad, size_t count, float* scaler) {
    pdl_trigger();
    pdl_wait();
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < count; i += (size_t)gridDim.x * 256)
        bad |= !isfinite(grad[i]);
    if (__syncthreads_or(bad) && threadIdx.x == 0) scaler[2] = 1.0f;      // idempotent flag write, no ordering needed
}
__global__ void __launch_bounds__(256) sgd_nesterov_kernel(float* param, const float* grad, float* mom, size_t count,
                                                           const float* scaler, float lr, float momentum, int nesterov) {
    pdl_trigger();
    pdl_wait();
    if (scaler[2] != 0.0f) return;                        // GradScaler.step skips the update on inf / nan
    const float inv = 1.0f / scaler[0];
    const bool first = scaler[3] == 0.0f;                 // momentum buffer starts as a copy of the first gradient
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < count; i += (size_t)gridDim.x * 256) {
        const float g = grad[i] * inv;
        const float buf = first ? g : fmaf(momentum, mom[i], g);
        mom[i] = buf;
        const float d = nesterov ? fmaf(momentum, buf, g) : buf;
        param[i] -= lr * d;
    }
}
__global__ void scaler_update_kernel(float* scaler, float growth, float backoff, float interval, int dynamic) {
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (scaler[2] != 0.0f) {
        if (dynamic) { scaler[0] *= backoff; scaler[1] = 0.0f; }
    } else {
        scaler[3] += 1.0f;
        if (dynamic) {
            scaler[1] += 1.0f;
            if (scaler[1] >= interval) { scaler[0] *= growth; scaler[1] = 0.0f; }
        }
    }
    scaler[2] = 0.0f;
}
// ModelEma.update (src/ema.py:49-57): ema = decay * ema + (1 - decay) * value, in the reference's float32 arithmetic
// (two rounded products, one rounded sum -- no fused multiply-add, so the result is bit-identical to torch)
__global__ void __launch_bounds__(256) ema_update_kernel(float* ema, const float* cur, size_t count, float decay, float one_minus) {
    pdl_trigger();
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < count; i += (size_t)gridDim.x * 256)
        ema[i] = __fadd_rn(__fmul_rn(decay, ema[i]), __fmul_rn(one_minus, cur[i]));
}
// fp32 master weight [R][Cc] -> fp16 copy and fp16 transposed copy [Cc][R] (operands of the forward / dgrad GEMMs);
// blockIdx.z selects the weight from a device table so that one launch refreshes all of them
struct CastJob { const float* src; __half* dst; __half* dst_t; int R, Cc; };
__global__ void __launch_bounds__(256) cast_transpose_kernel(const CastJob* jobs) {
    pdl_trigger();
    pdl_wait();
    __shared__ float tile[32][33];
    const CastJob j = jobs[blockIdx.z];
    const float* src = j.src; __half* dst = j.dst; __half* dst_t = j.dst_t;
    const int R = j.R, Cc = j.Cc;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    if (r0 >= R || c0 >= Cc) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        float v = 0.f;
        if (r < R && c < Cc) {
            v = src[(size_t)r * Cc + c];
            dst[(size_t)r * Cc + c] = __float2half_rn(v);
        }
        tile[i][tx] = v;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        if (r < R && c < Cc) dst_t[(size_t)c * R + r] = __float2half_rn(tile[tx][i]);
    }
}

// counter-based Bernoulli masks (DropPath / Dropout) when the caller does not supply them: splitmix64 of (seed, index)
__global__ void __launch_bounds__(256) bernoulli_mask_kernel(float* out, size_t count, float keep, unsigned long long seed) {
    pdl_trigger();
    pdl_wait();
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= count) return;
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u = (float)(z >> 40) * (1.0f / 16777216.0f);
    out[i] = u < keep ? 1.0f / keep : 0.0f;
}

}  // namespace mds
