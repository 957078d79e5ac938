// K4/K6 + K5: depthwise 3x3 (2D, stride 1/2, TF-SAME) and 3x3x3 (3D, pad 1) convolution + folded BN + SiLU, NHWC fp16,
// the SE squeeze AND the SE excitation MLP in one launch
// (timm InvertedResidual conv_dw/bn2/se; InvertedResidual3d + SqueezeExcite, multidim_stacker.py:110-114, 72-90).
//
// A CTA owns (image, [out plane t], 40 output columns, 64-channel slab, row chunk).  Input rows are staged by TMA
// (cp.async.bulk.tensor over the NHWC tensor as a 4-D / 5-D map, box = 64 channels x 42 (81) columns x rows [x 3 planes],
// hardware zero fill = the conv padding, rows beyond the image and the channel tail) into an mbarrier ring; one elected
// thread issues the loads two stages ahead, nobody computes an address.  Warp w computes output columns [5w, 5w+5), lane
// l owns channels (2l, 2l+1) of the slab as one packed f32x2: every shared-memory read is a conflict-free LDS.32, every
// global store a 128-byte line per pixel, the MACs are FFMA2 and bias / SiLU / squeeze use FADD2 / FMUL2.  Each input row
// is scattered into the three output rows it feeds; the three accumulator sets rotate their roles by unrolling the row
// loop three times (no register moves).
// SE: every CTA leaves one squeeze partial per channel; the CTA that finishes LAST for an image (device-scope counter)
// adds the partials in a fixed order, evaluates conv_reduce -> SiLU -> conv_expand -> sigmoid and writes the image's
// fp32 gate vector, which the projection GEMM applies to its A operand.  Row chunks depend on the layer shape only, so
// the result for an image does not depend on the rest of the batch, and there are no floating-point atomics.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "tc_helpers.cuh"
#include "gemm_tc.cuh"

namespace mds {

struct DwSeParams {
    __half* out;             // [n][T][Ho][Wo][C]
    const float* w;          // [taps][C]  tap = (dt*3 + r)*3 + s, BN scale folded
    const float* bias;       // [C]
    float* partials;         // [n][nparts][C]
    const float* se_w1;      // [rd][C]   (nullptr: no SE, partials only)
    const float* se_b1;      // [rd]
    const float* se_w2t;     // [rd][C]
    const float* se_b2;      // [C]
    float* gate;             // [n][C] fp32
    int* done;               // [n] arrival counters, zero on entry, left zero
    int n, T, H, W, C, Ho, Wo, rd;
    float inv_count;
    int rows_per_chunk, chunks, xtiles, slabs, nparts;
    int ctas_per_img;        // T * chunks * xtiles * slabs
};

constexpr int kDtCS = 64, kDtPXW = 5, kDtTWX = 40;

template <int KT, int STRIDE>
struct DwTmaCfg {
    static constexpr int IW = (STRIDE == 1) ? kDtTWX + 2 : 2 * kDtTWX + 1;
    static constexpr int NV = (STRIDE == 1) ? kDtPXW + 2 : 2 * kDtPXW + 1;
    static constexpr int RPS = (KT == 3) ? 1 : (STRIDE == 1 ? 3 : 2);         // input rows per stage
    static constexpr int ROWB = IW * kDtCS * 2;                                // bytes of one input row tile (one plane)
    static constexpr int STAGEB = KT * RPS * ROWB;
    static constexpr int NST = 3;
    // ring (re-used after the row loop as s_part[8][64] | s_mean[1152] | s_hid[64]) | s_w[27][64] (3D) | barriers
    static constexpr size_t SMEM = 128 + (size_t)NST * STAGEB + (KT == 3 ? 27 * 64 : 0) * sizeof(float) + 64;
    static_assert((size_t)NST * STAGEB >= (8 * 64 + 1152 + 64) * sizeof(float), "the SE scratch aliases the ring");
};

__device__ __forceinline__ void dt_tma_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ float dt_ld_cg(const float* p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 dt_lds_half2(uint32_t saddr) {
    uint32_t u;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(saddr));
    return unpack_half2(u);
}
// SiLU of two channels on the packed fp32 pipe: FMUL2, 2 x MUFU.EX2, FADD2, 2 x MUFU.RCP, FMUL2
__device__ __forceinline__ float2 dt_silu2(float2 x) {
    const float2 t = __fmul2_rn(x, make_float2(-1.4426950408889634f, -1.4426950408889634f));
    float2 e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(t.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(t.y));
    const float2 d = __fadd2_rn(e, make_float2(1.0f, 1.0f));
    float2 r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(d.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(d.y));
    return __fmul2_rn(x, r);
}

template <int KT, int STRIDE>
__global__ void __launch_bounds__(256, 3) dwconv_tma_kernel(const __grid_constant__ CUtensorMap tmIn, DwSeParams p) {
    using Cfg = DwTmaCfg<KT, STRIDE>;
    extern __shared__ unsigned char dt_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dt_smem_raw) + 127) & ~uintptr_t(127));
    unsigned char* s_ring = smem;
    float* s_part = reinterpret_cast<float*>(s_ring);      // [8][64]   } alias the ring: only touched after the row loop, when every
    float* s_mean = s_part + 8 * 64;                       // [1152]    } TMA load has landed and every warp has consumed its rows
    float* s_hid = s_mean + 1152;                          // [64]      }
    float* s_w = reinterpret_cast<float*>(s_ring + (size_t)Cfg::NST * Cfg::STAGEB);         // [27][64] (KT == 3)
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + (KT == 3 ? 27 * kDtCS : 0));
    uint64_t* full = bars;                // [3]
    uint64_t* empty = bars + 3;           // [3]
    int* s_last = reinterpret_cast<int*>(bars + 6);

    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int xt = blockIdx.x % p.xtiles, slab = blockIdx.x / p.xtiles;
    const int t = blockIdx.y / p.chunks, chunk = blockIdx.y - t * p.chunks;
    const int n = blockIdx.z;
    const int yo0 = chunk * p.rows_per_chunk, yo1 = min(p.Ho, yo0 + p.rows_per_chunk);
    const int c_slab = slab * kDtCS;
    const int c = c_slab + 2 * lane;
    const bool c_ok = c < p.C;
    const int c_ld = c_ok ? c : 0;
    const int yi0 = (STRIDE == 1) ? yo0 - 1 : 2 * yo0;
    const int NR = (STRIDE == 1) ? (yo1 - yo0) + 2 : 2 * (yo1 - yo0) + 1;
    const int xi0 = (STRIDE == 1) ? xt * kDtTWX - 1 : 2 * xt * kDtTWX;
    const int nstg = (NR + Cfg::RPS - 1) / Cfg::RPS;

    if (tid == 0) {
        for (int i = 0; i < Cfg::NST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmIn) : "memory");
    }
    // taps: 2D: 9 x f32x2 in registers.  3D: 27 do not fit next to the accumulators; they are staged in shared memory once per
    // CTA and read per plane (LDS.64).  (Measured: reading the 2D taps from shared memory as well brings the kernel to 64
    // registers and 4 CTAs / SM, but it is 10 % slower: the kernel is bound by issue slots and the FMA / SFU pipes, not by
    // latency, see profiles/experiments_r02.md.)
    float2 wreg[9];
    if constexpr (KT == 1) {
#pragma unroll
        for (int i = 0; i < 9; ++i) wreg[i] = __ldg(reinterpret_cast<const float2*>(p.w + (size_t)i * p.C + c_ld));
    } else {
        for (int i = tid; i < KT * 9 * kDtCS; i += 256) {
            const int tap = i / kDtCS, cc = i - tap * kDtCS;
            s_w[i] = (c_slab + cc < p.C) ? __ldg(p.w + (size_t)tap * p.C + c_slab + cc) : 0.f;
        }
    }
    const uint32_t sw_lane = smem_u32(s_w) + (uint32_t)(2 * lane) * 4u;
    auto lds_w = [&](int tap) {      // this lane's two channels of one tap
        if constexpr (KT == 1) {
            return wreg[tap];
        } else {
            float2 r;
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(sw_lane + (uint32_t)(tap * kDtCS) * 4u));
            return r;
        }
    };
    const float2 bias = __ldg(reinterpret_cast<const float2*>(p.bias + c_ld));
    __syncthreads();
    pdl_wait();       // weights above are constants; the input rows below come from the previous kernel

    auto issue = [&](int js) {      // elected lane of warp 0: stage use js -> ring slot js % NST
        const int st = js % Cfg::NST;
        mbar_wait(&empty[st], ((js / Cfg::NST) & 1) ^ 1);
        mbar_expect_tx(&full[st], (uint32_t)Cfg::STAGEB);
        if constexpr (KT == 3)
            tma_load_5d(s_ring + (size_t)st * Cfg::STAGEB, &tmIn, &full[st], c_slab, xi0, yi0 + js, t - 1, n);
        else
            dt_tma_4d(s_ring + (size_t)st * Cfg::STAGEB, &tmIn, &full[st], c_slab, xi0, yi0 + js * Cfg::RPS, n);
    };
    if (warp == 0) {              // one elected lane of the converged warp issues (no ptxas election loop around UTMALDG)
        if (elect_one()) {
#pragma unroll
            for (int j = 0; j < Cfg::NST - 1; ++j)
                if (j < nstg) issue(j);
        }
        __syncwarp();
    }

    float2 lsum = make_float2(0.f, 0.f);
    float2 aX[kDtPXW], aY[kDtPXW], aZ[kDtPXW];          // three output rows in flight; they start from the bias
#pragma unroll
    for (int j = 0; j < kDtPXW; ++j) aX[j] = aY[j] = aZ[j] = bias;
    const int xw = xt * kDtTWX + warp * kDtPXW;
    const int npx = c_ok ? min(kDtPXW, p.Wo - xw) : 0;
    __half* o_ptr = p.out + ((((size_t)n * p.T + t) * p.Ho + yo0) * (size_t)p.Wo + xw) * p.C + c;
    const uint32_t lds_lane = smem_u32(s_ring) + (uint32_t)(((STRIDE == 1) ? warp * kDtPXW : 2 * warp * kDtPXW) * kDtCS + 2 * lane) * 2u;

    auto emit = [&](float2 (&acc)[kDtPXW]) {           // emits one output row and re-arms the accumulators with the bias
#pragma unroll
        for (int j = 0; j < kDtPXW; ++j) {           // all SiLU chains first, unconditionally: independent, so they interleave
            acc[j] = dt_silu2(acc[j]);               // (a branch per pixel serialised them: ~6 dependent SFU / FMA steps each)
        }
#pragma unroll
        for (int j = 0; j < kDtPXW; ++j) {
            if (j < npx) {
                lsum = __fadd2_rn(lsum, acc[j]);
                *reinterpret_cast<uint32_t*>(o_ptr + j * p.C) = pack_half2(acc[j].x, acc[j].y);
            }
            acc[j] = bias;
        }
        o_ptr += (size_t)p.Wo * p.C;
    };
    // stride 1: input row k (0-based inside the chunk) is kernel row 2 of out row k-2 (aOld, complete afterwards), kernel row 1 of
    // out row k-1 (aMid) and kernel row 0 of out row k (aNew)
    auto row_s1 = [&](uint32_t srow, int k, float2 (&aOld)[kDtPXW], float2 (&aMid)[kDtPXW], float2 (&aNew)[kDtPXW]) {
#pragma unroll
        for (int dt = 0; dt < KT; ++dt) {
            float2 v[Cfg::NV];
#pragma unroll
            for (int i = 0; i < Cfg::NV; ++i) v[i] = dt_lds_half2(srow + (uint32_t)(dt * Cfg::ROWB + i * kDtCS * 2));
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const float2 w0 = lds_w(dt * 9 + s), w1 = lds_w(dt * 9 + 3 + s), w2 = lds_w(dt * 9 + 6 + s);
#pragma unroll
                for (int j = 0; j < kDtPXW; ++j) {
                    aNew[j] = __ffma2_rn(w0, v[j + s], aNew[j]);
                    aMid[j] = __ffma2_rn(w1, v[j + s], aMid[j]);
                    aOld[j] = __ffma2_rn(w2, v[j + s], aOld[j]);
                }
            }
        }
        if (k >= 2) emit(aOld);
        else {
#pragma unroll
            for (int j = 0; j < kDtPXW; ++j) aOld[j] = bias;       // rows above the chunk: discard
        }
    };

    int k = 0;
    for (int js = 0; js < nstg; ++js) {
        if (warp == 0 && js + Cfg::NST - 1 < nstg) {
            if (elect_one()) issue(js + Cfg::NST - 1);
            __syncwarp();
        }
        const int st = js % Cfg::NST;
        mbar_wait(&full[st], (js / Cfg::NST) & 1);
        const uint32_t sbase = lds_lane + (uint32_t)st * Cfg::STAGEB;
        if constexpr (STRIDE == 1 && KT == 1) {
            // three rows per stage = one full rotation of the accumulator roles
            if (k < NR) row_s1(sbase, k, aX, aY, aZ);
            if (k + 1 < NR) row_s1(sbase + Cfg::ROWB, k + 1, aY, aZ, aX);
            if (k + 2 < NR) row_s1(sbase + 2 * Cfg::ROWB, k + 2, aZ, aX, aY);
            k += 3;
        } else if constexpr (STRIDE == 1) {
            // 3D: one row (three planes) per stage; the ring slot index is the rotation phase (NST == 3)
            if (st == 0) row_s1(sbase, k, aX, aY, aZ);
            else if (st == 1) row_s1(sbase, k, aY, aZ, aX);
            else row_s1(sbase, k, aZ, aX, aY);
            k += 1;
        } else {
            // stride 2: even input row 2yo is kernel row 0 of out yo and kernel row 2 of out yo-1; odd row 2yo+1 is kernel row 1
#pragma unroll
            for (int rr = 0; rr < Cfg::RPS; ++rr, ++k) {
                if (k >= NR) break;
                float2 v[Cfg::NV];
#pragma unroll
                for (int i = 0; i < Cfg::NV; ++i) v[i] = dt_lds_half2(sbase + (uint32_t)(rr * Cfg::ROWB + i * kDtCS * 2));
                if ((k & 1) == 0) {
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        const float2 w6 = lds_w(6 + s), w0 = lds_w(s);
#pragma unroll
                        for (int j = 0; j < kDtPXW; ++j) {
                            aX[j] = __ffma2_rn(w6, v[2 * j + s], aX[j]);     // closes out row k/2 - 1
                            aY[j] = __ffma2_rn(w0, v[2 * j + s], aY[j]);     // opens out row k/2
                        }
                    }
                    if (k > 0) emit(aX);
#pragma unroll
                    for (int j = 0; j < kDtPXW; ++j) { aX[j] = aY[j]; aY[j] = bias; }
                } else {
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        const float2 w3 = lds_w(3 + s);
#pragma unroll
                        for (int j = 0; j < kDtPXW; ++j) aX[j] = __ffma2_rn(w3, v[2 * j + s], aX[j]);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
    }

    // ---- SE squeeze: fixed-order sum of the 8 column strips, one plain store per channel per CTA ----
    __syncthreads();                  // every warp is done reading the ring (s_part aliases it)
    s_part[warp * kDtCS + 2 * lane] = lsum.x;
    s_part[warp * kDtCS + 2 * lane + 1] = lsum.y;
    __syncthreads();
    if (tid < kDtCS && c_slab + tid < p.C) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += s_part[i * kDtCS + tid];
        const int part = blockIdx.y * p.xtiles + xt;       // (t, row chunk, column tile)
        p.partials[((size_t)n * p.nparts + part) * p.C + c_slab + tid] = s;
    }
    if (p.se_w1 == nullptr) return;

    // ---- SE excitation by the last CTA of the image (multidim_stacker.py:86-90 / timm SqueezeExcite) ----
    __threadfence();                  // this thread's partial store is visible device-wide before the arrival below
    __syncthreads();
    if (tid == 0) {
        const int old = atomicAdd(&p.done[n], 1);
        *s_last = (old == p.ctas_per_img - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!*s_last) return;
    __threadfence();
    const float* part = p.partials + (size_t)n * p.nparts * p.C;
    for (int cc = tid; cc < p.C; cc += 256) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int q = 0;
        for (; q + 3 < p.nparts; q += 4) {
            s0 += dt_ld_cg(part + (size_t)q * p.C + cc);       s1 += dt_ld_cg(part + (size_t)(q + 1) * p.C + cc);
            s2 += dt_ld_cg(part + (size_t)(q + 2) * p.C + cc); s3 += dt_ld_cg(part + (size_t)(q + 3) * p.C + cc);
        }
        for (; q < p.nparts; ++q) s0 += dt_ld_cg(part + (size_t)q * p.C + cc);
        s_mean[cc] = ((s0 + s1) + (s2 + s3)) * p.inv_count;
    }
    __syncthreads();
    {
        const float4* m = reinterpret_cast<const float4*>(s_mean);
        const int c4n = p.C >> 2;
        for (int j0 = warp; j0 < p.rd; j0 += 8 * 3) {
            float acc[3];
            const float4* wrow[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                acc[u] = 0.f;
                const int j = j0 + u * 8;
                wrow[u] = reinterpret_cast<const float4*>(p.se_w1 + (size_t)(j < p.rd ? j : j0) * p.C);
            }
#pragma unroll 3
            for (int c4 = lane; c4 < c4n; c4 += 32) {
                const float4 mv = m[c4];
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const float4 wv = __ldg(wrow[u] + c4);
                    acc[u] = fmaf(wv.x, mv.x, fmaf(wv.y, mv.y, fmaf(wv.z, mv.z, fmaf(wv.w, mv.w, acc[u]))));
                }
            }
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int j = j0 + u * 8;
                const float a = warp_sum(acc[u]);
                if (lane == 0 && j < p.rd) s_hid[j] = silu_f(a + __ldg(p.se_b1 + j));
            }
        }
    }
    __syncthreads();
    for (int cc = tid; cc < p.C; cc += 256) {
        float a0s = 0.f, a1s = 0.f;
        int j = 0;
        for (; j + 1 < p.rd; j += 2) {
            a0s = fmaf(__ldg(p.se_w2t + (size_t)j * p.C + cc), s_hid[j], a0s);
            a1s = fmaf(__ldg(p.se_w2t + (size_t)(j + 1) * p.C + cc), s_hid[j + 1], a1s);
        }
        if (j < p.rd) a0s = fmaf(__ldg(p.se_w2t + (size_t)j * p.C + cc), s_hid[j], a0s);
        p.gate[(size_t)n * p.C + cc] = sigmoid_f(a0s + a1s + __ldg(p.se_b2 + cc));
    }
    if (tid == 0) p.done[n] = 0;      // clean for the next launch (kernel boundaries order it)
}

}  // namespace mds
