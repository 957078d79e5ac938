// K4/K6: depthwise 3x3 (2D, stride 1/2, TF-SAME) and 3x3x3 (3D, pad 1) convolution + folded BN + SiLU, NHWC fp16,
// with the SE squeeze (per-image channel sums over the output, fp32, as per-CTA partials) produced in the same pass
// (timm InvertedResidual conv_dw/bn2/se; multidim_stacker.py:110-114,86).
//
// HBM-bound streaming kernel.  A CTA owns (image, [out plane t], 40 output columns, 64-channel slab, row chunk) and
// streams the input rows it needs through a ring of shared-memory stages filled by cp.async (16 B per request,
// zero-fill for the padding halo, rows beyond the image and channel tails), several rows ahead of the math, so the
// bytes in flight per SM do not depend on registers or occupancy.  Warp w computes output columns [5w, 5w+5); lane l
// owns channels (2l, 2l+1) of the slab as one packed f32x2, so every shared-memory read is a conflict-free LDS.32,
// every global store is a 128-byte line per pixel, and the MACs are packed FFMA2 (fma.rn.f32x2).  Each input row is
// scattered into the (up to) three output rows it contributes to; the window never lives in registers.
#pragma once
#include "common.cuh"

namespace mds {

struct DwParams {
    const __half* in;    // [n][T][H][W][C]
    __half* out;         // [n][T][Ho][Wo][C]
    const float* w;      // [taps][C]  tap = (dt*3 + r)*3 + s, BN scale folded
    const float* bias;   // [C]   (LIN mode: the shift reference of the BatchNorm statistics, see below)
    float* partials;     // [n][nparts][C]  per-CTA sums over its part of (T,Ho,Wo) of the fp32 SiLU output (deterministic SE squeeze)
                         // LIN mode (training): [n * nparts][2][C] sums of (y - ref) and (y - ref)^2 of the stored fp16
                         // output, or nullptr
    int n, T, H, W, C, Ho, Wo;
    int rows_per_chunk, chunks, xtiles, slabs, nparts;
};

constexpr int kDwCS = 64;      // channels per slab (lane = 2 channels)
constexpr int kDwPXW = 5;      // output columns per warp
constexpr int kDwTWX = 40;     // output columns per CTA (8 warps)

template <int KT, int STRIDE>
struct DwCfg {
    static constexpr int IW = (STRIDE == 1) ? kDwTWX + 2 : 2 * kDwTWX + 1;   // input columns per row tile
    static constexpr int NV = (STRIDE == 1) ? kDwPXW + 2 : 2 * kDwPXW + 1;   // input columns per warp strip
    static constexpr int ROW_HALVES = IW * kDwCS;
    static constexpr int RPS = (KT == 3) ? 1 : 2;                            // input rows per stage (one barrier per stage)
    static constexpr int STEP_HALVES = KT * ROW_HALVES;                      // one input row (x KT planes)
    static constexpr int STAGE_HALVES = RPS * STEP_HALVES;
    static constexpr int NST = (KT == 3) ? 4 : (STRIDE == 2 ? 3 : 4);
    static constexpr int CHUNKS = KT * IW * 8;                               // 16-byte requests per input row
    static constexpr int SLOTS = (CHUNKS + 255) / 256;
    static constexpr size_t SMEM = (size_t)NST * STAGE_HALVES * 2 + 2 * 8 * kDwCS * sizeof(float);
};

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 lds_half2(uint32_t saddr) {
    uint32_t u;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(saddr));
    return unpack_half2(u);
}

// LIN = true (training, conv_dw alone): no bias / activation; the epilogue instead accumulates the BatchNorm batch
// statistics of the stored output (sums of y - ref and (y - ref)^2, ref = p.bias) as per-CTA partials.
template <int KT, int STRIDE, bool LIN = false>
__global__ void __launch_bounds__(256) dwconv_kernel(DwParams p) {
    using Cfg = DwCfg<KT, STRIDE>;
    static_assert(KT == 1 || STRIDE == 1, "3D depthwise is stride 1");
    extern __shared__ __align__(16) unsigned char dw_smem[];
    __half* s_ring = reinterpret_cast<__half*>(dw_smem);
    float* s_part = reinterpret_cast<float*>(dw_smem + (size_t)Cfg::NST * Cfg::STAGE_HALVES * 2);   // [8][64]

    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int xt = blockIdx.x % p.xtiles, slab = blockIdx.x / p.xtiles;
    const int t = blockIdx.y / p.chunks, chunk = blockIdx.y - t * p.chunks;
    const int n = blockIdx.z;
    const int yo0 = chunk * p.rows_per_chunk, yo1 = min(p.Ho, yo0 + p.rows_per_chunk);
    const int xo0 = xt * kDwTWX;
    const int c_slab = slab * kDwCS;
    const int c = c_slab + 2 * lane;
    const bool c_ok = c < p.C;
    const int c_ld = c_ok ? c : 0;

    // input rows consumed by this chunk and the first input column of the tile (TF-SAME: pad 1 / pad right-bottom only)
    const int yi0 = (STRIDE == 1) ? yo0 - 1 : 2 * yo0;
    const int NR = (STRIDE == 1) ? (yo1 - yo0) + 2 : 2 * (yo1 - yo0) + 1;
    const int xi0 = (STRIDE == 1) ? xo0 - 1 : 2 * xo0;

    const size_t plane = (size_t)p.H * p.W * p.C;
    const __half* in_n = p.in + (size_t)n * p.T * plane;

    // cp.async bookkeeping that does not depend on the row is computed once per thread; the row-dependent part is a
    // running pointer, so the steady-state loop has no address arithmetic beyond pointer increments
    uint32_t s_off[Cfg::SLOTS];       // bytes, inside one row step (smem)
    const __half* g_ptr[Cfg::SLOTS];  // source of this thread's request for the NEXT row to be issued
    bool s_ok[Cfg::SLOTS];
    const long long row_pitch = (long long)p.W * p.C;
#pragma unroll
    for (int sl = 0; sl < Cfg::SLOTS; ++sl) {
        const int idx = tid + sl * 256;
        const int pl = idx / (Cfg::IW * 8);
        const int rem = idx - pl * (Cfg::IW * 8);
        const int px = rem >> 3, c16 = rem & 7;
        const int xi = xi0 + px;
        const int ti = (KT == 3) ? t + pl - 1 : t;
        const int cc = c_slab + c16 * 8;
        s_ok[sl] = (idx < Cfg::CHUNKS) && (xi >= 0) && (xi < p.W) && (ti >= 0) && (ti < p.T) && (cc < p.C);
        s_off[sl] = (uint32_t)(pl * Cfg::ROW_HALVES + px * kDwCS + c16 * 8) * 2u;
        g_ptr[sl] = s_ok[sl] ? in_n + (long long)ti * (long long)plane + (long long)yi0 * row_pitch + (long long)xi * p.C + cc : p.in;
        if (idx >= Cfg::CHUNKS) s_off[sl] = 0xffffffffu;
    }
    const uint32_t ring_u32 = smem_u32(s_ring);
    uint32_t iss_off = 0;             // byte offset of the stage the producer side fills next
    int iss_y = yi0;                  // input row the producer side requests next

    auto issue = [&]() {
#pragma unroll
        for (int rr = 0; rr < Cfg::RPS; ++rr) {
            const bool yok = (unsigned)iss_y < (unsigned)p.H;
#pragma unroll
            for (int sl = 0; sl < Cfg::SLOTS; ++sl) {
                if (s_off[sl] != 0xffffffffu) {
                    const bool ok = yok && s_ok[sl];
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(ring_u32 + iss_off + rr * (Cfg::STEP_HALVES * 2) + s_off[sl]),
                                 "l"(ok ? g_ptr[sl] : p.in), "r"(ok ? 16 : 0));
                    if (s_ok[sl]) g_ptr[sl] += row_pitch;
                }
            }
            ++iss_y;
        }
        iss_off += Cfg::STAGE_HALVES * 2;
        if (iss_off == Cfg::NST * Cfg::STAGE_HALVES * 2) iss_off = 0;
    };

    // per-thread weights: KT*9 taps x 2 channels
    float2 w[KT * 9];
#pragma unroll
    for (int i = 0; i < KT * 9; ++i) w[i] = __ldg(reinterpret_cast<const float2*>(p.w + (size_t)i * p.C + c_ld));
    const float2 bias = __ldg(reinterpret_cast<const float2*>(p.bias + c_ld));

    pdl_wait();       // weights above are constants; the input rows below come from the previous kernel
    const int NSTG = (NR + Cfg::RPS - 1) / Cfg::RPS;     // stages to stream (rows past NR are loaded but unused)
#pragma unroll
    for (int j = 0; j < Cfg::NST - 1; ++j) {
        if (j < NSTG) issue();
        cp_async_commit();
    }

    float2 lsum = make_float2(0.f, 0.f), lsq = make_float2(0.f, 0.f);
    float2 a0[kDwPXW], a1[kDwPXW], a2[kDwPXW];
#pragma unroll
    for (int j = 0; j < kDwPXW; ++j) a0[j] = a1[j] = a2[j] = make_float2(0.f, 0.f);

    const int xw = xo0 + warp * kDwPXW;               // first output column of this warp
    const int px_base = (STRIDE == 1) ? warp * kDwPXW : 2 * warp * kDwPXW;
    const long long out_pitch = (long long)p.Wo * p.C;
    // running output pointer: row yo0 of this warp's strip; every emitted row advances it by one output row
    __half* o_ptr = p.out + ((size_t)n * p.T + t) * (size_t)p.Ho * out_pitch + (long long)yo0 * out_pitch + (long long)xw * p.C + c;
    bool px_ok[kDwPXW];
#pragma unroll
    for (int j = 0; j < kDwPXW; ++j) px_ok[j] = c_ok && (xw + j < p.Wo);
    const int pxC = p.C;                             // halves between adjacent output columns

    auto emit = [&](const float2 (&acc)[kDwPXW]) {   // rows are emitted in order yo0, yo0+1, ...
#pragma unroll
        for (int j = 0; j < kDwPXW; ++j) {
            if (px_ok[j]) {
                if constexpr (LIN) {
                    const uint32_t packed = pack_half2(acc[j].x, acc[j].y);
                    const float2 r = unpack_half2(packed);
                    const float dx = r.x - bias.x, dy = r.y - bias.y;
                    lsum.x += dx; lsum.y += dy;
                    lsq.x = fmaf(dx, dx, lsq.x); lsq.y = fmaf(dy, dy, lsq.y);
                    *reinterpret_cast<uint32_t*>(o_ptr + j * pxC) = packed;
                } else {
                    const float ox = silu_f(acc[j].x + bias.x), oy = silu_f(acc[j].y + bias.y);
                    lsum.x += ox; lsum.y += oy;
                    *reinterpret_cast<uint32_t*>(o_ptr + j * pxC) = pack_half2(ox, oy);
                }
            }
        }
        o_ptr += out_pitch;
    };

    uint32_t cons_off = 0;            // byte offset of the stage being consumed
    const uint32_t lds_base = ring_u32 + (uint32_t)(px_base * kDwCS + 2 * lane) * 2u;
    for (int js = 0; js < NSTG; ++js) {
        cp_async_wait<Cfg::NST - 2>();
        __syncthreads();
        if (js + Cfg::NST - 1 < NSTG) issue();
        cp_async_commit();
#pragma unroll
        for (int rr = 0; rr < Cfg::RPS; ++rr) {
            const int k = js * Cfg::RPS + rr;
            if (k >= NR) break;
            const uint32_t st = lds_base + cons_off + rr * (Cfg::STEP_HALVES * 2);
            const int yi = yi0 + k;

            if constexpr (STRIDE == 1) {
                // input row yi feeds out rows yi+1 (kernel row 0), yi (row 1), yi-1 (row 2)
#pragma unroll
                for (int j = 0; j < kDwPXW; ++j) a2[j] = make_float2(0.f, 0.f);
#pragma unroll
                for (int dt = 0; dt < KT; ++dt) {
                    float2 v[Cfg::NV];
#pragma unroll
                    for (int i = 0; i < Cfg::NV; ++i)
                        v[i] = lds_half2(st + (dt * Cfg::ROW_HALVES + i * kDwCS) * 2);
#pragma unroll
                    for (int j = 0; j < kDwPXW; ++j)
#pragma unroll
                        for (int s = 0; s < 3; ++s) {
                            a2[j] = ffma2(w[dt * 9 + 0 + s], v[j + s], a2[j]);
                            a1[j] = ffma2(w[dt * 9 + 3 + s], v[j + s], a1[j]);
                            a0[j] = ffma2(w[dt * 9 + 6 + s], v[j + s], a0[j]);
                        }
                }
                if (yi - 1 >= yo0) emit(a0);        // row yi - 1 (< yo1 always holds since yi <= yo1)
#pragma unroll
                for (int j = 0; j < kDwPXW; ++j) { a0[j] = a1[j]; a1[j] = a2[j]; }
            } else {
                // stride 2: even input row 2yo is kernel row 0 of out yo and kernel row 2 of out yo-1; odd row 2yo+1 is row 1
                float2 v[Cfg::NV];
#pragma unroll
                for (int i = 0; i < Cfg::NV; ++i) v[i] = lds_half2(st + i * kDwCS * 2);
                if ((k & 1) == 0) {       // yi = 2*yo0 + k is even
#pragma unroll
                    for (int j = 0; j < kDwPXW; ++j) {
                        a2[j] = make_float2(0.f, 0.f);
#pragma unroll
                        for (int s = 0; s < 3; ++s) {
                            a1[j] = ffma2(w[6 + s], v[2 * j + s], a1[j]);     // closes out row (yi/2 - 1)
                            a2[j] = ffma2(w[0 + s], v[2 * j + s], a2[j]);     // opens out row yi/2
                        }
                    }
                    if (k > 0) emit(a1);           // row yi/2 - 1
#pragma unroll
                    for (int j = 0; j < kDwPXW; ++j) a1[j] = a2[j];
                } else {
#pragma unroll
                    for (int j = 0; j < kDwPXW; ++j)
#pragma unroll
                        for (int s = 0; s < 3; ++s) a1[j] = ffma2(w[3 + s], v[2 * j + s], a1[j]);
                }
            }
        }
        cons_off += Cfg::STAGE_HALVES * 2;
        if (cons_off == Cfg::NST * Cfg::STAGE_HALVES * 2) cons_off = 0;
    }
    cp_async_wait<0>();

    // ---- SE squeeze: reduce the 8 warps' partial sums in a fixed order; one plain store per channel per CTA ----
    if (LIN && p.partials == nullptr) return;
    s_part[warp * kDwCS + 2 * lane] = lsum.x;
    s_part[warp * kDwCS + 2 * lane + 1] = lsum.y;
    if constexpr (LIN) {
        s_part[8 * kDwCS + warp * kDwCS + 2 * lane] = lsq.x;
        s_part[8 * kDwCS + warp * kDwCS + 2 * lane + 1] = lsq.y;
    }
    __syncthreads();
    const int part = blockIdx.y * p.xtiles + xt;       // (t, row chunk, column tile)
    if (tid < kDwCS && c_slab + tid < p.C) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += s_part[i * kDwCS + tid];
        if constexpr (LIN) p.partials[(((size_t)n * p.nparts + part) * 2) * p.C + c_slab + tid] = s;
        else p.partials[((size_t)n * p.nparts + part) * p.C + c_slab + tid] = s;
    }
    if constexpr (LIN) {
        if (tid >= kDwCS && tid < 2 * kDwCS && c_slab + tid - kDwCS < p.C) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s += s_part[8 * kDwCS + i * kDwCS + tid - kDwCS];
            p.partials[(((size_t)n * p.nparts + part) * 2 + 1) * p.C + c_slab + tid - kDwCS] = s;
        }
    }
}

}  // namespace mds
