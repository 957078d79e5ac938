// K3/K5: 1x1 (pointwise) convolution as a GEMM on tensor cores.  NHWC fp16 activations are a row-major
// [M = pixels][K = cin] matrix, PyTorch's conv weight [cout][cin] is the K-major B operand as is.
//   C[m][n] = act( sum_k (A[m][k] * gate[img(m)][k]) * W[n][k] + bias[n] ) (+ res[m][n])
// gate = SE excitation (timm SqueezeExcite / multidim_stacker.py:86-90) applied to the A fragments in
// registers, so the gated tensor is never materialised.  M is tiled per image when gated.
#pragma once
#include "common.cuh"

namespace mds {

struct GemmParams {
    const __half* A;      // [n_img * rows_per_img][K]
    const __half* W;      // [N][K], BN scale folded
    const float* bias;    // [N]
    const __half* res;    // [M][N] or nullptr
    const __half* gate;   // [n_img][K] or nullptr
    __half* C;            // [M][N]
    int rows_per_img, n_img, N, K;
    int act;              // 1 = SiLU
};

constexpr int kGemmBM = 128, kGemmBK = 32, kGemmStages = 3, kGemmPitch = kGemmBK + 8;

template <int BN>
struct GemmCfg {
    static constexpr int A_HALVES = kGemmBM * kGemmPitch;
    static constexpr int B_HALVES = BN * kGemmPitch;
    static constexpr int STAGE = A_HALVES + B_HALVES;
    static constexpr int CP = BN + 8;   // epilogue staging pitch
    static constexpr size_t PIPE_BYTES = (size_t)kGemmStages * STAGE * 2;
    static constexpr size_t EPI_BYTES = (size_t)kGemmBM * CP * 2;
    static constexpr size_t GATE_BYTES = 1152 * 2;
    static constexpr size_t SMEM = (PIPE_BYTES > EPI_BYTES ? PIPE_BYTES : EPI_BYTES) + GATE_BYTES;
};

template <int BN, bool GATED>
__global__ void __launch_bounds__(256, 2) gemm1x1_kernel(GemmParams p) {
    using Cfg = GemmCfg<BN>;
    constexpr int WN = BN / 2;       // warp tile N
    constexpr int NT = WN / 8;       // n8 tiles per warp
    static_assert(WN % 8 == 0, "BN must be a multiple of 16");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* s_pipe = reinterpret_cast<__half*>(smem_raw);
    __half* s_gate = reinterpret_cast<__half*>(smem_raw + (Cfg::PIPE_BYTES > Cfg::EPI_BYTES ? Cfg::PIPE_BYTES : Cfg::EPI_BYTES));

    pdl_trigger();
    pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;
    const int tiles_per_img = (p.rows_per_img + kGemmBM - 1) / kGemmBM;
    const int img = blockIdx.y / tiles_per_img;   // N tiles vary fastest so CTAs sharing an A tile run together (L2 reuse)
    const int tile = blockIdx.y - img * tiles_per_img;
    const int rows_valid = min(kGemmBM, p.rows_per_img - tile * kGemmBM);
    const size_t row0 = (size_t)img * p.rows_per_img + (size_t)tile * kGemmBM;
    const int n0 = blockIdx.x * BN;
    const int K = p.K;
    const int KT = (K + kGemmBK - 1) / kGemmBK;

    if constexpr (GATED) {
        for (int i = tid; i < K / 8; i += 256)
            reinterpret_cast<uint4*>(s_gate)[i] = __ldg(reinterpret_cast<const uint4*>(p.gate + (size_t)img * K) + i);
    }

    auto load_stage = [&](int kt, int st) {
        __half* sa = s_pipe + st * Cfg::STAGE;
        __half* sb = sa + Cfg::A_HALVES;
        const int k0 = kt * kGemmBK;
        for (int i = tid; i < kGemmBM * 4; i += 256) {
            int r = i >> 2, c = i & 3;
            int k = k0 + c * 8;
            bool ok = (r < rows_valid) && (k < K);
            const __half* src = ok ? p.A + (row0 + r) * K + k : p.A;
            cp_async16(sa + r * kGemmPitch + c * 8, src, ok ? 16 : 0);
        }
        for (int i = tid; i < BN * 4; i += 256) {
            int r = i >> 2, c = i & 3;
            int k = k0 + c * 8;
            bool ok = (n0 + r < p.N) && (k < K);
            const __half* src = ok ? p.W + (size_t)(n0 + r) * K + k : p.W;
            cp_async16(sb + r * kGemmPitch + c * 8, src, ok ? 16 : 0);
        }
    };

#pragma unroll
    for (int s = 0; s < kGemmStages - 1; ++s) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }

    float acc[2][NT][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

    const int lm = lane >> 3, lr = lane & 7;
    const int a_row = wm * 32 + lr + (lm & 1) * 8, a_kof = (lm >> 1) * 8;
    const int b_row = wn * WN + (lm >> 1) * 8 + lr, b_kof = (lm & 1) * 8;
    const int g = lane >> 2, tq = lane & 3;

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<kGemmStages - 2>();
        __syncthreads();
        {   // prefetch tile kt + S - 1 into the stage consumed in iteration kt - 1
            int nk = kt + kGemmStages - 1;
            if (nk < KT) load_stage(nk, nk % kGemmStages);
            cp_async_commit();
        }
        const __half* sa = s_pipe + (kt % kGemmStages) * Cfg::STAGE;
        const __half* sb = sa + Cfg::A_HALVES;
        const uint32_t a_addr = smem_u32(sa + a_row * kGemmPitch + a_kof);
        const uint32_t b_addr = smem_u32(sb + b_row * kGemmPitch + b_kof);
#pragma unroll
        for (int ks = 0; ks < kGemmBK / 16; ++ks) {
            uint32_t a[2][4];
            ldmatrix_x4(a[0], a_addr + ks * 32);
            ldmatrix_x4(a[1], a_addr + 16 * kGemmPitch * 2 + ks * 32);
            if constexpr (GATED) {
                const int kk = kt * kGemmBK + ks * 16 + tq * 2;   // < K rounded up to 32; s_gate is zero-safe below
                uint32_t g0 = *reinterpret_cast<const uint32_t*>(s_gate + min(kk, K - 2));
                uint32_t g1 = *reinterpret_cast<const uint32_t*>(s_gate + min(kk + 8, K - 2));
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    a[i][0] = hmul2_u32(a[i][0], g0); a[i][1] = hmul2_u32(a[i][1], g0);
                    a[i][2] = hmul2_u32(a[i][2], g1); a[i][3] = hmul2_u32(a[i][3], g1);
                }
            }
#pragma unroll
            for (int j = 0; j < NT / 2; ++j) {
                uint32_t b[4];
                ldmatrix_x4(b, b_addr + j * 16 * kGemmPitch * 2 + ks * 32);
                mma16816(acc[0][2 * j], a[0], b[0], b[1]);
                mma16816(acc[0][2 * j + 1], a[0], b[2], b[3]);
                mma16816(acc[1][2 * j], a[1], b[0], b[1]);
                mma16816(acc[1][2 * j + 1], a[1], b[2], b[3]);
            }
            if constexpr (NT & 1) {
                uint32_t b[2];
                // x2: lanes 0-15 supply addresses (rows 0-7 at k, rows 0-7 at k+8)
                const int l = lane & 15;
                const uint32_t addr = smem_u32(sb + (wn * WN + (NT - 1) * 8 + (l & 7)) * kGemmPitch + (l >> 3) * 8) + ks * 32;
                ldmatrix_x2(b, addr);
                mma16816(acc[0][NT - 1], a[0], b[0], b[1]);
                mma16816(acc[1][NT - 1], a[1], b[0], b[1]);
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();

    // ---- epilogue: bias / SiLU / residual in fp32, one rounding to fp16, stage through smem for 16 B stores ----
    __half* s_c = s_pipe;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int r_lo = wm * 32 + i * 16 + g, r_hi = r_lo + 8;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int cl = wn * WN + j * 8 + tq * 2;
            const int cgl = n0 + cl;
            float bx = 0.f, by = 0.f;
            if (cgl < p.N) { bx = __ldg(p.bias + cgl); by = __ldg(p.bias + cgl + 1); }
            float v0 = acc[i][j][0] + bx, v1 = acc[i][j][1] + by, v2 = acc[i][j][2] + bx, v3 = acc[i][j][3] + by;
            if (p.act) { v0 = silu_f(v0); v1 = silu_f(v1); v2 = silu_f(v2); v3 = silu_f(v3); }
            if (p.res != nullptr && cgl < p.N) {
                if (r_lo < rows_valid) {
                    float2 f = __half22float2(*reinterpret_cast<const __half2*>(p.res + (row0 + r_lo) * p.N + cgl));
                    v0 += f.x; v1 += f.y;
                }
                if (r_hi < rows_valid) {
                    float2 f = __half22float2(*reinterpret_cast<const __half2*>(p.res + (row0 + r_hi) * p.N + cgl));
                    v2 += f.x; v3 += f.y;
                }
            }
            *reinterpret_cast<uint32_t*>(s_c + r_lo * Cfg::CP + cl) = pack_half2(v0, v1);
            *reinterpret_cast<uint32_t*>(s_c + r_hi * Cfg::CP + cl) = pack_half2(v2, v3);
        }
    }
    __syncthreads();
    constexpr int CPR = BN / 8;   // 16 B chunks per row
    for (int i = tid; i < kGemmBM * CPR; i += 256) {
        int r = i / CPR, c = i - r * CPR;
        if (r < rows_valid && n0 + c * 8 < p.N)
            *reinterpret_cast<uint4*>(p.C + (row0 + r) * p.N + n0 + c * 8) =
                *reinterpret_cast<const uint4*>(s_c + r * Cfg::CP + c * 8);
    }
}

}  // namespace mds
