// K0+K1 on the 5th-gen tensor cores: frame pre-processing (frames.py:7-31: zero-pad H, /255) fused into the encoder stem
// (timm conv_stem 3x3 s2 TF-SAME + bn1 + SiLU; multidim_stacker.py:214 treats `stack_size` = 3 frames as the input channels).
// uint8 planar frames -> NHWC fp16 [n][H/2][W/2][32], one persistent warp-specialised kernel:
//   warp 0      : TMA producer: [3 planes][17 rows][144 bytes] uint8 input tiles (4-D tensor map over the stored frames; rows
//                 above / below the stored 720 and columns beyond W are hardware zero fill = the H padding and TF-SAME)
//   warps 2-5   : im2col builders: a thread converts the 27 taps of two neighbouring output pixels to fp16 (PRMT against the
//                 0x6400 exponent, exact) and writes the K-major A operand [6 k-planes][128 pixels][16 B]; the K order is
//                 k = (ci*3 + r)*4 + s with a zero-weight fourth slot, so every (ci, r) is one aligned 8-byte group
//   warp 1      : MMA issuer: per 128-pixel M tile 3 K steps x (weights hi, weights lo) tcgen05.mma, N = 32, fp32 in TMEM
//                 (uint8 pixels are exact in fp16 and the folded weights are an fp16 hi/lo pair: fp32-level accuracy)
//   warps 6-9   : epilogue (two CTAs per SM): TMEM -> * 1/255 + bias -> SiLU -> fp16 -> two 32-byte stores
// The mma.sync stem gathered its A fragments from shared memory with ~890 thread instructions per output pixel; here the
// builders spend ~60 and the epilogue ~230.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "conv_tc.cuh"
#include "mbconv_tail.cuh"     // tma_load_4d

namespace mds {

struct StemTcParams {
    int n, H, W, Ho, Wo;
    int pad_top;           // logical row y maps to stored row y - pad_top (frames.py:19)
    int hflip;             // TTA: read columns mirrored (predictors.py:63)
    float scale;           // 1/255 (frames.py:8), applied to the fp32 accumulator
    const __half* wh;      // [2][32][32]  (hi, lo) x cout x k,  k = (ci*3 + r)*3 + s (27..31 zero), BN scale folded
    const float* bias;     // [32]
    __half* out;           // [n][Ho][Wo][32]
    int tiles_x, tiles_y;
};

constexpr int kStcTW = 64, kStcTH = 8;                       // output tile: 4 M tiles of 2 rows x 64 pixels
constexpr int kStcIW = 144, kStcIH = 2 * kStcTH + 1;         // input tile row: 129 bytes needed, 144 = multiple of 16 for TMA
constexpr int kStcInBytes = 3 * kStcIH * kStcIW;             // 7344
constexpr int kStcInAlloc = ((kStcInBytes + 127) / 128) * 128;
constexpr int kStcKP = 6;                                    // 8-element K planes (K = 48: 36 used)
constexpr int kStcABytes = kStcKP * 128 * 16;                // one M tile of A: 12 KB
constexpr int kStcSlots = 4;                                 // A ring: the four M tiles of one input tile
constexpr int kStcWBytes = kStcKP * 32 * 16;                 // one weight tile (hi or lo): 3 KB
constexpr int kStcND = 4;                                    // accumulators (32 columns each)
constexpr int kStcBuilders = 4, kStcEpi = 4;                 // 320 threads, two CTAs per SM: one CTA's builders / epilogue fill the
                                                             // other's waits, and the small CTAs interleave with neighbouring kernels
constexpr int kStcThreads = 32 * (2 + kStcBuilders + kStcEpi);
constexpr size_t kStcSmem = 128 + 2 * kStcInAlloc + (size_t)kStcSlots * kStcABytes + 2 * kStcWBytes + 32 * 4 + 256;

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
// two uint8 (bytes picked by `sel` from w) -> half2, exact: 0x6400 | b is the fp16 1024 + b
__device__ __forceinline__ uint32_t u8x2_to_half2(uint32_t w, uint32_t sel) {
    const uint32_t k = 0x64006400u;
    uint32_t x = prmt(w, 0x64646464u, sel);
    __half2 r = __hsub2(*reinterpret_cast<__half2*>(&x), *reinterpret_cast<const __half2*>(&k));
    return *reinterpret_cast<uint32_t*>(&r);
}

__global__ void __launch_bounds__(kStcThreads, 2) stem_tc_kernel(const __grid_constant__ CUtensorMap tmIn, StemTcParams p) {
    extern __shared__ unsigned char stc_smem_raw[];
    pdl_trigger();
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(stc_smem_raw) + 127) & ~uintptr_t(127));
    unsigned char* s_in = smem;                                      // [2][3][17][144] uint8
    unsigned char* s_a = s_in + 2 * kStcInAlloc;                     // [8 slots][6 planes][128 px][16 B]
    unsigned char* s_w = s_a + kStcSlots * kStcABytes;               // [hi, lo][6 planes][32 n][16 B]
    float* s_bias = reinterpret_cast<float*>(s_w + 2 * kStcWBytes);  // [32]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 32);
    uint64_t* in_full = bars;            // [2]
    uint64_t* in_empty = bars + 2;       // [2]
    uint64_t* a_full = bars + 4;         // [8]
    uint64_t* a_empty = bars + 12;       // [8]
    uint64_t* d_full = bars + 20;        // [4]
    uint64_t* d_empty = bars + 24;       // [4]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 28);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int ntiles = tiles_per_img * p.n;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], kStcBuilders); }
        for (int i = 0; i < kStcSlots; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < kStcND; ++i) { mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], kStcEpi); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmIn) : "memory");
    }
    // weights -> [part][k-plane][n][8 halves] with k' = (ci*3 + r)*4 + s; the fourth slot of every group and k' >= 36 are zero
    for (int i = tid; i < 2 * kStcKP * 32 * 8; i += kStcThreads) {
        const int e = i & 7, n = (i >> 3) & 31, pl = (i >> 8) % kStcKP, part = i / (kStcKP * 256);
        const int kq = pl * 8 + e, combo = kq >> 2, s = kq & 3;
        __half v = __float2half(0.f);
        if (combo < 9 && s < 3) v = p.wh[(part * 32 + n) * 32 + combo * 3 + s];
        reinterpret_cast<__half*>(s_w)[i] = v;
    }
    if (tid < 32) s_bias[tid] = p.bias[tid];
    // the sixth K plane of every A slot stays zero for the whole kernel (k' = 40..47)
    for (int i = tid; i < kStcSlots * 128; i += kStcThreads)
        reinterpret_cast<uint4*>(s_a + (size_t)(i >> 7) * kStcABytes + 5 * 2048)[i & 127] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    pdl_wait();       // the output buffer may still be read by the previous forward's last kernels

    auto tile_geom = [&](int tile, int& n, int& oy0, int& ox0, int& nm) {
        n = tile / tiles_per_img;
        const int rem = tile - n * tiles_per_img;
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        oy0 = ty * kStcTH; ox0 = tx * kStcTW;
        nm = (min(kStcTH, p.Ho - oy0) + 1) >> 1;                 // M tiles (2 output rows each) with valid pixels
    };

    if (warp == 0) {
        // ================= TMA producer =================
        int i = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
            int n, oy0, ox0, nm;
            tile_geom(tile, n, oy0, ox0, nm);
            const int buf = i & 1;
            mbar_wait(&in_empty[buf], (((uint32_t)i >> 1) & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&in_full[buf], (uint32_t)kStcInBytes);
                // mirrored: tile byte j = logical column 2*ox0 + 143 - j (the inner start coordinate stays a multiple of 16 bytes)
                const int xs = p.hflip ? p.W - kStcIW - 2 * ox0 : 2 * ox0;
                tma_load_4d(s_in + (size_t)buf * kStcInAlloc, &tmIn, &in_full[buf], xs, 2 * oy0 - p.pad_top, 0, n);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const uint32_t idesc = tc_idesc(128, 32);
        const uint32_t aa = smem_u32(s_a), wa = smem_u32(s_w);
        int i = 0, T = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
            int n, oy0, ox0, nm;
            tile_geom(tile, n, oy0, ox0, nm);
            for (int m = 0; m < nm; ++m, ++T) {
                const int slot = m, d = T % kStcND;
                mbar_wait(&a_full[slot], (uint32_t)i & 1);
                mbar_wait(&d_empty[d], (((uint32_t)T / kStcND) & 1) ^ 1);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int pi = 0; pi < 2; ++pi)
#pragma unroll
                        for (int kc = 0; kc < 3; ++kc) {
                            const int part = (MDS_NUMERICS_VARIANT & 1) ? 1 - pi : pi;
                            tc_mma_f16(tmem_base + d * 32, tc_desc_nosw(aa + slot * kStcABytes + kc * 2 * 2048, 2048),
                                       tc_desc_nosw(wa + part * kStcWBytes + kc * 2 * 512, 512), idesc, (pi | kc) != 0);
                        }
                    tc_commit(&a_empty[slot]);
                    tc_commit(&d_full[d]);
                }
                __syncwarp();
            }
            // slots of M tiles without valid pixels are still handed back to the builders (they wait for every slot of a tile)
            for (int m = nm; m < 4; ++m) {
                const int slot = m;
                mbar_wait(&a_full[slot], (uint32_t)i & 1);
                if (elect_one()) tc_commit(&a_empty[slot]);
                __syncwarp();
            }
        }
    } else if (warp < 2 + kStcBuilders) {
        // ================= im2col builders: warp bw converts M tile bw of the tile (output rows 2*bw, 2*bw + 1; two pixels per
        // thread and row) =================
        const int bw = warp - 2, slot = bw;
        int i = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i) {
            const int buf = i & 1;
            mbar_wait(&in_full[buf], ((uint32_t)i >> 1) & 1);
            mbar_wait(&a_empty[slot], ((uint32_t)i & 1) ^ 1);
#pragma unroll 1
            for (int rr = 0; rr < 2; ++rr) {
                const int trow = 2 * bw + rr;                    // output row inside the tile
                const unsigned char* src = s_in + (size_t)buf * kStcInAlloc + (2 * trow) * kStcIW + (p.hflip ? 136 - 4 * lane : 4 * lane);
                unsigned char* dst = s_a + (size_t)slot * kStcABytes + (rr * 64 + 2 * lane) * 16;      // pixel A (pixel B = + 16 bytes)
                uint32_t ca[4], cb[4];                           // the 16-byte chunk (two (ci, r) groups) being assembled for pixel A / B
#pragma unroll
                for (int combo = 0; combo < 9; ++combo) {
                    const int ci = combo / 3, r = combo - ci * 3;
                    const unsigned char* q = src + (ci * kStcIH + r) * kStcIW;
                    uint32_t w0 = *reinterpret_cast<const uint32_t*>(q), w1 = *reinterpret_cast<const uint32_t*>(q + 4);
                    if (p.hflip) {                               // w0 = bytes (x5, x4 at byte 3 ...), w1 = (x3, x2, x1, x0) -> logical order
                        const uint32_t lo = w0, hi = w1;
                        w0 = prmt(hi, hi, 0x0123u);              // (x0, x1, x2, x3)
                        w1 = prmt(lo, lo, 0x0023u);              // (x4, x5, .., ..)
                    }
                    const uint32_t h01 = u8x2_to_half2(w0, 0x4140u), h23 = u8x2_to_half2(w0, 0x4342u), h45 = u8x2_to_half2(w1, 0x4140u);
                    const int o = (combo & 1) * 2;
                    ca[o] = h01; ca[o + 1] = h23;                // pixel A: x0 x1 x2 (x3: zero weight)
                    cb[o] = h23; cb[o + 1] = h45;                // pixel B: x2 x3 x4 (x5: zero weight)
                    if ((combo & 1) || combo == 8) {
                        if (combo == 8) { ca[2] = ca[3] = cb[2] = cb[3] = 0u; }
                        unsigned char* d = dst + (combo >> 1) * 2048;
                        *reinterpret_cast<uint4*>(d) = make_uint4(ca[0], ca[1], ca[2], ca[3]);
                        *reinterpret_cast<uint4*>(d + 16) = make_uint4(cb[0], cb[1], cb[2], cb[3]);
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to tcgen05.mma
            __syncwarp();
            if (lane == 0) { mbar_arrive(&a_full[slot]); mbar_arrive(&in_empty[buf]); }
        }
    } else {
        // ================= epilogue: four warps, one per TMEM lane quadrant =================
        const int q = warp & 3;
        const int prow = q * 32 + lane;                          // pixel of the M tile: tile row 2m + (prow >> 6), column prow & 63
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        int T = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int n, oy0, ox0, nm;
            tile_geom(tile, n, oy0, ox0, nm);
            for (int m = 0; m < nm; ++m, ++T) {
                const int d = T % kStcND;
                const int oy = oy0 + 2 * m + (prow >> 6), ox = ox0 + (prow & 63);
                const bool ok = oy < p.Ho && ox < p.Wo;
                __half* optr = p.out + (((size_t)n * p.Ho + (ok ? oy : 0)) * p.Wo + (ok ? ox : 0)) * 32;
                mbar_wait(&d_full[d], ((uint32_t)T / kStcND) & 1);
                tc_fence_after();
                uint32_t v[2][16];
                tc_ld16(t_row + (uint32_t)(d * 32), v[0]);
                tc_ld16(t_row + (uint32_t)(d * 32 + 16), v[1]);
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[d]);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    uint32_t pk[8];
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const float4 b = *reinterpret_cast<const float4*>(s_bias + j * 16 + h * 4);
                        float x4[4] = {fmaf(__uint_as_float(v[j][4 * h]), p.scale, b.x), fmaf(__uint_as_float(v[j][4 * h + 1]), p.scale, b.y),
                                       fmaf(__uint_as_float(v[j][4 * h + 2]), p.scale, b.z), fmaf(__uint_as_float(v[j][4 * h + 3]), p.scale, b.w)};
                        silu4(x4);
                        pk[2 * h] = pack_half2(x4[0], x4[1]);
                        pk[2 * h + 1] = pack_half2(x4[2], x4[3]);
                    }
                    if (ok) st_global_v8(optr + j * 16, pk);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
    }
}

}  // namespace mds
