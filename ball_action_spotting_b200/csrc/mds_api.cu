// C-ABI of libmds_b200.so: packed-weight handle, the MultiDimStacker layer program, and per-kernel entry points.
// See include/mds_b200.h for the contract of every exported function.
#include "../../include/mds_b200.h"

#include <cudaTypedefs.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "tc_helpers.cuh"
#include "conv_tc.cuh"
#include "conv_tc_ws.cuh"
#include "dwconv.cuh"
#include "dwconv_tma.cuh"
#include "gemm1x1.cuh"
#include "gemm_tc.cuh"
#include "mbconv_tail.cuh"
#include "se_head.cuh"
#include "stem.cuh"
#include "stem_tc.cuh"
#include "train.cuh"
#include "postproc.cuh"

using namespace mds;

// --------------------------------------------------------------------------------------------------------------
// error plumbing
// --------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static thread_local long long g_launches = 0;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CUDA_TRY(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t e_ = (expr);                                                                            \
        if (e_ != cudaSuccess) return fail(MDS_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define LAUNCH_CHECK(name)                                                                                  \
    do {                                                                                                    \
        ++g_launches;                                                                                       \
        cudaError_t e_ = cudaGetLastError();                                                                \
        if (e_ != cudaSuccess) return fail(MDS_ERR_CUDA, "launch %s: %s", name, cudaGetErrorString(e_));    \
    } while (0)
#define TRY(expr)                \
    do {                         \
        int rc_ = (expr);        \
        if (rc_ != 0) return rc_; \
    } while (0)

extern "C" const char* mds_last_error(void) { return g_err.c_str(); }
extern "C" long long mds_launch_count(int reset) {
    long long v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

// ---- optional per-launch CUDA-event profiling (bench.py roofline; off by default) ----
struct ProfRec { int kind, tag; cudaEvent_t a, b; };
static thread_local bool g_prof_on = false;
static thread_local int g_prof_tag = -1;
static thread_local std::vector<ProfRec> g_prof;
struct ProfScope {
    cudaStream_t st; bool on;
    ProfScope(int kind, cudaStream_t s) : st(s), on(g_prof_on) {
        if (!on) return;
        ProfRec r; r.kind = kind; r.tag = g_prof_tag;
        cudaEventCreate(&r.a); cudaEventCreate(&r.b);
        cudaEventRecord(r.a, st);
        g_prof.push_back(r);
    }
    ~ProfScope() { if (on) cudaEventRecord(g_prof.back().b, st); }
};
extern "C" int mds_profile_begin(void) {
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_on = true;
    return MDS_OK;
}
extern "C" int mds_profile_end(int* kinds, int* tags, float* ms, int capacity, int* count) {
    g_prof_on = false;
    int n = 0;
    for (auto& r : g_prof) {
        cudaError_t e = cudaEventSynchronize(r.b);
        float t = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&t, r.a, r.b);
        if (e != cudaSuccess) return fail(MDS_ERR_CUDA, "profile: %s", cudaGetErrorString(e));
        if (n < capacity && kinds && tags && ms) { kinds[n] = r.kind; tags[n] = r.tag; ms[n] = t; }
        ++n;
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    g_prof.clear();
    if (count) *count = n;
    return MDS_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute of a kernel: set it once per (kernel, device)
constexpr int kMaxDevices = 64;
#define ENSURE_SMEM_ATTR(kern, bytes)                                                                           \
    do {                                                                                                        \
        static bool done_[kMaxDevices] = {};                                                                    \
        int dev_ = 0;                                                                                           \
        cudaGetDevice(&dev_);                                                                                   \
        if (dev_ >= 0 && dev_ < kMaxDevices && !done_[dev_]) {                                                  \
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));    \
            done_[dev_] = true;                                                                                 \
        }                                                                                                       \
    } while (0)

// Launch with programmatic stream serialization (see pdl_trigger / pdl_wait in common.cuh)
static bool g_pdl = !(getenv("MDS_PDL") && getenv("MDS_PDL")[0] == '0');     // MDS_PDL=0: plain stream serialization
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
extern "C" int mds_set_pdl(int enabled) { g_pdl = enabled != 0; return MDS_OK; }
// MBConv tails (depthwise + SE + projection), selectable for A/B measurements (profiles/experiments_r02.md):
//   3 (default): dwconv_tma_kernel (TMA-staged depthwise + squeeze partials), se_fc_kernel (excitation + per-image gated
//                projection weights), tcgen05 GEMM on the pre-gated weights                               -> 3 launches
//   2: dwconv_tma_kernel incl. the SE MLP (last CTA of an image), tcgen05 GEMM gating its A operand in smem -> 2 launches
//   1: mbconv_tail_kernel, everything in ONE persistent launch (depthwise and projection streams on disjoint warps)
//   0: round-1 path: cp.async depthwise kernel (batch-dependent row chunks), se_fc_kernel, GEMM            -> 3 launches
static int g_tail_mode = getenv("MDS_TAIL_MODE") ? atoi(getenv("MDS_TAIL_MODE")) : 3;
extern "C" int mds_set_tail_mode(int mode) {
    if (mode < 0 || mode > 3) return fail(MDS_ERR_INVALID, "tail mode must be 0..3");
    g_tail_mode = mode;
    return MDS_OK;
}
// Dense 3x3 blocks (blocks.0.0 - 2.0), selectable for A/B measurements (both parity-tested):
//   2 (default): conv_tc_kernel keeps the expanded tensor in tensor memory (A operand of the projection from TMEM); blocks.0.0
//                with the column taps folded into N
//   1: conv_tc_kernel stages the expanded tensor in shared memory; blocks.0.0 without the fold
// blocks.2.1 (conv_tc_ws_kernel) and the stem (stem_tc_kernel) have one implementation.  The round-1 mma.sync kernels are gone.
static int g_conv_mode = getenv("MDS_CONV_MODE") ? atoi(getenv("MDS_CONV_MODE")) : 2;
extern "C" int mds_set_conv_mode(int mode) {
    if (mode < 1 || mode > 2) return fail(MDS_ERR_INVALID, "conv mode must be 1 or 2");
    g_conv_mode = mode;
    return MDS_OK;
}
// Encoder (mds_forward, mds_forward_2d) as S equal parts of the images on S streams (default 2), see forward_2d_streams
constexpr int kMaxStreams = 4;
static int g_streams = getenv("MDS_STREAMS") ? atoi(getenv("MDS_STREAMS")) : 2;
extern "C" int mds_set_streams(int n) {
    if (n < 1 || n > kMaxStreams) return fail(MDS_ERR_INVALID, "streams must be in 1..%d", kMaxStreams);
    g_streams = n;
    return MDS_OK;
}
// measurement only (bench.py roofline_dw): 1 = the fused tails run their depthwise + SE items alone (outputs are NOT valid)
static bool g_tail_dw_only = false;
extern "C" int mds_set_tail_dw_only(int enabled) { g_tail_dw_only = enabled != 0; return MDS_OK; }

static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

// --------------------------------------------------------------------------------------------------------------
// kernel launchers
// --------------------------------------------------------------------------------------------------------------
// ---- tcgen05 path: TMA tensor maps are encoded on the host through the driver entry point (no libcuda link) ----
static PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}

static int launch_stem(const MdsFrames& f, int n, const __half* wh, const float* bias, __half* out, cudaStream_t st) {
    if (f.H % 2 || f.W % 4 || f.H <= 0 || f.W <= 0) return fail(MDS_ERR_INVALID, "stem: H must be even, W a multiple of 4");
    if (f.img_stride % 4 || f.plane_stride % 4) return fail(MDS_ERR_INVALID, "stem: image / plane strides must be multiples of 4 elements");
    if (n <= 0) return MDS_OK;
    StemParams p;
    p.in = f.data; p.img_stride = f.img_stride; p.plane_stride = f.plane_stride;
    p.stored_h = f.stored_h; p.pad_top = f.pad_top; p.H = f.H; p.W = f.W; p.hflip = f.hflip;
    p.scale = f.dtype == 0 ? 1.0f / 255.0f : 1.0f;
    p.wh = wh; p.bias = bias; p.out = out;
    // uint8 frames (the path of the predictor, the sweep and bench.py): TMA + tcgen05 stem (stem_tc.cuh).  Float input (the nn.Module
    // called with already normalised frames) keeps the mma.sync gather kernel of stem.cuh.
    if (f.dtype == 0) {
        if (f.W < kStcIW || f.W % 16 || f.plane_stride % 16 || f.img_stride % 16 || (reinterpret_cast<uintptr_t>(f.data) & 15) || f.stored_h < 1)
            return fail(MDS_ERR_INVALID, "stem: uint8 frames need W >= %d, W / plane / image strides multiples of 16 and a 16-byte aligned base "
                        "(W=%d plane_stride=%lld img_stride=%lld)", kStcIW, f.W, (long long)f.plane_stride, (long long)f.img_stride);
        auto enc = tensor_map_encoder();
        if (!enc) return fail(MDS_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
        CUtensorMap tm;
        cuuint64_t dims[4] = {(cuuint64_t)f.W, (cuuint64_t)f.stored_h, 3, (cuuint64_t)n};
        cuuint64_t strides[3] = {(cuuint64_t)f.W, (cuuint64_t)f.plane_stride, (cuuint64_t)f.img_stride};
        cuuint32_t box[4] = {(cuuint32_t)kStcIW, (cuuint32_t)kStcIH, 3, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void*>(f.data), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
            return fail(MDS_ERR_CUDA, "cuTensorMapEncodeTiled(stem) failed (%d) n=%d stored_h=%d W=%d", (int)r, n, f.stored_h, f.W);
        StemTcParams q;
        q.n = n; q.H = f.H; q.W = f.W; q.Ho = f.H / 2; q.Wo = f.W / 2; q.pad_top = f.pad_top; q.hflip = f.hflip;
        q.scale = 1.0f / 255.0f; q.wh = wh; q.bias = bias; q.out = out;
        q.tiles_x = (q.Wo + kStcTW - 1) / kStcTW; q.tiles_y = (q.Ho + kStcTH - 1) / kStcTH;
        const long long tiles = (long long)q.tiles_x * q.tiles_y * n;
        ENSURE_SMEM_ATTR(stem_tc_kernel, kStcSmem);
        int grid = 2 * num_sms();
        if (tiles < grid) grid = (int)tiles;
        ProfScope ps(MDS_KIND_STEM, st);
        launch_pdl(stem_tc_kernel, dim3(grid), dim3(kStcThreads), kStcSmem, st, tm, q);
        LAUNCH_CHECK("stem_tc");
        return MDS_OK;
    }
    dim3 grid((f.W / 2 + kStemTW - 1) / kStemTW, (f.H / 2 + kStemTH - 1) / kStemTH, n);
    ProfScope ps(MDS_KIND_STEM, st);
    if (f.dtype == 1) launch_pdl(stem_kernel<float>, grid, dim3(256), 0, st, p);
    else return fail(MDS_ERR_INVALID, "stem: dtype must be 0 (uint8) or 1 (float32)");
    LAUNCH_CHECK("stem");
    return MDS_OK;
}

// every dense 3x3 block on tcgen05 (conv_tc.cuh): halo tiles through 5-D tensor maps over NHWC seen as [n][C/8][H][W][8];
// stride 2 reads the four (row, column) parity phases of the input through four maps
template <int CIN, int CMID, int CPROJ, int STRIDE, bool RES, int TH, bool PT, int MINB, bool FOLD = false>
static int launch_conv_tc(const __half* in, __half* out, const __half* w1, const float* b1, const __half* w2, const float* b2,
                          int n, int H, int W, cudaStream_t st) {
    using Cfg = ConvTcCfg<CIN, CMID, CPROJ, STRIDE, RES, TH, PT, FOLD>;
    auto enc = tensor_map_encoder();
    if (!enc) return fail(MDS_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    if (STRIDE == 2 && (H % 2 || W % 2)) return fail(MDS_ERR_INVALID, "conv_tc: stride-2 input must be even");
    ConvTcMaps maps;
    memset(&maps, 0, sizeof(maps));
    for (int ph = 0; ph < Cfg::NPH; ++ph) {
        const int py = ph >> 1, px = ph & 1;
        cuuint64_t dims[5] = {8, (cuuint64_t)(W / STRIDE), (cuuint64_t)(H / STRIDE), (cuuint64_t)(CIN / 8), (cuuint64_t)n};
        cuuint64_t strides[4] = {(cuuint64_t)STRIDE * CIN * 2, (cuuint64_t)STRIDE * W * CIN * 2, 16, (cuuint64_t)H * W * CIN * 2};
        cuuint32_t box[5] = {8, (cuuint32_t)Cfg::PW, (cuuint32_t)Cfg::PH, (cuuint32_t)(CIN / 8), 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        const __half* base = in + ((size_t)py * W + px) * CIN;
        CUresult r = enc(&maps.m[ph], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
            return fail(MDS_ERR_CUDA, "cuTensorMapEncodeTiled(conv_tc) failed (%d) n=%d H=%d W=%d C=%d stride=%d", (int)r, n, H, W, CIN, STRIDE);
    }
    ConvTcParams p;
    p.in = in; p.out = out; p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.n = n; p.H = H; p.W = W;
    p.Ho = H / STRIDE; p.Wo = W / STRIDE;
    p.tiles_x = (p.Wo + Cfg::TW - 1) / Cfg::TW;
    p.tiles_y = (p.Ho + Cfg::TH - 1) / Cfg::TH;
    p.trace = nullptr;
#ifdef MDS_CONV_TRACE
    static long long* tr = nullptr;
    if (!tr) cudaMalloc(&tr, 64 * 16 * sizeof(long long));
    cudaMemsetAsync(tr, 0, 64 * 16 * sizeof(long long), st);
    p.trace = tr;
#endif
    const long long tiles = (long long)p.tiles_x * p.tiles_y * n;
    if (tiles <= 0) return MDS_OK;
    auto kern = conv_tc_kernel<CIN, CMID, CPROJ, STRIDE, RES, TH, PT, MINB, FOLD>;
    ENSURE_SMEM_ATTR(kern, Cfg::SMEM);
    int grid = num_sms() * MINB;
    if (tiles < grid) grid = (int)tiles;
    ProfScope ps(MDS_KIND_CONV3X3, st);
    launch_pdl(kern, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM, st, maps, p);
    LAUNCH_CHECK("conv_tc");
#ifdef MDS_CONV_TRACE
    {   // columns: MMA warp 0 loop top, 1 D1 free, 2 conv MMAs issued, 3 before tile wait, 4 proj start, 5 P ready, 6 D2 free, 7 proj issued;
        // epilogue warp 2: 8 before D1 wait, 9 D1 ready, 10 E1 start, 11 E1 done, 12 next chunk requested, 13 E2(prev) done, 14 D2 ready
        static long long h[64 * 16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost);
        const long long t0 = h[3] ? h[3] : h[0];
        fprintf(stderr, "TRACE cin=%d cmid=%d stride=%d\n", CIN, CMID, STRIDE);
        for (int t = 0; t < 40; ++t) {
            fprintf(stderr, "T%02d", t);
            for (int e = 0; e < 15; ++e) fprintf(stderr, " %6lld", h[t * 16 + e] ? h[t * 16 + e] - t0 : -1);
            fprintf(stderr, "\n");
        }
    }
#endif
    return MDS_OK;
}

// blocks.2.1 on tcgen05 with the 3x3 weights streamed from L2 (conv_tc_ws.cuh).  w1t = the tap-major copy of w1 made by
// conv_w1_tapmajor_kernel (the handle keeps one per block; the stand-alone entry point repacks into a per-thread scratch buffer).
template <int CIN, int CMID, int CPROJ>
static int launch_conv_tc_ws(const __half* in, __half* out, const __half* w1, const __half* w1t, const float* b1, const __half* w2,
                             const float* b2, int n, int H, int W, cudaStream_t st) {
    using Cfg = ConvWsCfg<CIN, CMID, CPROJ>;
    auto enc = tensor_map_encoder();
    if (!enc) return fail(MDS_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    if (w1t == nullptr) {
        static thread_local __half* scratch[kMaxDevices] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= kMaxDevices) return fail(MDS_ERR_INVALID, "conv_tc_ws: device index");
        if (!scratch[dev]) CUDA_TRY(cudaMalloc(&scratch[dev], (size_t)9 * CIN * CMID * sizeof(__half)));
        conv_w1_tapmajor_kernel<<<64, 256, 0, st>>>(w1, scratch[dev], CIN, CMID);
        LAUNCH_CHECK("conv_w1_tapmajor");
        w1t = scratch[dev];
    }
    CUtensorMap tm;
    cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(CIN / 8), (cuuint64_t)n};
    cuuint64_t strides[4] = {(cuuint64_t)CIN * 2, (cuuint64_t)W * CIN * 2, 16, (cuuint64_t)H * W * CIN * 2};
    cuuint32_t box[5] = {8, (cuuint32_t)Cfg::PW, (cuuint32_t)Cfg::PH, (cuuint32_t)(CIN / 8), 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half*>(in), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MDS_ERR_CUDA, "cuTensorMapEncodeTiled(conv_tc_ws) failed (%d) n=%d H=%d W=%d C=%d", (int)r, n, H, W, CIN);
    ConvWsParams p;
    p.in = in; p.out = out; p.w1t = w1t; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.n = n; p.H = H; p.W = W;
    p.tiles_x = (W + Cfg::TW - 1) / Cfg::TW;
    p.tiles_y = (H + Cfg::TH - 1) / Cfg::TH;
    const long long tiles = (long long)p.tiles_x * p.tiles_y * n;
    if (tiles <= 0) return MDS_OK;
    auto kern = conv_tc_ws_kernel<CIN, CMID, CPROJ>;
    ENSURE_SMEM_ATTR(kern, Cfg::SMEM);
    int grid = num_sms();
    if (tiles < grid) grid = (int)tiles;
    ProfScope ps(MDS_KIND_CONV3X3, st);
    launch_pdl(kern, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM, st, tm, p);
    LAUNCH_CHECK("conv_tc_ws");
    return MDS_OK;
}

static int launch_conv3(const __half* in, __half* out, const __half* w1, const float* b1, const __half* w2,
                        const float* b2, int n, int H, int W, int cin, int cmid, int stride, int cproj, int res,
                        cudaStream_t st, const __half* w1t = nullptr) {
    if (H <= 0 || W <= 0 || n < 0) return fail(MDS_ERR_INVALID, "conv3x3: bad size n=%d H=%d W=%d", n, H, W);
    if (stride == 2 && (H % 2 || W % 2)) return fail(MDS_ERR_INVALID, "conv3x3: stride-2 input must be even");
    const bool pt = g_conv_mode == 2;
#define CTCASE(CI, CM, CP, S, R, TH_S, TH_T, MB)                                                            \
    if (cin == CI && cmid == CM && stride == S && cproj == CP && res == (R ? 1 : 0))                          \
        return pt ? launch_conv_tc<CI, CM, CP, S, R, TH_T, true, MB>(in, out, w1, b1, w2, b2, n, H, W, st)   \
                  : launch_conv_tc<CI, CM, CP, S, R, TH_S, false, MB>(in, out, w1, b1, w2, b2, n, H, W, st);
    if (cin == 32 && cmid == 16 && stride == 1 && cproj == 0 && res == 0)      // blocks.0.0 ConvBnAct: two CTAs per SM
        return pt ? launch_conv_tc<32, 16, 0, 1, false, 16, true, 2, true>(in, out, w1, b1, w2, b2, n, H, W, st)     // column taps folded into N
                  : launch_conv_tc<32, 16, 0, 1, false, 15, false, 2, false>(in, out, w1, b1, w2, b2, n, H, W, st);
    CTCASE(16, 64, 32, 2, false, 15, 15, 1)    // blocks.1.0  EdgeResidual s2
    CTCASE(32, 128, 32, 1, true, 15, 15, 1)    // blocks.1.1
    CTCASE(32, 128, 48, 2, false, 3, 7, 1)     // blocks.2.0  (P in shared memory only fits with 3-row tiles)
#undef CTCASE
    if (cin == 48 && cmid == 192 && stride == 1 && cproj == 48 && res == 1)      // blocks.2.1: 3x3 weights streamed from L2
        return launch_conv_tc_ws<48, 192, 48>(in, out, w1, w1t, b1, w2, b2, n, H, W, st);
    return fail(MDS_ERR_INVALID, "conv3x3: unsupported shape cin=%d cmid=%d stride=%d cproj=%d res=%d", cin, cmid, stride, cproj, res);
}

template <int BN, bool GATED>
static int launch_gemm_t(const GemmParams& p, cudaStream_t st) {
    using Cfg = GemmCfg<BN>;
    auto kern = gemm1x1_kernel<BN, GATED>;
    ENSURE_SMEM_ATTR(kern, Cfg::SMEM);
    const int tiles_per_img = (p.rows_per_img + kGemmBM - 1) / kGemmBM;
    dim3 grid((p.N + BN - 1) / BN, tiles_per_img * p.n_img);
    if (grid.y > 65535) return fail(MDS_ERR_INVALID, "gemm1x1: too many M tiles (%u)", grid.y);
    ProfScope ps(MDS_KIND_GEMM1X1, st);
    launch_pdl(kern, grid, dim3(256), Cfg::SMEM, st, p);
    LAUNCH_CHECK("gemm1x1");
    return MDS_OK;
}

// fp16 row-major [rows][K] matrix, box = 64 (K) x box_rows, 128-byte swizzle, out-of-bounds elements read as zero
static int make_tmap_2d(CUtensorMap* map, const void* ptr, long long rows, int K, int box_rows) {
    auto enc = tensor_map_encoder();
    if (!enc) return fail(MDS_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MDS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld K=%d box_rows=%d", (int)r, rows, K, box_rows);
    return MDS_OK;
}

// per-image gated weights [n_img][N][K] fp16: box = 64 (K) x BN x 1
static int make_tmap_3d(CUtensorMap* map, const void* ptr, int n_img, int N, int K, int box_rows) {
    auto enc = tensor_map_encoder();
    if (!enc) return fail(MDS_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, (cuuint64_t)n_img};
    cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)N * K * 2};
    cuuint32_t box[3] = {(cuuint32_t)kTcBK, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(MDS_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed (%d) n=%d N=%d K=%d", (int)r, n_img, N, K);
    return MDS_OK;
}

static int tc_pick_bn(int N) {
    if (N <= 256) return N;
    const int cand[] = {256, 224, 192, 160, 128, 112, 96, 64};
    for (int c : cand) if (N % c == 0) return c;
    return 0;
}

static int tc_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBias, TcGemmParams& p, cudaStream_t st) {
    int cols = 32;
    while (cols < 2 * p.BN) cols <<= 1;
    p.tmem_cols = cols;
    p.stages = tc_stages(p.BN, p.streamed);
    const size_t smem = tc_smem_bytes(p.BN, p.streamed);
    if (smem > 227 * 1024) return fail(MDS_ERR_INVALID, "gemm_tc: BN=%d needs %zu bytes of shared memory", p.BN, smem);
    const bool act = p.act != 0, res = p.res != nullptr;
    auto kern = act ? (res ? gemm_tc_kernel<true, true> : gemm_tc_kernel<true, false>)
                    : (res ? gemm_tc_kernel<false, true> : gemm_tc_kernel<false, false>);
    {   // the opt-in limit is a per-device attribute of each instantiation: track the largest value set on each device
        static size_t smem_set[4][kMaxDevices] = {};
        const int ki = (act ? 2 : 0) + (res ? 1 : 0);
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < kMaxDevices && smem > smem_set[ki][dev]) {
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            smem_set[ki][dev] = smem;
        }
    }
    int grid = num_sms();
    if (p.m_tiles < grid) grid = p.m_tiles;
    p.trace = nullptr;
#ifdef MDS_GEMM_TRACE
    static long long* tr = nullptr;
    if (!tr) cudaMalloc(&tr, 64 * 16 * sizeof(long long));
    cudaMemsetAsync(tr, 0, 64 * 16 * sizeof(long long), st);
    p.trace = tr;
#endif
    ProfScope ps(MDS_KIND_GEMM1X1, st);
    launch_pdl(kern, dim3(grid), dim3(kTcThreads), smem, st, tmA, tmB, tmBias, p);
    LAUNCH_CHECK("gemm_tc");
#ifdef MDS_GEMM_TRACE
    {   // per accumulator tile t.  MMA warp: 0 loop top, 1 accumulator free, 2-5 k-block 0-3 in smem, 6 all MMAs issued;
        // epilogue warps 2 / 6 (group 0 / 1): 8 before acc_full, 9 accumulator ready, 10 first TMEM pass loaded, 11 last pass loaded, 12 done
        static long long h[64 * 16];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost);
        const long long t0 = h[0];
        fprintf(stderr, "GTRACE N=%d K=%d BN=%d streamed=%d act=%d m_tiles=%d n_tiles=%d\n", p.N, p.K, p.BN, p.streamed, p.act, p.m_tiles, p.n_tiles);
        for (int t = 0; t < 24; ++t) {
            fprintf(stderr, "T%02d", t);
            for (int e = 0; e < 13; ++e) fprintf(stderr, " %6lld", h[t * 16 + e] ? h[t * 16 + e] - t0 : -1);
            fprintf(stderr, "\n");
        }
    }
#endif
    return MDS_OK;
}

// resident-A mode: C = act(A W^T + bias), K <= 192, all N tiles swept per row tile
static int launch_gemm_tc(const __half* A, const __half* W, const __half* bias_mat, __half* C, long long M, int N, int K, int act,
                          cudaStream_t st) {
    const int BN = tc_pick_bn(N);
    if (BN < 32 || BN % 16 || K > kTcMaxKB * kTcBK) return fail(MDS_ERR_INVALID, "gemm_tc: unsupported N=%d K=%d", N, K);
    CUtensorMap tmA, tmB, tmBias;
    TRY(make_tmap_2d(&tmA, A, M, K, kTcBM));
    TRY(make_tmap_2d(&tmB, W, N, K, BN));
    TRY(make_tmap_2d(&tmBias, bias_mat, N, kTcBK, BN));
    TcGemmParams p;
    p.C = C; p.res = nullptr; p.M = M; p.rows_per_img = (int)M; p.tiles_per_img = (int)((M + kTcBM - 1) / kTcBM);
    p.N = N; p.K = K; p.BN = BN; p.act = act; p.streamed = 0; p.gate = nullptr;
    p.m_tiles = p.tiles_per_img;
    p.n_tiles = N / BN;
    return tc_launch(tmA, tmB, tmBias, p, st);
}

// streamed mode: C[img] = act(A[img] Wg[img]^T + bias) (+ res); Wg = per-image SE-gated weights written by se_fc_kernel
static int launch_gemm_tc_stream(const __half* A, const __half* Wg, const __half* bias_mat, const __half* res, __half* C,
                                 int rows_per_img, int n_img, int N, int K, int act, cudaStream_t st, const float* gate = nullptr) {
    if (N < 32 || N > 256 || N % 16 || K % 16) return fail(MDS_ERR_INVALID, "gemm_tc_stream: unsupported N=%d K=%d", N, K);
    if (rows_per_img <= 0 || n_img <= 0) return MDS_OK;
    const long long M = (long long)rows_per_img * n_img;
    if (M >= (1LL << 31)) return fail(MDS_ERR_INVALID, "gemm_tc_stream: too many rows");
    CUtensorMap tmA, tmB, tmBias;
    TRY(make_tmap_2d(&tmA, A, M, K, kTcBM));
    if (gate) {       // Wg = the shared [N][K] weights; the kernel applies gate[img] to the A blocks
        if (K > kTcMaxGateK) return fail(MDS_ERR_INVALID, "gemm_tc_stream: gated K=%d exceeds %d", K, kTcMaxGateK);
        TRY(make_tmap_2d(&tmB, Wg, N, K, N));
    } else {
        TRY(make_tmap_3d(&tmB, Wg, n_img, N, K, N));
    }
    TRY(make_tmap_2d(&tmBias, bias_mat, N, kTcBK, N));
    TcGemmParams p;
    p.C = C; p.res = res; p.M = M; p.rows_per_img = rows_per_img; p.tiles_per_img = (rows_per_img + kTcBM - 1) / kTcBM;
    p.N = N; p.K = K; p.BN = N; p.act = act; p.streamed = 1; p.gate = gate;
    p.m_tiles = p.tiles_per_img * n_img;
    p.n_tiles = 1;
    return tc_launch(tmA, tmB, tmBias, p, st);
}

// bias_mat: [N][64] fp16, column 0 = fp16(bias), column 1 = fp16(bias - column 0), rest 0 (packer.py): the tcgen05 path.
// The forward passes ONLY run the tcgen05 kernel: a shape outside its limits is an error, not a silent switch to another
// kernel.  legacy = true (training GEMMs with a residual and N > 256; the mds_k_gemm1x1 cross-check entry) selects the
// mma.sync kernel of gemm1x1.cuh explicitly.
static int launch_gemm(const __half* A, const __half* W, const float* bias, const __half* res, const __half* gate,
                       __half* C, long long rows_per_img, int n_img, int N, int K, int act, cudaStream_t st,
                       const __half* bias_mat = nullptr, bool legacy = false) {
    if (K % 16 || N % 16 || K > 1152) return fail(MDS_ERR_INVALID, "gemm1x1: N, K must be multiples of 16, K <= 1152 (N=%d K=%d)", N, K);
    if (rows_per_img <= 0 || n_img <= 0) return MDS_OK;
    if (bias_mat != nullptr && gate == nullptr && res == nullptr && K <= kTcMaxKB * kTcBK && rows_per_img * n_img < (1LL << 31) && tc_pick_bn(N) >= 32)
        return launch_gemm_tc(A, W, bias_mat, C, rows_per_img * n_img, N, K, act, st);
    if (!legacy)
        return fail(MDS_ERR_INVALID, "gemm: N=%d K=%d (gate %d, residual %d) is outside the tcgen05 kernel's limits (ungated, K <= %d, "
                    "N a multiple of 64..256)", N, K, gate != nullptr, res != nullptr, kTcMaxKB * kTcBK);
    // The M-tile index lives in gridDim.y (<= 65535): split very large ungated problems into row slabs.
    const long long max_rows = 65535LL * kGemmBM;
    if (gate == nullptr && rows_per_img * n_img > max_rows) {
        long long total = rows_per_img * n_img;
        for (long long r0 = 0; r0 < total; r0 += max_rows) {
            long long rows = total - r0 < max_rows ? total - r0 : max_rows;
            TRY(launch_gemm(A + r0 * K, W, bias, res ? res + r0 * N : nullptr, nullptr, C + r0 * N, rows, 1, N, K, act, st, nullptr, true));
        }
        return MDS_OK;
    }
    GemmParams p;
    p.A = A; p.W = W; p.bias = bias; p.res = res; p.gate = gate; p.C = C;
    p.N = N; p.K = K; p.act = act;
    if (gate == nullptr) { p.rows_per_img = (int)(rows_per_img * n_img); p.n_img = 1; }
    else { p.rows_per_img = (int)rows_per_img; p.n_img = n_img; }
#define GCASE(BN_)                                                      \
    if (N % BN_ == 0) {                                                 \
        if (gate) return launch_gemm_t<BN_, true>(p, st);               \
        return launch_gemm_t<BN_, false>(p, st);                        \
    }
    GCASE(128) GCASE(112) GCASE(96) GCASE(64)
#undef GCASE
    return fail(MDS_ERR_INVALID, "gemm1x1: N=%d is not a multiple of 128/112/96/64", N);
}

template <int KT, int STRIDE>
static int launch_dw_t(const DwParams& p, dim3 grid, cudaStream_t st) {
    using Cfg = DwCfg<KT, STRIDE>;
    auto kern = dwconv_kernel<KT, STRIDE>;
    ENSURE_SMEM_ATTR(kern, Cfg::SMEM);
    ProfScope ps(KT == 1 ? MDS_KIND_DWCONV2D : MDS_KIND_DWCONV3D, st);
    launch_pdl(kern, grid, dim3(256), Cfg::SMEM, st, p);
    LAUNCH_CHECK("dwconv");
    return MDS_OK;
}

constexpr int kDwMaxParts = 64;     // upper bound on the squeeze partials per image (workspace sizing)
static int launch_dw(const __half* in, __half* out, const float* w, const float* bias, float* partials, int* nparts_out, int n,
                     int T, int H, int W, int C, int kt, int stride, cudaStream_t st) {
    if (C % 8 || C > 4096) return fail(MDS_ERR_INVALID, "dwconv: C must be a multiple of 8");
    if (!((kt == 1 && (stride == 1 || stride == 2)) || (kt == 3 && stride == 1)))
        return fail(MDS_ERR_INVALID, "dwconv: unsupported kt=%d stride=%d", kt, stride);
    if (kt == 1 && T != 1) return fail(MDS_ERR_INVALID, "dwconv 2D: T must be 1");
    if (stride == 2 && (H % 2 || W % 2)) return fail(MDS_ERR_INVALID, "dwconv: stride-2 input must be even");
    if (n <= 0) return MDS_OK;
    if (n > 65535) return fail(MDS_ERR_INVALID, "dwconv: n too large");
    DwParams p;
    p.in = in; p.out = out; p.w = w; p.bias = bias; p.partials = partials;
    p.n = n; p.T = T; p.H = H; p.W = W; p.C = C;
    p.Ho = H / stride; p.Wo = W / stride;
    p.xtiles = (p.Wo + kDwTWX - 1) / kDwTWX;
    p.slabs = (C + kDwCS - 1) / kDwCS;
    // enough CTAs to fill the machine a few times over: split rows when the batch is small (halo rows are re-read)
    int chunks = 1;
    const long long base = (long long)p.xtiles * p.slabs * n * T;
    const long long want = (long long)num_sms() * 4;     // measured at batch 4 / 8: 4 CTAs per SM beat 6 and 8 by ~0.8 %
    if (base < want) {
        chunks = (int)((want + base - 1) / base);
        const int max_chunks = p.Ho / 6 > 0 ? p.Ho / 6 : 1;
        if (chunks > max_chunks) chunks = max_chunks;
    }
    while (chunks > 1 && (long long)chunks * T * p.xtiles > kDwMaxParts) --chunks;
    p.rows_per_chunk = (p.Ho + chunks - 1) / chunks;
    p.chunks = (p.Ho + p.rows_per_chunk - 1) / p.rows_per_chunk;
    p.nparts = p.chunks * T * p.xtiles;
    if (p.nparts > kDwMaxParts) return fail(MDS_ERR_INVALID, "dwconv: %d squeeze partials per image exceed %d (T=%d, W=%d)", p.nparts, kDwMaxParts, T, W);
    if (nparts_out) *nparts_out = p.nparts;
    dim3 grid(p.xtiles * p.slabs, p.chunks * T, n);
    if (kt == 3) return launch_dw_t<3, 1>(p, grid, st);
    if (stride == 1) return launch_dw_t<1, 1>(p, grid, st);
    return launch_dw_t<1, 2>(p, grid, st);
}

static int launch_se(const float* partials, int nparts, const float* w1, const float* b1, const float* w2t, const float* b2,
                     __half* gate, const float* w32, __half* wg, int n, int C, int rd, int N, float inv_count, cudaStream_t st) {
    if (n <= 0) return MDS_OK;
    if (C > 4096 || rd > 64 || C % 8) return fail(MDS_ERR_INVALID, "se_fc: C must be a multiple of 8, C <= 4096, rd <= 64");
    if (n > 65535 || nparts <= 0) return fail(MDS_ERR_INVALID, "se_fc: bad n / nparts");
    SeParams p;
    p.partials = partials; p.nparts = nparts; p.w1 = w1; p.b1 = b1; p.w2t = w2t; p.b2 = b2; p.gate = gate;
    p.w32 = w32; p.wg = (w32 != nullptr) ? wg : nullptr; p.C = C; p.rd = rd; p.N = N; p.inv_count = inv_count;
    const size_t smem = se_smem_bytes(C, rd);
    if (smem > 227 * 1024) return fail(MDS_ERR_INVALID, "se_fc: C=%d rd=%d need %zu bytes of shared memory", C, rd, smem);
    {
        static size_t smem_set[kMaxDevices] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < kMaxDevices && smem > smem_set[dev]) {
            CUDA_TRY(cudaFuncSetAttribute(se_fc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            smem_set[dev] = smem;
        }
    }
    ProfScope ps(MDS_KIND_SE_FC, st);
    {   // one cluster of kSeSlices CTAs per image (+ programmatic stream serialization)
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(kSeSlices, n); cfg.blockDim = dim3(kSeThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kSeSlices; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = g_pdl ? 2 : 1;
        cudaLaunchKernelEx(&cfg, se_fc_kernel, p);
    }
    LAUNCH_CHECK("se_fc");
    return MDS_OK;
}


// ---- fused MBConv tail (mbconv_tail.cuh): depthwise + SE + gated projection in one persistent launch ----
// input halo rows for the depthwise items: 4-D ([n][H][W][C]) or 5-D ([n][T][H][W][C]) map, no swizzle, zero fill
static int make_tmap_dw(CUtensorMap* map, const void* ptr, int n, int T, int H, int W, int C, int kt, int stride) {
    auto enc = tensor_map_encoder();
    if (!enc) return fail(MDS_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    const cuuint32_t iw = stride == 1 ? kTlTWX + 2 : 2 * kTlTWX + 1;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r;
    if (kt == 1) {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)kTlCS, iw, (cuuint32_t)(stride == 1 ? 2 : 1), 1};      // rows per stage: TailCfg::RPS
        r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)n};
        cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)T * H * W * C * 2};
        cuuint32_t box[5] = {(cuuint32_t)kTlCS, iw, 1, 3, 1};
        r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return fail(MDS_ERR_CUDA, "cuTensorMapEncodeTiled(dw) failed (%d) n=%d T=%d H=%d W=%d C=%d", (int)r, n, T, H, W, C);
    return MDS_OK;
}

struct TailArgs {
    const __half* m1;        // [n][T][H][W][C] expanded input
    __half* m2;              // [n][T][Ho][Wo][C]
    const float *dw_w, *dw_b, *se_w1, *se_b1, *se_w2t, *se_b2;
    float* partials;         // [n][kDwMaxParts][C]
    float* gate;             // [n][C] f32
    int* sync;               // [3][sync_stride] zero
    int sync_stride;
    const __half* wpwl;      // [N][C] fp16 (N == 0: depthwise + SE only)
    const __half* bm;        // [N][64] bias matrix
    const __half* res;
    __half* out;
    int n, T, H, W, C, kt, stride, rd, N;
    int rows_per_chunk;      // 0: default
};

static int g_tail_rows = getenv("MDS_TAIL_ROWS") ? atoi(getenv("MDS_TAIL_ROWS")) : 12;

template <int KT, int STRIDE>
static int launch_tail_t(const CUtensorMap& tmDw, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBias,
                         TailParams& p, cudaStream_t st) {
    using Cfg = TailCfg<KT, STRIDE>;
    auto kern = mbconv_tail_kernel<KT, STRIDE>;
    p.g_stages = 4;
    size_t smem = tl_smem_bytes(Cfg::DW_RING, p.N, p.g_stages, KT);
    while (smem > 227 * 1024 && p.g_stages > 2) smem = tl_smem_bytes(Cfg::DW_RING, p.N, --p.g_stages, KT);
    if (smem > 227 * 1024) return fail(MDS_ERR_INVALID, "mbconv_tail: N=%d needs %zu bytes of shared memory", p.N, smem);
    {
        static size_t smem_set[kMaxDevices] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < kMaxDevices && smem > smem_set[dev]) {
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            smem_set[dev] = smem;
        }
    }
    int grid = num_sms();
    const int n_dw = p.n * p.dw_per_img, n_tiles = p.n * p.tiles_per_img;
    if (n_dw < grid && n_tiles < grid) grid = n_dw > n_tiles ? n_dw : n_tiles;
    ProfScope ps(KT == 1 ? MDS_KIND_TAIL2D : MDS_KIND_TAIL3D, st);
    launch_pdl(kern, dim3(grid), dim3(kTlThreads), smem, st, tmDw, tmA, tmB, tmBias, p);
    LAUNCH_CHECK("mbconv_tail");
    return MDS_OK;
}

static int launch_tail(const TailArgs& a, cudaStream_t st) {
    if (a.n <= 0) return MDS_OK;
    if (a.C % 8 || a.C > kTlMaxC || a.C < kTlCS) return fail(MDS_ERR_INVALID, "mbconv_tail: C must be a multiple of 8 in [128, 1152] (C=%d)", a.C);
    if (a.rd <= 0 || a.rd > kTlMaxRd) return fail(MDS_ERR_INVALID, "mbconv_tail: rd must be in [1, 64]");
    if (!((a.kt == 1 && (a.stride == 1 || a.stride == 2)) || (a.kt == 3 && a.stride == 1)))
        return fail(MDS_ERR_INVALID, "mbconv_tail: unsupported kt=%d stride=%d", a.kt, a.stride);
    if (a.kt == 1 && a.T != 1) return fail(MDS_ERR_INVALID, "mbconv_tail 2D: T must be 1");
    if (a.stride == 2 && (a.H % 2 || a.W % 2)) return fail(MDS_ERR_INVALID, "mbconv_tail: stride-2 input must be even");
    if (a.N && (a.N % 16 || a.N < 32 || a.N > 256)) return fail(MDS_ERR_INVALID, "mbconv_tail: N must be a multiple of 16 in [32, 256]");
    if (a.n > a.sync_stride) return fail(MDS_ERR_INVALID, "mbconv_tail: %d images exceed the %d sync slots", a.n, a.sync_stride);
    TailParams p;
    memset(&p, 0, sizeof(p));
    p.m2 = a.m2; p.dw_w = a.dw_w; p.dw_b = a.dw_b; p.partials = a.partials;
    p.se_w1 = a.se_w1; p.se_b1 = a.se_b1; p.se_w2t = a.se_w2t; p.se_b2 = a.se_b2; p.gate = a.gate;
    p.sync = a.sync; p.sync_stride = a.sync_stride; p.out = a.out; p.res = a.res;
    p.n = a.n; p.T = a.T; p.H = a.H; p.W = a.W; p.C = a.C; p.Ho = a.H / a.stride; p.Wo = a.W / a.stride; p.rd = a.rd;
    p.inv_count = 1.0f / (float)((size_t)a.T * p.Ho * p.Wo);
    // row chunks are a function of the layer shape only, so the squeeze partial sums (and with them every logit) do not
    // depend on what else is in the batch
    const int want_rows = a.rows_per_chunk > 0 ? a.rows_per_chunk : g_tail_rows;
    int chunks = (p.Ho + want_rows / 2) / want_rows;
    if (chunks < 1) chunks = 1;
    p.xtiles = (p.Wo + kTlTWX - 1) / kTlTWX;
    while (chunks > 1 && (long long)chunks * a.T * p.xtiles > kDwMaxParts) --chunks;
    p.rows_per_chunk = (p.Ho + chunks - 1) / chunks;
    p.chunks = (p.Ho + p.rows_per_chunk - 1) / p.rows_per_chunk;
    p.slab_pairs = (a.C + kTlCS - 1) / kTlCS;
    p.nparts = p.chunks * a.T * p.xtiles;
    if (p.nparts > kDwMaxParts) return fail(MDS_ERR_INVALID, "mbconv_tail: %d squeeze partials per image exceed %d", p.nparts, kDwMaxParts);
    p.dw_per_img = a.T * p.chunks * p.xtiles * p.slab_pairs;
    p.rows_per_img = a.T * p.Ho * p.Wo;
    p.N = a.N;
    p.tiles_per_img = (a.N && !g_tail_dw_only) ? (p.rows_per_img + kTcBM - 1) / kTcBM : 0;
    p.num_kb = (a.C + kTcBK - 1) / kTcBK;
    if ((long long)a.n * p.dw_per_img >= (1LL << 31) || (long long)a.n * p.tiles_per_img >= (1LL << 31))
        return fail(MDS_ERR_INVALID, "mbconv_tail: too many work items");
    int cols = 32;
    while (cols < 2 * (a.N ? a.N : 16)) cols <<= 1;
    p.tmem_cols = cols;

    CUtensorMap tmDw, tmA, tmB, tmBias;
    TRY(make_tmap_dw(&tmDw, a.m1, a.n, a.T, a.H, a.W, a.C, a.kt, a.stride));
    if (a.N) {
        TRY(make_tmap_2d(&tmA, a.m2, (long long)a.n * p.rows_per_img, a.C, kTcBM));
        TRY(make_tmap_2d(&tmB, a.wpwl, a.N, a.C, a.N));
        TRY(make_tmap_2d(&tmBias, a.bm, a.N, kTcBK, a.N));
    } else {
        tmA = tmDw; tmB = tmDw; tmBias = tmDw;      // never dereferenced
    }
    if (a.kt == 3) return launch_tail_t<3, 1>(tmDw, tmA, tmB, tmBias, p, st);
    if (a.stride == 1) return launch_tail_t<1, 1>(tmDw, tmA, tmB, tmBias, p, st);
    return launch_tail_t<1, 2>(tmDw, tmA, tmB, tmBias, p, st);
}


// ---- TMA-staged depthwise + SE (dwconv_tma.cuh) ----
static int g_dw_rows2d = getenv("MDS_DW_ROWS") ? atoi(getenv("MDS_DW_ROWS")) : 12;
static int g_dw_rows3d = getenv("MDS_DW_ROWS3D") ? atoi(getenv("MDS_DW_ROWS3D")) : 8;
template <int KT, int STRIDE>
static int launch_dw_se_t(const CUtensorMap& tm, const DwSeParams& p, dim3 grid, cudaStream_t st) {
    using Cfg = DwTmaCfg<KT, STRIDE>;
    auto kern = dwconv_tma_kernel<KT, STRIDE>;
    ENSURE_SMEM_ATTR(kern, Cfg::SMEM);
    ProfScope ps(KT == 1 ? MDS_KIND_DWCONV2D : MDS_KIND_DWCONV3D, st);
    launch_pdl(kern, grid, dim3(256), Cfg::SMEM, st, tm, p);
    LAUNCH_CHECK("dwconv_tma");
    return MDS_OK;
}
static int launch_dw_se(const __half* in, __half* out, const float* w, const float* bias, float* partials, int* nparts_out,
                        const float* se_w1, const float* se_b1, const float* se_w2t, const float* se_b2, float* gate, int* done,
                        int n, int T, int H, int W, int C, int kt, int stride, int rd, int rows_per_chunk, cudaStream_t st) {
    if (C % 8 || C > 1152) return fail(MDS_ERR_INVALID, "dwconv_tma: C must be a multiple of 8, <= 1152");
    if (!((kt == 1 && (stride == 1 || stride == 2)) || (kt == 3 && stride == 1)))
        return fail(MDS_ERR_INVALID, "dwconv_tma: unsupported kt=%d stride=%d", kt, stride);
    if (kt == 1 && T != 1) return fail(MDS_ERR_INVALID, "dwconv_tma 2D: T must be 1");
    if (stride == 2 && (H % 2 || W % 2)) return fail(MDS_ERR_INVALID, "dwconv_tma: stride-2 input must be even");
    if (se_w1 && (rd <= 0 || rd > 64 || !gate || !done)) return fail(MDS_ERR_INVALID, "dwconv_tma: bad SE arguments");
    if (n <= 0) return MDS_OK;
    if (n > 65535) return fail(MDS_ERR_INVALID, "dwconv_tma: n too large");
    DwSeParams p;
    memset(&p, 0, sizeof(p));
    p.out = out; p.w = w; p.bias = bias; p.partials = partials; p.se_w1 = se_w1; p.se_b1 = se_b1; p.se_w2t = se_w2t; p.se_b2 = se_b2;
    p.gate = gate; p.done = done; p.n = n; p.T = T; p.H = H; p.W = W; p.C = C; p.Ho = H / stride; p.Wo = W / stride; p.rd = rd;
    p.inv_count = 1.0f / (float)((size_t)T * p.Ho * p.Wo);
    // row chunks depend on the layer shape only: the squeeze sums of an image never depend on the rest of the batch
    const int want = rows_per_chunk > 0 ? rows_per_chunk : (kt == 3 ? g_dw_rows3d : g_dw_rows2d);
    int chunks = (p.Ho + want / 2) / want;
    if (chunks < 1) chunks = 1;
    p.xtiles = (p.Wo + kDtTWX - 1) / kDtTWX;
    while (chunks > 1 && (long long)chunks * T * p.xtiles > kDwMaxParts) --chunks;
    p.rows_per_chunk = (p.Ho + chunks - 1) / chunks;
    p.chunks = (p.Ho + p.rows_per_chunk - 1) / p.rows_per_chunk;
    p.slabs = (C + kDtCS - 1) / kDtCS;
    p.nparts = p.chunks * T * p.xtiles;
    if (p.nparts > kDwMaxParts) return fail(MDS_ERR_INVALID, "dwconv_tma: %d squeeze partials per image exceed %d", p.nparts, kDwMaxParts);
    if (nparts_out) *nparts_out = p.nparts;
    p.ctas_per_img = T * p.chunks * p.xtiles * p.slabs;
    if ((long long)p.chunks * T > 65535) return fail(MDS_ERR_INVALID, "dwconv_tma: too many row chunks");

    auto enc = tensor_map_encoder();
    if (!enc) return fail(MDS_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    CUtensorMap tm;
    const cuuint32_t iw = stride == 1 ? kDtTWX + 2 : 2 * kDtTWX + 1;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r;
    if (kt == 1) {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
        cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)kDtCS, iw, (cuuint32_t)(stride == 1 ? 3 : 2), 1};       // rows per stage: DwTmaCfg::RPS
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)n};
        cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)T * H * W * C * 2};
        cuuint32_t box[5] = {(cuuint32_t)kDtCS, iw, 1, 3, 1};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half*>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return fail(MDS_ERR_CUDA, "cuTensorMapEncodeTiled(dwconv_tma) failed (%d) n=%d T=%d H=%d W=%d C=%d", (int)r, n, T, H, W, C);
    dim3 grid(p.xtiles * p.slabs, p.chunks * T, n);
    if (kt == 3) return launch_dw_se_t<3, 1>(tm, p, grid, st);
    if (stride == 1) return launch_dw_se_t<1, 1>(tm, p, grid, st);
    return launch_dw_se_t<1, 2>(tm, p, grid, st);
}

static size_t gem_part_floats(int b, int T, int C) { return (size_t)b * T * kGemSplit * C; }
static int launch_gem(const __half* x, float* feat, float* part, int b, int T, int P, int C, float pw, float eps, cudaStream_t st) {
    if (C % 8 || C > 256 || C < 32) return fail(MDS_ERR_INVALID, "gem: C must be a multiple of 8 in [32, 256]");
    if (b <= 0) return MDS_OK;
    if (b > 65535) return fail(MDS_ERR_INVALID, "gem: b too large");
    GemParams g;
    g.x = x; g.part = part; g.feat = feat; g.T = T; g.P = P; g.C = C; g.p = pw; g.eps = eps;
    ProfScope ps(MDS_KIND_HEAD, st);
    launch_pdl(gem_kernel, dim3(T, b, kGemSplit), dim3(256), 0, st, g);
    LAUNCH_CHECK("gem");
    const int total = b * T * C;
    launch_pdl(gem_finish_kernel, dim3((total + 255) / 256), dim3(256), 0, st, g, total);
    LAUNCH_CHECK("gem_finish");
    return MDS_OK;
}

static int launch_linear(const float* feat, const float* w, const float* bias, float* out, int b, int F, int k, int sig,
                         cudaStream_t st) {
    if (b <= 0) return MDS_OK;
    ProfScope ps(MDS_KIND_HEAD, st);
    launch_pdl(linear_head_kernel, dim3(b), dim3(256), 0, st, feat, w, bias, out, F, k, sig);
    LAUNCH_CHECK("linear_head");
    return MDS_OK;
}

// --------------------------------------------------------------------------------------------------------------
// handle: packed weights + layer program
// --------------------------------------------------------------------------------------------------------------
struct DevBuf {
    void* ptr = nullptr;
    size_t bytes = 0;
};

struct Block2d {
    char kind;   // 'c' ConvBnAct, 'e' EdgeResidual, 'i' InvertedResidual
    int cin, mid, cout, stride, rd;
    bool skip;
    const __half *w1 = nullptr, *w1t = nullptr, *w2 = nullptr, *wpw = nullptr, *wpwl = nullptr, *bmpw = nullptr, *bmpwl = nullptr;
    const float *b1 = nullptr, *b2 = nullptr, *bpw = nullptr, *bpwl = nullptr;
    const float *wdw = nullptr, *bdw = nullptr, *se_w1 = nullptr, *se_b1 = nullptr, *se_w2t = nullptr, *se_b2 = nullptr, *w32pwl = nullptr;
};
struct Block3d {
    const __half *wpw, *wpwl, *bmpw, *bmpwl;
    const float *bpw, *bpwl, *wdw, *bdw, *se_w1, *se_b1, *se_w2t, *se_b2, *w32pwl;
};

struct MdsHandle {
    MdsConfig cfg;
    std::map<std::string, DevBuf> tensors;
    bool committed = false;
    const __half* stem_w = nullptr;
    const float* stem_b = nullptr;
    std::vector<Block2d> blocks;
    const __half *proj2d_w = nullptr, *proj3d_w = nullptr, *proj2d_bm = nullptr, *proj3d_bm = nullptr;
    const float *proj2d_b = nullptr, *proj3d_b = nullptr;
    std::vector<Block3d> blocks3d;
    float gem_p = 3.0f;
    const float *cls_w = nullptr, *cls_b = nullptr;
    int* sync = nullptr;          // [3][kSyncSlots] inter-CTA words of the fused MBConv tail; zero between launches
    cudaStream_t aux_stream[kMaxStreams - 1] = {};      // extra streams of the multi-stream encoder (mds_set_streams)
    cudaEvent_t ev_fork = nullptr, ev_join[kMaxStreams - 1] = {};
    int T() const { return cfg.num_frames / cfg.stack_size; }
    int mid3d() const { return cfg.num_3d_features * cfg.expansion_3d_ratio; }
    int rd3d() const { return mid3d() / cfg.se_reduce_3d_ratio; }
};

struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) { cudaSetDevice(dev); changed = true; }
    }
    ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
};

constexpr int kSyncSlots = 4096;      // images per pass of the fused MBConv tail (>= chunk_images)
static const int kStageDefs[6][6] = {
    // kind(0 c,1 e,2 i), repeats, stride, expand, cout, se (x100)
    {0, 1, 1, 1, 16, 0}, {1, 2, 2, 4, 32, 0}, {1, 2, 2, 4, 48, 0}, {2, 3, 2, 4, 96, 25}, {2, 5, 1, 6, 112, 25}, {2, 8, 2, 6, 192, 25}};

extern "C" int mds_create(const MdsConfig* cfg, MdsHandle** out) {
    if (!cfg || !out) return fail(MDS_ERR_INVALID, "mds_create: null argument");
    if (cfg->stack_size != 3) return fail(MDS_ERR_INVALID, "stack_size must be 3 (got %d)", cfg->stack_size);
    if (cfg->num_frames <= 0 || cfg->num_frames % cfg->stack_size)
        return fail(MDS_ERR_INVALID, "num_frames (%d) must be a positive multiple of stack_size", cfg->num_frames);
    if (cfg->num_3d_features != 192) return fail(MDS_ERR_INVALID, "num_3d_features must be 192");
    if (cfg->num_3d_stack_proj % 64 || cfg->num_3d_stack_proj > 256 || cfg->num_3d_stack_proj <= 0)
        return fail(MDS_ERR_INVALID, "num_3d_stack_proj must be a multiple of 64, <= 256");
    const int mid = cfg->num_3d_features * cfg->expansion_3d_ratio;
    if (mid % 64 || mid > 1152 || cfg->se_reduce_3d_ratio <= 0 || mid / cfg->se_reduce_3d_ratio <= 0)
        return fail(MDS_ERR_INVALID, "unsupported 3D expansion / SE ratio");
    if (cfg->num_classes <= 0 || cfg->num_3d_blocks < 0) return fail(MDS_ERR_INVALID, "bad num_classes / num_3d_blocks");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(MDS_ERR_INVALID, "device %d out of range", cfg->device);
    MdsHandle* h = new MdsHandle();
    h->cfg = *cfg;
    if (h->cfg.chunk_images <= 0) h->cfg.chunk_images = 160;   // one pass for up to 32 stacks; ~41 MB of scratch per image
    if (h->cfg.chunk_images > kSyncSlots) h->cfg.chunk_images = kSyncSlots;
    {
        DeviceGuard g(cfg->device);
        cudaError_t e = cudaMalloc(&h->sync, 3 * kSyncSlots * sizeof(int));
        if (e == cudaSuccess) e = cudaMemset(h->sync, 0, 3 * kSyncSlots * sizeof(int));
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
        for (int i = 0; i < kMaxStreams - 1 && e == cudaSuccess; ++i) {
            e = cudaStreamCreateWithFlags(&h->aux_stream[i], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) { delete h; return fail(MDS_ERR_CUDA, "mds_create: %s", cudaGetErrorString(e)); }
    }
    *out = h;
    return MDS_OK;
}

extern "C" int mds_destroy(MdsHandle* h) {
    if (!h) return MDS_OK;
    DeviceGuard g(h->cfg.device);
    for (auto& kv : h->tensors) cudaFree(kv.second.ptr);
    cudaFree(h->sync);
    for (int i = 0; i < kMaxStreams - 1; ++i) {
        if (h->aux_stream[i]) cudaStreamDestroy(h->aux_stream[i]);
        if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    delete h;
    return MDS_OK;
}

extern "C" int mds_weights_add(MdsHandle* h, const char* name, const void* host_data, size_t nbytes) {
    if (!h || !name || !host_data || nbytes == 0) return fail(MDS_ERR_INVALID, "mds_weights_add: bad argument");
    DeviceGuard g(h->cfg.device);
    DevBuf& b = h->tensors[name];
    if (b.ptr && b.bytes != nbytes) { cudaFree(b.ptr); b.ptr = nullptr; }
    if (!b.ptr) CUDA_TRY(cudaMalloc(&b.ptr, nbytes));
    b.bytes = nbytes;
    CUDA_TRY(cudaMemcpy(b.ptr, host_data, nbytes, cudaMemcpyHostToDevice));
    h->committed = false;
    return MDS_OK;
}

template <typename T>
static int get_tensor(MdsHandle* h, const std::string& name, size_t count, const T** out) {
    auto it = h->tensors.find(name);
    if (it == h->tensors.end()) return fail(MDS_ERR_WEIGHTS, "missing tensor '%s'", name.c_str());
    if (it->second.bytes != count * sizeof(T))
        return fail(MDS_ERR_WEIGHTS, "tensor '%s': expected %zu bytes, got %zu", name.c_str(), count * sizeof(T), it->second.bytes);
    *out = reinterpret_cast<const T*>(it->second.ptr);
    return MDS_OK;
}

extern "C" int mds_weights_commit(MdsHandle* h) {
    if (!h) return fail(MDS_ERR_INVALID, "null handle");
    DeviceGuard g(h->cfg.device);
    h->blocks.clear();
    h->blocks3d.clear();
    TRY(get_tensor(h, "stem.wh", 2 * 32 * 32, &h->stem_w));
    TRY(get_tensor(h, "stem.b", 32, &h->stem_b));
    int cin = 32;
    for (int si = 0; si < 6; ++si) {
        const int* d = kStageDefs[si];
        for (int bi = 0; bi < d[1]; ++bi) {
            Block2d b;
            b.kind = d[0] == 0 ? 'c' : d[0] == 1 ? 'e' : 'i';
            b.cin = cin; b.cout = d[4]; b.stride = bi == 0 ? d[2] : 1; b.mid = cin * d[3];
            b.rd = d[5] ? (cin * d[5] + 50) / 100 : 0;
            b.skip = b.kind != 'c' && b.stride == 1 && b.cin == b.cout;
            char pre[32];
            snprintf(pre, sizeof(pre), "b%d.%d.", si, bi);
            const std::string p(pre);
            if (b.kind == 'c') {
                TRY(get_tensor(h, p + "c3.w", (size_t)b.cout * 9 * b.cin, &b.w1));
                TRY(get_tensor(h, p + "c3.b", b.cout, &b.b1));
            } else if (b.kind == 'e') {
                TRY(get_tensor(h, p + "c3.w", (size_t)b.mid * 9 * b.cin, &b.w1));
                TRY(get_tensor(h, p + "c3.b", b.mid, &b.b1));
                TRY(get_tensor(h, p + "pwl.w", (size_t)b.cout * b.mid, &b.w2));
                TRY(get_tensor(h, p + "pwl.b", b.cout, &b.b2));
                if (b.cin == 48 && b.mid == 192 && b.stride == 1) {      // conv_tc_ws_kernel streams a tap-major copy of the 3x3 weights
                    DevBuf& wt = h->tensors[p + "c3.wt"];
                    const size_t bytes = (size_t)b.mid * 9 * b.cin * sizeof(__half);
                    if (wt.bytes != bytes) {
                        if (wt.ptr) cudaFree(wt.ptr);
                        wt.ptr = nullptr; wt.bytes = 0;
                        CUDA_TRY(cudaMalloc(&wt.ptr, bytes));
                        wt.bytes = bytes;
                    }
                    conv_w1_tapmajor_kernel<<<64, 256>>>(b.w1, reinterpret_cast<__half*>(wt.ptr), b.cin, b.mid);
                    CUDA_TRY(cudaGetLastError());
                    CUDA_TRY(cudaDeviceSynchronize());
                    b.w1t = reinterpret_cast<const __half*>(wt.ptr);
                }
            } else {
                TRY(get_tensor(h, p + "pw.w", (size_t)b.mid * b.cin, &b.wpw));
                TRY(get_tensor(h, p + "pw.b", b.mid, &b.bpw));
                TRY(get_tensor(h, p + "pw.bm", (size_t)b.mid * 64, &b.bmpw));
                TRY(get_tensor(h, p + "dw.w", (size_t)9 * b.mid, &b.wdw));
                TRY(get_tensor(h, p + "dw.b", b.mid, &b.bdw));
                TRY(get_tensor(h, p + "se.w1", (size_t)b.rd * b.mid, &b.se_w1));
                TRY(get_tensor(h, p + "se.b1", b.rd, &b.se_b1));
                TRY(get_tensor(h, p + "se.w2t", (size_t)b.rd * b.mid, &b.se_w2t));
                TRY(get_tensor(h, p + "se.b2", b.mid, &b.se_b2));
                TRY(get_tensor(h, p + "pwl.w", (size_t)b.cout * b.mid, &b.wpwl));
                TRY(get_tensor(h, p + "pwl.b", b.cout, &b.bpwl));
                TRY(get_tensor(h, p + "pwl.w32", (size_t)b.cout * b.mid, &b.w32pwl));
                TRY(get_tensor(h, p + "pwl.bm", (size_t)b.cout * 64, &b.bmpwl));
            }
            h->blocks.push_back(b);
            cin = b.cout;
        }
    }
    const int c3 = h->cfg.num_3d_features, mid = h->mid3d(), rd = h->rd3d(), pj = h->cfg.num_3d_stack_proj;
    TRY(get_tensor(h, "proj2d.w", (size_t)c3 * 192, &h->proj2d_w));
    TRY(get_tensor(h, "proj2d.b", c3, &h->proj2d_b));
    TRY(get_tensor(h, "proj2d.bm", (size_t)c3 * 64, &h->proj2d_bm));
    for (int i = 0; i < h->cfg.num_3d_blocks; ++i) {
        Block3d b;
        char pre[32];
        snprintf(pre, sizeof(pre), "c3d.%d.", i);
        const std::string p(pre);
        TRY(get_tensor(h, p + "pw.w", (size_t)mid * c3, &b.wpw));
        TRY(get_tensor(h, p + "pw.b", mid, &b.bpw));
        TRY(get_tensor(h, p + "pw.bm", (size_t)mid * 64, &b.bmpw));
        TRY(get_tensor(h, p + "dw.w", (size_t)27 * mid, &b.wdw));
        TRY(get_tensor(h, p + "dw.b", mid, &b.bdw));
        TRY(get_tensor(h, p + "se.w1", (size_t)rd * mid, &b.se_w1));
        TRY(get_tensor(h, p + "se.b1", rd, &b.se_b1));
        TRY(get_tensor(h, p + "se.w2t", (size_t)rd * mid, &b.se_w2t));
        TRY(get_tensor(h, p + "se.b2", mid, &b.se_b2));
        TRY(get_tensor(h, p + "pwl.w", (size_t)c3 * mid, &b.wpwl));
        TRY(get_tensor(h, p + "pwl.b", c3, &b.bpwl));
        TRY(get_tensor(h, p + "pwl.w32", (size_t)c3 * mid, &b.w32pwl));
        TRY(get_tensor(h, p + "pwl.bm", (size_t)c3 * 64, &b.bmpwl));
        h->blocks3d.push_back(b);
    }
    TRY(get_tensor(h, "proj3d.w", (size_t)pj * c3, &h->proj3d_w));
    TRY(get_tensor(h, "proj3d.b", pj, &h->proj3d_b));
    TRY(get_tensor(h, "proj3d.bm", (size_t)pj * 64, &h->proj3d_bm));
    const float* gp = nullptr;
    TRY(get_tensor(h, "gem.p", 1, &gp));
    CUDA_TRY(cudaMemcpy(&h->gem_p, gp, sizeof(float), cudaMemcpyDeviceToHost));
    if (!(h->gem_p > 0.f)) return fail(MDS_ERR_WEIGHTS, "gem.p must be > 0");
    TRY(get_tensor(h, "cls.w", (size_t)h->cfg.num_classes * pj * h->T(), &h->cls_w));
    TRY(get_tensor(h, "cls.b", h->cfg.num_classes, &h->cls_b));
    h->committed = true;
    return MDS_OK;
}

// --------------------------------------------------------------------------------------------------------------
// workspace
// --------------------------------------------------------------------------------------------------------------
struct Arena {
    char* base;
    size_t cap, off = 0;
    bool overflow = false;
    Arena(void* p, size_t c) : base(reinterpret_cast<char*>(p)), cap(c) {}
    template <typename T>
    T* take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
        if (off + bytes > cap) { overflow = true; return nullptr; }
        T* r = reinterpret_cast<T*>(base + off);
        off += bytes;
        return r;
    }
};
static size_t al256(size_t b) { return (b + 255) & ~size_t(255); }

struct Sizes2d {
    size_t stream_elems, mid1_elems, mid2_elems;   // per image
};
static Sizes2d sizes2d(int H, int W) {
    Sizes2d s{0, 0, 0};
    int h = H / 2, w = W / 2;
    auto upd = [](size_t& a, size_t b) { if (b > a) a = b; };
    upd(s.stream_elems, (size_t)h * w * 32);
    int cin = 32;
    for (int si = 0; si < 6; ++si) {
        const int* d = kStageDefs[si];
        for (int bi = 0; bi < d[1]; ++bi) {
            int stride = bi == 0 ? d[2] : 1, mid = cin * d[3], cout = d[4];
            int ho = h / stride, wo = w / stride;
            if (d[0] == 2) {
                upd(s.mid1_elems, (size_t)h * w * mid);
                upd(s.mid2_elems, (size_t)ho * wo * mid);
            }
            upd(s.stream_elems, (size_t)ho * wo * cout);
            h = ho; w = wo; cin = cout;
        }
    }
    return s;
}

constexpr size_t kMaxGatedW = 192 * 1152;   // largest encoder projection (blocks.5.x conv_pwl)
static size_t ws2d_bytes(const MdsHandle* h, int H, int W, int n_images) {
    int cs = n_images < h->cfg.chunk_images ? n_images : h->cfg.chunk_images;
    if (cs <= 0) cs = 1;
    Sizes2d s = sizes2d(H, W);
    return 2 * al256(s.stream_elems * cs * 2) + al256(s.mid1_elems * cs * 2) + al256(s.mid2_elems * cs * 2) +
           al256((size_t)cs * kDwMaxParts * 1152 * 4) + al256((size_t)cs * 1152 * 2) + al256((size_t)cs * 1152 * 4) +
           al256((size_t)cs * kMaxGatedW * 2);
}
static size_t ws3d_bytes(const MdsHandle* h, int b, int P) {
    const size_t rows = (size_t)b * h->T() * P;
    return 2 * al256(rows * h->cfg.num_3d_features * 2) + 2 * al256(rows * h->mid3d() * 2) +
           al256((size_t)b * kDwMaxParts * h->mid3d() * 4) + al256((size_t)b * h->mid3d() * 2) + al256((size_t)b * h->mid3d() * 4) +
           al256((size_t)b * h->cfg.num_3d_features * h->mid3d() * 2);
}
static size_t wshead_bytes(const MdsHandle* h, int b) {
    return al256((size_t)b * h->cfg.num_3d_stack_proj * h->T() * 4) + al256(gem_part_floats(b, h->T(), h->cfg.num_3d_stack_proj) * 4);
}

extern "C" size_t mds_workspace_bytes(const MdsHandle* h, int H, int W, int n_images, int n_stacks) {
    if (!h) return 0;
    const int P = (H / 32) * (W / 32);
    const size_t rows = (size_t)n_stacks * h->T() * P;
    size_t w2d = ws2d_bytes(h, H, W, n_images);
    for (int S = 2; S <= kMaxStreams && S <= n_images; ++S) {        // multi-stream encoder: each part of the images has its own scratch
        size_t parts = 0;
        for (int k = 0; k < S; ++k) parts += ws2d_bytes(h, H, W, n_images / S + (k < n_images % S ? 1 : 0));
        if (parts > w2d) w2d = parts;
    }
    size_t full = w2d + al256(rows * 192 * 2) + ws3d_bytes(h, n_stacks, P) +
                  al256(rows * h->cfg.num_3d_stack_proj * 2) + wshead_bytes(h, n_stacks);
    return full + 4096;
}

// --------------------------------------------------------------------------------------------------------------
// forward passes
// --------------------------------------------------------------------------------------------------------------
static int check_ready(MdsHandle* h) {
    if (!h) return fail(MDS_ERR_INVALID, "null handle");
    if (!h->committed) return fail(MDS_ERR_WEIGHTS, "weights not committed");
    return MDS_OK;
}

static int forward_2d_impl(MdsHandle* h, const MdsFrames& fr, int n_images, __half* feats_out, Arena& ar, cudaStream_t st,
                           bool project = true) {
    if (fr.H % 32 || fr.W % 32 || fr.H <= 0 || fr.W <= 0)
        return fail(MDS_ERR_INVALID, "forward_2d: H (%d) and W (%d) must be positive multiples of 32", fr.H, fr.W);
    if (fr.stored_h + fr.pad_top > fr.H || fr.pad_top < 0 || fr.stored_h <= 0)
        return fail(MDS_ERR_INVALID, "forward_2d: frame rows (%d + pad %d) exceed H=%d", fr.stored_h, fr.pad_top, fr.H);
    if (n_images <= 0) return MDS_OK;
    const int cs_max = n_images < h->cfg.chunk_images ? n_images : h->cfg.chunk_images;
    const Sizes2d sz = sizes2d(fr.H, fr.W);
    __half* X[2];
    X[0] = ar.take<__half>(sz.stream_elems * cs_max);
    X[1] = ar.take<__half>(sz.stream_elems * cs_max);
    __half* M1 = ar.take<__half>(sz.mid1_elems * cs_max);
    __half* M2 = ar.take<__half>(sz.mid2_elems * cs_max);
    float* partials = ar.take<float>((size_t)cs_max * kDwMaxParts * 1152);
    __half* gate = ar.take<__half>((size_t)cs_max * 1152);
    float* gate32 = ar.take<float>((size_t)cs_max * 1152);
    __half* wg = ar.take<__half>((size_t)cs_max * kMaxGatedW);
    if (ar.overflow) return fail(MDS_ERR_WORKSPACE, "forward_2d: workspace too small");

    const int elem = fr.dtype == 0 ? 1 : 4;
    const int P = (fr.H / 32) * (fr.W / 32);

    for (int i0 = 0; i0 < n_images; i0 += cs_max) {
        const int cs = n_images - i0 < cs_max ? n_images - i0 : cs_max;
        MdsFrames f = fr;
        f.data = reinterpret_cast<const char*>(fr.data) + (size_t)i0 * fr.img_stride * elem;
        int hh = fr.H / 2, ww = fr.W / 2, cur = 0;
        g_prof_tag = 0;
        TRY(launch_stem(f, cs, h->stem_w, h->stem_b, X[0], st));
        for (const Block2d& b : h->blocks) {
            ++g_prof_tag;
            const int ho = hh / b.stride, wo = ww / b.stride;
            if (b.kind == 'c') {
                TRY(launch_conv3(X[cur], X[cur ^ 1], b.w1, b.b1, nullptr, nullptr, cs, hh, ww, b.cin, b.cout, b.stride, 0, 0, st));
            } else if (b.kind == 'e') {
                TRY(launch_conv3(X[cur], X[cur ^ 1], b.w1, b.b1, b.w2, b.b2, cs, hh, ww, b.cin, b.mid, b.stride, b.cout, b.skip, st, b.w1t));
            } else {
                TRY(launch_gemm(X[cur], b.wpw, b.bpw, nullptr, nullptr, M1, (long long)hh * ww, cs, b.mid, b.cin, 1, st, b.bmpw));
                if (g_tail_mode == 1) {
                    TailArgs a;
                    a.m1 = M1; a.m2 = M2; a.dw_w = b.wdw; a.dw_b = b.bdw; a.se_w1 = b.se_w1; a.se_b1 = b.se_b1; a.se_w2t = b.se_w2t;
                    a.se_b2 = b.se_b2; a.partials = partials; a.gate = gate32; a.sync = h->sync; a.sync_stride = kSyncSlots;
                    a.wpwl = b.wpwl; a.bm = b.bmpwl; a.res = b.skip ? X[cur] : nullptr; a.out = X[cur ^ 1];
                    a.n = cs; a.T = 1; a.H = hh; a.W = ww; a.C = b.mid; a.kt = 1; a.stride = b.stride; a.rd = b.rd; a.N = b.cout;
                    a.rows_per_chunk = 0;
                    TRY(launch_tail(a, st));
                } else if (g_tail_mode == 2) {
                    TRY(launch_dw_se(M1, M2, b.wdw, b.bdw, partials, nullptr, b.se_w1, b.se_b1, b.se_w2t, b.se_b2, gate32, h->sync,
                                     cs, 1, hh, ww, b.mid, 1, b.stride, b.rd, 0, st));
                    TRY(launch_gemm_tc_stream(M2, b.wpwl, b.bmpwl, b.skip ? X[cur] : nullptr, X[cur ^ 1], ho * wo, cs, b.cout, b.mid, 0, st,
                                              gate32));
                } else {
                    int nparts = 0;
                    if (g_tail_mode == 3)
                        TRY(launch_dw_se(M1, M2, b.wdw, b.bdw, partials, &nparts, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                         cs, 1, hh, ww, b.mid, 1, b.stride, b.rd, 0, st));
                    else
                    TRY(launch_dw(M1, M2, b.wdw, b.bdw, partials, &nparts, cs, 1, hh, ww, b.mid, 1, b.stride, st));
                    TRY(launch_se(partials, nparts, b.se_w1, b.se_b1, b.se_w2t, b.se_b2, gate, b.w32pwl, wg, cs, b.mid, b.rd, b.cout,
                                  1.0f / (float)(ho * wo), st));
                    TRY(launch_gemm_tc_stream(M2, wg, b.bmpwl, b.skip ? X[cur] : nullptr, X[cur ^ 1], ho * wo, cs, b.cout, b.mid, 0, st));
                }
            }
            cur ^= 1; hh = ho; ww = wo;
        }
        ++g_prof_tag;
        if (project) {
            TRY(launch_gemm(X[cur], h->proj2d_w, h->proj2d_b, nullptr, nullptr, feats_out + (size_t)i0 * P * 192,
                            (long long)P, cs, h->cfg.num_3d_features, 192, 1, st, h->proj2d_bm));
        } else {      // frozen-encoder training: hand out the encoder output itself (input of conv2d_projection)
            CUDA_TRY(cudaMemcpyAsync(feats_out + (size_t)i0 * P * 192, X[cur], (size_t)cs * P * 192 * sizeof(__half),
                                     cudaMemcpyDeviceToDevice, st));
        }
    }
    return MDS_OK;
}

// forward_2d over one stream or, when mds_set_streams(S >= 2), as S equal parts of the images on S streams, each with its own
// scratch: the images are independent, and the kernels of one part fill the ramps and tails of the others (measured +1.8 % at
// batch 4, +1.1 % at batch 32 for S = 2; 3 and 4 are slower).  Results are bit-identical: every kernel is batch-invariant.
static int forward_2d_streams(MdsHandle* h, const MdsFrames& frames, int n_images, __half* feats, Arena& ar, cudaStream_t st) {
    const int S = g_streams < n_images ? g_streams : n_images;
    // long buffers (the sweep: thousands of images in chunk_images passes) keep the machine full on one stream: measured 2 % slower there
    if (S < 2 || n_images > 2 * h->cfg.chunk_images || !(g_tail_mode == 0 || g_tail_mode == 3) || g_prof_on)
        return forward_2d_impl(h, frames, n_images, feats, ar, st);
    const int P = (frames.H / 32) * (frames.W / 32);
    CUDA_TRY(cudaEventRecord(h->ev_fork, st));
    size_t off = ar.off;
    int i0 = 0;
    for (int k = 0; k < S; ++k) {
        const int nk = n_images / S + (k < n_images % S ? 1 : 0);
        const size_t wk = ws2d_bytes(h, frames.H, frames.W, nk);
        if (off + wk > ar.cap) return fail(MDS_ERR_WORKSPACE, "forward_2d: workspace too small for the %d-stream encoder", S);
        Arena ak(ar.base + off, wk);
        off += wk;
        MdsFrames fk = frames;
        fk.data = reinterpret_cast<const char*>(frames.data) + (size_t)i0 * frames.img_stride * (frames.dtype == 0 ? 1 : 4);
        cudaStream_t sk = k == 0 ? st : h->aux_stream[k - 1];
        if (k > 0) CUDA_TRY(cudaStreamWaitEvent(sk, h->ev_fork, 0));
        TRY(forward_2d_impl(h, fk, nk, feats + (size_t)i0 * P * 192, ak, sk));
        if (k > 0) {
            CUDA_TRY(cudaEventRecord(h->ev_join[k - 1], sk));
            CUDA_TRY(cudaStreamWaitEvent(st, h->ev_join[k - 1], 0));
        }
        i0 += nk;
    }
    return MDS_OK;
}

static int forward_3d_impl(MdsHandle* h, const __half* feats, int b, int fh, int fw, __half* out, Arena& ar, cudaStream_t st) {
    if (b <= 0) return MDS_OK;
    if (fh <= 0 || fw <= 0) return fail(MDS_ERR_INVALID, "forward_3d: bad feature map size %dx%d", fh, fw);
    const int P = fh * fw;
    const int T = h->T(), c3 = h->cfg.num_3d_features, mid = h->mid3d(), rd = h->rd3d();
    const size_t rows = (size_t)b * T * P;
    __half* Y[2];
    Y[0] = ar.take<__half>(rows * c3);
    Y[1] = ar.take<__half>(rows * c3);
    __half* M1 = ar.take<__half>(rows * mid);
    __half* M2 = ar.take<__half>(rows * mid);
    float* partials = ar.take<float>((size_t)b * kDwMaxParts * mid);
    __half* gate = ar.take<__half>((size_t)b * mid);
    float* gate32 = ar.take<float>((size_t)b * mid);
    __half* wg = ar.take<__half>((size_t)b * c3 * mid);
    if (ar.overflow) return fail(MDS_ERR_WORKSPACE, "forward_3d: workspace too small");

    const __half* x = feats;
    int nxt = 0;
    // the 2D features are [b][T][h][w][C] already, i.e. the reference's transpose(1,2) is a no-op in NDHWC
    // (multidim_stacker.py:224,226)
    g_prof_tag = 100;
    for (const Block3d& blk : h->blocks3d) {
        ++g_prof_tag;
        TRY(launch_gemm(x, blk.wpw, blk.bpw, nullptr, nullptr, M1, (long long)T * P, b, mid, c3, 1, st, blk.bmpw));
        if (g_tail_mode == 1 && b <= kSyncSlots) {
            TailArgs a;
            a.m1 = M1; a.m2 = M2; a.dw_w = blk.wdw; a.dw_b = blk.bdw; a.se_w1 = blk.se_w1; a.se_b1 = blk.se_b1; a.se_w2t = blk.se_w2t;
            a.se_b2 = blk.se_b2; a.partials = partials; a.gate = gate32; a.sync = h->sync; a.sync_stride = kSyncSlots;
            a.wpwl = blk.wpwl; a.bm = blk.bmpwl; a.res = x; a.out = Y[nxt];
            a.n = b; a.T = T; a.H = fh; a.W = fw; a.C = mid; a.kt = 3; a.stride = 1; a.rd = rd; a.N = c3;
            a.rows_per_chunk = 0;
            TRY(launch_tail(a, st));
        } else if (g_tail_mode == 2 && b <= kSyncSlots) {
            TRY(launch_dw_se(M1, M2, blk.wdw, blk.bdw, partials, nullptr, blk.se_w1, blk.se_b1, blk.se_w2t, blk.se_b2, gate32, h->sync,
                             b, T, fh, fw, mid, 3, 1, rd, 0, st));
            TRY(launch_gemm_tc_stream(M2, blk.wpwl, blk.bmpwl, x, Y[nxt], T * P, b, c3, mid, 0, st, gate32));
        } else {
            int nparts = 0;
            if (g_tail_mode == 3)
                TRY(launch_dw_se(M1, M2, blk.wdw, blk.bdw, partials, &nparts, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                 b, T, fh, fw, mid, 3, 1, rd, 0, st));
            else
            TRY(launch_dw(M1, M2, blk.wdw, blk.bdw, partials, &nparts, b, T, fh, fw, mid, 3, 1, st));
            TRY(launch_se(partials, nparts, blk.se_w1, blk.se_b1, blk.se_w2t, blk.se_b2, gate, blk.w32pwl, wg, b, mid, rd, c3,
                          1.0f / (float)((size_t)T * P), st));
            TRY(launch_gemm_tc_stream(M2, wg, blk.bmpwl, x, Y[nxt], T * P, b, c3, mid, 0, st));
        }
        x = Y[nxt];
        nxt ^= 1;
    }
    g_prof_tag = 150;
    TRY(launch_gemm(x, h->proj3d_w, h->proj3d_b, nullptr, nullptr, out, (long long)T * P, b, h->cfg.num_3d_stack_proj, c3, 1, st, h->proj3d_bm));
    return MDS_OK;
}

static int forward_head_impl(MdsHandle* h, const __half* x, int b, int P, float* logits, int sig, Arena& ar, cudaStream_t st) {
    if (b <= 0) return MDS_OK;
    const int T = h->T(), pj = h->cfg.num_3d_stack_proj;
    float* feat = ar.take<float>((size_t)b * T * pj);
    float* part = ar.take<float>(gem_part_floats(b, T, pj));
    if (ar.overflow) return fail(MDS_ERR_WORKSPACE, "forward_head: workspace too small");
    g_prof_tag = 200;
    TRY(launch_gem(x, feat, part, b, T, P, pj, h->gem_p, 1e-6f, st));
    TRY(launch_linear(feat, h->cls_w, h->cls_b, logits, b, T * pj, h->cfg.num_classes, sig, st));
    return MDS_OK;
}

extern "C" int mds_forward_2d(MdsHandle* h, const MdsFrames* frames, int n_images, void* feats_out, void* ws,
                              size_t ws_bytes, void* stream) {
    TRY(check_ready(h));
    if (!frames || !frames->data || !feats_out || !ws) return fail(MDS_ERR_INVALID, "forward_2d: null argument");
    DeviceGuard g(h->cfg.device);
    Arena ar(ws, ws_bytes);
    return forward_2d_streams(h, *frames, n_images, reinterpret_cast<__half*>(feats_out), ar, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mds_forward_encoder(MdsHandle* h, const MdsFrames* frames, int n_images, void* feats_out, void* ws,
                                   size_t ws_bytes, void* stream) {
    TRY(check_ready(h));
    if (!frames || !frames->data || !feats_out || !ws) return fail(MDS_ERR_INVALID, "forward_encoder: null argument");
    DeviceGuard g(h->cfg.device);
    Arena ar(ws, ws_bytes);
    return forward_2d_impl(h, *frames, n_images, reinterpret_cast<__half*>(feats_out), ar, reinterpret_cast<cudaStream_t>(stream), false);
}

extern "C" int mds_forward_3d(MdsHandle* h, const void* feats, int b, int fh, int fw, void* out, void* ws, size_t ws_bytes,
                              void* stream) {
    TRY(check_ready(h));
    if (!feats || !out || !ws) return fail(MDS_ERR_INVALID, "forward_3d: null argument");
    DeviceGuard g(h->cfg.device);
    Arena ar(ws, ws_bytes);
    return forward_3d_impl(h, reinterpret_cast<const __half*>(feats), b, fh, fw, reinterpret_cast<__half*>(out), ar,
                           reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mds_forward_head(MdsHandle* h, const void* x, int b, int P, float* logits, int apply_sigmoid, void* ws,
                                size_t ws_bytes, void* stream) {
    TRY(check_ready(h));
    if (!x || !logits || !ws) return fail(MDS_ERR_INVALID, "forward_head: null argument");
    DeviceGuard g(h->cfg.device);
    Arena ar(ws, ws_bytes);
    return forward_head_impl(h, reinterpret_cast<const __half*>(x), b, P, logits, apply_sigmoid, ar,
                             reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mds_forward(MdsHandle* h, const MdsFrames* frames, int b, float* logits, int apply_sigmoid, void* ws,
                           size_t ws_bytes, void* stream) {
    TRY(check_ready(h));
    if (!frames || !frames->data || !logits || !ws) return fail(MDS_ERR_INVALID, "forward: null argument");
    DeviceGuard g(h->cfg.device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Arena ar(ws, ws_bytes);
    const int T = h->T();
    const int h3 = frames->H / 32, w3 = frames->W / 32, P = h3 * w3;
    const size_t rows = (size_t)b * T * P;
    __half* feats = ar.take<__half>(rows * 192);
    __half* out3d = ar.take<__half>(rows * h->cfg.num_3d_stack_proj);
    if (ar.overflow) return fail(MDS_ERR_WORKSPACE, "forward: workspace too small");
    TRY(forward_2d_streams(h, *frames, b * T, feats, ar, st));
    // the 2D scratch is dead now: re-use the arena from the same mark for the 3D stage
    Arena ar3(ar.base + al256(rows * 192 * 2) + al256(rows * h->cfg.num_3d_stack_proj * 2),
              ws_bytes - al256(rows * 192 * 2) - al256(rows * h->cfg.num_3d_stack_proj * 2));
    TRY(forward_3d_impl(h, feats, b, h3, w3, out3d, ar3, st));
    TRY(forward_head_impl(h, out3d, b, P, logits, apply_sigmoid, ar3, st));
    return MDS_OK;
}

// --------------------------------------------------------------------------------------------------------------
// converters + per-kernel entry points
// --------------------------------------------------------------------------------------------------------------
extern "C" int mds_nchw32_to_nhwc16(const float* src, void* dst, int n, int C, int P, void* stream) {
    if (n <= 0) return MDS_OK;
    if (n > 65535) return fail(MDS_ERR_INVALID, "converter: n too large");
    dim3 grid((P + 31) / 32, (C + 31) / 32, n);
    nchw32_to_nhwc16_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, reinterpret_cast<__half*>(dst), C, P);
    LAUNCH_CHECK("nchw32_to_nhwc16");
    return MDS_OK;
}
extern "C" int mds_nhwc16_to_nchw32(const void* src, float* dst, int n, int C, int P, void* stream) {
    if (n <= 0) return MDS_OK;
    if (n > 65535) return fail(MDS_ERR_INVALID, "converter: n too large");
    dim3 grid((P + 31) / 32, (C + 31) / 32, n);
    nhwc16_to_nchw32_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const __half*>(src), dst, C, P);
    LAUNCH_CHECK("nhwc16_to_nchw32");
    return MDS_OK;
}

extern "C" int mds_gather_stacks(const void* feats, void* out, long long first_image, int hop, int n_pred, int T, long long plane_elems,
                                 void* stream) {
    if (!feats || !out) return fail(MDS_ERR_INVALID, "gather_stacks: null argument");
    if (n_pred <= 0 || T <= 0) return MDS_OK;
    if (plane_elems <= 0 || plane_elems % 8 || first_image < 0 || hop < 0 || n_pred > 65535)
        return fail(MDS_ERR_INVALID, "gather_stacks: plane must be a multiple of 8 fp16 elements, n_pred <= 65535");
    gather_stacks_kernel<<<dim3(T, n_pred), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const uint4*>(feats), reinterpret_cast<uint4*>(out), first_image, hop, T, plane_elems / 8);
    LAUNCH_CHECK("gather_stacks");
    return MDS_OK;
}
extern "C" int mds_axpby(float* y, const float* x, float a, float b, long long n, void* stream) {
    if (!y || !x) return fail(MDS_ERR_INVALID, "axpby: null argument");
    if (n <= 0) return MDS_OK;
    axpby_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(y, x, a, b, n);
    LAUNCH_CHECK("axpby");
    return MDS_OK;
}

extern "C" int mds_k_stem(const MdsFrames* frames, int n_images, const void* wh, const float* bias, void* out, void* stream) {
    if (!frames) return fail(MDS_ERR_INVALID, "null frames");
    return launch_stem(*frames, n_images, reinterpret_cast<const __half*>(wh), bias, reinterpret_cast<__half*>(out),
                       reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int mds_k_conv3x3(const void* in, void* out, const void* w1, const float* b1, const void* w2, const float* b2,
                             int n, int H, int W, int cin, int cmid, int stride, int cproj, int res, void* stream) {
    return launch_conv3(reinterpret_cast<const __half*>(in), reinterpret_cast<__half*>(out), reinterpret_cast<const __half*>(w1),
                        b1, reinterpret_cast<const __half*>(w2), b2, n, H, W, cin, cmid, stride, cproj, res,
                        reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int mds_k_gemm1x1(const void* A, const void* W, const float* bias, const void* bias_mat, const void* res,
                             const void* gate, void* C, int rows_per_img, int n_img, int N, int K, int act, void* stream) {
    return launch_gemm(reinterpret_cast<const __half*>(A), reinterpret_cast<const __half*>(W), bias,
                       reinterpret_cast<const __half*>(res), reinterpret_cast<const __half*>(gate),
                       reinterpret_cast<__half*>(C), rows_per_img, n_img, N, K, act, reinterpret_cast<cudaStream_t>(stream),
                       reinterpret_cast<const __half*>(bias_mat), true);
}
extern "C" int mds_k_dwconv(const void* in, void* out, const float* w, const float* bias, float* partials, int* nparts, int n,
                            int T, int H, int W, int C, int kt, int stride, void* stream) {
    return launch_dw(reinterpret_cast<const __half*>(in), reinterpret_cast<__half*>(out), w, bias, partials, nparts, n, T, H, W, C,
                     kt, stride, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int mds_k_se_fc(const float* partials, int nparts, const float* w1, const float* b1, const float* w2t, const float* b2,
                           void* gate, const float* w32, void* wg, int n, int C, int rd, int N, float inv_count, void* stream) {
    return launch_se(partials, nparts, w1, b1, w2t, b2, reinterpret_cast<__half*>(gate), w32, reinterpret_cast<__half*>(wg), n, C,
                     rd, N, inv_count, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int mds_k_gemm_gated(const void* A, const void* wg, const void* bias_mat, const void* res, void* C, int rows_per_img,
                                int n_img, int N, int K, int act, void* stream) {
    return launch_gemm_tc_stream(reinterpret_cast<const __half*>(A), reinterpret_cast<const __half*>(wg),
                                 reinterpret_cast<const __half*>(bias_mat), reinterpret_cast<const __half*>(res),
                                 reinterpret_cast<__half*>(C), rows_per_img, n_img, N, K, act, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int mds_k_dwconv_se(const void* in, void* out, const float* w, const float* bias, float* partials, int* nparts,
                               const float* se_w1, const float* se_b1, const float* se_w2t, const float* se_b2, float* gate, int* done,
                               int n, int T, int H, int W, int C, int kt, int stride, int rd, int rows_per_chunk, void* stream) {
    if (!in || !out || !w || !bias || !partials) return fail(MDS_ERR_INVALID, "dwconv_se: null argument");
    return launch_dw_se(reinterpret_cast<const __half*>(in), reinterpret_cast<__half*>(out), w, bias, partials, nparts, se_w1, se_b1,
                        se_w2t, se_b2, gate, done, n, T, H, W, C, kt, stride, rd, rows_per_chunk, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int mds_k_gemm_gate(const void* A, const void* W, const float* gate, const void* bias_mat, const void* res, void* C,
                               int rows_per_img, int n_img, int N, int K, int act, void* stream) {
    if (!gate) return fail(MDS_ERR_INVALID, "gemm_gate: null gate");
    return launch_gemm_tc_stream(reinterpret_cast<const __half*>(A), reinterpret_cast<const __half*>(W),
                                 reinterpret_cast<const __half*>(bias_mat), reinterpret_cast<const __half*>(res),
                                 reinterpret_cast<__half*>(C), rows_per_img, n_img, N, K, act, reinterpret_cast<cudaStream_t>(stream), gate);
}
extern "C" int mds_k_mbconv_tail(const void* m1, void* m2, const float* dw_w, const float* dw_b, float* partials, const float* se_w1,
                                 const float* se_b1, const float* se_w2t, const float* se_b2, float* gate, int* sync, const void* wpwl,
                                 const void* bias_mat, const void* res, void* out, int n, int T, int H, int W, int C, int kt, int stride,
                                 int rd, int N, int rows_per_chunk, void* stream) {
    if (!m1 || !m2 || !dw_w || !dw_b || !partials || !se_w1 || !se_b1 || !se_w2t || !se_b2 || !gate || !sync)
        return fail(MDS_ERR_INVALID, "mbconv_tail: null argument");
    if (N && (!wpwl || !bias_mat || !out)) return fail(MDS_ERR_INVALID, "mbconv_tail: null projection argument");
    TailArgs a;
    a.m1 = reinterpret_cast<const __half*>(m1); a.m2 = reinterpret_cast<__half*>(m2); a.dw_w = dw_w; a.dw_b = dw_b;
    a.se_w1 = se_w1; a.se_b1 = se_b1; a.se_w2t = se_w2t; a.se_b2 = se_b2; a.partials = partials; a.gate = gate;
    a.sync = sync; a.sync_stride = n; a.wpwl = reinterpret_cast<const __half*>(wpwl); a.bm = reinterpret_cast<const __half*>(bias_mat);
    a.res = reinterpret_cast<const __half*>(res); a.out = reinterpret_cast<__half*>(out);
    a.n = n; a.T = T; a.H = H; a.W = W; a.C = C; a.kt = kt; a.stride = stride; a.rd = rd; a.N = N;
    a.rows_per_chunk = rows_per_chunk;
    return launch_tail(a, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" int mds_k_gem(const void* x, float* feat, int b, int T, int P, int C, float p, float eps, void* stream) {
    // per-kernel entry point (tests): the slice sums need scratch, allocated here stream-ordered
    float* part = nullptr;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (b <= 0) return MDS_OK;
    CUDA_TRY(cudaMallocAsync(&part, gem_part_floats(b, T, C) * sizeof(float), st));
    const int rc = launch_gem(reinterpret_cast<const __half*>(x), feat, part, b, T, P, C, p, eps, st);
    CUDA_TRY(cudaFreeAsync(part, st));
    return rc;
}
extern "C" int mds_k_linear(const float* feat, const float* w, const float* bias, float* out, int b, int F, int num_classes,
                            int apply_sigmoid, void* stream) {
    return launch_linear(feat, w, bias, out, b, F, num_classes, apply_sigmoid, reinterpret_cast<cudaStream_t>(stream));
}

#include "train_api.inl"
#include "postproc_api.inl"
