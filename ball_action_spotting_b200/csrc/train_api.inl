// Host side of the frozen-encoder training step (included at the end of mds_api.cu; see include/mds_b200.h).
// Reference: BallActionModel.train_step (src/argus_models.py:41-74) with freeze_conv2d_encoder (:104-110).

// --------------------------------------------------------------------------------------------------------------
// trainer handle: fp32 master parameters, gradients, momentum, BatchNorm running statistics, fp16 GEMM operands
// --------------------------------------------------------------------------------------------------------------
struct TrainTensor {
    std::string name;
    size_t off, numel;
};
struct TrainBn {
    int C;
    size_t gamma, beta;          // offsets into the parameter buffer
    size_t rm, rv;               // offsets into the statistics buffer
    float* scratch;              // scale, shift, mean, rstd, c1, c2, gr: 7 x C floats
    float* scale() const { return scratch; }
    float* shift() const { return scratch + C; }
    float* mean() const { return scratch + 2 * C; }
    float* rstd() const { return scratch + 3 * C; }
    float* c1() const { return scratch + 4 * C; }
    float* c2() const { return scratch + 5 * C; }
    float* gr() const { return scratch + 6 * C; }
};
struct TrainGemmW {
    size_t off;                  // fp32 master [N][K]
    int N, K;
    __half *w16, *w16t;          // [N][K] and [K][N]
};
struct TrainBlock {
    TrainGemmW pw, pwl;
    TrainBn bn1, bn2, bn3;
    size_t dw, se_w1, se_b1, se_w2, se_b2;
    float *dw27, *dw27f;         // tap-major [27][mid] copy of conv_dw.weight and its mirror (data gradient)
};

struct MdsTrainer {
    MdsTrainConfig cfg;
    std::vector<TrainTensor> params, buffers;
    std::map<std::string, int> pindex, bindex;
    size_t n_params = 0, n_stats = 0;
    float *Pema = nullptr, *Sema = nullptr;      // ModelEma copies of the parameters / BatchNorm statistics (src/ema.py)
    float *P = nullptr, *G = nullptr, *Mom = nullptr, *S = nullptr, *scaler = nullptr, *bn_scratch = nullptr, *zeros = nullptr, *dw27 = nullptr;
    __half* w16 = nullptr;
    CastJob* cast_jobs = nullptr;   // device table of the fp16 operand refreshes (one launch)
    int n_cast = 0, cast_gx = 0, cast_gy = 0;
    __half* zeros16 = nullptr;   // [1152][64] zero bias matrix of the tcgen05 GEMMs (no conv in this path has a bias)
    TrainGemmW proj2d, proj3d;
    TrainBn bn_p2d, bn_p3d;
    std::vector<TrainBlock> blocks;
    size_t gem_p = 0, cls_w = 0, cls_b = 0;
    long long batches_tracked = 0;
    bool committed = false;
    int T() const { return cfg.num_frames / cfg.stack_size; }
    int mid() const { return cfg.num_3d_features * cfg.expansion_3d_ratio; }
    int rd() const { return mid() / cfg.se_reduce_3d_ratio; }
    int F() const { return cfg.num_3d_stack_proj * T(); }
};

static size_t train_add(std::vector<TrainTensor>& v, std::map<std::string, int>& idx, size_t& total, const std::string& name, size_t numel) {
    const size_t off = total;
    idx[name] = (int)v.size();
    v.push_back({name, off, numel});
    total += (numel + 3) & ~size_t(3);      // 16-byte aligned slices
    return off;
}

extern "C" int mds_train_create(const MdsTrainConfig* cfg, MdsTrainer** out) {
    if (!cfg || !out) return fail(MDS_ERR_INVALID, "mds_train_create: null argument");
    if (cfg->stack_size != 3 || cfg->num_frames <= 0 || cfg->num_frames % 3) return fail(MDS_ERR_INVALID, "train: num_frames must be a positive multiple of stack_size = 3");
    if (cfg->num_3d_features != 192) return fail(MDS_ERR_INVALID, "train: num_3d_features must be 192");
    if (cfg->num_3d_stack_proj % 64 || cfg->num_3d_stack_proj > 256 || cfg->num_3d_stack_proj <= 0)
        return fail(MDS_ERR_INVALID, "train: num_3d_stack_proj must be a multiple of 64, <= 256");
    const int mid = cfg->num_3d_features * cfg->expansion_3d_ratio;
    if (mid % 64 || mid > 1152 || cfg->se_reduce_3d_ratio <= 0 || mid / cfg->se_reduce_3d_ratio <= 0 || mid / cfg->se_reduce_3d_ratio > 256)
        return fail(MDS_ERR_INVALID, "train: unsupported 3D expansion / SE ratio");
    if (cfg->num_classes <= 0 || cfg->num_classes > 64 || cfg->num_3d_blocks < 0) return fail(MDS_ERR_INVALID, "train: bad num_classes / num_3d_blocks");
    if (cfg->drop_rate < 0.f || cfg->drop_rate >= 1.f || cfg->drop_path_rate < 0.f || cfg->drop_path_rate >= 1.f)
        return fail(MDS_ERR_INVALID, "train: drop rates must be in [0, 1)");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(MDS_ERR_INVALID, "device %d out of range", cfg->device);
    DeviceGuard g(cfg->device);
    MdsTrainer* t = new MdsTrainer();
    t->cfg = *cfg;
    const int c3 = cfg->num_3d_features, rd = mid / cfg->se_reduce_3d_ratio, pj = cfg->num_3d_stack_proj;
    size_t bn_floats = 0, w16_halves = 0;
    auto bn = [&](TrainBn& b, const std::string& prefix, int C) {
        b.C = C;
        b.gamma = train_add(t->params, t->pindex, t->n_params, prefix + ".weight", C);
        b.beta = train_add(t->params, t->pindex, t->n_params, prefix + ".bias", C);
        b.rm = train_add(t->buffers, t->bindex, t->n_stats, prefix + ".running_mean", C);
        b.rv = train_add(t->buffers, t->bindex, t->n_stats, prefix + ".running_var", C);
        b.scratch = reinterpret_cast<float*>(bn_floats);      // offset for now, rebased after allocation
        bn_floats += 7 * (size_t)C;
    };
    auto gw = [&](TrainGemmW& w, const std::string& name, int N, int K) {
        w.N = N; w.K = K;
        w.off = train_add(t->params, t->pindex, t->n_params, name, (size_t)N * K);
        w.w16 = reinterpret_cast<__half*>(w16_halves);
        w.w16t = reinterpret_cast<__half*>(w16_halves + (size_t)N * K);
        w16_halves += 2 * (size_t)N * K;
    };
    // order = the reference's parameter order for the unfrozen part (multidim_stacker.py:178-208)
    gw(t->proj2d, "conv2d_projection.0.weight", c3, 192);
    bn(t->bn_p2d, "conv2d_projection.1", c3);
    for (int i = 0; i < cfg->num_3d_blocks; ++i) {
        TrainBlock b;
        const std::string p = "conv3d_encoder." + std::to_string(i) + ".";
        gw(b.pw, p + "conv_pw.weight", mid, c3);
        bn(b.bn1, p + "bn1.bn3d", mid);
        b.dw = train_add(t->params, t->pindex, t->n_params, p + "conv_dw.weight", (size_t)mid * 27);
        bn(b.bn2, p + "bn2.bn3d", mid);
        b.se_w1 = train_add(t->params, t->pindex, t->n_params, p + "se.conv_reduce.weight", (size_t)rd * mid);
        b.se_b1 = train_add(t->params, t->pindex, t->n_params, p + "se.conv_reduce.bias", rd);
        b.se_w2 = train_add(t->params, t->pindex, t->n_params, p + "se.conv_expand.weight", (size_t)mid * rd);
        b.se_b2 = train_add(t->params, t->pindex, t->n_params, p + "se.conv_expand.bias", mid);
        gw(b.pwl, p + "conv_pwl.weight", c3, mid);
        bn(b.bn3, p + "bn3.bn3d", c3);
        t->blocks.push_back(b);
    }
    gw(t->proj3d, "conv3d_projection.0.weight", pj, c3);
    bn(t->bn_p3d, "conv3d_projection.1", pj);
    t->gem_p = train_add(t->params, t->pindex, t->n_params, "global_pool.p", 1);
    t->cls_w = train_add(t->params, t->pindex, t->n_params, "classifier.weight", (size_t)cfg->num_classes * pj * t->T());
    t->cls_b = train_add(t->params, t->pindex, t->n_params, "classifier.bias", cfg->num_classes);

    auto alloc = [&](void** p, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc(p, bytes);
        if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes);
        return e;
    };
    cudaError_t e = alloc((void**)&t->P, t->n_params * 4);
    if (e == cudaSuccess) e = alloc((void**)&t->G, t->n_params * 4);
    if (e == cudaSuccess) e = alloc((void**)&t->Mom, t->n_params * 4);
    if (e == cudaSuccess) e = alloc((void**)&t->S, t->n_stats * 4);
    if (e == cudaSuccess) e = alloc((void**)&t->Pema, t->n_params * 4);
    if (e == cudaSuccess) e = alloc((void**)&t->Sema, t->n_stats * 4);
    if (e == cudaSuccess) e = alloc((void**)&t->scaler, 16);
    if (e == cudaSuccess) e = alloc((void**)&t->bn_scratch, bn_floats * 4);
    if (e == cudaSuccess) e = alloc((void**)&t->zeros, 1152 * 4);
    if (e == cudaSuccess) e = alloc((void**)&t->w16, w16_halves * 2);
    if (e == cudaSuccess) e = alloc((void**)&t->zeros16, 1152 * 64 * 2);
    if (e == cudaSuccess) e = alloc((void**)&t->dw27, (size_t)(cfg->num_3d_blocks > 0 ? cfg->num_3d_blocks : 1) * 2 * 27 * mid * 4);
    if (e != cudaSuccess) {
        mds_train_destroy(t);
        return fail(MDS_ERR_CUDA, "mds_train_create: %s", cudaGetErrorString(e));
    }
    auto rebase_bn = [&](TrainBn& b) { b.scratch = t->bn_scratch + reinterpret_cast<size_t>(b.scratch); };
    auto rebase_w = [&](TrainGemmW& w) {
        w.w16 = t->w16 + reinterpret_cast<size_t>(w.w16);
        w.w16t = t->w16 + reinterpret_cast<size_t>(w.w16t);
    };
    rebase_bn(t->bn_p2d); rebase_bn(t->bn_p3d); rebase_w(t->proj2d); rebase_w(t->proj3d);
    for (size_t i = 0; i < t->blocks.size(); ++i) {
        TrainBlock& b = t->blocks[i];
        rebase_bn(b.bn1); rebase_bn(b.bn2); rebase_bn(b.bn3); rebase_w(b.pw); rebase_w(b.pwl);
        b.dw27 = t->dw27 + i * 2 * 27 * (size_t)mid;
        b.dw27f = b.dw27 + 27 * (size_t)mid;
    }
    {   // table of fp16 operand refreshes
        std::vector<CastJob> jobs;
        auto add = [&](const TrainGemmW& w) {
            jobs.push_back({t->P + w.off, w.w16, w.w16t, w.N, w.K});
            if ((w.K + 31) / 32 > t->cast_gx) t->cast_gx = (w.K + 31) / 32;
            if ((w.N + 31) / 32 > t->cast_gy) t->cast_gy = (w.N + 31) / 32;
        };
        add(t->proj2d);
        for (auto& b : t->blocks) { add(b.pw); add(b.pwl); }
        add(t->proj3d);
        t->n_cast = (int)jobs.size();
        if (cudaMalloc((void**)&t->cast_jobs, jobs.size() * sizeof(CastJob)) != cudaSuccess ||
            cudaMemcpy(t->cast_jobs, jobs.data(), jobs.size() * sizeof(CastJob), cudaMemcpyHostToDevice) != cudaSuccess) {
            mds_train_destroy(t);
            return fail(MDS_ERR_CUDA, "mds_train_create: cast table");
        }
    }
    const float init[4] = {cfg->amp ? (cfg->init_scale > 0.f ? cfg->init_scale : 65536.0f) : 1.0f, 0.f, 0.f, 0.f};
    cudaMemcpy(t->scaler, init, sizeof(init), cudaMemcpyHostToDevice);
    *out = t;
    return MDS_OK;
}

extern "C" int mds_train_destroy(MdsTrainer* t) {
    if (!t) return MDS_OK;
    DeviceGuard g(t->cfg.device);
    cudaFree(t->P); cudaFree(t->G); cudaFree(t->Mom); cudaFree(t->S); cudaFree(t->Pema); cudaFree(t->Sema); cudaFree(t->scaler); cudaFree(t->bn_scratch);
    cudaFree(t->zeros); cudaFree(t->w16); cudaFree(t->dw27); cudaFree(t->zeros16); cudaFree(t->cast_jobs);
    delete t;
    return MDS_OK;
}

extern "C" int mds_train_num_tensors(const MdsTrainer* t, int kind) {
    if (!t) return 0;
    return (int)(kind == 0 ? t->params.size() : t->buffers.size());
}
extern "C" int mds_train_tensor_info(const MdsTrainer* t, int kind, int i, const char** name, long long* numel) {
    if (!t) return fail(MDS_ERR_INVALID, "null trainer");
    const auto& v = kind == 0 ? t->params : t->buffers;
    if (i < 0 || i >= (int)v.size()) return fail(MDS_ERR_INVALID, "tensor index %d out of range", i);
    if (name) *name = v[i].name.c_str();
    if (numel) *numel = (long long)v[i].numel;
    return MDS_OK;
}

static int train_find(MdsTrainer* t, const char* name, long long numel, bool* is_param, const TrainTensor** out) {
    if (!t || !name) return fail(MDS_ERR_INVALID, "train: null argument");
    auto ip = t->pindex.find(name);
    if (ip != t->pindex.end()) { *is_param = true; *out = &t->params[ip->second]; }
    else {
        auto ib = t->bindex.find(name);
        if (ib == t->bindex.end()) return fail(MDS_ERR_WEIGHTS, "train: unknown tensor '%s'", name);
        *is_param = false; *out = &t->buffers[ib->second];
    }
    if ((long long)(*out)->numel != numel) return fail(MDS_ERR_WEIGHTS, "train: tensor '%s' has %zu elements, got %lld", name, (*out)->numel, numel);
    return MDS_OK;
}

extern "C" int mds_train_set(MdsTrainer* t, const char* name, const float* host, long long numel) {
    bool is_param; const TrainTensor* tt;
    TRY(train_find(t, name, numel, &is_param, &tt));
    if (!host) return fail(MDS_ERR_INVALID, "train_set: null data");
    DeviceGuard g(t->cfg.device);
    CUDA_TRY(cudaMemcpy((is_param ? t->P : t->S) + tt->off, host, (size_t)numel * 4, cudaMemcpyHostToDevice));
    t->committed = false;
    return MDS_OK;
}

extern "C" int mds_train_get(MdsTrainer* t, const char* name, int what, float* host, long long numel) {
    bool is_param; const TrainTensor* tt;
    TRY(train_find(t, name, numel, &is_param, &tt));
    if (!host) return fail(MDS_ERR_INVALID, "train_get: null data");
    if (!is_param && what != 0 && what != 3) return fail(MDS_ERR_INVALID, "train_get: buffers have no gradient / momentum");
    if (what < 0 || what > 3) return fail(MDS_ERR_INVALID, "train_get: what must be 0..3");
    DeviceGuard g(t->cfg.device);
    CUDA_TRY(cudaDeviceSynchronize());
    const float* src = !is_param ? (what == 3 ? t->Sema : t->S) : what == 0 ? t->P : what == 1 ? t->G : what == 2 ? t->Mom : t->Pema;
    CUDA_TRY(cudaMemcpy(host, src + tt->off, (size_t)numel * 4, cudaMemcpyDeviceToHost));
    if (is_param && what == 1) {       // gradients are stored multiplied by the loss scale
        float sc[4];
        CUDA_TRY(cudaMemcpy(sc, t->scaler, sizeof(sc), cudaMemcpyDeviceToHost));
        const float inv = 1.0f / sc[0];
        for (long long i = 0; i < numel; ++i) host[i] *= inv;
    }
    return MDS_OK;
}

extern "C" int mds_train_scaler_state(MdsTrainer* t, float* host4) {
    if (!t || !host4) return fail(MDS_ERR_INVALID, "train_scaler_state: null argument");
    DeviceGuard g(t->cfg.device);
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpy(host4, t->scaler, 16, cudaMemcpyDeviceToHost));
    return MDS_OK;
}

static int train_derive(MdsTrainer* t, cudaStream_t st) {
    ProfScope ps(MDS_KIND_TRAIN_SMALL, st);
    launch_pdl(cast_transpose_kernel, dim3(t->cast_gx, t->cast_gy, t->n_cast), dim3(256), 0, st, t->cast_jobs);
    LAUNCH_CHECK("cast_transpose");
    for (auto& b : t->blocks) {
        launch_pdl(dw3_weights_kernel, dim3((27 * t->mid() + 255) / 256), dim3(256), 0, st, t->P + b.dw, b.dw27, b.dw27f, t->mid());
        LAUNCH_CHECK("dw3_weights");
    }
    return MDS_OK;
}

extern "C" int mds_train_commit(MdsTrainer* t, void* stream) {
    if (!t) return fail(MDS_ERR_INVALID, "null trainer");
    DeviceGuard g(t->cfg.device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    TRY(train_derive(t, st));
    // ModelEma(model) deep-copies the model it is given (src/ema.py:38-41): the averages start from the committed values
    CUDA_TRY(cudaMemcpyAsync(t->Pema, t->P, t->n_params * 4, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(t->Sema, t->S, t->n_stats * 4, cudaMemcpyDeviceToDevice, st));
    t->committed = true;
    return MDS_OK;
}

extern "C" int mds_train_ema_update(MdsTrainer* t, double decay, void* stream) {
    if (!t) return fail(MDS_ERR_INVALID, "null trainer");
    if (!t->committed) return fail(MDS_ERR_WEIGHTS, "train_ema_update: parameters not committed");
    if (!(decay >= 0.0 && decay <= 1.0)) return fail(MDS_ERR_INVALID, "train_ema_update: decay must be in [0, 1]");
    DeviceGuard g(t->cfg.device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // the reference evaluates `self.decay * e + (1. - self.decay) * m` with a Python (double) decay on float32 tensors:
    // both scalars are rounded to float32 separately
    const float decay32 = (float)decay, one_minus = (float)(1.0 - decay);
    ProfScope ps(MDS_KIND_TRAIN_SMALL, st);
    const int blocks = (int)((t->n_params + 255) / 256) < 4 * num_sms() ? (int)((t->n_params + 255) / 256) : 4 * num_sms();
    ema_update_kernel<<<blocks, 256, 0, st>>>(t->Pema, t->P, t->n_params, decay32, one_minus);
    LAUNCH_CHECK("ema_update");
    ema_update_kernel<<<(unsigned)((t->n_stats + 255) / 256), 256, 0, st>>>(t->Sema, t->S, t->n_stats, decay32, one_minus);
    LAUNCH_CHECK("ema_update");
    return MDS_OK;
}

// --------------------------------------------------------------------------------------------------------------
// launchers
// --------------------------------------------------------------------------------------------------------------
// C[M][N] = A[M][K] W[N][K]^T (+ res): forward and data-gradient GEMMs of the 1x1x1 convolutions.  tcgen05 path of the
// inference engine (gemm_tc.cuh): resident-A mode for K <= 192, streamed mode (one "image") for N <= 256; mma.sync otherwise.
static int train_gemm(MdsTrainer* t, const __half* A, const __half* W, const __half* res, __half* C, long long M, int N, int K, cudaStream_t st) {
    if (res == nullptr && K <= kTcMaxKB * kTcBK && tc_pick_bn(N) >= 32 && M < (1LL << 31))
        return launch_gemm(A, W, t->zeros, nullptr, nullptr, C, M, 1, N, K, 0, st, t->zeros16);
    if (N >= 32 && N <= 256 && N % 16 == 0 && K % 16 == 0 && M < (1LL << 31))
        return launch_gemm_tc_stream(A, W, t->zeros16, res, C, (int)M, 1, N, K, 0, st);
    return launch_gemm(A, W, t->zeros, res, nullptr, C, M, 1, N, K, 0, st, nullptr, true);      // mma.sync kernel, explicitly
}

// rows per CTA of the column kernels: about four CTAs per SM in total, so that the grid is whole waves at 2 CTAs / SM
static int train_chunk_rows(int rows_per_sample, int b) {
    int per_sample = (4 * num_sms()) / (b > 0 ? b : 1);
    if (per_sample < 1) per_sample = 1;
    int rows = (rows_per_sample + per_sample - 1) / per_sample;
    return rows < 16 ? 16 : rows;
}
static int train_chunks(int rows_per_sample, int b) {
    const int r = train_chunk_rows(rows_per_sample, b);
    return (rows_per_sample + r - 1) / r;
}

static EwParams ew_base(const __half* y, int C, int rows_per_sample, int b) {
    EwParams p;
    memset(&p, 0, sizeof(p));
    p.y = y; p.C = C; p.rows_per_sample = rows_per_sample; p.rows_per_chunk = train_chunk_rows(rows_per_sample, b);
    return p;
}
static dim3 ew_grid(int rows_per_sample, int b) { return dim3(train_chunks(rows_per_sample, b), b); }

// batch statistics partials -> scale / shift (+ running-stat update).  `ref` = the per-channel shift the partial sums
// were taken against (the running mean *before* this update: a cheap, data-independent guess of the batch mean that
// keeps the one-pass variance well conditioned); the finalize kernel reads it before overwriting the running mean.
static int train_bn_finalize(MdsTrainer* t, const TrainBn& bn, const float* partials, int nparts, double count, const float* ref, cudaStream_t st) {
    ProfScope ps(MDS_KIND_TRAIN_BN, st);
    BnFwdFin f;
    f.partials = partials; f.nparts = nparts; f.ref = ref;
    f.gamma = t->P + bn.gamma; f.beta = t->P + bn.beta;
    f.running_mean = t->S + bn.rm; f.running_var = t->S + bn.rv;
    f.scale = bn.scale(); f.shift = bn.shift(); f.mean = bn.mean(); f.rstd = bn.rstd();
    f.C = bn.C; f.count = (float)count; f.eps = 1e-5f; f.momentum = 0.1f;
    launch_pdl(bn_fwd_finalize_kernel, dim3((bn.C + kFinCh - 1) / kFinCh), dim3(kFinThreads), 0, st, f);
    LAUNCH_CHECK("bn_fwd_finalize");
    return MDS_OK;
}
// conv output y -> batch statistics -> scale / shift
static int train_bn_stats(MdsTrainer* t, const TrainBn& bn, const __half* y, int b, int rows_per_sample, float* partials, cudaStream_t st) {
    {
        ProfScope ps(MDS_KIND_TRAIN_BN, st);
        EwParams p = ew_base(y, bn.C, rows_per_sample, b);
        p.partials = partials; p.mean = t->S + bn.rm;
        launch_pdl(bn_stats_kernel, ew_grid(rows_per_sample, b), dim3(kEwThreads), 0, st, p);
        LAUNCH_CHECK("bn_stats");
    }
    return train_bn_finalize(t, bn, partials, train_chunks(rows_per_sample, b) * b, (double)b * rows_per_sample, t->S + bn.rm, st);
}

template <int MODE>
static int train_bn_fwd(const TrainBn& bn, const __half* y, const __half* res, __half* out, float* partials, const float* smul,
                        const float* bmul, int b, int rows_per_sample, cudaStream_t st) {
    ProfScope ps(MDS_KIND_TRAIN_BN, st);
    EwParams p = ew_base(y, bn.C, rows_per_sample, b);
    p.g = res; p.out = out; p.partials = partials; p.scale = bn.scale(); p.shift = bn.shift(); p.smul = smul; p.bmul = bmul;
    launch_pdl(bn_fwd_kernel<MODE>, ew_grid(rows_per_sample, b), dim3(kEwThreads), 0, st, p);
    LAUNCH_CHECK("bn_fwd");
    return MDS_OK;
}

// BatchNorm(+SiLU) backward: da (= g * smul + sadd, * bmul) -> dy, and d gamma / d beta into the gradient buffer
static int train_bn_bwd(MdsTrainer* t, const TrainBn& bn, bool act, const __half* y, const __half* g, __half* dy, const float* smul,
                        const float* sadd, const float* bmul, int b, int rows_per_sample, float* partials, cudaStream_t st) {
    ProfScope ps(MDS_KIND_TRAIN_BN, st);
    EwParams p = ew_base(y, bn.C, rows_per_sample, b);
    p.g = g; p.out = dy; p.partials = partials; p.scale = bn.scale(); p.shift = bn.shift(); p.mean = bn.mean(); p.rstd = bn.rstd();
    p.smul = smul; p.sadd = sadd; p.bmul = bmul; p.c1 = bn.c1(); p.c2 = bn.c2(); p.gr = bn.gr();
    const dim3 grid = ew_grid(rows_per_sample, b);
    if (act) launch_pdl(bn_bwd_reduce_kernel<true>, dim3(grid), dim3(kEwThreads), 0, st, p);
    else launch_pdl(bn_bwd_reduce_kernel<false>, dim3(grid), dim3(kEwThreads), 0, st, p);
    LAUNCH_CHECK("bn_bwd_reduce");
    BnBwdFin f;
    f.partials = partials; f.nparts = grid.x * b; f.gamma = t->P + bn.gamma; f.rstd = bn.rstd(); f.mean = bn.mean();
    f.dgamma = t->G + bn.gamma; f.dbeta = t->G + bn.beta; f.c1 = bn.c1(); f.c2 = bn.c2(); f.gr = bn.gr();
    f.C = bn.C; f.count = (float)((double)b * rows_per_sample);
    launch_pdl(bn_bwd_finalize_kernel, dim3((bn.C + kFinCh - 1) / kFinCh), dim3(kFinThreads), 0, st, f);
    LAUNCH_CHECK("bn_bwd_finalize");
    if (act) launch_pdl(bn_bwd_apply_kernel<true>, dim3(grid), dim3(kEwThreads), 0, st, p);
    else launch_pdl(bn_bwd_apply_kernel<false>, dim3(grid), dim3(kEwThreads), 0, st, p);
    LAUNCH_CHECK("bn_bwd_apply");
    return MDS_OK;
}

static size_t wgrad_partial_floats(long long M, int N, int K, int* splits_out, int* rows_out) {
    const int tiles = (N / 64) * (K / 64);
    int want = (4 * num_sms() + tiles - 1) / tiles;       // ~4 CTAs of 128 threads per SM hide the smem / L2 latency
    if (want < 1) want = 1;
    long long rows = (M + want - 1) / want;
    rows = (rows + 31) / 32 * 32;
    if (rows < 256) rows = 256;
    const int splits = (int)((M + rows - 1) / rows);
    if (splits_out) *splits_out = splits > 0 ? splits : 1;
    if (rows_out) *rows_out = (int)rows;
    return (size_t)(splits > 0 ? splits : 1) * N * K;
}
// grad[N][K] = dY^T X
static int train_wgrad(const __half* dY, const __half* X, long long M, int N, int K, float* partials, float* grad, cudaStream_t st) {
    if (N % 64 || K % 64) return fail(MDS_ERR_INVALID, "wgrad: N and K must be multiples of 64 (N=%d K=%d)", N, K);
    ProfScope ps(MDS_KIND_TRAIN_WGRAD, st);
    WgradParams p;
    p.dY = dY; p.X = X; p.partials = partials; p.M = M; p.N = N; p.K = K;
    int splits = 1;
    wgrad_partial_floats(M, N, K, &splits, &p.rows_per_split);
    launch_pdl(wgrad_gemm_kernel, dim3(K / 64, N / 64, splits), dim3(128), 0, st, p);
    LAUNCH_CHECK("wgrad_gemm");
    const size_t count = (size_t)N * K;
    launch_pdl(sum_partials_kernel, dim3((unsigned)((count + 255) / 256)), dim3(256), 0, st, partials, splits, count, grad);
    LAUNCH_CHECK("sum_partials");
    return MDS_OK;
}

// conv_dw forward / data gradient on the streaming inference kernel in LIN mode: w27 tap-major weights (mirrored for the
// data gradient); stats (optional): per-CTA BatchNorm partial sums relative to `ref`, *nparts_total of them
static int train_dw3_conv(const __half* in, __half* out, const float* w27, const float* ref, float* stats, int* nparts_total, int b, int T,
                          int H, int W, int C, cudaStream_t st) {
    if (C % 8 || b > 65535) return fail(MDS_ERR_INVALID, "dw3 (train): C must be a multiple of 8, b <= 65535");
    DwParams p;
    p.in = in; p.out = out; p.w = w27; p.bias = ref; p.partials = stats;
    p.n = b; p.T = T; p.H = H; p.W = W; p.C = C; p.Ho = H; p.Wo = W;
    p.xtiles = (W + kDwTWX - 1) / kDwTWX;
    p.slabs = (C + kDwCS - 1) / kDwCS;
    int chunks = 1;
    const long long base = (long long)p.xtiles * p.slabs * b * T;
    const long long want = (long long)num_sms() * 6;
    if (base < want) {
        chunks = (int)((want + base - 1) / base);
        const int max_chunks = H / 6 > 0 ? H / 6 : 1;
        if (chunks > max_chunks) chunks = max_chunks;
    }
    p.rows_per_chunk = (H + chunks - 1) / chunks;
    p.chunks = (H + p.rows_per_chunk - 1) / p.rows_per_chunk;
    p.nparts = p.chunks * T * p.xtiles;
    if (p.nparts > kDwMaxParts * 4) return fail(MDS_ERR_INVALID, "dw3 (train): %d statistics partials per sample exceed %d", p.nparts, kDwMaxParts * 4);
    if (nparts_total) *nparts_total = p.nparts * b;
    using Cfg = DwCfg<3, 1>;
    auto kern = dwconv_kernel<3, 1, true>;
    ENSURE_SMEM_ATTR(kern, Cfg::SMEM);
    ProfScope ps(MDS_KIND_TRAIN_DW, st);
    launch_pdl(kern, dim3(p.xtiles * p.slabs, p.chunks * T, b), dim3(256), Cfg::SMEM, st, p);
    LAUNCH_CHECK("dwconv_lin");
    return MDS_OK;
}
static int dw3_wgrad_geometry(int b, int T, int H, int W, int C, int* chunks, int* rows) {
    const int xtiles = (W + kDwTWX - 1) / kDwTWX, slabs = (C + kDwCS - 1) / kDwCS;
    const long long base = (long long)xtiles * slabs * b * T;
    int ch = (int)((2LL * num_sms() + base - 1) / base);
    const int max_ch = H / 4 > 0 ? H / 4 : 1;
    if (ch > max_ch) ch = max_ch;
    if (ch < 1) ch = 1;
    const int r = (H + ch - 1) / ch;
    ch = (H + r - 1) / r;
    if (chunks) *chunks = ch;
    if (rows) *rows = r;
    return b * T * ch * xtiles;       // partials per channel
}
static int train_dw3_wgrad(const __half* in, const __half* dy, float* partials, float* grad, int b, int T, int H, int W, int C, cudaStream_t st) {
    if (C % 8 || b > 65535) return fail(MDS_ERR_INVALID, "dw3 wgrad: C must be a multiple of 8, b <= 65535");
    Dw3WgParams p;
    p.in = in; p.dy = dy; p.partials = partials; p.n = b; p.T = T; p.H = H; p.W = W; p.C = C;
    p.xtiles = (W + kDwTWX - 1) / kDwTWX;
    const int nparts = dw3_wgrad_geometry(b, T, H, W, C, &p.chunks, &p.rows_per_chunk);
    ENSURE_SMEM_ATTR(dw3_wgrad_kernel, Dw3WgCfg::SMEM);
    ProfScope ps(MDS_KIND_TRAIN_DW, st);
    launch_pdl(dw3_wgrad_kernel, dim3(p.xtiles * ((C + kDwCS - 1) / kDwCS), p.chunks * T, b), dim3(256), Dw3WgCfg::SMEM, st, p);
    LAUNCH_CHECK("dw3_wgrad");
    launch_pdl(dw3_wgrad_reduce_kernel, dim3((27 * C + 255) / 256), dim3(256), 0, st, partials, nparts, C, grad);
    LAUNCH_CHECK("dw3_wgrad_reduce");
    return MDS_OK;
}

// --------------------------------------------------------------------------------------------------------------
// workspace + step
// --------------------------------------------------------------------------------------------------------------
struct TrainWs {
    __half *y0, *yp, *ap, *dP, *dP2, *dXa, *dXb, *D3, *DM1, *DM2;
    std::vector<__half*> x, y1, a1, y2, a2g, y3;
    float *partials, *wpart, *dwpart, *se_s, *se_h, *se_g, *se_dg, *se_dh, *se_sadd, *dp_mask, *do_mask;
    float *feat, *pooled, *mlog, *coef, *gem_part, *dp_part, *se_sum, *logit_part;
};
static int train_ws_take(const MdsTrainer* t, int b, int fh, int fw, Arena& ar, TrainWs& w) {
    const int P = fh * fw;
    const int T = t->T(), c3 = t->cfg.num_3d_features, mid = t->mid(), rd = t->rd(), pj = t->cfg.num_3d_stack_proj;
    const int nb = (int)t->blocks.size();
    const size_t M = (size_t)b * T * P;
    const int cmax = mid > pj ? mid : pj;
    w.y0 = ar.take<__half>(M * c3);
    w.x.resize(nb + 1);
    w.y1.resize(nb); w.a1.resize(nb); w.y2.resize(nb); w.a2g.resize(nb); w.y3.resize(nb);
    for (int i = 0; i <= nb; ++i) w.x[i] = ar.take<__half>(M * c3);
    for (int i = 0; i < nb; ++i) {
        w.y1[i] = ar.take<__half>(M * mid); w.a1[i] = ar.take<__half>(M * mid);
        w.y2[i] = ar.take<__half>(M * mid); w.a2g[i] = ar.take<__half>(M * mid);
        w.y3[i] = ar.take<__half>(M * c3);
    }
    w.yp = ar.take<__half>(M * pj); w.ap = ar.take<__half>(M * pj);
    w.dP = ar.take<__half>(M * pj); w.dP2 = ar.take<__half>(M * pj);
    w.dXa = ar.take<__half>(M * c3); w.dXb = ar.take<__half>(M * c3); w.D3 = ar.take<__half>(M * c3);
    w.DM1 = ar.take<__half>(M * mid); w.DM2 = ar.take<__half>(M * mid);
    size_t npart = (size_t)b * train_chunks(T * P, b);
    {   // the depthwise kernel's fused statistics use its own CTA count
        const size_t dwp = (size_t)b * kDwMaxParts * 4;
        if (dwp > npart) npart = dwp;
    }
    w.partials = ar.take<float>(npart * 2 * cmax);
    size_t wp = 0;
    auto upd = [&](int N, int K) { size_t f = wgrad_partial_floats((long long)M, N, K, nullptr, nullptr); if (f > wp) wp = f; };
    upd(c3, 192); upd(mid, c3); upd(c3, mid); upd(pj, c3);
    w.wpart = ar.take<float>(wp);
    w.dwpart = ar.take<float>((size_t)dw3_wgrad_geometry(b, T, fh, fw, mid, nullptr, nullptr) * 27 * mid);
    w.se_s = ar.take<float>((size_t)nb * b * mid); w.se_h = ar.take<float>((size_t)nb * b * rd); w.se_g = ar.take<float>((size_t)nb * b * mid);
    w.se_dg = ar.take<float>((size_t)b * mid); w.se_dh = ar.take<float>((size_t)b * rd); w.se_sadd = ar.take<float>((size_t)b * mid);
    w.dp_mask = ar.take<float>((size_t)(nb > 0 ? nb : 1) * b); w.do_mask = ar.take<float>((size_t)b * t->F());
    w.feat = ar.take<float>((size_t)b * t->F()); w.pooled = ar.take<float>((size_t)b * t->F());
    w.mlog = ar.take<float>((size_t)b * t->F()); w.coef = ar.take<float>((size_t)b * t->F());
    w.gem_part = ar.take<float>((size_t)b * t->F() * kGemChunks * 2); w.dp_part = ar.take<float>((size_t)(t->F() + 255) / 256);
    w.se_sum = ar.take<float>((size_t)b * mid);
    w.logit_part = ar.take<float>((size_t)b * ((t->F() + 255) / 256) * t->cfg.num_classes);
    return ar.overflow ? 1 : 0;
}

extern "C" size_t mds_train_workspace_bytes(const MdsTrainer* t, int b, int fh, int fw) {
    if (!t || b <= 0 || fh <= 0 || fw <= 0) return 0;
    Arena ar(nullptr, ~size_t(0) >> 1);
    TrainWs w;
    train_ws_take(t, b, fh, fw, ar, w);
    return ar.off + 4096;
}

extern "C" int mds_train_step(MdsTrainer* t, const MdsTrainStepArgs* a, void* ws, size_t ws_bytes, void* stream) {
    if (!t || !a || !ws) return fail(MDS_ERR_INVALID, "train_step: null argument");
    if (!t->committed) return fail(MDS_ERR_WEIGHTS, "train_step: parameters not committed (mds_train_commit)");
    if (!a->enc_feats || !a->targets) return fail(MDS_ERR_INVALID, "train_step: enc_feats and targets are required");
    if (a->b <= 0 || a->b > 65535 || a->fh <= 0 || a->fw <= 0) return fail(MDS_ERR_INVALID, "train_step: bad batch / feature-map size");
    if ((long long)a->b * t->cfg.num_classes > 1024) return fail(MDS_ERR_INVALID, "train_step: b * num_classes must be <= 1024");
    DeviceGuard dg(t->cfg.device);
    // measured: programmatic dependent launch slows this chain down (3.83 vs 3.66 ms per step), so it runs plainly serialized
    struct PdlOff { bool prev; PdlOff() : prev(g_pdl) { g_pdl = false; } ~PdlOff() { g_pdl = prev; } } pdl_off;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int b = a->b, T = t->T(), P = a->fh * a->fw, rows = T * P;
    const int c3 = t->cfg.num_3d_features, mid = t->mid(), rd = t->rd(), pj = t->cfg.num_3d_stack_proj, nb = (int)t->blocks.size();
    const long long M = (long long)b * rows;
    const int F = t->F(), K = t->cfg.num_classes;
    Arena ar(ws, ws_bytes);
    TrainWs w;
    if (train_ws_take(t, b, a->fh, a->fw, ar, w)) return fail(MDS_ERR_WORKSPACE, "train_step: workspace too small (%zu bytes)", ws_bytes);
    const __half* xe = reinterpret_cast<const __half*>(a->enc_feats);
    const float inv_rows = 1.0f / (float)rows;
    const size_t se_smem = (size_t)(2 * mid + 2 * rd) * sizeof(float);

    // ---- stochastic masks: supplied by the caller (parity tests) or drawn here from (seed, step) ----
    const float* dp_mask = a->dp_masks;
    const float* do_mask = a->dropout_mask;
    if (!dp_mask && t->cfg.drop_path_rate > 0.f && nb > 0) {
        launch_pdl(bernoulli_mask_kernel, dim3((nb * b + 255) / 256), dim3(256), 0, st, w.dp_mask, (size_t)nb * b, 1.0f - t->cfg.drop_path_rate, a->seed * 2 + 1);
        LAUNCH_CHECK("bernoulli_mask");
        dp_mask = w.dp_mask;
    }
    if (!do_mask && t->cfg.drop_rate > 0.f) {
        launch_pdl(bernoulli_mask_kernel, dim3((unsigned)(((size_t)b * F + 255) / 256)), dim3(256), 0, st, w.do_mask, (size_t)b * F, 1.0f - t->cfg.drop_rate, a->seed * 2 + 2);
        LAUNCH_CHECK("bernoulli_mask");
        do_mask = w.do_mask;
    }

    // ================================================= forward =================================================
    g_prof_tag = 300;
    TRY(train_gemm(t, xe, t->proj2d.w16, nullptr, w.y0, M, c3, 192, st));            // multidim_stacker.py:216
    TRY(train_bn_stats(t, t->bn_p2d, w.y0, b, rows, w.partials, st));
    TRY(train_bn_fwd<0>(t->bn_p2d, w.y0, nullptr, w.x[0], nullptr, nullptr, nullptr, b, rows, st));
    for (int i = 0; i < nb; ++i) {                                                                           // InvertedResidual3d.forward (:124-134)
        const TrainBlock& B = t->blocks[i];
        g_prof_tag = 301 + i;
        float* se_s = w.se_s + (size_t)i * b * mid; float* se_h = w.se_h + (size_t)i * b * rd; float* se_g = w.se_g + (size_t)i * b * mid;
        TRY(train_gemm(t, w.x[i], B.pw.w16, nullptr, w.y1[i], M, mid, c3, st));
        TRY(train_bn_stats(t, B.bn1, w.y1[i], b, rows, w.partials, st));
        TRY(train_bn_fwd<0>(B.bn1, w.y1[i], nullptr, w.a1[i], nullptr, nullptr, nullptr, b, rows, st));
        {   // conv_dw with the bn2 batch statistics accumulated in its epilogue
            int nparts = 0;
            TRY(train_dw3_conv(w.a1[i], w.y2[i], B.dw27, t->S + B.bn2.rm, w.partials, &nparts, b, T, a->fh, a->fw, mid, st));
            TRY(train_bn_finalize(t, B.bn2, w.partials, nparts, (double)M, t->S + B.bn2.rm, st));
        }
        TRY(train_bn_fwd<1>(B.bn2, w.y2[i], nullptr, nullptr, w.partials, nullptr, nullptr, b, rows, st));   // SE squeeze sums
        {
            ProfScope ps(MDS_KIND_TRAIN_SMALL, st);
            launch_pdl(colsum_finalize_kernel, dim3((mid + kFinCh - 1) / kFinCh, b), dim3(kFinThreads), 0, st, w.partials, train_chunks(rows, b), mid, inv_rows, se_s);
            LAUNCH_CHECK("colsum_finalize");
            SeTrainParams sp;
            memset(&sp, 0, sizeof(sp));
            sp.sums = se_s;
            sp.w1 = t->P + B.se_w1; sp.b1 = t->P + B.se_b1; sp.w2 = t->P + B.se_w2; sp.b2 = t->P + B.se_b2;
            sp.hpre = se_h; sp.gate = se_g; sp.C = mid; sp.rd = rd; sp.inv_count = inv_rows;
            launch_pdl(se_train_fwd_kernel, dim3(b), dim3(256), se_smem, st, sp);
            LAUNCH_CHECK("se_train_fwd");
        }
        TRY(train_bn_fwd<2>(B.bn2, w.y2[i], nullptr, w.a2g[i], nullptr, se_g, nullptr, b, rows, st));
        TRY(train_gemm(t, w.a2g[i], B.pwl.w16, nullptr, w.y3[i], M, c3, mid, st));
        TRY(train_bn_stats(t, B.bn3, w.y3[i], b, rows, w.partials, st));
        TRY(train_bn_fwd<3>(B.bn3, w.y3[i], w.x[i], w.x[i + 1], nullptr, nullptr, dp_mask ? dp_mask + (size_t)i * b : nullptr, b, rows, st));
    }
    g_prof_tag = 350;
    TRY(train_gemm(t, w.x[nb], t->proj3d.w16, nullptr, w.yp, M, pj, c3, st));         // :227
    TRY(train_bn_stats(t, t->bn_p3d, w.yp, b, rows, w.partials, st));
    TRY(train_bn_fwd<0>(t->bn_p3d, w.yp, nullptr, w.ap, nullptr, nullptr, nullptr, b, rows, st));
    g_prof_tag = 360;
    {
        ProfScope ps(MDS_KIND_TRAIN_SMALL, st);
        GemTrainParams gp;
        gp.x = w.ap; gp.p = t->P + t->gem_p; gp.partials = w.gem_part; gp.T = T; gp.P = P; gp.C = pj; gp.eps = 1e-6f;
        launch_pdl(gem_train_fwd_kernel, dim3(T, b, kGemChunks), dim3(256), 0, st, gp);
        LAUNCH_CHECK("gem_train_fwd");
        HeadTrainParams hp;
        hp.gem_partials = w.gem_part; hp.feat = w.feat; hp.pooled = w.pooled; hp.mlog = w.mlog; hp.dmask = do_mask;
        hp.w = t->P + t->cls_w; hp.bias = t->P + t->cls_b;
        hp.targets = a->targets; hp.gem_p = t->P + t->gem_p; hp.scaler = t->scaler;
        hp.logits = a->logits_out ? a->logits_out : w.se_dg;      // scratch when the caller does not want them
        hp.loss = a->loss_out ? a->loss_out : w.se_dh;
        hp.dw = t->G + t->cls_w; hp.dbias = t->G + t->cls_b; hp.dgem_p = t->G + t->gem_p; hp.dp_partials = w.dp_part; hp.coef = w.coef;
        hp.b = b; hp.F = F; hp.K = K; hp.P = P; hp.C = pj; hp.alpha = t->cfg.focal_alpha; hp.gamma = t->cfg.focal_gamma;
        const int hblocks = (F + 255) / 256;
        hp.logit_partials = w.logit_part;
        launch_pdl(head_logits_kernel, dim3(hblocks, b), dim3(256), 0, st, hp);
        LAUNCH_CHECK("head_logits");
        launch_pdl(head_grad_kernel, dim3(hblocks), dim3(256), (size_t)2 * b * K * sizeof(float), st, hp);
        LAUNCH_CHECK("head_grad");
        launch_pdl(head_dp_reduce_kernel, dim3(1), dim3(32), 0, st, w.dp_part, hblocks, t->G + t->gem_p);
        LAUNCH_CHECK("head_dp_reduce");
        // ============================================= backward =============================================
        GemBwdParams gb;
        gb.x = w.ap; gb.coef = w.coef; gb.p = t->P + t->gem_p; gb.dx = w.dP; gb.T = T; gb.P = P; gb.C = pj; gb.eps = 1e-6f;
        launch_pdl(gem_bwd_kernel, dim3(T, b, kGemChunks), dim3(256), 0, st, gb);
        LAUNCH_CHECK("gem_bwd");
    }
    g_prof_tag = 450;
    TRY(train_bn_bwd(t, t->bn_p3d, true, w.yp, w.dP, w.dP2, nullptr, nullptr, nullptr, b, rows, w.partials, st));
    TRY(train_wgrad(w.dP2, w.x[nb], M, pj, c3, w.wpart, t->G + t->proj3d.off, st));
    __half* dX = w.dXa;      // gradient with respect to the current block output
    __half* dXn = w.dXb;
    TRY(train_gemm(t, w.dP2, t->proj3d.w16t, nullptr, dX, M, c3, pj, st));
    for (int i = nb - 1; i >= 0; --i) {
        const TrainBlock& B = t->blocks[i];
        g_prof_tag = 401 + i;
        float* se_s = w.se_s + (size_t)i * b * mid; float* se_h = w.se_h + (size_t)i * b * rd; float* se_g = w.se_g + (size_t)i * b * mid;
        // bn3 (no activation), DropPath mask on the branch
        TRY(train_bn_bwd(t, B.bn3, false, w.y3[i], dX, w.D3, nullptr, nullptr, dp_mask ? dp_mask + (size_t)i * b : nullptr, b, rows, w.partials, st));
        TRY(train_wgrad(w.D3, w.a2g[i], M, c3, mid, w.wpart, t->G + B.pwl.off, st));
        TRY(train_gemm(t, w.D3, B.pwl.w16t, nullptr, w.DM1, M, mid, c3, st));         // d (a2 * gate)
        {   // SE backward: d gate -> d squeeze, parameter gradients
            ProfScope ps(MDS_KIND_TRAIN_BN, st);
            EwParams p = ew_base(w.y2[i], mid, rows, b);
            p.g = w.DM1; p.partials = w.partials; p.scale = B.bn2.scale(); p.shift = B.bn2.shift();
            launch_pdl(dgate_kernel, ew_grid(rows, b), dim3(kEwThreads), 0, st, p);
            LAUNCH_CHECK("dgate");
        }
        {
            ProfScope ps(MDS_KIND_TRAIN_SMALL, st);
            launch_pdl(colsum_finalize_kernel, dim3((mid + kFinCh - 1) / kFinCh, b), dim3(kFinThreads), 0, st, w.partials, train_chunks(rows, b), mid, 1.0f, w.se_sum);
            LAUNCH_CHECK("colsum_finalize");
            SeTrainParams sp;
            memset(&sp, 0, sizeof(sp));
            sp.sums = w.se_sum;
            sp.w1 = t->P + B.se_w1; sp.b1 = t->P + B.se_b1; sp.w2 = t->P + B.se_w2; sp.b2 = t->P + B.se_b2;
            sp.hpre = se_h; sp.gate = se_g; sp.dgpre = w.se_dg; sp.dhpre = w.se_dh; sp.sadd = w.se_sadd;
            sp.C = mid; sp.rd = rd; sp.inv_count = inv_rows;
            launch_pdl(se_train_bwd_kernel, dim3(b), dim3(256), se_smem, st, sp);
            LAUNCH_CHECK("se_train_bwd");
            SeGradParams sg;
            sg.s = se_s; sg.hpre = se_h; sg.dgpre = w.se_dg; sg.dhpre = w.se_dh;
            sg.dw1 = t->G + B.se_w1; sg.db1 = t->G + B.se_b1; sg.dw2 = t->G + B.se_w2; sg.db2 = t->G + B.se_b2;
            sg.b = b; sg.C = mid; sg.rd = rd;
            launch_pdl(se_train_wgrad_kernel, dim3((mid * rd + 255) / 256), dim3(256), 0, st, sg);
            LAUNCH_CHECK("se_train_wgrad");
        }
        TRY(train_bn_bwd(t, B.bn2, true, w.y2[i], w.DM1, w.DM2, se_g, w.se_sadd, nullptr, b, rows, w.partials, st));   // -> d y2
        TRY(train_dw3_wgrad(w.a1[i], w.DM2, w.dwpart, t->G + B.dw, b, T, a->fh, a->fw, mid, st));
        TRY(train_dw3_conv(w.DM2, w.DM1, B.dw27f, t->zeros, nullptr, nullptr, b, T, a->fh, a->fw, mid, st));           // -> d a1
        TRY(train_bn_bwd(t, B.bn1, true, w.y1[i], w.DM1, w.DM2, nullptr, nullptr, nullptr, b, rows, w.partials, st));  // -> d y1
        TRY(train_wgrad(w.DM2, w.x[i], M, mid, c3, w.wpart, t->G + B.pw.off, st));
        TRY(train_gemm(t, w.DM2, B.pw.w16t, dX, dXn, M, c3, mid, st));                // + shortcut gradient
        __half* tmp = dX; dX = dXn; dXn = tmp;
    }
    g_prof_tag = 400;
    TRY(train_bn_bwd(t, t->bn_p2d, true, w.y0, dX, w.D3, nullptr, nullptr, nullptr, b, rows, w.partials, st));
    TRY(train_wgrad(w.D3, xe, M, c3, 192, w.wpart, t->G + t->proj2d.off, st));

    // ================================================= optimizer ================================================
    if (a->apply_update) {
        g_prof_tag = 500;
        {
            ProfScope ps(MDS_KIND_TRAIN_SMALL, st);
            const int blocks = (int)((t->n_params + 255) / 256) < 4 * num_sms() ? (int)((t->n_params + 255) / 256) : 4 * num_sms();
            launch_pdl(grad_check_kernel, dim3(blocks), dim3(256), 0, st, t->G, t->n_params, t->scaler);
            LAUNCH_CHECK("grad_check");
            launch_pdl(sgd_nesterov_kernel, dim3(blocks), dim3(256), 0, st, t->P, t->G, t->Mom, t->n_params, t->scaler, a->lr, t->cfg.momentum, t->cfg.nesterov);
            LAUNCH_CHECK("sgd_nesterov");
            launch_pdl(scaler_update_kernel, dim3(1), dim3(32), 0, st, t->scaler, 2.0f, 0.5f, 2000.0f, t->cfg.amp);
            LAUNCH_CHECK("scaler_update");
        }
        TRY(train_derive(t, st));
    }
    ++t->batches_tracked;
    return MDS_OK;
}

extern "C" long long mds_train_batches_tracked(const MdsTrainer* t) { return t ? t->batches_tracked : 0; }

extern "C" int mds_focal_loss(const float* logits, const float* targets, int n, float alpha, float gamma, float* loss_out,
                              float* probs_out, void* stream) {
    if (!logits || !targets || !loss_out || !probs_out) return fail(MDS_ERR_INVALID, "focal_loss: null argument");
    if (n <= 0 || n > (1 << 20)) return fail(MDS_ERR_INVALID, "focal_loss: n must be in [1, 2^20]");
    focal_loss_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(logits, targets, n, alpha, gamma, loss_out, probs_out);
    LAUNCH_CHECK("focal_loss");
    return MDS_OK;
}
