/*
 * mds_b200.h — C ABI of the B200-native MultiDimStacker forward path (libmds_b200.so).
 *
 * The reference (lRomul/ball-action-spotting) is pure Python/PyTorch and has no FFI for this path; the
 * functions below are what a binding for the path would have to offer.  Each entry point names the reference
 * interface it replaces (paths relative to the reference repo).  All pointers are plain device pointers unless
 * a parameter says "host"; no torch types cross this boundary.  Every function returns 0 on success and a
 * negative MdsStatus on failure; mds_last_error() returns a thread-local message.  Launches are asynchronous
 * on the caller's stream (the reference runs on the current stream without syncs, src/predictors.py:50).
 * Ownership: the caller owns every input/output/workspace buffer; the handle owns only packed weights and a few KB of
 * inter-CTA synchronisation words.
 * A handle is bound to one device and is not thread-safe (the reference is single-threaded per device).
 */
#ifndef MDS_B200_H_
#define MDS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct MdsHandle MdsHandle;

typedef enum MdsStatus {
    MDS_OK = 0,
    MDS_ERR_INVALID = -1,   /* bad argument / unsupported shape (mirrors the asserts at multidim_stacker.py:155,212,223) */
    MDS_ERR_WEIGHTS = -2,   /* missing / mis-sized tensor at commit (mirrors load_state_dict strictness) */
    MDS_ERR_WORKSPACE = -3, /* workspace too small */
    MDS_ERR_CUDA = -4       /* CUDA runtime error */
} MdsStatus;

/* kwargs of MultiDimStacker.__init__ that change shapes (src/models/multidim_stacker.py:138-153).
 * model_name is fixed to tf_efficientnetv2_b0 (configs/ball_action/sampling_weights_001.py:31). */
typedef struct MdsConfig {
    int num_classes;
    int num_frames;
    int stack_size;          /* must be 3 */
    int num_3d_blocks;
    int num_3d_features;     /* must be 192 (== encoder feature width) */
    int num_3d_stack_proj;   /* <= 256 */
    int expansion_3d_ratio;
    int se_reduce_3d_ratio;
    int device;              /* CUDA ordinal */
    int chunk_images;        /* images per pass through the 2D encoder (L2 blocking); 0 = default */
} MdsConfig;

/* Input frames for the 2D encoder: planar, `stack_size` consecutive planes are the channels of one image
 * (multidim_stacker.py:214).  dtype 0: uint8 raw frames, zero-padded to (H, W) and divided by 255 inside the
 * stem kernel (src/frames.py:7-31, PadNormalizeFramesProcessor); dtype 1: float32 already-normalised input
 * (what MultiDimStacker.forward receives). */
typedef struct MdsFrames {
    const void* data;
    int dtype;
    long long img_stride;     /* elements between images */
    long long plane_stride;   /* elements between the channel planes of one image */
    int stored_h;             /* rows stored per plane (e.g. 720) */
    int pad_top;              /* (H - stored_h) / 2, frames.py:19 */
    int H, W;                 /* padded size, multiples of 32 */
    int hflip;                /* 1: horizontal-flip TTA (kornia hflip, src/predictors.py:63) */
} MdsFrames;

const char* mds_last_error(void);

/* argus.load_model -> BallActionModel(params) -> MultiDimStacker(**kwargs) (src/predictors.py:22,
 * src/argus_models.py:17-21): create, then add every packed tensor, then commit. */
int mds_create(const MdsConfig* cfg, MdsHandle** out);
int mds_destroy(MdsHandle* h);
/* nn.Module.load_state_dict: host_data is copied to the device. Names are listed in packer.py. */
int mds_weights_add(MdsHandle* h, const char* name, const void* host_data, size_t nbytes);
int mds_weights_commit(MdsHandle* h);

size_t mds_workspace_bytes(const MdsHandle* h, int H, int W, int n_images, int n_stacks);

/* MultiDimStacker.forward_2d (multidim_stacker.py:210-219): n_images 3-frame images -> feats fp16
 * [n_images][H/32][W/32][192] (NHWC; the reference's (b, k, 192, h, w) with b*k = n_images). */
int mds_forward_2d(MdsHandle* h, const MdsFrames* frames, int n_images, void* feats_out,
                   void* ws, size_t ws_bytes, void* stream);
/* The frozen conv2d_encoder alone (multidim_stacker.py:215, `self.conv2d_encoder(x)[-1]`, eval mode): the input of
 * conv2d_projection, fp16 [n_images][H/32][W/32][192].  Feeds the training step below (freeze_conv2d_encoder,
 * src/argus_models.py:104-110). */
int mds_forward_encoder(MdsHandle* h, const MdsFrames* frames, int n_images, void* feats_out, void* ws, size_t ws_bytes,
                        void* stream);
/* MultiDimStacker.forward_3d (multidim_stacker.py:221-230): feats fp16 [b][T][fh][fw][192] -> out fp16
 * [b][T][fh][fw][proj] (the reference's (b, proj*T, h, w) with channel = t*proj + c). */
int mds_forward_3d(MdsHandle* h, const void* feats, int b, int fh, int fw, void* out, void* ws, size_t ws_bytes,
                   void* stream);
/* MultiDimStacker.forward_head (multidim_stacker.py:232-237) + optional nn.Sigmoid (src/argus_models.py:26). */
int mds_forward_head(MdsHandle* h, const void* x, int b, int P, float* logits, int apply_sigmoid,
                     void* ws, size_t ws_bytes, void* stream);
/* MultiDimStacker.forward (multidim_stacker.py:239-243): b stacks of num_frames frames -> logits f32 [b][classes].
 * frames->img_stride addresses image i = stack (i / T), triple (i % T). */
int mds_forward(MdsHandle* h, const MdsFrames* frames, int b, float* logits, int apply_sigmoid,
                void* ws, size_t ws_bytes, void* stream);

/* Sliding-window assembly (src/predictors.py:66-68: torch.cat of the per-triple cached features): window p of n_pred
 * consecutive predictions stacks images first_image + p + hop * t, t = 0..T-1, of a cached fp16 feature buffer
 * [n_img][plane_elems] into out [n_pred][T][plane_elems]. */
int mds_gather_stacks(const void* feats, void* out, long long first_image, int hop, int n_pred, int T, long long plane_elems,
                      void* stream);
/* y = a * y + b * x over n floats: the mean over the TTA branches (predictors.py:72). */
int mds_axpby(float* y, const float* x, float a, float b, long long n, void* stream);

/* Boundary layout converters (reference tensors are NCHW float32, the engine is NHWC fp16). */
int mds_nchw32_to_nhwc16(const float* src, void* dst, int n, int C, int P, void* stream);
int mds_nhwc16_to_nchw32(const void* src, float* dst, int n, int C, int P, void* stream);

/* ---- per-kernel entry points (parity tests, ncu) ---- */
/* Stem (frames.py:7-31 pad + /255, timm conv_stem + bn1 + SiLU; multidim_stacker.py:214): out fp16 [n][H/2][W/2][32].
 * wh: fp16 [2][32][32] = (hi, lo) x cout x k, k = (ci*3 + r)*3 + s, columns 27..31 zero (packer.stem_weights).
 * uint8 frames run stem_tc_kernel (TMA + tcgen05) and need W >= 144, W / plane_stride / img_stride multiples of 16 and a 16-byte
 * aligned base; float32 frames run the mma.sync gather kernel. */
int mds_k_stem(const MdsFrames* frames, int n_images, const void* wh, const float* bias, void* out, void* stream);
/* Dense 3x3 blocks of timm tf_efficientnetv2_b0 (built at multidim_stacker.py:166-176), NHWC fp16, BN folded:
 * ConvBnAct (cproj = 0): out = SiLU(conv3x3(in) + b1); EdgeResidual: out = conv1x1(SiLU(conv3x3_stride(in) + b1)) + b2 (+ in if res).
 * w1 [cmid][9*cin] (k = (r*3+s)*cin + ci), w2 [cproj][cmid].  Supported (cin, cmid, stride, cproj, res): (32,16,1,0,0), (16,64,2,32,0),
 * (32,128,1,32,1), (32,128,2,48,0) -> conv_tc_kernel; (48,192,1,48,1) -> conv_tc_ws_kernel; anything else is MDS_ERR_INVALID.
 * stride 2 = TF-SAME on an even H, W (pad bottom / right). */
int mds_k_conv3x3(const void* in, void* out, const void* w1, const float* b1, const void* w2, const float* b2,
                  int n, int H, int W, int cin, int cmid, int stride, int cproj, int res, void* stream);
/* bias_mat (optional): [N][64] fp16, col 0 = fp16(bias), col 1 = fp16(bias - col 0); selects the tcgen05 kernel for
 * ungated GEMMs with K <= 192 (the bias is then added by the tensor core). */
int mds_k_gemm1x1(const void* A, const void* W, const float* bias, const void* bias_mat, const void* res, const void* gate,
                  void* C, int rows_per_img, int n_img, int N, int K, int act, void* stream);
/* partials: f32 [n][*nparts][C] (room for 64 parts per image): per-CTA sums of the output for the SE squeeze, every
 * element written (no atomics, nothing to clear); *nparts receives the number of parts this launch produced. */
int mds_k_dwconv(const void* in, void* out, const float* w, const float* bias, float* partials, int* nparts,
                 int n, int T, int H, int W, int C, int kt, int stride, void* stream);
/* SE excitation from the depthwise kernel's partial sums (added in a fixed order: deterministic); w32 [N][C] fp32 +
 * wg [n][N][C] fp16 (optional): per-image gated projection weights wg = fp16(w32 * gate) for mds_k_gemm_gated. */
int mds_k_se_fc(const float* partials, int nparts, const float* w1, const float* b1, const float* w2t, const float* b2,
                void* gate, const float* w32, void* wg, int n, int C, int rd, int N, float inv_count, void* stream);
/* SE-gated projection GEMM on tcgen05: C[img] = act(A[img] . wg[img]^T + bias) (+ res); N <= 256. */
int mds_k_gemm_gated(const void* A, const void* wg, const void* bias_mat, const void* res, void* C, int rows_per_img,
                     int n_img, int N, int K, int act, void* stream);
/* Depthwise conv + BN + SiLU with TMA-staged input rows, the SE squeeze and the SE excitation MLP in one launch (timm
 * InvertedResidual conv_dw/bn2/se.conv_reduce/se.conv_expand; multidim_stacker.py:110-114, 72-90): out fp16 [n][T][H/s][W/s][C],
 * gate f32 [n][C] = sigmoid(W2 SiLU(W1 mean + b1) + b2), evaluated by the CTA that finishes an image last.  partials: f32 scratch
 * [n][64][C] (*nparts receives the number used); done: int scratch [n], ZERO on entry, left zero.  se_w1 = NULL: no SE (partials
 * only).  rows_per_chunk: output rows per CTA, 0 = default (a function of the layer shape only). */
int mds_k_dwconv_se(const void* in, void* out, const float* w, const float* bias, float* partials, int* nparts,
                    const float* se_w1, const float* se_b1, const float* se_w2t, const float* se_b2, float* gate, int* done,
                    int n, int T, int H, int W, int C, int kt, int stride, int rd, int rows_per_chunk, void* stream);
/* SE-gated projection GEMM on tcgen05: C[img] = act((A[img] * gate[img]) . W^T + bias) (+ res) with the gate applied to the A
 * blocks in shared memory (x * gate then conv_pwl + bn3 (+ shortcut), multidim_stacker.py:90,129-134); W fp16 [N][K], N <= 256,
 * K <= 1152, gate f32 [n_img][K]. */
int mds_k_gemm_gate(const void* A, const void* W, const float* gate, const void* bias_mat, const void* res, void* C,
                    int rows_per_img, int n_img, int N, int K, int act, void* stream);
/* Fused MBConv tail, ONE launch (timm InvertedResidual conv_dw/bn2/se/conv_pwl/bn3 (+ shortcut), built at
 * multidim_stacker.py:166-176; InvertedResidual3d multidim_stacker.py:110-134): m1 fp16 [n][T][H][W][C] (expanded input) ->
 * m2 fp16 [n][T][H/s][W/s][C] (depthwise + BN + SiLU, before gating), gate f32 [n][C] (SE excitation, SqueezeExcite :72-90),
 * out fp16 [n][T*(H/s)*(W/s)][N] = (m2 * gate) . wpwl^T + bias (+ res).  kt = 1 (2D, stride 1 or 2, T = 1) or 3 (3x3x3).
 * partials: f32 scratch [n][64][C]; sync: int scratch [3][n], must be ZERO on entry and is left zero.
 * N = 0: depthwise + SE only (wpwl / bias_mat / res / out unused).  rows_per_chunk: output rows per depthwise item, 0 = default. */
int mds_k_mbconv_tail(const void* m1, void* m2, const float* dw_w, const float* dw_b, float* partials, const float* se_w1,
                      const float* se_b1, const float* se_w2t, const float* se_b2, float* gate, int* sync, const void* wpwl,
                      const void* bias_mat, const void* res, void* out, int n, int T, int H, int W, int C, int kt, int stride,
                      int rd, int N, int rows_per_chunk, void* stream);
int mds_k_gem(const void* x, float* feat, int b, int T, int P, int C, float p, float eps, void* stream);
int mds_k_linear(const float* feat, const float* w, const float* bias, float* out, int b, int F, int num_classes,
                 int apply_sigmoid, void* stream);

/* Optional per-launch CUDA-event timing of this library's kernels (bench.py's live roofline measurement).
 * begin: start recording (events on the launching stream); end: synchronise the events, return one record per
 * launch (kind = MdsKernelKind, tag = layer id: 0 stem, 1..22 encoder blocks, 23 conv2d_projection, 101.. 3D
 * blocks, 150 conv3d_projection, 200 head, -1 direct kernel call), stop recording. */
typedef enum MdsKernelKind {
    MDS_KIND_STEM = 0, MDS_KIND_CONV3X3 = 1, MDS_KIND_GEMM1X1 = 2, MDS_KIND_DWCONV2D = 3, MDS_KIND_DWCONV3D = 4,
    MDS_KIND_SE_FC = 5, MDS_KIND_HEAD = 6,
    /* training step: weight-gradient GEMMs, BatchNorm column kernels, depthwise 3x3x3 fwd/bwd, small kernels */
    MDS_KIND_TRAIN_WGRAD = 7, MDS_KIND_TRAIN_BN = 8, MDS_KIND_TRAIN_DW = 9, MDS_KIND_TRAIN_SMALL = 10,
    /* fused MBConv tail (depthwise + SE + gated projection in one launch), 2D blocks / 3D blocks */
    MDS_KIND_TAIL2D = 11, MDS_KIND_TAIL3D = 12
} MdsKernelKind;
int mds_profile_begin(void);
int mds_profile_end(int* kinds, int* tags, float* ms, int capacity, int* count);

/* ---- training step with a frozen 2D encoder (BASELINE.json configs[4]) -------------------------------------------
 * Replaces BallActionModel.train_step (src/argus_models.py:41-74) for configs with freeze_conv2d_encoder
 * (configs/ball_action/ball_finetune_long_004.py:72): train-mode forward of conv2d_projection, the InvertedResidual3d
 * blocks, conv3d_projection, GeM, dropout and the classifier (src/models/multidim_stacker.py:216-237), sigmoid focal
 * loss (src/losses.py:31-48), backward, GradScaler (amp) and torch.optim.SGD with Nesterov momentum.  The trainer owns
 * fp32 master parameters, gradients, momentum buffers and BatchNorm running statistics, all addressed by the reference's
 * state_dict names; the caller owns inputs, outputs and the workspace. */
typedef struct MdsTrainer MdsTrainer;
typedef struct MdsTrainConfig {
    int num_classes, num_frames, stack_size, num_3d_blocks, num_3d_features, num_3d_stack_proj, expansion_3d_ratio,
        se_reduce_3d_ratio, device;
    int amp;                 /* 1: dynamic loss scaling like torch.cuda.amp.GradScaler (argus_models.py:36), 0: scale 1 */
    int nesterov;
    float drop_rate, drop_path_rate;       /* multidim_stacker.py:150-151 */
    float focal_alpha, focal_gamma;        /* src/losses.py:54-57 */
    float momentum;
    float init_scale;        /* <= 0: GradScaler default 65536 */
} MdsTrainConfig;
typedef struct MdsTrainStepArgs {
    const void* enc_feats;       /* fp16 [b][T][fh][fw][192]: mds_forward_encoder output */
    const float* targets;        /* f32 [b][num_classes] (soft labels allowed) */
    const float* dp_masks;       /* f32 [num_3d_blocks][b], 0 or 1/keep (timm drop_path); NULL: drawn from seed */
    const float* dropout_mask;   /* f32 [b][proj*T], 0 or 1/(1-p) (F.dropout); NULL: drawn from seed */
    unsigned long long seed;
    int b, fh, fw;
    float lr;
    int apply_update;            /* 0: forward + backward only (gradients stay readable through mds_train_get) */
    float* loss_out;             /* device f32 [1], optional */
    float* logits_out;           /* device f32 [b][num_classes], optional */
} MdsTrainStepArgs;
int mds_train_create(const MdsTrainConfig* cfg, MdsTrainer** out);
int mds_train_destroy(MdsTrainer* t);
/* kind 0: trainable parameters (reference parameter order), kind 1: BatchNorm running statistics */
int mds_train_num_tensors(const MdsTrainer* t, int kind);
int mds_train_tensor_info(const MdsTrainer* t, int kind, int i, const char** name, long long* numel);
int mds_train_set(MdsTrainer* t, const char* name, const float* host, long long numel);
/* what 0: value, 1: gradient of the last step (unscaled), 2: momentum buffer, 3: ModelEma average; synchronises the device */
int mds_train_get(MdsTrainer* t, const char* name, int what, float* host, long long numel);
int mds_train_commit(MdsTrainer* t, void* stream);     /* after mds_train_set: derive the fp16 GEMM operands, re-seed the EMA */
/* ModelEma.update (src/ema.py:49-57; argus_models.py:68-69): ema = decay * ema + (1 - decay) * value for every trainable
 * parameter and BatchNorm statistic, in the reference's float32 arithmetic. */
int mds_train_ema_update(MdsTrainer* t, double decay, void* stream);
size_t mds_train_workspace_bytes(const MdsTrainer* t, int b, int fh, int fw);
int mds_train_step(MdsTrainer* t, const MdsTrainStepArgs* args, void* ws, size_t ws_bytes, void* stream);
/* host4: loss scale, growth tracker, found_inf flag, optimizer steps performed; synchronises the device */
int mds_train_scaler_state(MdsTrainer* t, float* host4);
long long mds_train_batches_tracked(const MdsTrainer* t);   /* BatchNorm num_batches_tracked increment */
/* Validation step (BallActionModel.val_step, src/argus_models.py:76-91): sigmoid focal loss (src/losses.py:31-48, mean
 * reduction) of n = b * num_classes eval-mode logits and their sigmoid (prediction_transform); device pointers. */
int mds_focal_loss(const float* logits, const float* targets, int n, float alpha, float gamma, float* loss_out,
                   float* probs_out, void* stream);

/* ---- post-processing of raw predictions into action spots ----------------------------------------------------------
 * Replaces post_processing (src/utils.py:55-64) for every class at once: scipy.ndimage.gaussian_filter (1-D, 'reflect',
 * weights = scipy's _gaussian_kernel1d(sigma, 0, radius) passed in by the host, radius = int(4 * sigma + 0.5)) followed by
 * scipy.signal.find_peaks(height=height, distance=distance).  raw: device f32 [n_frames][num_classes]
 * (`raw_predictions`, scripts/ball_action/predict.py:50-55).  Outputs (device): out_index / out_conf [num_classes][n_frames]
 * hold, per class, the ascending peak positions (add frame_indexes[0], utils.py:63) and the smoothed value there;
 * out_count [num_classes].  One launch, no host synchronisation. */
size_t mds_post_processing_workspace_bytes(int n_frames, int num_classes);
int mds_post_processing(const float* raw, int n_frames, int num_classes, const double* weights_host, int radius, float height,
                        int distance, int* out_index, float* out_conf, int* out_count, void* ws, size_t ws_bytes, void* stream);

/* Programmatic dependent launch of the forward chain (default on): each kernel's launch latency, CTA scheduling and
 * constant set-up overlap the previous kernel's tail.  0 restores plain stream serialization (A/B measurements). */
int mds_set_pdl(int enabled);
/* How the MBConv tails (depthwise + SE + projection) are launched (A/B measurements; all four are parity-tested).
 * 3 (default): mds_k_dwconv_se without SE weights (TMA-staged depthwise + squeeze) + mds_k_se_fc + mds_k_gemm_gated;
 * 2: mds_k_dwconv_se (incl. the SE MLP) + mds_k_gemm_gate; 1: mds_k_mbconv_tail (ONE persistent launch);
 * 0: round-1 path mds_k_dwconv + mds_k_se_fc + mds_k_gemm_gated. */
int mds_set_tail_mode(int mode);
/* How the dense 3x3 blocks (blocks.0.0 - 2.0) keep their expanded tensor (A/B measurements, both parity-tested).  2 (default):
 * conv_tc_kernel (TMA + tcgen05) leaves it in tensor memory as the projection's A operand and folds the column taps of blocks.0.0
 * into N; 1: conv_tc_kernel stages it in shared memory. */
int mds_set_conv_mode(int mode);
/* The encoder of mds_forward / mds_forward_2d runs as n equal parts of the images on n streams (default 2: the kernels of one half
 * fill the ramps and tails of the other; handle-owned streams forked from / joined to the caller's stream; 1 = single stream). */
int mds_set_streams(int n);
/* Measurement only (bench.py `roofline_dw`): 1 = the fused tails execute their depthwise + SE items alone, so that the
 * depthwise stage can be timed by itself; forward outputs are NOT valid while this is set. */
int mds_set_tail_dw_only(int enabled);

/* number of kernels launched by this library in the calling thread since the last reset (bench "gpu_launches") */
long long mds_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* MDS_B200_H_ */
