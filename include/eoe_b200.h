/*
 * eoe_b200 -- C ABI of the B200-native anomaly-detection scoring / loss / AUC hot path.
 *
 * The reference (liznerski/eoe) is pure Python: its "FFI" for this path is the three ADTrainer hooks
 *   prepare_metric / compute_anomaly_score / loss      src/eoe/training/ad_trainer.py:624-662
 * implemented per objective in src/eoe/training/{hsc,bce,clip}.py, the image-encoder call
 *   image_features = model(imgs)                        ad_trainer.py:429,507
 * and the scikit-learn AUC call at ad_trainer.py:453-454,517-521.  Each entry point below names the
 * reference lines it replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all memory
 *   - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on it, never synchronise,
 *     never allocate, and are re-entrant as long as two in-flight calls do not share a workspace
 *   - return value 0 = success, negative = error (eoe_strerror); nothing throws across the ABI
 *   - NaN / Inf inputs propagate into scores and losses (the reference's NanGradientsError guard,
 *     ad_trainer.py:448-449, relies on that); eoe_auc reports non-finite scores in its status word
 *   - dtypes: EOE_F32 / EOE_F16 / EOE_BF16 for feature and score tensors; labels are int64
 *     (torch.long, as produced by the reference's loaders); accumulation is always fp32 (fp64 for AUC)
 */
#ifndef EOE_B200_H
#define EOE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EOE_ABI_VERSION 6

enum { EOE_F32 = 0, EOE_F16 = 1, EOE_BF16 = 2,
       /* Encoder operand dtype only ("split fp16", the precise mode): every 16-bit matrix [rows, C] is stored as
        * [rows, 2C] = [hi | lo] with hi = rn_fp16(x), lo = rn_fp16(x - hi), and every product of two stored tensors is
        * evaluated as hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM (3x the tensor work, ~2^-21 relative operand
        * error instead of 2^-12).  This is the mode whose end-to-end SCORES are within 1e-3 relative of the fp32
        * reference on every image (clip.py:66-79 on top of model.py:219-236); the softmax probabilities are pairs too.
        * Accepted by eoe_vit_* (LayerNorm-folded weights required), eoe_gemm*, eoe_vit_fold_layernorm, eoe_layernorm
        * (out_dtype) and eoe_attention (L == 197, or L <= 64 with an even head count). */
       EOE_F16X2 = 3 };

enum {
    EOE_OK = 0,
    EOE_ERR_ARG = -1,         /* null pointer / negative size / bad enum                        */
    EOE_ERR_DTYPE = -2,       /* unsupported dtype                                              */
    EOE_ERR_SHAPE = -3,       /* unsupported shape (e.g. d not a multiple of 4, K too large)     */
    EOE_ERR_ALIGN = -4,       /* pointer not aligned as required (16 B for feature rows)         */
    EOE_ERR_WORKSPACE = -5,   /* workspace missing or too small                                  */
    EOE_ERR_CUDA = -6,        /* a CUDA runtime / driver call or kernel launch failed            */
    EOE_ERR_ARCH = -7         /* device is not sm_100 (Blackwell B200)                            */
};

int eoe_abi_version(void);
const char* eoe_strerror(int code);
/* last CUDA error string recorded by this library on the calling thread ("" if none) */
const char* eoe_last_cuda_error(void);
/* number of CUDA kernels this library has launched in this process so far (bench.py's `gpu_launches`) */
long long eoe_launch_count(void);
/* identity of the kernel sources this library was built from (first 16 hex digits of a sha256 over csrc/ and this header;
 * eoe_b200/build.py:source_id).  bench.py only quotes ncu-derived figures (profiles/gemm_traffic.json) whose recorded id
 * equals this one. */
const char* eoe_build_id(void);

/* ------------------------------------------------------------------------------------------------
 * Loss / score heads.  `head_ws` is a zero-initialised device buffer of EOE_HEAD_WS_BYTES bytes
 * allocated once by the caller; kernels leave it zeroed again so it can be reused call after call.
 * ---------------------------------------------------------------------------------------------- */
#define EOE_HEAD_WS_BYTES 32768

/* HSCTrainer.loss + its autograd backward + HSCTrainer.compute_anomaly_score in ONE kernel.
 * Replaces src/eoe/training/hsc.py:17-21 (loss), :12-15 (score) and the 28-op autograd backward.
 *   z [n,d] row-major (z_dtype), labels [n] int64, nominal_label as in kwargs (ad_trainer.py:430)
 *   loss_out  [1] fp32   = mean_i( label_i==nominal ? dist_i : -log(score_i + 1e-9) )
 *   scores_out[n] fp32   = 1 - exp(-dist_i), dist_i = sqrt(||z_i||^2 + 1) - 1        (nullable)
 *   grad_z_out[n,d] z_dtype = d loss / d z   (for upstream gradient 1)                  (nullable)
 * d % 4 == 0 and 16-byte aligned rows take the vector path; any other d takes a scalar path. */
int eoe_hsc_fwd_bwd(const void* z, int z_dtype, const int64_t* labels, int64_t n, int64_t d,
                    int64_t nominal_label, float* loss_out, float* scores_out, void* grad_z_out,
                    void* head_ws, void* stream);

/* HSCTrainer.compute_anomaly_score alone (test time, ad_trainer.py:508). hsc.py:12-15 */
int eoe_hsc_score(const void* z, int z_dtype, int64_t n, int64_t d, float* scores_out, void* stream);

/* BCETrainer.loss + backward + compute_anomaly_score. Replaces src/eoe/training/bce.py:15-20.
 *   x [n] (= features [n,1] squeezed), labels [n] int64 used raw as float targets (bce.py:20)
 *   loss_out [1] fp32 = mean( max(x,0) - x*y + log1p(exp(-|x|)) )
 *   scores_out [n] fp32 = sigmoid(x), or 1 - sigmoid(x) if nominal_label != 0 (bce.py:17)  (nullable)
 *   grad_x_out [n] x_dtype = (sigmoid(x) - y)/n                                           (nullable) */
int eoe_bce_fwd_bwd(const void* x, int x_dtype, const int64_t* labels, int64_t n, int64_t nominal_label,
                    float* loss_out, float* scores_out, void* grad_x_out, void* head_ws, void* stream);

int eoe_bce_score(const void* x, int x_dtype, int64_t n, int64_t nominal_label, float* scores_out,
                  void* stream);

/* DSADTrainer.loss + backward + compute_anomaly_score. Replaces src/eoe/training/dsad.py:13-22.
 *   s_i = norm(z_i)^2 (square root then square, as the reference evaluates it)
 *   loss_out [1] = mean_i( label_i==nominal ? s_i : 1/(s_i + 1e-9) );  grad = 2 z/n  |  -2 z/((s+1e-9)^2 n)
 *   scores_out [n] = 1 - exp(-(sqrt(s_i + 1) - 1))   (dsad.py:13-16; identical to eoe_hsc_score)      (nullable) */
int eoe_dsad_fwd_bwd(const void* z, int z_dtype, const int64_t* labels, int64_t n, int64_t d,
                     int64_t nominal_label, float* loss_out, float* scores_out, void* grad_z_out,
                     void* head_ws, void* stream);

/* DSVDDTrainer.loss + backward + compute_anomaly_score. Replaces src/eoe/training/dsvdd.py:23-27.
 *   center [d] fp32 (the [1,d] tensor returned by prepare_metric, dsvdd.py:11-21), 16-byte aligned for the vector path
 *   scores_out [n] = sum_j (z_ij - c_j)^2;  loss_out [1] = mean(scores);  grad = 2 (z - c)/n
 *   any of loss_out / scores_out / grad_z_out may be null (score only: test time, ad_trainer.py:508) */
int eoe_dsvdd_fwd_bwd(const void* z, int z_dtype, const float* center, int64_t n, int64_t d,
                      float* loss_out, float* scores_out, void* grad_z_out, void* head_ws, void* stream);

/* FocalTrainer.loss (FocalLoss, gamma = 2, eps = 1e-7) + backward + compute_anomaly_score.
 * Replaces src/eoe/training/focal.py:11-39.
 *   bce_i = binary_cross_entropy_with_logits(x_i, y_i); pt = clamp(exp(-bce), eps, 1-eps)
 *   loss_out [1] = mean( (1 - pt)^gamma * bce );  scores_out [n] = sigmoid(x) or 1 - sigmoid(x) (focal.py:33-35)
 *   grad_x_out [n] = d loss / d x (clamp passes the gradient on the closed interval, as torch.clamp does) */
int eoe_focal_fwd_bwd(const void* x, int x_dtype, const int64_t* labels, int64_t n, int64_t nominal_label,
                      float gamma, float eps, float* loss_out, float* scores_out, void* grad_x_out,
                      void* head_ws, void* stream);

/* ADClipTrainer.compute_anomaly_score. Replaces src/eoe/training/clip.py:66-79.
 *   z [n,d] image features, text [K,d] fp32 (`center`); text rows are re-normalised (clip.py:69)
 *   scores_out [n] fp32 = softmax_k(scale * z^_i . T^_k)[K-1]     (scale = 100, clip.py:71)
 * Limits: d % 4 == 0, d <= 1024, K <= 256 prompts (leave_one_out on cub builds 200, clip.py:53-54).  K <= 32 and
 * n >= 2048 run on tensor cores, K <= 64 keeps the text rows in shared memory, larger prompt sets read them through L2. */
int eoe_clip_score(const void* z, int z_dtype, const float* text, int64_t n, int64_t d, int64_t K,
                   float scale, float* scores_out, void* stream);

/* ADClipTrainer.loss + backward w.r.t. image features. Replaces src/eoe/training/clip.py:81-103.
 *   text used as given (not re-normalised, clip.py:82,86); leave_one_out != 0 selects clip.py:93-98
 *   loss_out [1] fp32; grad_z_out [n,d] z_dtype (nullable). Rows whose label is neither nominal nor
 *   1-nominal contribute 0 but count in the mean (clip.py:90,96,102). */
int eoe_clip_oe_loss_fwd_bwd(const void* z, int z_dtype, const float* text, const int64_t* labels,
                             int64_t n, int64_t d, int64_t K, float scale, int64_t nominal_label,
                             int leave_one_out, float* loss_out, void* grad_z_out, void* head_ws,
                             void* stream);

/* ------------------------------------------------------------------------------------------------
 * ROC-AUC (and PRC / AP) on the device, bit-exact with scikit-learn's
 *   fpr, tpr, thr = roc_curve(labels, scores); auc(fpr, tpr)         ad_trainer.py:453-454,517-518
 *   precision_recall_curve / average_precision_score                  ad_trainer.py:520-521
 * Device radix sort of (descending score key, label bit) + tie / corner scans + fp64 trapezoid terms
 * summed in numpy's pairwise order.
 * ---------------------------------------------------------------------------------------------- */
enum {
    EOE_AUC_IGNORE_NEGATIVE_LABELS = 1,   /* drop rows with label < 0 (ad_trainer.py:517 filter)   */
    EOE_AUC_WITH_PRC = 2,                 /* also compute average precision (+ PRC arrays if given) */
    EOE_AUC_FORCE_TILED = 4,              /* n <= EOE_AUC_SINGLE_LAUNCH_MAX: use the multi-kernel pipeline anyway (tests) */
    EOE_AUC_FORCE_SINGLE_CTA = 8,         /* n <= 16 384: the one-CTA kernel (tests, A/B timing)                          */
    EOE_AUC_FORCE_CLUSTER = 16            /* n <= 131 072: the 8-CTA cluster kernel (tests, A/B timing)                   */
};
#define EOE_AUC_SINGLE_LAUNCH_MAX 49152   /* up to here eoe_auc is ONE kernel launch: one CTA up to 12 288 scores (the
                                             reference's sizes: 3 000 - 10 000 per class and epoch, ad_trainer.py:452-455,
                                             516-522), a thread-block cluster of 8 CTAs sorting through distributed shared
                                             memory above that; larger inputs run the tiled radix-sort pipeline (11+ launches) */
enum {                                    /* bits of info_out[4]                                    */
    EOE_AUC_STATUS_NONFINITE = 1,         /* a kept score is NaN/Inf (sklearn raises ValueError)     */
    EOE_AUC_STATUS_SINGLE_CLASS = 2       /* only one class present: AUC undefined (NaN)             */
};
size_t eoe_auc_workspace_bytes(int64_t n);
/*   scores [n] (score_dtype), labels [n] int64 (positive class == 1)
 *   auc_out  [2] fp64 device: [0] = ROC-AUC, [1] = average precision (if EOE_AUC_WITH_PRC)
 *   info_out [8] int64 device: [0] n kept, [1] n positives, [2] n distinct scores,
 *                               [3] n ROC points (incl. the prepended origin), [4] status bits
 *   fpr_out, tpr_out [n+1] fp64, thr_out [n+1] fp32 (thr[0] = +inf): ROC curve, nullable (all three or none)
 *   prec_out, rec_out [n+1] fp64: PRC in sklearn's (reversed, (1,0)-terminated) order, nullable
 *   prc_thr_out [n] fp32: precision_recall_curve's thresholds (the distinct scores, increasing; info_out[2] of them),
 *                         nullable; needs EOE_AUC_WITH_PRC (what eoe.utils.logger.PRC.ths holds, logger.py:65-91) */
int eoe_auc(const void* scores, int score_dtype, const int64_t* labels, int64_t n, int flags,
            void* workspace, size_t workspace_bytes, double* auc_out, int64_t* info_out,
            double* fpr_out, double* tpr_out, float* thr_out, double* prec_out, double* rec_out,
            float* prc_thr_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * CLIP ViT-B image encoder forward (+ optional fused zero-shot score head).
 * Replaces VisualTransformer.forward  src/eoe/models/clip_official/clip/model.py:219-236
 * (ResidualAttentionBlock :167-188, LayerNorm :153-159, QuickGELU :162-164, encode_image :336-337)
 * followed, when `text` is given, by ADClipTrainer.compute_anomaly_score (training/clip.py:66-79).
 * GEMMs run on tcgen05 tensor cores (16-bit operands, fp32 TMEM accumulators, TMA-fed); the
 * residual stream, LayerNorm and softmax statistics are fp32.
 * ---------------------------------------------------------------------------------------------- */
typedef struct eoe_vit_layer {
    const float* ln_1_w; const float* ln_1_b;          /* [width]                                  */
    const void*  in_proj_w;  const float* in_proj_b;   /* [3*width, width] operand dtype, [3*width] */
    const void*  out_proj_w; const float* out_proj_b;  /* [width, width], [width]                  */
    const float* ln_2_w; const float* ln_2_b;
    const void*  c_fc_w;   const float* c_fc_b;        /* [4*width, width], [4*width]              */
    const void*  c_proj_w; const float* c_proj_b;      /* [width, 4*width], [width]                */
    /* Optional LayerNorm-folded copies produced by eoe_vit_fold_layernorm (all six or none, and for every
     * layer or for none).  When present the encoder drops the stand-alone ln_1 / ln_2 kernels: the residual
     * GEMM epilogues emit a 16-bit copy of the residual stream plus per-row (sum, sum of squares) and the
     * QKV / c_fc GEMMs apply  rstd*(acc - mean*c1) + c2  in their epilogue (see DESIGN.md "LayerNorm fold"). */
    const void*  in_proj_wf; const float* in_proj_c1; const float* in_proj_c2;   /* [3*width,width], [3*width] x2 */
    const void*  c_fc_wf;    const float* c_fc_c1;    const float* c_fc_c2;      /* [4*width,width], [4*width] x2 */
    /* Optional (only used together with the folded tensors): c_proj weights pre-divided by 1.702, operand dtype,
     * rounded once from the fp32 master.  When present the c_fc epilogue emits 1.702 * QuickGELU (see
     * EOE_EPI_LNFOLD_QUICKGELU_X1702) and c_proj multiplies by these weights -- same product, fewer instructions. */
    const void*  c_proj_w_div1702;
} eoe_vit_layer;

typedef struct eoe_vit_weights {
    int32_t patch;            /* 32 or 16                                                           */
    int32_t resolution;       /* 224                                                                */
    int32_t width;            /* 768                                                                */
    int32_t heads;            /* 12 (head dim must be 64)                                           */
    int32_t n_layers;         /* 12                                                                 */
    int32_t embed_dim;        /* 512                                                                */
    int32_t operand_dtype;    /* EOE_BF16, EOE_F16 or EOE_F16X2 (every *_w matrix [N, K] is then [N, 2K] = [hi | lo]) */
    int32_t reserved;
    const void*  conv1_w;     /* [width, 3*patch*patch]  (visual.conv1.weight flattened)            */
    const float* class_embedding;       /* [width]                                                  */
    const float* positional_embedding;  /* [L, width], L = (resolution/patch)^2 + 1                 */
    const float* ln_pre_w;  const float* ln_pre_b;
    const float* ln_post_w; const float* ln_post_b;
    const float* proj;        /* [width, embed_dim] fp32 (visual.proj)                              */
    const eoe_vit_layer* layers_host;   /* HOST array of n_layers entries holding DEVICE pointers   */
} eoe_vit_weights;

typedef struct eoe_vit_plan eoe_vit_plan;

/* Builds TMA descriptors / launch geometry for batches of up to max_batch images whose activations
 * live in `workspace` (device, eoe_vit_workspace_bytes(...) bytes, 1024-byte aligned). The plan keeps
 * pointers to the weights and the workspace; both must outlive it. Host-side only, no GPU work. */
size_t eoe_vit_workspace_bytes(const eoe_vit_weights* w_host, int64_t max_batch);
int eoe_vit_plan_create(const eoe_vit_weights* w_host, int64_t max_batch, void* workspace,
                        size_t workspace_bytes, eoe_vit_plan** plan_out);
void eoe_vit_plan_destroy(eoe_vit_plan* plan);

/*   imgs [B,3,R,R] NCHW fp32 (already normalised, as handed to model(imgs) at ad_trainer.py:507), B <= max_batch
 *   feats_out [B, embed_dim] fp32                                        (nullable)
 *   text [K, embed_dim] fp32 + scores_out [B] fp32: fused clip.py:66-79   (nullable together) */
int eoe_vit_encode(eoe_vit_plan* plan, const float* imgs, int64_t B, float* feats_out,
                   const float* text, int64_t K, float scale, float* scores_out, void* stream);

/* Same forward pass from RAW uint8 pixels: torchvision's ToTensor (u8/255) and Normalize ((x-mean)/std), i.e. the tail of
 * CLIP's `_transform` (clip_official/clip/clip.py:58-65) and eoe's GPU Normalize (utils/transformations.py:126-138), are
 * fused into the patchify kernel -- bit-identical to normalising first and calling eoe_vit_encode, at a quarter of the
 * input bytes.  layout: EOE_LAYOUT_NCHW imgs [B,3,R,R] | EOE_LAYOUT_NHWC imgs [B,R,R,3]; mean/std: HOST float[3]. */
enum { EOE_LAYOUT_NCHW = 0, EOE_LAYOUT_NHWC = 1 };
int eoe_vit_encode_u8(eoe_vit_plan* plan, const uint8_t* imgs, int layout, const float* mean_host,
                      const float* std_host, int64_t B, float* feats_out, const float* text, int64_t K,
                      float scale, float* scores_out, void* stream);

/* The whole of CLIP's `_transform` (clip_official/clip/clip.py:58-65) in front of the encoder, on the device:
 *   Resize(R, BICUBIC) -> CenterCrop(R) -> ToTensor -> Normalize -> patchify,   imgs [B, H, W, 3] uint8 of ANY size H x W.
 * Resize reproduces Pillow's fixed-point two-pass resampler (libImaging/Resample.c) and torchvision's size / crop rules
 * bit for bit (oracle/resize.py is pinned against both), so the features equal those of the reference pipeline run on
 * the host followed by eoe_vit_encode_u8.  Limit: one row of patches may touch at most ~200 KB of resampled source rows
 * (down-scaling factors up to ~12). */
int eoe_vit_encode_u8_resize(eoe_vit_plan* plan, const uint8_t* imgs, int64_t H, int64_t W, const float* mean_host,
                             const float* std_host, int64_t B, float* feats_out, const float* text, int64_t K,
                             float scale, float* scores_out, void* stream);
/* torchvision's geometry for Resize(n_px) + CenterCrop(n_px) of an H x W image:
 * out6_host = {resized height, resized width, crop top, crop left, horizontal taps, vertical taps}. */
int eoe_resize_geometry(int64_t H, int64_t W, int n_px, int* out6_host);

/* Optional instrumentation for roofline reporting: while enabled, eoe_vit_encode brackets every GEMM launch with a
 * CUDA event pair on `stream` (no synchronisation). eoe_vit_profile_read waits for the recorded events and returns,
 * per GEMM kind (0 patch-embed, 1 qkv, 2 out-proj, 3 c_fc, 4 c_proj), accumulated milliseconds, launches and
 * algorithmic FLOPs (2*M*N*K) since the last read. HOST pointers to 5 entries each. */
int eoe_vit_profile_enable(eoe_vit_plan* plan, int enable);
int eoe_vit_profile_read(eoe_vit_plan* plan, double* ms_out_host, int64_t* launches_out_host, double* flops_out_host);

/* LayerNorm fold (model.py:153-159 applied in front of a Linear, model.py:171,174):
 *   LN(x) @ W^T + b  ==  rstd * ( x @ (W*ln_w)^T  -  mean * c1 )  +  c2
 *   w_folded_out [N,K] operand dtype = round(W[n,k] * ln_w[k]);  c1[n] = sum_k float(w_folded[n,k]);
 *   c2[n] = sum_k W[n,k] * ln_b[k] + bias[n].          w_f32 [N,K] fp32 master weights, K % 4 == 0. */
int eoe_vit_fold_layernorm(const float* w_f32, const float* ln_w, const float* ln_b, const float* bias,
                           int64_t N, int64_t K, int operand_dtype, void* w_folded_out, float* c1_out,
                           float* c2_out, void* stream);

/* Diagnostics for tools/gemm_probe.py (NOT part of the stable ABI; 0 in production): bit 0 GEMM epilogues only release
 * their accumulators, bit 1 no global stores, bit 3 the last block computes Q for every token again (A/B of the class-token-only
 * Q GEMM), bit 4 clusters of two CTA pairs with W multicast, bits 8.. grid size in CTA pairs. */
void eoe_debug_set(int flags);

/* Building blocks of the encoder, exported so that each kernel is parity-tested through the ABI. */
enum { EOE_EPI_BIAS = 0, EOE_EPI_BIAS_QUICKGELU = 1, EOE_EPI_BIAS_RESIDUAL_F32 = 2, EOE_EPI_PATCH_EMBED = 3,
       EOE_EPI_LNFOLD_BIAS = 4, EOE_EPI_LNFOLD_QUICKGELU = 5, EOE_EPI_RESIDUAL_STATS = 6,
       EOE_EPI_LNFOLD_QUICKGELU_X1702 = 8 /* 7 is internal */ };
/* out = epilogue(A[M,K] @ W[N,K]^T): tcgen05 GEMM, A/W operand dtype (BF16/F16), K % 64 == 0, N % 256 == 0.
 *   EOE_EPI_BIAS / _QUICKGELU: out [M,N] operand dtype;  _RESIDUAL_F32: out [M,N] fp32 += (in place);
 *   _PATCH_EMBED: out fp32 row (m/g2)*(g2+1)+1+(m%g2) = acc + pos_emb[1+m%g2] (aux = pos_emb, aux_i = g2). */
int eoe_gemm(const void* A, const void* W, const float* bias, void* out, int64_t M, int64_t N, int64_t K,
             int operand_dtype, int epilogue, const float* aux, int64_t aux_i, void* stream);
/* LayerNorm-folded GEMM (QKV / c_fc with ln_1 / ln_2 folded in): A [M,K] = 16-bit copy of the residual stream MINUS
 * shift[m] (null: no shift), Wf / c1 / c2 from eoe_vit_fold_layernorm, stats [M, K/128] float2 = per-row (sum, sum of
 * squares) of the fp32 residual stream over each 128-column chunk.
 * out [M,N] operand dtype = rstd*(A@Wf^T - (mean - shift)*c1) + c2, then
 * QuickGELU if quick_gelu == 1; quick_gelu == 2 emits 1.702 * QuickGELU (for a consumer whose weights are pre-divided by
 * 1.702: two FP32 multiplies fewer per element).  K % 256 == 0, K <= 768; c1, c2, stats, shift 16-byte aligned. */
int eoe_gemm_lnfold(const void* A, const void* Wf, const float* c1, const float* c2, const float* stats,
                    const float* shift, void* out, int64_t M, int64_t N, int64_t K, int operand_dtype, int quick_gelu,
                    void* stream);
/* Residual GEMM that also prepares the next folded LayerNorm: x [M,N] fp32 += A@W^T + bias (in place),
 * stats_out [M, N/128] float2 = per-row (sum, sum of squares) per chunk of the updated x,
 * shift_out [M] (nullable) = the row's mean BEFORE the update, taken from stats_in [M, N/128] (the previous producer's
 * sums; null: 0; must not alias stats_out), xb_out [M,N] operand dtype = round(x - shift): the 16-bit copy is rounded
 * around the row's (previous) mean, as LayerNorm's own output would be. */
int eoe_gemm_residual_stats(const void* A, const void* W, const float* bias, const float* stats_in, float* x, void* xb_out,
                            float* stats_out, float* shift_out, int64_t M, int64_t N, int64_t K, int operand_dtype,
                            void* stream);
/* y[M,width] (out_dtype) = LayerNorm_fp32(x[M,width]) * w + b, eps 1e-5 (model.py:153-159) */
int eoe_layernorm(const float* x, const float* w, const float* b, void* y, int out_dtype, int64_t M,
                  int64_t width, void* stream);
/* softmax(Q K^T / sqrt(64)) V per (image, head); qkv [B*L, 3*width] operand dtype (q;k;v column blocks,
 * model.py:171 in_proj layout), out [B*L, width] operand dtype */
int eoe_attention(const void* qkv, void* out, int64_t B, int64_t L, int64_t heads, int operand_dtype,
                  void* stream);

/* ---- CLIP text tower building blocks (CLIP.encode_text, clip_official/clip/model.py:339-352; run once per class by
 * ADClipTrainer.prepare_metric, training/clip.py:50-64).  The blocks between these calls are eoe_layernorm, eoe_gemm and
 * eoe_attention_causal; eoe_b200/text_encoder.py composes them. ---- */
/* eoe_attention with the text tower's causal mask (model.py:324-331: key j is visible to query i iff j <= i). L <= 208. */
int eoe_attention_causal(const void* qkv, void* out, int64_t B, int64_t L, int64_t heads, int operand_dtype,
                         void* stream);
/* x [n*ctx, width] fp32 = token_embedding[tokens] + positional_embedding (model.py:340-342).  tokens [n, ctx] int64;
 * ids outside [0, vocab) produce NaN rows.  width % 4 == 0, fp32 pointers 16-byte aligned. */
int eoe_text_embed(const int64_t* tokens, const float* token_embedding, const float* positional_embedding, float* x,
                   int64_t n, int64_t ctx, int64_t width, int64_t vocab, void* stream);
/* feats [n, embed] fp32 = ln_final(x[i, argmax_j tokens[i, j]]) @ text_projection [width, embed] (model.py:346-350;
 * first position of the largest id, as torch.argmax).  All fp32. */
int eoe_text_tail(const float* x, const int64_t* tokens, const float* ln_w, const float* ln_b, const float* proj,
                  float* feats, int64_t n, int64_t ctx, int64_t width, int64_t embed, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EOE_B200_H */
