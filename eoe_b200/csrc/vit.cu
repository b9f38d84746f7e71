// CLIP ViT-B image encoder forward on sm_100a: host orchestration + the non-GEMM kernels.
// Reference: VisualTransformer.forward, src/eoe/models/clip_official/clip/model.py:219-236
//   conv1 patchify (:220) -> im2col_kernel + tcgen05 GEMM with fused positional-embedding epilogue
//   class token + pos-emb + ln_pre (:223-225) -> layernorm_kernel (cls rows synthesised in place)
//   12 x ResidualAttentionBlock (:185-188): LN -> QKV GEMM -> attention_kernel -> out-proj GEMM (+residual)
//                                           LN -> c_fc GEMM (+QuickGELU) -> c_proj GEMM (+residual)
//   ln_post(x[:,0]) @ proj (:231-234) -> tail_kernel, then optionally the zero-shot score head (training/clip.py:66-79)
// Token layout is [B, L, width] (batch major), residual stream fp32, GEMM operands bf16/fp16.
#include <mutex>
#include <new>
#include <vector>

#include "attention_sm100.cuh"
#include "clip_head_sm100.cuh"
#include "gemm_sm100.cuh"

namespace eoe {

int clip_prompts_ok(int64_t K);
int clip_score_f32(const float* z, const float* text, int64_t n, int64_t d, int64_t K, float scale, float* scores,
                   cudaStream_t st);

// ------------------------------------------------------------------------------------------ TMA descriptors
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
        else set_cuda_error(e, "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)");
    });
    return fn;
}

// [rows, K] row-major 16-bit matrix, box = box_rows x 64 columns (128 bytes), SWIZZLE_128B, OOB rows read as zero
static int make_tmap(CUtensorMap* m, const void* base, int64_t rows, int64_t K, int box_rows, int dtype) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return EOE_ERR_CUDA;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, dtype == EOE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled");
        return EOE_ERR_CUDA;
    }
    return EOE_OK;
}

// the same for a [rows, cols] window of a wider row-major matrix (row stride `ld` elements): K and V columns of qkv
static int make_tmap_ld(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int dtype) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return EOE_ERR_CUDA;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, dtype == EOE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(ld)");
        return EOE_ERR_CUDA;
    }
    return EOE_OK;
}

// out [B, L, width] 16-bit as a 3-D map (width, token, image): a box of `box_rows` tokens x 64 columns that runs past the
// image's last token is clipped by the TMA unit (stores) / zero-filled (loads)
static int make_tmap_tokens(CUtensorMap* m, const void* base, int64_t B, int64_t L, int64_t width, int box_rows, int dtype) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return EOE_ERR_CUDA;
    cuuint64_t dims[3] = {(cuuint64_t)width, (cuuint64_t)L, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)width * 2, (cuuint64_t)L * width * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(m, dtype == EOE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled(3d)");
        return EOE_ERR_CUDA;
    }
    return EOE_OK;
}

static int num_sms();
static int g_gemm_debug = 0;            // eoe_debug_set(): diagnostics only, 0 in production

// CLIP score head for many 16-bit rows on tcgen05 (clip_head_sm100.cuh); called by heads.cu's dispatch.  z [n, d] fp16 / bf16
// (d % 64 == 0, d <= 512, 16-byte aligned), text [K, d] fp32 (K <= 32), scores [n] fp32.
int clip_score_tc16(const void* z, int dtype, const float* text, int64_t n, int64_t d, int64_t K, float scale, float* scores,
                    cudaStream_t st) {
    CUtensorMap tm;
    int rc = make_tmap(&tm, z, n, d, 128, dtype);
    if (rc) return rc;
    const int64_t tiles = (n + 127) / 128;
    const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
#define EOE_CLIP_TC(BF)                                                                                                   \
    {                                                                                                                     \
        auto kern = cliptc::clip_score_tc_kernel<BF>;                                                                     \
        static bool attr_done = false;                                                                                    \
        if (!attr_done) {                                                                                                 \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cliptc::SMEM_BYTES); \
            if (e != cudaSuccess) { set_cuda_error(e, "clip_score_tc smem attr"); return EOE_ERR_CUDA; }                  \
            attr_done = true;                                                                                             \
        }                                                                                                                 \
        kern<<<grid, cliptc::THREADS, cliptc::SMEM_BYTES, st>>>(tm, text, n, (int)d, (int)K, scale, scores);             \
    }
    if (dtype == EOE_BF16) EOE_CLIP_TC(true) else EOE_CLIP_TC(false)
#undef EOE_CLIP_TC
    return check_launch("clip_score_tc_kernel");
}

// CLIP OE loss + backward for many 16-bit rows on tcgen05 (clip_head_sm100.cuh); grad [n, d] in the dtype of z.
int clip_loss_tc16(const void* z, int dtype, const float* text, const int64_t* labels, int64_t n, int64_t d, int64_t K,
                   float scale, int64_t nominal, int loo, float* loss_out, void* grad, void* ws, cudaStream_t st) {
    CUtensorMap tm_z, tm_g;
    int rc = make_tmap(&tm_z, z, n, d, 128, dtype);
    if (!rc) rc = make_tmap(&tm_g, grad, n, d, 32, dtype);
    if (rc) return rc;
    const int64_t tiles = (n + 127) / 128;
    const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    const float inv_n_f = 1.0f / (float)n;
    const double inv_n = 1.0 / (double)n;
#define EOE_CLIPL_TC(BF)                                                                                                  \
    {                                                                                                                     \
        auto kern = cliptc::clip_oe_loss_tc_kernel<BF>;                                                                   \
        static bool attr_done = false;                                                                                    \
        if (!attr_done) {                                                                                                 \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cliptc::LOSS_SMEM_BYTES); \
            if (e != cudaSuccess) { set_cuda_error(e, "clip_oe_loss_tc smem attr"); return EOE_ERR_CUDA; }                \
            attr_done = true;                                                                                             \
        }                                                                                                                 \
        kern<<<grid, cliptc::LOSS_THREADS, cliptc::LOSS_SMEM_BYTES, st>>>(tm_z, tm_g, text, labels, n, (int)d, (int)K, scale,   \
                                                                     nominal, loo, (HeadWorkspace*)ws, loss_out, inv_n_f, inv_n, (g_gemm_debug >> 20) & 7); \
    }
    if (dtype == EOE_BF16) EOE_CLIPL_TC(true) else EOE_CLIPL_TC(false)
#undef EOE_CLIPL_TC
    return check_launch("clip_oe_loss_tc_kernel");
}

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = kNumSMs;
    }
    return n;
}


static int g_last_max_clusters[2] = {0, 0};

template <int EPI, bool BF16, int CLP, bool SPLIT = false>
static int gemm_launch_c(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const gemm::Params& p, cudaStream_t st) {
    auto kern = gemm::gemm_kernel<EPI, BF16, CLP, SPLIT>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm::Cfg<EPI, SPLIT>::kSmemBytes);
        if (e != cudaSuccess) { set_cuda_error(e, "gemm smem attr"); return EOE_ERR_CUDA; }
        attr_done = true;
    }
    const int64_t units = (((p.M + gemm::BM - 1) / gemm::BM + CLP - 1) / CLP) * (p.N / gemm::BN);
    // persistent grid = the number of clusters that can be co-resident (GPC boundaries can leave SMs unusable for
    // clusters of 4: the occupancy query knows)
    static int max_clusters = 0;
    if (max_clusters == 0) {
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3((unsigned)(num_sms() / (2 * CLP) * 2 * CLP), 1, 1);
        q.blockDim = dim3(gemm::THREADS, 1, 1);
        q.dynamicSmemBytes = gemm::Cfg<EPI, SPLIT>::kSmemBytes;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = 2 * CLP; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
        q.attrs = qa; q.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess || n <= 0) { cudaGetLastError(); n = num_sms() / (2 * CLP); }
        max_clusters = n < num_sms() / (2 * CLP) ? n : num_sms() / (2 * CLP);
        g_last_max_clusters[CLP - 1] = max_clusters;
    }
    int64_t clusters = max_clusters;
    if (g_gemm_debug >> 8) clusters = (g_gemm_debug >> 8) / CLP > 0 ? (g_gemm_debug >> 8) / CLP : 1;   // diagnostics: restrict the grid
    if (units < clusters) clusters = units;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * 2 * CLP), 1, 1);
    cfg.blockDim = dim3(gemm::THREADS, 1, 1);
    cfg.dynamicSmemBytes = gemm::Cfg<EPI, SPLIT>::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2 * CLP;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // prologue overlaps the previous kernel's tail
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (g_gemm_debug & 32) ? 1 : 2;                           // diagnostics bit 5: no dependent launch
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, p);
    if (e != cudaSuccess) { set_cuda_error(e, "gemm_kernel launch"); return EOE_ERR_CUDA; }
    return check_launch("gemm_kernel");
}

// Default: one CTA pair per cluster (all 148 SMs).  Diagnostics bit 4 (16) selects two pairs per cluster sharing the W
// tile by TMA multicast: -25 % operand traffic and +8 % per SM, but only 33 clusters of 4 fit the GPCs (132 SMs), so it
// does not pay on B200 (profiles/r1_gemm_probe_multicast.json).
template <int EPI, bool BF16, bool SPLIT = false>
static int gemm_launch_t(const CUtensorMap& ta, const CUtensorMap& tb, const gemm::Params& p_in, cudaStream_t st,
                         const CUtensorMap* tc) {
    gemm::Params p = p_in;
    p.dbg = g_gemm_debug & 7;
    // 16-bit outputs leave through TMA tile stores: box 32 rows x 64 columns over out [M, N] (rows >= M are clipped);
    // SPLIT: out is [M, 2N] = [hi | lo]
    CUtensorMap local;
    if (gemm::Cfg<EPI>::kOut16 && !tc) {
        int rc = make_tmap(&local, p.out, p.M, SPLIT ? 2 * p.N : p.N, 32, BF16 ? EOE_BF16 : EOE_F16);
        if (rc) return rc;
        tc = &local;
    }
    if (!tc) tc = &ta;                                       // unused by the fp32-output epilogues
    if (SPLIT) return gemm_launch_c<EPI, false, 1, SPLIT>(ta, tb, *tc, p, st);
    if (g_gemm_debug & 16) return gemm_launch_c<EPI, BF16, 2>(ta, tb, *tc, p, st);
    return gemm_launch_c<EPI, BF16, 1>(ta, tb, *tc, p, st);
}

// operand dtype EOE_F16X2: the same epilogues over split fp16 operands (gemm::Cfg)
static int gemm_launch_split(const CUtensorMap& ta, const CUtensorMap& tb, const gemm::Params& p, int epi, cudaStream_t st,
                             const CUtensorMap* tc) {
    switch (epi) {
        case EOE_EPI_BIAS: return gemm_launch_t<EOE_EPI_BIAS, false, true>(ta, tb, p, st, tc);
        case EOE_EPI_BIAS_QUICKGELU: return gemm_launch_t<EOE_EPI_BIAS_QUICKGELU, false, true>(ta, tb, p, st, tc);
        case EOE_EPI_BIAS_RESIDUAL_F32: return gemm_launch_t<EOE_EPI_BIAS_RESIDUAL_F32, false, true>(ta, tb, p, st, tc);
        case EOE_EPI_PATCH_EMBED: return gemm_launch_t<EOE_EPI_PATCH_EMBED, false, true>(ta, tb, p, st, tc);
        case EOE_EPI_LNFOLD_BIAS: return gemm_launch_t<EOE_EPI_LNFOLD_BIAS, false, true>(ta, tb, p, st, tc);
        case EOE_EPI_LNFOLD_QUICKGELU: return gemm_launch_t<EOE_EPI_LNFOLD_QUICKGELU, false, true>(ta, tb, p, st, tc);
        case EOE_EPI_LNFOLD_QUICKGELU_X1702: return gemm_launch_t<EOE_EPI_LNFOLD_QUICKGELU_X1702, false, true>(ta, tb, p, st, tc);
        case EOE_EPI_RESIDUAL_STATS:
            if (p.K <= 1024) return gemm_launch_t<gemm::EPI_RESIDUAL_STATS_ASYNC, false, true>(ta, tb, p, st, tc);
            return gemm_launch_t<EOE_EPI_RESIDUAL_STATS, false, true>(ta, tb, p, st, tc);
        default: return EOE_ERR_ARG;
    }
}

static int gemm_launch(const CUtensorMap& ta, const CUtensorMap& tb, const gemm::Params& p, int dtype, int epi,
                       cudaStream_t st, const CUtensorMap* tc = nullptr) {
    if (dtype == EOE_F16X2) return gemm_launch_split(ta, tb, p, epi, st, tc);
    const bool bf = dtype == EOE_BF16;
    switch (epi) {
        case EOE_EPI_BIAS: return bf ? gemm_launch_t<EOE_EPI_BIAS, true>(ta, tb, p, st, tc) : gemm_launch_t<EOE_EPI_BIAS, false>(ta, tb, p, st, tc);
        case EOE_EPI_BIAS_QUICKGELU: return bf ? gemm_launch_t<EOE_EPI_BIAS_QUICKGELU, true>(ta, tb, p, st, tc) : gemm_launch_t<EOE_EPI_BIAS_QUICKGELU, false>(ta, tb, p, st, tc);
        case EOE_EPI_BIAS_RESIDUAL_F32: return bf ? gemm_launch_t<EOE_EPI_BIAS_RESIDUAL_F32, true>(ta, tb, p, st, tc) : gemm_launch_t<EOE_EPI_BIAS_RESIDUAL_F32, false>(ta, tb, p, st, tc);
        case EOE_EPI_PATCH_EMBED: return bf ? gemm_launch_t<EOE_EPI_PATCH_EMBED, true>(ta, tb, p, st, tc) : gemm_launch_t<EOE_EPI_PATCH_EMBED, false>(ta, tb, p, st, tc);
        case EOE_EPI_LNFOLD_BIAS: return bf ? gemm_launch_t<EOE_EPI_LNFOLD_BIAS, true>(ta, tb, p, st, tc) : gemm_launch_t<EOE_EPI_LNFOLD_BIAS, false>(ta, tb, p, st, tc);
        case EOE_EPI_LNFOLD_QUICKGELU: return bf ? gemm_launch_t<EOE_EPI_LNFOLD_QUICKGELU, true>(ta, tb, p, st, tc) : gemm_launch_t<EOE_EPI_LNFOLD_QUICKGELU, false>(ta, tb, p, st, tc);
        case EOE_EPI_LNFOLD_QUICKGELU_X1702: return bf ? gemm_launch_t<EOE_EPI_LNFOLD_QUICKGELU_X1702, true>(ta, tb, p, st, tc) : gemm_launch_t<EOE_EPI_LNFOLD_QUICKGELU_X1702, false>(ta, tb, p, st, tc);
        case EOE_EPI_RESIDUAL_STATS:
            if (p.K <= 1024)      // HBM-bound shapes (out_proj): cp.async residual pipeline; see gemm::Cfg
                return bf ? gemm_launch_t<gemm::EPI_RESIDUAL_STATS_ASYNC, true>(ta, tb, p, st, tc) : gemm_launch_t<gemm::EPI_RESIDUAL_STATS_ASYNC, false>(ta, tb, p, st, tc);
            return bf ? gemm_launch_t<EOE_EPI_RESIDUAL_STATS, true>(ta, tb, p, st, tc) : gemm_launch_t<EOE_EPI_RESIDUAL_STATS, false>(ta, tb, p, st, tc);
        default: return EOE_ERR_ARG;
    }
}

static int gemm_check(int64_t M, int64_t N, int64_t K, int dtype) {
    if (M <= 0 || N <= 0 || K <= 0) return EOE_ERR_ARG;
    if (dtype != EOE_BF16 && dtype != EOE_F16 && dtype != EOE_F16X2) return EOE_ERR_DTYPE;
    if (N % gemm::BN != 0 || K % gemm::BK != 0) return EOE_ERR_SHAPE;
    return EOE_OK;
}

// ------------------------------------------------------------------------------------------ im2col
// imgs [B,3,R,R] fp32 NCHW -> patches [B*g*g, 3*P*P] 16-bit with k = c*P*P + py*P + px (= conv1.weight.view(width,-1))
// SPLIT (operand dtype EOE_F16X2): patches [B*g*g, 2 * 3*P*P] = [hi | lo] fp16 pairs (gemm::split2)
template <bool BF16, bool SPLIT = false>
__global__ void __launch_bounds__(256)
im2col_kernel(const float* __restrict__ imgs, uint16_t* __restrict__ patches, int64_t B, int R, int P) {
    // one thread: 8 consecutive pixels of a patch row (two 16-byte loads, one 16-byte store); 8 | P
    const int g = R / P;
    const int pq = P / 8;
    const int64_t total = B * 3 * (int64_t)R * R / 8;
    for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256) {
        int64_t t = idx;
        const int px8 = (int)(t % pq); t /= pq;
        const int py = (int)(t % P); t /= P;
        const int c = (int)(t % 3); t /= 3;
        const int gx = (int)(t % g); t /= g;
        const int gy = (int)(t % g); t /= g;
        const int64_t b = t;
        const float4* src = reinterpret_cast<const float4*>(
            imgs + ((b * 3 + c) * R + (gy * P + py)) * (int64_t)R + gx * P + px8 * 8);
        float v0[4], v1[4];
        load4_stream<float>(reinterpret_cast<const float*>(src), v0);
        load4_stream<float>(reinterpret_cast<const float*>(src + 1), v1);
        if (SPLIT) {
            uint4 hi, lo;
            gemm::split2(v0[0], v0[1], hi.x, lo.x);
            gemm::split2(v0[2], v0[3], hi.y, lo.y);
            gemm::split2(v1[0], v1[1], hi.z, lo.z);
            gemm::split2(v1[2], v1[3], hi.w, lo.w);
            const int64_t kp = 3 * (int64_t)P * P, row = idx * 8 / kp, k = idx * 8 % kp;
            *reinterpret_cast<uint4*>(patches + row * 2 * kp + k) = hi;
            *reinterpret_cast<uint4*>(patches + row * 2 * kp + kp + k) = lo;
            continue;
        }
        uint4 o;
        o.x = gemm::pack2<BF16>(v0[0], v0[1]);
        o.y = gemm::pack2<BF16>(v0[2], v0[3]);
        o.z = gemm::pack2<BF16>(v1[0], v1[1]);
        o.w = gemm::pack2<BF16>(v1[2], v1[3]);
        *reinterpret_cast<uint4*>(patches + idx * 8) = o;      // idx*8 == ((b*g+gy)*g+gx)*3*P*P + c*P*P + py*P + px8*8
    }
}

// uint8 images -> patches with torchvision's ToTensor + Normalize fused in (reference input pipeline:
// clip_official/clip/clip.py:58-65 `_transform` = ... ToTensor(), Normalize(mean, std); eoe's GPU Normalize,
// utils/transformations.py:126-138).  ToTensor is u8 / 255 and Normalize is (x - mean) / std, both fp32 with correctly
// rounded divisions, so there are only 3 x 256 possible results: they are tabulated once per CTA in shared memory,
// already rounded to the operand dtype -- bit-identical to normalising on the host and feeding the fp32 path.
// NHWC == false: imgs [B,3,R,R] (ToTensor's layout);  NHWC == true: imgs [B,R,R,3] (decoded image files).
// One thread converts 16 consecutive pixels of one patch row (16 | P).
struct NormParams { float mean[3], stdv[3]; };
template <bool BF16, bool NHWC, bool SPLIT = false>
__global__ void __launch_bounds__(256)
im2col_u8_kernel(const uint8_t* __restrict__ imgs, uint16_t* __restrict__ patches, int64_t B, int R, int P, NormParams np) {
    __shared__ uint16_t lut[3][256];
    __shared__ uint16_t lut_lo[SPLIT ? 3 : 1][256];     // SPLIT: the lo halves, written P*P*3 elements behind the hi halves
    for (int i = threadIdx.x; i < 768; i += 256) {
        const int c = i >> 8, v = i & 255;
        const float x = __fdiv_rn(__fdiv_rn((float)v, 255.0f) - np.mean[c], np.stdv[c]);
        if (SPLIT) {
            uint32_t hi, lo;
            gemm::split2(x, 0.f, hi, lo);
            lut[c][v] = (uint16_t)(hi & 0xffffu);
            lut_lo[c][v] = (uint16_t)(lo & 0xffffu);
        } else {
            lut[c][v] = (uint16_t)(gemm::pack2<BF16>(x, 0.f) & 0xffffu);
        }
    }
    __syncthreads();
    const int64_t rstride = (SPLIT ? 2 : 1) * 3 * (int64_t)(P * P);      // elements per patch row
    const int g = R / P, segs = P / 16;
    const int64_t total = NHWC ? B * (int64_t)R * (R / 16) : B * 3 * (int64_t)R * (R / 16);
    for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256) {
        int64_t t = idx;
        const int xs = (int)(t % (R / 16)); t /= (R / 16);          // 16-pixel segment along x
        const int y = (int)(t % R); t /= R;
        const int gx = xs / segs, sx = xs % segs, gy = y / P, py = y % P;
        if (NHWC) {
            const int64_t b = t;
            const uint4* src = reinterpret_cast<const uint4*>(imgs + ((b * R + y) * (int64_t)R + xs * 16) * 3);
            const uint4 q0 = __ldg(src), q1 = __ldg(src + 1), q2 = __ldg(src + 2);
            const uint32_t wds[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
#pragma unroll
            for (int part = 0; part < (SPLIT ? 2 : 1); ++part) {
                uint16_t o[3][16];
#pragma unroll
                for (int i = 0; i < 48; ++i) {
                    const uint32_t byte = (wds[i >> 2] >> ((i & 3) * 8)) & 0xffu;
                    o[i % 3][i / 3] = part ? lut_lo[SPLIT ? i % 3 : 0][byte] : lut[i % 3][byte];
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    uint16_t* dst = patches + ((b * g + gy) * g + gx) * rstride + (part * 3 + c) * (int64_t)(P * P) + py * P + sx * 16;
                    uint4 w0, w1;
                    w0.x = o[c][0] | ((uint32_t)o[c][1] << 16); w0.y = o[c][2] | ((uint32_t)o[c][3] << 16);
                    w0.z = o[c][4] | ((uint32_t)o[c][5] << 16); w0.w = o[c][6] | ((uint32_t)o[c][7] << 16);
                    w1.x = o[c][8] | ((uint32_t)o[c][9] << 16); w1.y = o[c][10] | ((uint32_t)o[c][11] << 16);
                    w1.z = o[c][12] | ((uint32_t)o[c][13] << 16); w1.w = o[c][14] | ((uint32_t)o[c][15] << 16);
                    reinterpret_cast<uint4*>(dst)[0] = w0;
                    reinterpret_cast<uint4*>(dst)[1] = w1;
                }
            }
        } else {
            const int c = (int)(t % 3); t /= 3;
            const int64_t b = t;
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(imgs + ((b * 3 + c) * R + y) * (int64_t)R + xs * 16));
            const uint32_t wds[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int part = 0; part < (SPLIT ? 2 : 1); ++part) {
                const uint16_t* lt = part ? lut_lo[SPLIT ? c : 0] : lut[c];
                uint32_t o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t b0 = (wds[i >> 1] >> ((i & 1) * 16)) & 0xffu, b1 = (wds[i >> 1] >> ((i & 1) * 16 + 8)) & 0xffu;
                    o[i] = lt[b0] | ((uint32_t)lt[b1] << 16);
                }
                uint16_t* dst = patches + ((b * g + gy) * g + gx) * rstride + (part * 3 + c) * (int64_t)(P * P) + py * P + sx * 16;
                reinterpret_cast<uint4*>(dst)[0] = make_uint4(o[0], o[1], o[2], o[3]);
                reinterpret_cast<uint4*>(dst)[1] = make_uint4(o[4], o[5], o[6], o[7]);
            }
        }
    }
}

// Raw image -> patches with the WHOLE of CLIP's `_transform` fused in (clip_official/clip/clip.py:58-65):
//   Resize(n_px, BICUBIC) -> CenterCrop(n_px) -> ToTensor -> Normalize
// Resize is Pillow's two-pass fixed-point resampler (libImaging/Resample.c) reproduced bit for bit: Keys bicubic a = -0.5,
// support 2 * max(scale, 1), coefficients normalised in double precision (explicit _rn intrinsics: no FMA contraction) and
// truncated to 22-bit fixed point, horizontal pass first into an 8-bit intermediate, clip8((2^21 + sum) >> 22); size and
// crop rules are torchvision's (host side, resize_geometry()).  One CTA = one image x one row of patches (P output rows):
// it tabulates the coefficients it needs, resamples the source rows those P output rows touch into shared memory, then
// runs the vertical pass, the uint8 -> normalised 16-bit LUT of im2col_u8_kernel and the patch scatter.
struct ResizeGeom { int H, W, nh, nw, top, left, ksh, ksv; double sch, scv; };   // source, resized, crop offset, taps, scales

__device__ __forceinline__ double pil_bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(a + 2.0, x), a + 3.0), x), x), 1.0);
    if (x < 2.0) return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), a);
    return 0.0;
}
// Resample.c precompute_coeffs + normalize_coeffs_8bpc for output index xx (box = whole axis): first tap, tap count, taps
__device__ void pil_coeffs(int xx, int in_size, double scale, int ksize, int* first, int* count, int* kk) {
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = __dmul_rn(2.0, filterscale);
    const double ss = __ddiv_rn(1.0, filterscale);
    const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
    int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x)
        ww = __dadd_rn(ww, pil_bicubic(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss)));
    for (int x = 0; x < ksize; ++x) {
        int v = 0;
        if (x < xmax) {
            double w = pil_bicubic(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
            if (ww != 0.0) w = __ddiv_rn(w, ww);
            const double f = __dmul_rn(w, (double)(1 << 22));
            v = w < 0 ? (int)__dadd_rn(-0.5, f) : (int)__dadd_rn(0.5, f);
        }
        kk[x] = v;
    }
    *first = xmin;
    *count = xmax;
}
__device__ __forceinline__ int pil_clip8(int acc) {
    const int v = acc >> 22;                       // arithmetic shift, as clip8_lookups is indexed
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

template <bool BF16, bool SPLIT = false>
__global__ void __launch_bounds__(256)
resize_patchify_kernel(const uint8_t* __restrict__ imgs, uint16_t* __restrict__ patches, int R, int P, ResizeGeom gm,
                       NormParams np, int max_rows) {
    extern __shared__ uint8_t rs_smem[];
    __shared__ uint16_t lut[3][256];
    __shared__ uint16_t lut_lo[SPLIT ? 3 : 1][256];
    __shared__ int s_vfirst[32], s_vcount[32];
    __shared__ int s_rlo, s_rhi;
    const int g = R / P, gy = blockIdx.x;
    const int64_t b = blockIdx.y;
    int* hfirst = reinterpret_cast<int*>(rs_smem);                 // [R]
    int* hcount = hfirst + R;                                      // [R]
    int* hk = hcount + R;                                          // [R][ksh]
    int* vk = hk + R * gm.ksh;                                     // [P][ksv]
    uint8_t* tmp = reinterpret_cast<uint8_t*>(vk + P * gm.ksv);    // [rows][R][3]
    for (int i = threadIdx.x; i < 768; i += 256) {
        const int c = i >> 8, v = i & 255;
        const float x = __fdiv_rn(__fdiv_rn((float)v, 255.0f) - np.mean[c], np.stdv[c]);
        if (SPLIT) {
            uint32_t hi, lo;
            gemm::split2(x, 0.f, hi, lo);
            lut[c][v] = (uint16_t)(hi & 0xffffu);
            lut_lo[c][v] = (uint16_t)(lo & 0xffffu);
        } else {
            lut[c][v] = (uint16_t)(gemm::pack2<BF16>(x, 0.f) & 0xffffu);
        }
    }
    // coefficient tables: R cropped output columns, P output rows of this patch row
    for (int i = threadIdx.x; i < R + P; i += 256) {
        if (i < R) {
            if (gm.nw != gm.W) pil_coeffs(gm.left + i, gm.W, gm.sch, gm.ksh, &hfirst[i], &hcount[i], hk + i * gm.ksh);
            else { hfirst[i] = gm.left + i; hcount[i] = 1; hk[i * gm.ksh] = 1 << 22; for (int k = 1; k < gm.ksh; ++k) hk[i * gm.ksh + k] = 0; }
        } else {
            const int j = i - R;
            if (gm.nh != gm.H) pil_coeffs(gm.top + gy * P + j, gm.H, gm.scv, gm.ksv, &s_vfirst[j], &s_vcount[j], vk + j * gm.ksv);
            else { s_vfirst[j] = gm.top + gy * P + j; s_vcount[j] = 1; vk[j * gm.ksv] = 1 << 22; for (int k = 1; k < gm.ksv; ++k) vk[j * gm.ksv + k] = 0; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) { s_rlo = s_vfirst[0]; s_rhi = s_vfirst[P - 1] + s_vcount[P - 1]; }
    __syncthreads();
    const int rlo = s_rlo, rows = min(s_rhi - s_rlo, max_rows);
    const uint8_t* img = imgs + b * (int64_t)gm.H * gm.W * 3;
    // horizontal pass (identity copy when the width is unchanged: hk = {2^22}) -> 8-bit intermediate, as Pillow keeps it
    for (int i = threadIdx.x; i < rows * R; i += 256) {
        const int r = i / R, xx = i % R;
        const uint8_t* src = img + ((int64_t)(rlo + r) * gm.W + hfirst[xx]) * 3;
        const int n = hcount[xx];
        const int* k = hk + xx * gm.ksh;
        int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
        for (int t = 0; t < n; ++t) {
            const int w = k[t];
            a0 += (int)src[3 * t] * w; a1 += (int)src[3 * t + 1] * w; a2 += (int)src[3 * t + 2] * w;
        }
        uint8_t* d = tmp + (r * R + xx) * 3;
        d[0] = (uint8_t)pil_clip8(a0); d[1] = (uint8_t)pil_clip8(a1); d[2] = (uint8_t)pil_clip8(a2);
    }
    __syncthreads();
    // vertical pass + ToTensor/Normalize LUT + patch scatter: one task = 8 consecutive output pixels of one row and channel
    const int segs = R / 8;
    for (int i = threadIdx.x; i < P * segs * 3; i += 256) {
        const int seg = i % segs, c = (i / segs) % 3, py = i / (segs * 3);
        const int n = s_vcount[py], r0 = s_vfirst[py] - rlo;
        const int* k = vk + py * gm.ksv;
        int acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 1 << 21;
        for (int t = 0; t < n; ++t) {
            const int w = k[t];
            const uint8_t* srow = tmp + ((r0 + t) * R + seg * 8) * 3 + c;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += (int)srow[3 * e] * w;
        }
        const int xx = seg * 8, gx = xx / P, px = xx % P;
#pragma unroll
        for (int part = 0; part < (SPLIT ? 2 : 1); ++part) {
            const uint16_t* lt = part ? lut_lo[SPLIT ? c : 0] : lut[c];
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                o[e] = lt[pil_clip8(acc[2 * e])] | ((uint32_t)lt[pil_clip8(acc[2 * e + 1])] << 16);
            uint16_t* dst = patches + ((b * g + gy) * g + gx) * (int64_t)((SPLIT ? 6 : 3) * P * P) + (part * 3 + c) * (int64_t)(P * P) + py * P + px;
            *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// torchvision's size rules: Resize(int) scales the shorter side to n_px and the longer to int(n_px * long / short);
// CenterCrop offsets are int(round((size - n_px) / 2.0)) with Python's round-half-even.
static int round_half_even_half(int d) {           // round(d / 2.0) for integer d
    if (d % 2 == 0) return d / 2;
    const int lo = (d - 1) / 2;                    // floor for positive and negative odd d ((d-1)/2 is exact)
    return (lo % 2 == 0) ? lo : lo + 1;
}
static int resize_geometry(int64_t H, int64_t W, int n_px, ResizeGeom* gm) {
    if (H <= 0 || W <= 0 || H > (1 << 20) || W > (1 << 20)) return EOE_ERR_SHAPE;
    const int64_t short_ = W <= H ? W : H, long_ = W <= H ? H : W;
    const int new_long = (int)((double)((int64_t)n_px * long_) / (double)short_);
    gm->H = (int)H; gm->W = (int)W;
    gm->nh = W <= H ? new_long : n_px;
    gm->nw = W <= H ? n_px : new_long;
    if (gm->nh < n_px || gm->nw < n_px) return EOE_ERR_SHAPE;
    gm->top = round_half_even_half(gm->nh - n_px);
    gm->left = round_half_even_half(gm->nw - n_px);
    gm->sch = (double)gm->W / (double)gm->nw;
    gm->scv = (double)gm->H / (double)gm->nh;
    const double fh = gm->sch < 1.0 ? 1.0 : gm->sch, fv = gm->scv < 1.0 ? 1.0 : gm->scv;
    gm->ksh = (int)ceil(2.0 * fh) * 2 + 1;
    gm->ksv = (int)ceil(2.0 * fv) * 2 + 1;
    return EOE_OK;
}

// ------------------------------------------------------------------------------------------ LayerNorm
// fp32 statistics over `width` (model.py:153-159, eps 1e-5); one warp per row, row kept in registers.
// cls_emb != null: rows with row % L == 0 are synthesised as class_embedding + positional_embedding[0]
// (model.py:223-224) -- used by ln_pre, whose other rows already hold patch_embed + pos_emb from the GEMM epilogue.
// SPLIT (out_dtype EOE_F16X2): y [M, 2 * width] = [hi | lo] fp16 pairs.
template <typename OutT, int ITERS, bool SPLIT = false>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                 OutT* __restrict__ y, int64_t M, int width, const float* __restrict__ cls_emb,
                 const float* __restrict__ pos0, int L) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    const int nvec = width >> 2;
    float v[ITERS][4];
    const bool is_cls = cls_emb != nullptr && (row % L) == 0;
    float s = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int vi = it * 32 + lane;
        if (vi < nvec) {
            if (is_cls) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(cls_emb) + vi);
                const float4 c = __ldg(reinterpret_cast<const float4*>(pos0) + vi);
                v[it][0] = a.x + c.x; v[it][1] = a.y + c.y; v[it][2] = a.z + c.z; v[it][3] = a.w + c.w;
            } else {
                const float4 a = *(reinterpret_cast<const float4*>(x + row * width) + vi);
                v[it][0] = a.x; v[it][1] = a.y; v[it][2] = a.z; v[it][3] = a.w;
            }
        } else {
            v[it][0] = v[it][1] = v[it][2] = v[it][3] = 0.f;
        }
        s += v[it][0] + v[it][1] + v[it][2] + v[it][3];
    }
    const float mean = warp_sum(s) / (float)width;
    float q = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int vi = it * 32 + lane;
        if (vi < nvec) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float dlt = v[it][j] - mean; q += dlt * dlt; }
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)width + 1e-5f);
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int vi = it * 32 + lane;
        if (vi < nvec) {
            const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + vi);
            const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + vi);
            float o[4] = {(v[it][0] - mean) * rstd * ww.x + bb.x, (v[it][1] - mean) * rstd * ww.y + bb.y,
                          (v[it][2] - mean) * rstd * ww.z + bb.z, (v[it][3] - mean) * rstd * ww.w + bb.w};
            OutT* dst = y + row * (SPLIT ? 2 : 1) * width + vi * 4;
            if (SPLIT) {
                uint2 hi, lo;
                gemm::split2(o[0], o[1], hi.x, lo.x);
                gemm::split2(o[2], o[3], hi.y, lo.y);
                *reinterpret_cast<uint2*>(dst) = hi;
                *reinterpret_cast<uint2*>(dst + width) = lo;
            } else if (sizeof(OutT) == 4) {
                *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
                uint2 pk;
                constexpr bool bf = std::is_same<OutT, __nv_bfloat16>::value;
                pk.x = gemm::pack2<bf>(o[0], o[1]);
                pk.y = gemm::pack2<bf>(o[2], o[3]);
                *reinterpret_cast<uint2*>(dst) = pk;
            }
        }
    }
}

template <typename OutT, bool SPLIT = false>
static int layernorm_launch(const float* x, const float* w, const float* b, void* y, int64_t M, int64_t width,
                            const float* cls_emb, const float* pos0, int L, cudaStream_t st) {
    const int grid = (int)((M + 7) / 8);
    if (width <= 768)
        layernorm_kernel<OutT, 6, SPLIT><<<grid, 256, 0, st>>>(x, w, b, (OutT*)y, M, (int)width, cls_emb, pos0, L);
    else
        layernorm_kernel<OutT, 8, SPLIT><<<grid, 256, 0, st>>>(x, w, b, (OutT*)y, M, (int)width, cls_emb, pos0, L);
    return check_launch("layernorm_kernel");
}

static int layernorm_dispatch(const float* x, const float* w, const float* b, void* y, int out_dtype, int64_t M,
                              int64_t width, const float* cls_emb, const float* pos0, int L, cudaStream_t st) {
    if (width % 4 != 0 || width > 1024) return EOE_ERR_SHAPE;
    switch (out_dtype) {
        case EOE_F32: return layernorm_launch<float>(x, w, b, y, M, width, cls_emb, pos0, L, st);
        case EOE_F16: return layernorm_launch<__half>(x, w, b, y, M, width, cls_emb, pos0, L, st);
        case EOE_BF16: return layernorm_launch<__nv_bfloat16>(x, w, b, y, M, width, cls_emb, pos0, L, st);
        case EOE_F16X2: return layernorm_launch<__half, true>(x, w, b, y, M, width, cls_emb, pos0, L, st);
        default: return EOE_ERR_DTYPE;
    }
}

// ln_pre for the LayerNorm-folded encoder: y = LN(x) in place (fp32, class-token rows synthesised as above) plus what the
// first folded GEMM needs: xb = round16(y) and stats[row][0] = (sum y, sum y^2), stats[row][1..nch) = 0.
template <bool BF16, int ITERS, bool SPLIT = false>
__global__ void __launch_bounds__(256)
ln_pre_stats_kernel(float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                    uint16_t* __restrict__ xb, float2* __restrict__ stats, float* __restrict__ shift, int64_t M, int width,
                    const float* __restrict__ cls_emb, const float* __restrict__ pos0, int L) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    const int nvec = width >> 2;
    float v[ITERS][4];
    const bool is_cls = (row % L) == 0;
    float s = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int vi = it * 32 + lane;
        if (vi < nvec) {
            if (is_cls) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(cls_emb) + vi);
                const float4 c = __ldg(reinterpret_cast<const float4*>(pos0) + vi);
                v[it][0] = a.x + c.x; v[it][1] = a.y + c.y; v[it][2] = a.z + c.z; v[it][3] = a.w + c.w;
            } else {
                const float4 a = *(reinterpret_cast<const float4*>(x + row * width) + vi);
                v[it][0] = a.x; v[it][1] = a.y; v[it][2] = a.z; v[it][3] = a.w;
            }
        } else {
            v[it][0] = v[it][1] = v[it][2] = v[it][3] = 0.f;
        }
        s += v[it][0] + v[it][1] + v[it][2] + v[it][3];
    }
    const float mean = warp_sum(s) / (float)width;
    float q = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int vi = it * 32 + lane;
        if (vi < nvec) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float dlt = v[it][j] - mean; q += dlt * dlt; }
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)width + 1e-5f);
    float ys = 0.f, yq = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int vi = it * 32 + lane;
        if (vi < nvec) {
            const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + vi);
            const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + vi);
            v[it][0] = (v[it][0] - mean) * rstd * ww.x + bb.x; v[it][1] = (v[it][1] - mean) * rstd * ww.y + bb.y;
            v[it][2] = (v[it][2] - mean) * rstd * ww.z + bb.z; v[it][3] = (v[it][3] - mean) * rstd * ww.w + bb.w;
            *reinterpret_cast<float4*>(x + row * width + vi * 4) = make_float4(v[it][0], v[it][1], v[it][2], v[it][3]);
            ys += (v[it][0] + v[it][1]) + (v[it][2] + v[it][3]);
            yq += (v[it][0] * v[it][0] + v[it][1] * v[it][1]) + (v[it][2] * v[it][2] + v[it][3] * v[it][3]);
        }
    }
    ys = warp_sum(ys);
    yq = warp_sum(yq);
    // the 16-bit copy is centred on the row's own mean (exact here: the whole row is in registers); the folded GEMM is told
    const float ymean = ys / (float)width;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int vi = it * 32 + lane;
        if (vi < nvec && SPLIT) {
            uint2 hi, lo;
            gemm::split2(v[it][0] - ymean, v[it][1] - ymean, hi.x, lo.x);
            gemm::split2(v[it][2] - ymean, v[it][3] - ymean, hi.y, lo.y);
            *reinterpret_cast<uint2*>(xb + row * 2 * width + vi * 4) = hi;
            *reinterpret_cast<uint2*>(xb + row * 2 * width + width + vi * 4) = lo;
        } else if (vi < nvec) {
            *reinterpret_cast<uint2*>(xb + row * width + vi * 4) =
                make_uint2(gemm::pack2<BF16>(v[it][0] - ymean, v[it][1] - ymean), gemm::pack2<BF16>(v[it][2] - ymean, v[it][3] - ymean));
        }
    }
    const int nch = width >> 7;
    if (lane < nch) stats[row * nch + lane] = lane == 0 ? make_float2(ys, yq) : make_float2(0.f, 0.f);
    if (lane == 0) shift[row] = ymean;
}

// LayerNorm fold of one Linear (eoe_vit_fold_layernorm): one warp per output row n.
// SPLIT (EOE_F16X2): wf [N, 2K] = [hi | lo] fp16 pairs of W * ln_w, c1 = rowsum(hi + lo).
template <bool BF16, bool SPLIT = false>
__global__ void __launch_bounds__(256)
fold_ln_kernel(const float* __restrict__ w, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
               const float* __restrict__ bias, int64_t N, int64_t K, uint16_t* __restrict__ wf,
               float* __restrict__ c1, float* __restrict__ c2) {
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (n >= N) return;
    float s1 = 0.f, s2 = 0.f;
    for (int64_t k = lane * 4; k < K; k += 128) {
        const float4 a = *reinterpret_cast<const float4*>(w + n * K + k);
        const float4 g = __ldg(reinterpret_cast<const float4*>(ln_w + k));
        const float4 be = __ldg(reinterpret_cast<const float4*>(ln_b + k));
        uint32_t p0, p1, q0 = 0, q1 = 0;
        if (SPLIT) {
            gemm::split2(a.x * g.x, a.y * g.y, p0, q0);
            gemm::split2(a.z * g.z, a.w * g.w, p1, q1);
            *reinterpret_cast<uint2*>(wf + n * 2 * K + k) = make_uint2(p0, p1);
            *reinterpret_cast<uint2*>(wf + n * 2 * K + K + k) = make_uint2(q0, q1);
        } else {
            p0 = gemm::pack2<BF16>(a.x * g.x, a.y * g.y);
            p1 = gemm::pack2<BF16>(a.z * g.z, a.w * g.w);
            *reinterpret_cast<uint2*>(wf + n * K + k) = make_uint2(p0, p1);
        }
        float r[4];
        if (BF16) {
            r[0] = __uint_as_float(p0 << 16); r[1] = __uint_as_float(p0 & 0xffff0000u);
            r[2] = __uint_as_float(p1 << 16); r[3] = __uint_as_float(p1 & 0xffff0000u);
        } else {
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&p0));
            const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&p1));
            r[0] = f0.x; r[1] = f0.y; r[2] = f1.x; r[3] = f1.y;
            if (SPLIT) {
                const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&q0));
                const float2 l1 = __half22float2(*reinterpret_cast<const __half2*>(&q1));
                r[0] += l0.x; r[1] += l0.y; r[2] += l1.x; r[3] += l1.y;
            }
        }
        s1 += (r[0] + r[1]) + (r[2] + r[3]);
        s2 += (a.x * be.x + a.y * be.y) + (a.z * be.z + a.w * be.w);
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
        c1[n] = s1;
        c2[n] = s2 + (bias ? bias[n] : 0.f);
    }
}

// ------------------------------------------------------------------------------------------ attention
// softmax(Q K^T / 8) V for one (image, head) per CTA: L <= LP keys live in shared memory, every warp owns 16 query
// rows at a time, the whole score row block stays in registers (no online rescaling needed at L <= 208).
// Tensor work uses warp-level mma.sync m16n8k16 (attention is 4 % of the encoder FLOPs; the GEMMs carry tcgen05).
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
template <bool BF16>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if (BF16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kAttnThreads = 128;

// CAUSAL: key j is visible to query i only if j <= i (the text tower's additive -inf mask, model.py:324-331).
template <bool BF16, int LP, bool CAUSAL = false>
__global__ void __launch_bounds__(kAttnThreads)
attention_kernel(const uint16_t* __restrict__ qkv, uint16_t* __restrict__ out, int L, int heads) {
    extern __shared__ __align__(128) uint8_t sm[];
    const int width = heads * 64;
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* sQ = sm;
    uint8_t* sK = sm + LP * 128;
    uint8_t* sV = sm + 2 * LP * 128;
    // stage Q, K, V head slices: row r = 128 bytes = 8 chunks of 16 B, chunk c stored at c ^ (r & 7)
    for (int i = tid; i < 3 * LP * 8; i += kAttnThreads) {
        const int mat = i / (LP * 8), rem = i % (LP * 8), r = rem >> 3, c = rem & 7;
        const uint32_t dst = ptx::smem_u32(sm + mat * LP * 128 + r * 128 + ((c ^ (r & 7)) << 4));
        if (r < L) {
            const uint16_t* src = qkv + ((int64_t)b * L + r) * (3 * width) + mat * width + h * 64 + c * 8;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        } else {
            asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(0u) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    constexpr int NT = LP / 8;                      // key tiles of 8
    const float sl2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    const uint32_t q_base = ptx::smem_u32(sQ), k_base = ptx::smem_u32(sK), v_base = ptx::smem_u32(sV);
    for (int qt = warp; qt * 16 < L; qt += kAttnThreads / 32) {
        const int q0 = qt * 16;
        uint32_t qf[4][4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int r = q0 + (lane & 7) + ((lane >> 3) & 1) * 8;
            const int c = kk * 2 + (lane >> 4);
            ldmatrix_x4(q_base + r * 128 + ((c ^ (r & 7)) << 4), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
        }
        float s[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int np = 0; np < NT / 2; ++np) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int mi = lane >> 3;
                const int r = np * 16 + (lane & 7) + (mi >> 1) * 8;
                const int c = kk * 2 + (mi & 1);
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4(k_base + r * 128 + ((c ^ (r & 7)) << 4), b0, b1, b2, b3);
                mma_16816<BF16>(s[2 * np], qf[kk], b0, b1);
                mma_16816<BF16>(s[2 * np + 1], qf[kk], b2, b3);
            }
        }
        // row-wise softmax statistics: this thread holds rows (lane/4) [e=0,1] and (lane/4 + 8) [e=2,3]
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int col = nt * 8 + (lane & 3) * 2;
            if (col >= L) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
            if (col + 1 >= L) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
            if (CAUSAL) {                            // rows q0 + lane/4 (e = 0, 1) and + 8 (e = 2, 3); key 0 is always visible
                const int r0 = q0 + (lane >> 2), r1 = r0 + 8;
                if (col > r0) s[nt][0] = -INFINITY;
                if (col + 1 > r0) s[nt][1] = -INFINITY;
                if (col > r1) s[nt][2] = -INFINITY;
                if (col + 1 > r1) s[nt][3] = -INFINITY;
            }
            m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
            m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
        }
        m0 = fmaxf(m0, __shfl_xor_sync(kFullMask, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(kFullMask, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(kFullMask, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(kFullMask, m1, 2));
        float l0 = 0.f, l1 = 0.f;
        const float ms0 = m0 * sl2, ms1 = m1 * sl2;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            s[nt][0] = exp2f(s[nt][0] * sl2 - ms0); s[nt][1] = exp2f(s[nt][1] * sl2 - ms0);
            s[nt][2] = exp2f(s[nt][2] * sl2 - ms1); s[nt][3] = exp2f(s[nt][3] * sl2 - ms1);
            l0 += s[nt][0] + s[nt][1];
            l1 += s[nt][2] + s[nt][3];
        }
        l0 += __shfl_xor_sync(kFullMask, l0, 1); l0 += __shfl_xor_sync(kFullMask, l0, 2);
        l1 += __shfl_xor_sync(kFullMask, l1, 1); l1 += __shfl_xor_sync(kFullMask, l1, 2);
        float o[8][4];
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < NT / 2; ++kk) {
            uint32_t pa[4];
            pa[0] = gemm::pack2<BF16>(s[2 * kk][0], s[2 * kk][1]);
            pa[1] = gemm::pack2<BF16>(s[2 * kk][2], s[2 * kk][3]);
            pa[2] = gemm::pack2<BF16>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
            pa[3] = gemm::pack2<BF16>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {
                const int mi = lane >> 3;
                const int r = kk * 16 + (lane & 7) + (mi & 1) * 8;
                const int c = dp * 2 + (mi >> 1);
                uint32_t b0, b1, b2, b3;
                ldmatrix_x4_trans(v_base + r * 128 + ((c ^ (r & 7)) << 4), b0, b1, b2, b3);
                mma_16816<BF16>(o[2 * dp], pa, b0, b1);
                mma_16816<BF16>(o[2 * dp + 1], pa, b2, b3);
            }
        }
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
        const int r0 = q0 + (lane >> 2), r1 = r0 + 8;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
            const int d = dt * 8 + (lane & 3) * 2;
            if (r0 < L)
                *reinterpret_cast<uint32_t*>(out + ((int64_t)b * L + r0) * width + h * 64 + d) =
                    gemm::pack2<BF16>(o[dt][0] * i0, o[dt][1] * i0);
            if (r1 < L)
                *reinterpret_cast<uint32_t*>(out + ((int64_t)b * L + r1) * width + h * 64 + d) =
                    gemm::pack2<BF16>(o[dt][2] * i1, o[dt][3] * i1);
        }
    }
}

template <bool BF16, int LP, bool CAUSAL = false>
static int attention_launch_t(const void* qkv, void* out, int64_t B, int L, int heads, cudaStream_t st) {
    auto kern = attention_kernel<BF16, LP, CAUSAL>;
    const int smem = 3 * LP * 128;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { set_cuda_error(e, "attention smem attr"); return EOE_ERR_CUDA; }
        attr_done = true;
    }
    kern<<<(unsigned)(B * heads), kAttnThreads, smem, st>>>((const uint16_t*)qkv, (uint16_t*)out, L, heads);
    return check_launch("attention_kernel");
}

// tcgen05 path (L == 197): qkv is read through a TMA descriptor with 128-row x 64-column boxes
template <bool BF16, bool SPLIT = false>
static int attention_tc_launch(const CUtensorMap& tm_qkv, void* out, int64_t B, int heads, int dtype, cudaStream_t st,
                               const CUtensorMap* tm_o128 = nullptr, const CUtensorMap* tm_o72 = nullptr) {
    CUtensorMap l128, l72;
    if (!tm_o128) {
        int rc = make_tmap_tokens(&l128, out, B, 197, (SPLIT ? 2 : 1) * heads * 64, 128, dtype);
        if (!rc) rc = make_tmap_tokens(&l72, out, B, 197, (SPLIT ? 2 : 1) * heads * 64, 72, dtype);
        if (rc) return rc;
        tm_o128 = &l128;
        tm_o72 = &l72;
    }
    auto kern = attn::attention_tc_kernel<BF16, 197, SPLIT>;
    constexpr uint32_t smem_bytes = SPLIT ? attn::SMEM_BYTES_SPLIT : attn::SMEM_BYTES;
    constexpr int per_sm = SPLIT ? 1 : 2;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) { set_cuda_error(e, "attention_tc smem attr"); return EOE_ERR_CUDA; }
        attr_done = true;
    }
    const int64_t items = B * heads;
    const int grid = (int)(items < per_sm * (int64_t)num_sms() ? items : per_sm * (int64_t)num_sms());
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(attn::THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (g_gemm_debug & 32) ? 0 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tm_qkv, *tm_o128, *tm_o72, (int)items, heads);
    if (e != cudaSuccess) { set_cuda_error(e, "attention_tc_kernel launch"); return EOE_ERR_CUDA; }
    return check_launch("attention_tc_kernel");
}

// tcgen05 path for L <= 64 (ViT-B/32: 50 tokens): two heads of an image per 128-row tile, 64-token TMA boxes
template <bool BF16, bool SPLIT = false>
static int attention_tc64_launch(const void* qkv, void* out, int64_t B, int L, int heads, int dtype, cudaStream_t st,
                                 const CUtensorMap* tm_qkv64 = nullptr, const CUtensorMap* tm_o64 = nullptr) {
    CUtensorMap lq, lo;
    if (!tm_qkv64) {
        int rc = make_tmap_tokens(&lq, qkv, B, L, (SPLIT ? 6 : 3) * heads * 64, 64, dtype);
        if (!rc) rc = make_tmap_tokens(&lo, out, B, L, (SPLIT ? 2 : 1) * heads * 64, 64, dtype);
        if (rc) return rc;
        tm_qkv64 = &lq;
        tm_o64 = &lo;
    }
    auto kern = attn::attention_tc64_kernel<BF16, SPLIT>;
    constexpr uint32_t smem_bytes = SPLIT ? attn::tc64::SMEM_BYTES_SPLIT : attn::tc64::SMEM_BYTES;
    constexpr int per_sm = SPLIT ? 2 : 4;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) { set_cuda_error(e, "attention_tc64 smem attr"); return EOE_ERR_CUDA; }
        attr_done = true;
    }
    const int64_t pairs = B * (heads / 2);
    const int grid = (int)(pairs < per_sm * (int64_t)num_sms() ? pairs : per_sm * (int64_t)num_sms());
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(attn::tc64::THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (g_gemm_debug & 32) ? 0 : 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, *tm_qkv64, *tm_o64, (int)pairs, heads, L);
    if (e != cudaSuccess) { set_cuda_error(e, "attention_tc64_kernel launch"); return EOE_ERR_CUDA; }
    return check_launch("attention_tc64_kernel");
}

static int attention_dispatch(const void* qkv, void* out, int64_t B, int64_t L, int64_t heads, int dtype, cudaStream_t st,
                              const CUtensorMap* tm_qkv = nullptr, const CUtensorMap* tm_o128 = nullptr,
                              const CUtensorMap* tm_o72 = nullptr, bool causal = false) {
    if (dtype != EOE_BF16 && dtype != EOE_F16 && dtype != EOE_F16X2) return EOE_ERR_DTYPE;
    if (B <= 0 || L <= 0 || heads <= 0 || B * heads > 0x7fffffff) return EOE_ERR_ARG;
    if (dtype == EOE_F16X2) {
        // split fp16 operands: qkv [B*L, 6*width] = [q k v | lo halves], out [B*L, 2*width]; the two tcgen05 kernels only
        if (causal || (uintptr_t)qkv % 16 != 0 || (uintptr_t)out % 16 != 0) return EOE_ERR_SHAPE;
        if (L == 197) {
            CUtensorMap local;
            if (!tm_qkv) {
                int rc = make_tmap_tokens(&local, qkv, B, L, 6 * heads * 64, 128, EOE_F16);
                if (rc) return rc;
                tm_qkv = &local;
            }
            return attention_tc_launch<false, true>(*tm_qkv, out, B, (int)heads, EOE_F16, st, tm_o128, tm_o72);
        }
        if (L <= 64 && heads % 2 == 0)
            return attention_tc64_launch<false, true>(qkv, out, B, (int)L, (int)heads, EOE_F16, st, tm_qkv, tm_o128);
        return EOE_ERR_SHAPE;
    }
    const bool bf = dtype == EOE_BF16;
    if (causal) {                                    // text tower (context 77): warp-level kernel, keys in shared memory
        if (L <= 80) return bf ? attention_launch_t<true, 80, true>(qkv, out, B, (int)L, (int)heads, st) : attention_launch_t<false, 80, true>(qkv, out, B, (int)L, (int)heads, st);
        if (L <= 208) return bf ? attention_launch_t<true, 208, true>(qkv, out, B, (int)L, (int)heads, st) : attention_launch_t<false, 208, true>(qkv, out, B, (int)L, (int)heads, st);
        return EOE_ERR_SHAPE;
    }
    if (L == 197) {
        CUtensorMap local;
        if (!tm_qkv) {
            int rc = make_tmap_tokens(&local, qkv, B, L, 3 * heads * 64, 128, dtype);
            if (rc) return rc;
            tm_qkv = &local;
        }
        return bf ? attention_tc_launch<true>(*tm_qkv, out, B, (int)heads, dtype, st, tm_o128, tm_o72)
                  : attention_tc_launch<false>(*tm_qkv, out, B, (int)heads, dtype, st, tm_o128, tm_o72);
    }
    if (L <= 64 && heads % 2 == 0 && (uintptr_t)qkv % 16 == 0 && (uintptr_t)out % 16 == 0 && !(g_gemm_debug & 128)) {
        // (diagnostics bit 7 keeps the warp-level kernel for A/B timing)
        return bf ? attention_tc64_launch<true>(qkv, out, B, (int)L, (int)heads, dtype, st, tm_qkv, tm_o128)
                  : attention_tc64_launch<false>(qkv, out, B, (int)L, (int)heads, dtype, st, tm_qkv, tm_o128);
    }
    if (L <= 64) return bf ? attention_launch_t<true, 64>(qkv, out, B, (int)L, (int)heads, st) : attention_launch_t<false, 64>(qkv, out, B, (int)L, (int)heads, st);
    if (L <= 208) return bf ? attention_launch_t<true, 208>(qkv, out, B, (int)L, (int)heads, st) : attention_launch_t<false, 208>(qkv, out, B, (int)L, (int)heads, st);
    return EOE_ERR_SHAPE;
}

// ------------------------------------------------------------------------------------------ last block, class token only
// Only token 0 of the last ResidualAttentionBlock reaches ln_post (model.py:231), so for the last block the attention
// output, out-proj, ln_2 and the MLP are evaluated for the class-token rows alone (the reference computes and discards
// the other 196 rows).  K and V still come from all tokens.  One warp per (image, head): lanes own keys l, l+32, ...;
// also gathers the class-token rows of the fp32 residual stream into the compact buffer x_cls [B, width].
// Last block, LayerNorm-folded path: the class-token rows' 16-bit residual copy, chunk sums and shifts, compacted to
// [B, ...] so that the folded QKV GEMM can produce Q for those B rows only (K and V are still needed for every token).
__global__ void __launch_bounds__(256)
gather_cls_rows_kernel(const uint16_t* __restrict__ xb, const float2* __restrict__ stats, const float* __restrict__ shift,
                       uint16_t* __restrict__ xb_cls, float2* __restrict__ stats_cls, float* __restrict__ shift_cls,
                       int64_t B, int L, int width, int xw) {      // xw: elements per row of xb (2 * width for split operands)
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    const int64_t row = b * L;
    const int nvec = xw / 8, nch = width / 128;
    for (int i = lane; i < nvec; i += 32)
        reinterpret_cast<uint4*>(xb_cls + b * xw)[i] = __ldg(reinterpret_cast<const uint4*>(xb + row * xw) + i);
    if (lane < nch) stats_cls[b * nch + lane] = stats[row * nch + lane];
    if (lane == 0) shift_cls[b] = shift[row];
}

// SPLIT (EOE_F16X2): qkv rows are [q k v | q_lo k_lo v_lo], q_cls and h_cls rows [hi | lo]; values are hi + lo in fp32.
template <bool BF16, bool SPLIT = false>
__global__ void __launch_bounds__(128)
attention_cls_kernel(const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ q_cls, const float* __restrict__ x,
                     uint16_t* __restrict__ h_cls, float* __restrict__ x_cls, int64_t B, int L, int heads) {
    __shared__ float s_red[4][32][33];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t item = (int64_t)blockIdx.x * 4 + wib;
    if (item >= B * heads) return;
    const int width = heads * 64;
    const int64_t b = item / heads;
    const int h = (int)(item % heads);
    auto cvt2 = [](uint32_t u, float& lo, float& hi) {
        if (BF16) { lo = __uint_as_float(u << 16); hi = __uint_as_float(u & 0xffff0000u); }
        else { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u)); lo = f.x; hi = f.y; }
    };
    if (h == 0)                                     // gather the residual-stream row of the class token
        for (int i = lane; i < width / 4; i += 32)
            reinterpret_cast<float4*>(x_cls + b * width)[i] = reinterpret_cast<const float4*>(x + b * L * (int64_t)width)[i];
    constexpr int S = SPLIT ? 2 : 1;
    const int64_t rs = (int64_t)S * 3 * width;       // elements per qkv row
    const uint16_t* base = qkv + b * L * rs;
    // 8 consecutive 16-bit values (+ their lo halves `lo_off` elements further on) -> fp32
    auto load8 = [&](const uint16_t* src, int lo_off, float* f) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
        cvt2(v.x, f[0], f[1]); cvt2(v.y, f[2], f[3]); cvt2(v.z, f[4], f[5]); cvt2(v.w, f[6], f[7]);
        if (SPLIT) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + lo_off));
            float g[8];
            cvt2(u.x, g[0], g[1]); cvt2(u.y, g[2], g[3]); cvt2(u.z, g[4], g[5]); cvt2(u.w, g[6], g[7]);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] += g[e];
        }
    };
    float q[64];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        // the class token's query: from its own [B, width] buffer when the last block computed Q for those rows only
        if (q_cls) load8(q_cls + b * S * width + h * 64 + i * 8, width, q + i * 8);
        else load8(base + h * 64 + i * 8, 3 * width, q + i * 8);
    }
    constexpr int MAXK = 7;                         // keys per lane: L <= 224
    float sc[MAXK];
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < MAXK; ++t) {
        const int j = t * 32 + lane;
        float d = -INFINITY;
        if (j < L) {
            const uint16_t* kr = base + (int64_t)j * rs + width + h * 64;
            d = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float a[8];
                load8(kr + i * 8, 3 * width, a);
                d += q[i * 8] * a[0] + q[i * 8 + 1] * a[1];
                d += q[i * 8 + 2] * a[2] + q[i * 8 + 3] * a[3];
                d += q[i * 8 + 4] * a[4] + q[i * 8 + 5] * a[5];
                d += q[i * 8 + 6] * a[6] + q[i * 8 + 7] * a[7];
            }
            d *= 0.125f;
        }
        sc[t] = d;
        m = fmaxf(m, d);
    }
    m = warp_max(m);
    float sum = 0.f;
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;
#pragma unroll
    for (int t = 0; t < MAXK; ++t) {
        const int j = t * 32 + lane;
        if (j < L) {
            const float pj = __expf(sc[t] - m);
            sum += pj;
            const uint16_t* vr = base + (int64_t)j * rs + 2 * width + h * 64;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float a[8];
                load8(vr + i * 8, 3 * width, a);
                o[i * 8] += pj * a[0]; o[i * 8 + 1] += pj * a[1];
                o[i * 8 + 2] += pj * a[2]; o[i * 8 + 3] += pj * a[3];
                o[i * 8 + 4] += pj * a[4]; o[i * 8 + 5] += pj * a[5];
                o[i * 8 + 6] += pj * a[6]; o[i * 8 + 7] += pj * a[7];
            }
        }
    }
    sum = warp_sum(sum);
    // reduce the 64 per-lane partial outputs across the warp through shared memory (two 32-column halves)
    float res[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int i = 0; i < 32; ++i) s_red[wib][lane][i] = o[half * 32 + i];
        __syncwarp();
        float acc = 0.f;
#pragma unroll
        for (int l2 = 0; l2 < 32; ++l2) acc += s_red[wib][l2][lane];
        res[half] = acc / sum;
        __syncwarp();
    }
    uint16_t* dst = h_cls + b * S * width + h * 64;
    if (SPLIT) {
        const __half h0 = __float2half_rn(res[0]), h1 = __float2half_rn(res[1]);
        dst[lane] = __half_as_ushort(h0);
        dst[32 + lane] = __half_as_ushort(h1);
        dst[width + lane] = __half_as_ushort(__float2half_rn(res[0] - __half2float(h0)));
        dst[width + 32 + lane] = __half_as_ushort(__float2half_rn(res[1] - __half2float(h1)));
    } else if (BF16) {
        dst[lane] = __bfloat16_as_ushort(__float2bfloat16_rn(res[0]));
        dst[32 + lane] = __bfloat16_as_ushort(__float2bfloat16_rn(res[1]));
    } else {
        dst[lane] = __half_as_ushort(__float2half_rn(res[0]));
        dst[32 + lane] = __half_as_ushort(__float2half_rn(res[1]));
    }
}

// Causal attention over split fp16 operands (EOE_F16X2), text tower only (model.py:324-331: key j is visible to query i
// iff j <= i): 2 310 rows once per class, so plain fp32 CUDA-core arithmetic on hi + lo values, one warp per
// (prompt, head, query row), lanes own keys l, l + 32, ... -- no 16-bit rounding anywhere inside (P stays fp32).
// qkv [B*L, 6*width] = [q k v | lo halves], out [B*L, 2*width] = [hi | lo].
__global__ void __launch_bounds__(128)
attention_causal_split_kernel(const uint16_t* __restrict__ qkv, uint16_t* __restrict__ out, int64_t B, int L, int heads) {
    __shared__ float s_red[4][32][33];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t item = (int64_t)blockIdx.x * 4 + wib;
    if (item >= B * heads * L) return;
    const int width = heads * 64;
    const int i = (int)(item % L);
    const int h = (int)((item / L) % heads);
    const int64_t b = item / ((int64_t)L * heads);
    const int64_t rs = 6 * (int64_t)width;
    const uint16_t* base = qkv + b * L * rs;
    auto load8 = [&](const uint16_t* src, float* f) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + 3 * width));
        const uint32_t vw[4] = {v.x, v.y, v.z, v.w}, uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&vw[e]));
            const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&uw[e]));
            f[2 * e] = a.x + c.x;
            f[2 * e + 1] = a.y + c.y;
        }
    };
    float q[64];
#pragma unroll
    for (int c = 0; c < 8; ++c) load8(base + (int64_t)i * rs + h * 64 + c * 8, q + c * 8);
    constexpr int MAXK = 7;                         // keys per lane: L <= 224
    float sc[MAXK];
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < MAXK; ++t) {
        const int j = t * 32 + lane;
        float d = -INFINITY;
        if (j <= i) {
            const uint16_t* kr = base + (int64_t)j * rs + width + h * 64;
            d = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float a[8];
                load8(kr + c * 8, a);
#pragma unroll
                for (int e = 0; e < 8; ++e) d = fmaf(q[c * 8 + e], a[e], d);
            }
            d *= 0.125f;
        }
        sc[t] = d;
        m = fmaxf(m, d);
    }
    m = warp_max(m);
    float sum = 0.f;
    float o[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) o[c] = 0.f;
#pragma unroll
    for (int t = 0; t < MAXK; ++t) {
        const int j = t * 32 + lane;
        if (j <= i) {
            const float pj = expf(sc[t] - m);
            sum += pj;
            const uint16_t* vr = base + (int64_t)j * rs + 2 * width + h * 64;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float a[8];
                load8(vr + c * 8, a);
#pragma unroll
                for (int e = 0; e < 8; ++e) o[c * 8 + e] = fmaf(pj, a[e], o[c * 8 + e]);
            }
        }
    }
    sum = warp_sum(sum);
    float res[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int c = 0; c < 32; ++c) s_red[wib][lane][c] = o[half * 32 + c];
        __syncwarp();
        float acc = 0.f;
#pragma unroll
        for (int l2 = 0; l2 < 32; ++l2) acc += s_red[wib][l2][lane];
        res[half] = acc / sum;
        __syncwarp();
    }
    uint16_t* dst = out + (b * L + i) * 2 * (int64_t)width + h * 64;
    const __half h0 = __float2half_rn(res[0]), h1 = __float2half_rn(res[1]);
    dst[lane] = __half_as_ushort(h0);
    dst[32 + lane] = __half_as_ushort(h1);
    dst[width + lane] = __half_as_ushort(__float2half_rn(res[0] - __half2float(h0)));
    dst[width + 32 + lane] = __half_as_ushort(__float2half_rn(res[1] - __half2float(h1)));
}

// ------------------------------------------------------------------------------------------ tail
// feats[b] = ln_post(x[b, 0, :]) @ proj   (model.py:231-234), fp32 throughout (it feeds a 100x cosine logit).
// Block = 4 images x 128 output columns; proj (1.5 MB, L2 resident) is streamed once per 4 images.  Thread
// (col = tid & 127, kh = tid >> 7) accumulates half of the K range for its column; halves are combined in smem.
constexpr int kTailImgs = 4;
constexpr int kTailCols = 128;
__global__ void __launch_bounds__(256)
tail_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
            const float* __restrict__ proj, float* __restrict__ feats, int64_t B, int L, int width, int embed,
            const int64_t* __restrict__ tokens = nullptr) {
    extern __shared__ float s_h[];        // [kTailImgs][width] then [kTailImgs][kTailCols] partials
    float* s_part = s_h + kTailImgs * width;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t img0 = (int64_t)blockIdx.x * kTailImgs;
    const int col = blockIdx.y * kTailCols + (threadIdx.x & (kTailCols - 1));
    const int kh = threadIdx.x >> 7;
    if (warp < kTailImgs) {
        const int64_t img = img0 + warp;
        float* h = s_h + warp * width;
        if (img < B) {
            const float* xr = x + img * L * (int64_t)width;
            if (tokens) {
                // text tower (model.py:350): the row of the <eot> token = first position of the largest token id
                long long best = -1;
                int pos = 0;
                for (int j = lane; j < L; j += 32) {
                    const long long t = tokens[img * L + j];
                    if (t > best) { best = t; pos = j; }         // strictly greater: keeps this lane's first maximum
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const long long ob = __shfl_xor_sync(kFullMask, best, o);
                    const int op = __shfl_xor_sync(kFullMask, pos, o);
                    if (ob > best || (ob == best && op < pos)) { best = ob; pos = op; }
                }
                xr += (int64_t)pos * width;
            }
            float s = 0.f;
            for (int i = lane; i < width; i += 32) s += xr[i];
            const float mean = warp_sum(s) / (float)width;
            float q = 0.f;
            for (int i = lane; i < width; i += 32) { const float d = xr[i] - mean; q += d * d; }
            const float rstd = rsqrtf(warp_sum(q) / (float)width + 1e-5f);
            for (int i = lane; i < width; i += 32) h[i] = (xr[i] - mean) * rstd * w[i] + b[i];
        } else {
            for (int i = lane; i < width; i += 32) h[i] = 0.f;
        }
    }
    __syncthreads();
    float acc[kTailImgs];
#pragma unroll
    for (int g = 0; g < kTailImgs; ++g) acc[g] = 0.f;
    const int k0 = kh * (width / 2), k1 = k0 + width / 2;
    if (col < embed) {
#pragma unroll 4
        for (int i = k0; i < k1; ++i) {
            const float pv = __ldg(proj + (int64_t)i * embed + col);
#pragma unroll
            for (int g = 0; g < kTailImgs; ++g) acc[g] += s_h[g * width + i] * pv;
        }
    }
    if (kh == 1) {
#pragma unroll
        for (int g = 0; g < kTailImgs; ++g) s_part[g * kTailCols + (threadIdx.x & (kTailCols - 1))] = acc[g];
    }
    __syncthreads();
    if (kh == 0 && col < embed) {
#pragma unroll
        for (int g = 0; g < kTailImgs; ++g)
            if (img0 + g < B) feats[(img0 + g) * embed + col] = acc[g] + s_part[g * kTailCols + (threadIdx.x & (kTailCols - 1))];
    }
}

// Text tower input (model.py:340-342): x[row] = token_embedding[tokens[row]] + positional_embedding[row % ctx], fp32.
// One warp per row, 16-byte accesses; ids outside [0, vocab) give NaN rows (the host wrapper rejects them first).
__global__ void __launch_bounds__(256)
text_embed_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ tok_emb, const float* __restrict__ pos,
                  float* __restrict__ x, int64_t rows, int ctx, int width, int64_t vocab) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int64_t t = tokens[row];
    const bool ok = t >= 0 && t < vocab;
    const float4* e = reinterpret_cast<const float4*>(tok_emb + (ok ? t : 0) * width);
    const float4* pe = reinterpret_cast<const float4*>(pos + (row % ctx) * width);
    float4* xo = reinterpret_cast<float4*>(x + row * width);
    const float nan = __int_as_float(0x7fc00000);
    for (int i = lane; i < width / 4; i += 32) {
        const float4 a = __ldg(e + i), q = __ldg(pe + i);
        xo[i] = ok ? make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w) : make_float4(nan, nan, nan, nan);
    }
}

}  // namespace eoe

// ================================================================================================ plan / encode
using namespace eoe;

struct eoe_vit_plan {
    eoe_vit_weights w;
    eoe_vit_layer* layers;      // host copy
    int64_t max_batch;
    int L, g2, kpatch;
    char* ws;
    size_t ws_bytes;
    // workspace carve-up
    uint16_t* patches;   // [B*g2, 3*P*P]
    float* x;            // [B*L, width] fp32 residual stream
    uint16_t* h;         // [B*L, width] LN output / attention output
    uint16_t* qkv;       // [B*L, 3*width]
    uint16_t* u;         // [B*L, 4*width]
    float* feats;        // [B, embed]
    float* x_cls;        // [B, width]   class-token rows of the residual stream (last block)
    uint16_t* h_cls;     // [B, width]
    uint16_t* u_cls;     // [B, 4*width]
    uint16_t* xb;        // [B*L, width] 16-bit copy of the residual stream (LayerNorm-folded path)
    float2* stats;       // [2][B*L, width/128] per-row chunk (sum, sum of squares) of the residual stream, ping-pong:
                         // a residual GEMM reads the previous producer's sums (row means) while it writes its own
    float2* stats2;
    float* shift;        // [B*L] what was subtracted from the row of xb (the row's mean before the last update)
    uint16_t* xb_cls;    // [B, width]   last block: class-token rows of xb / stats / shift, and their queries
    float2* stats_cls;   // [B, width/128]
    float* shift_cls;    // [B]
    uint16_t* q_cls;     // [B, width]
    bool fused_ln;       // every layer carries folded in_proj / c_fc weights: no stand-alone ln_1 / ln_2 launches
    CUtensorMap tm_patches, tm_h, tm_u, tm_conv, tm_qkv, tm_hc, tm_uc, tm_xb, tm_ho128, tm_ho72;
    CUtensorMap tm_qkv_st, tm_u_st, tm_uc_st;          // 32-row store boxes over the 16-bit GEMM outputs
    CUtensorMap tm_kv_w, tm_kv_st, tm_xbc, tm_qc_st;   // last block: K;V rows of the folded in_proj weights, K;V columns of
                                                       // qkv (stores), class-token rows of xb (A), their queries (stores)
    CUtensorMap *tm_in, *tm_out, *tm_fc, *tm_proj;     // per layer
    CUtensorMap *tm_inf, *tm_fcf;                      // per layer, LayerNorm-folded weights
    CUtensorMap* tm_projs;                             // per layer, c_proj weights / 1.702 (optional)
    // optional instrumentation (eoe_vit_profile_*): CUDA event pairs around every GEMM launch
    bool profile;
    struct Span { cudaEvent_t a, b; int kind; double flops; };
    std::vector<Span> spans;
    size_t spans_used;
};

enum { KIND_PATCH = 0, KIND_QKV = 1, KIND_OUT = 2, KIND_FC = 3, KIND_PROJ = 4 };

static int timed_gemm(eoe_vit_plan* p, int kind, const CUtensorMap& ta, const CUtensorMap& tb, const gemm::Params& gp,
                      int dt, int epi, cudaStream_t st, const CUtensorMap* tc = nullptr) {
    if (!p->profile) return gemm_launch(ta, tb, gp, dt, epi, st, tc);
    if (p->spans_used == p->spans.size()) {
        eoe_vit_plan::Span s;
        if (cudaEventCreate(&s.a) != cudaSuccess || cudaEventCreate(&s.b) != cudaSuccess) return EOE_ERR_CUDA;
        p->spans.push_back(s);
    }
    eoe_vit_plan::Span& s = p->spans[p->spans_used++];
    s.kind = kind;
    s.flops = 2.0 * (double)gp.M * (double)gp.N * (double)gp.K;
    cudaEventRecord(s.a, st);
    int rc = gemm_launch(ta, tb, gp, dt, epi, st, tc);
    cudaEventRecord(s.b, st);
    return rc;
}

static size_t rup(size_t x) { return (x + 1023) / 1024 * 1024; }

static int vit_check(const eoe_vit_weights* w) {
    if (!w || !w->layers_host) return EOE_ERR_ARG;
    if (w->operand_dtype != EOE_BF16 && w->operand_dtype != EOE_F16 && w->operand_dtype != EOE_F16X2) return EOE_ERR_DTYPE;
    if (w->patch <= 0 || w->resolution % w->patch != 0 || w->patch % 8 != 0) return EOE_ERR_SHAPE;
    if (w->width != w->heads * 64 || w->width % 256 != 0 || w->width > 1024) return EOE_ERR_SHAPE;
    if ((3 * w->patch * w->patch) % 64 != 0 || w->embed_dim > 1024 || w->embed_dim % 4 != 0 || w->width % 2 != 0) return EOE_ERR_SHAPE;
    const int g = w->resolution / w->patch;
    if (g * g + 1 > 208) return EOE_ERR_SHAPE;            // attention kernels: L <= 208 (and 7 keys per lane in the cls path)
    if (w->n_layers <= 0) return EOE_ERR_ARG;
    int folded = 0;
    for (int i = 0; i < w->n_layers; ++i) {
        const eoe_vit_layer& l = w->layers_host[i];
        const int have = (l.in_proj_wf != nullptr) + (l.in_proj_c1 != nullptr) + (l.in_proj_c2 != nullptr) +
                         (l.c_fc_wf != nullptr) + (l.c_fc_c1 != nullptr) + (l.c_fc_c2 != nullptr);
        if (have != 0 && have != 6) return EOE_ERR_ARG;         // all six folded tensors of a layer or none
        folded += have == 6;
    }
    if (folded != 0 && folded != w->n_layers) return EOE_ERR_ARG;   // every layer or none
    if (w->operand_dtype == EOE_F16X2) {
        // split fp16 operands: LayerNorm-folded path only (width <= 768), and the two tcgen05 attention shapes
        if (folded != w->n_layers || w->width > 128 * gemm::MAX_NCH) return EOE_ERR_ARG;
        if (g * g + 1 != 197 && !(g * g + 1 <= 64 && w->heads % 2 == 0)) return EOE_ERR_SHAPE;
    }
    for (int i = 0; i < w->n_layers; ++i) {                         // epilogue parameter vectors travel by 16-byte cp.async
        const eoe_vit_layer& l = w->layers_host[i];
        const void* v[] = {l.in_proj_b, l.out_proj_b, l.c_fc_b, l.c_proj_b, l.in_proj_c1, l.in_proj_c2, l.c_fc_c1, l.c_fc_c2};
        for (const void* q : v)
            if ((uintptr_t)q % 16 != 0) return EOE_ERR_ALIGN;
    }
    return EOE_OK;
}
static bool vit_fused_ln(const eoe_vit_weights* w) { return w->layers_host[0].in_proj_wf != nullptr; }

struct VitLayout { size_t patches, x, h, qkv, u, feats, x_cls, h_cls, u_cls, xb, stats, stats2, shift, xb_cls, stats_cls,
                   shift_cls, q_cls, total; };
static VitLayout vit_layout(const eoe_vit_weights* w, int64_t B) {
    const int g = w->resolution / w->patch;
    const int64_t g2 = g * g, L = g2 + 1, W = w->width;
    // +256 rows of slack: TMA boxes of the last M tile may start below M but never beyond the allocation
    VitLayout l;
    size_t o = 0;
    const size_t eb = w->operand_dtype == EOE_F16X2 ? 4 : 2;   // bytes per stored operand element (split: hi + lo)
    l.patches = o; o += rup((size_t)(B * g2 + 256) * 3 * w->patch * w->patch * eb);
    l.x = o; o += rup((size_t)(B * L + 256) * W * 4);
    l.h = o; o += rup((size_t)(B * L + 256) * W * eb);
    l.qkv = o; o += rup((size_t)(B * L + 256) * 3 * W * eb);
    l.u = o; o += rup((size_t)(B * L + 256) * 4 * W * eb);
    l.feats = o; o += rup((size_t)B * w->embed_dim * 4);
    l.x_cls = o; o += rup((size_t)B * W * 4);                  // last block, class-token rows only
    l.h_cls = o; o += rup((size_t)(B + 256) * W * eb);
    l.u_cls = o; o += rup((size_t)(B + 256) * 4 * W * eb);
    l.xb = o; l.stats = o; l.stats2 = o; l.shift = o;
    if (vit_fused_ln(w)) {
        o += rup((size_t)(B * L + 256) * W * eb);
        l.stats = o; o += rup((size_t)(B * L) * (W / 128) * sizeof(float2));
        l.stats2 = o; o += rup((size_t)(B * L) * (W / 128) * sizeof(float2));
        l.shift = o; o += rup((size_t)(B * L + 256) * sizeof(float));
    }
    l.xb_cls = o; l.stats_cls = o; l.shift_cls = o; l.q_cls = o;
    if (vit_fused_ln(w)) {
        l.xb_cls = o; o += rup((size_t)(B + 256) * W * eb);
        l.stats_cls = o; o += rup((size_t)(B + 256) * (W / 128) * sizeof(float2));
        l.shift_cls = o; o += rup((size_t)(B + 256) * sizeof(float));
        l.q_cls = o; o += rup((size_t)(B + 256) * W * eb);
    }
    l.total = o;
    return l;
}

extern "C" size_t eoe_vit_workspace_bytes(const eoe_vit_weights* w, int64_t max_batch) {
    if (vit_check(w) != EOE_OK || max_batch <= 0) return 0;
    return vit_layout(w, max_batch).total;
}

extern "C" int eoe_vit_plan_create(const eoe_vit_weights* w, int64_t max_batch, void* workspace, size_t workspace_bytes,
                                   eoe_vit_plan** plan_out) {
    int rc = vit_check(w);
    if (rc) return rc;
    if (!plan_out || max_batch <= 0) return EOE_ERR_ARG;
    const VitLayout lay = vit_layout(w, max_batch);
    if (!workspace || workspace_bytes < lay.total) return EOE_ERR_WORKSPACE;
    if ((uintptr_t)workspace % 1024 != 0) return EOE_ERR_ALIGN;
    eoe_vit_plan* p = new (std::nothrow) eoe_vit_plan();
    if (!p) return EOE_ERR_ARG;
    p->w = *w;
    p->layers = new eoe_vit_layer[w->n_layers];
    for (int i = 0; i < w->n_layers; ++i) p->layers[i] = w->layers_host[i];
    p->w.layers_host = p->layers;
    p->max_batch = max_batch;
    p->profile = false;
    p->spans_used = 0;
    const int g = w->resolution / w->patch;
    p->g2 = g * g;
    p->L = p->g2 + 1;
    p->kpatch = 3 * w->patch * w->patch;
    p->ws = (char*)workspace;
    p->ws_bytes = workspace_bytes;
    p->patches = (uint16_t*)(p->ws + lay.patches);
    p->x = (float*)(p->ws + lay.x);
    p->h = (uint16_t*)(p->ws + lay.h);
    p->qkv = (uint16_t*)(p->ws + lay.qkv);
    p->u = (uint16_t*)(p->ws + lay.u);
    p->feats = (float*)(p->ws + lay.feats);
    p->x_cls = (float*)(p->ws + lay.x_cls);
    p->h_cls = (uint16_t*)(p->ws + lay.h_cls);
    p->u_cls = (uint16_t*)(p->ws + lay.u_cls);
    p->fused_ln = vit_fused_ln(w);
    p->xb = (uint16_t*)(p->ws + lay.xb);
    p->stats = (float2*)(p->ws + lay.stats);
    p->stats2 = (float2*)(p->ws + lay.stats2);
    p->shift = (float*)(p->ws + lay.shift);
    p->xb_cls = (uint16_t*)(p->ws + lay.xb_cls);
    p->stats_cls = (float2*)(p->ws + lay.stats_cls);
    p->shift_cls = (float*)(p->ws + lay.shift_cls);
    p->q_cls = (uint16_t*)(p->ws + lay.q_cls);
    const int W = w->width, dt = w->operand_dtype;
    const int64_t rows = max_batch * p->L;
    p->tm_in = new CUtensorMap[w->n_layers];
    p->tm_out = new CUtensorMap[w->n_layers];
    p->tm_fc = new CUtensorMap[w->n_layers];
    p->tm_proj = new CUtensorMap[w->n_layers];
    p->tm_inf = new CUtensorMap[w->n_layers];
    p->tm_fcf = new CUtensorMap[w->n_layers];
    p->tm_projs = new CUtensorMap[w->n_layers];
    // split fp16 operands (EOE_F16X2): every 16-bit matrix [rows, C] is [rows, S * C] = [hi | lo]
    const int S = dt == EOE_F16X2 ? 2 : 1;
    rc = make_tmap(&p->tm_patches, p->patches, max_batch * p->g2, S * p->kpatch, gemm::CTA_M, dt);
    if (!rc) rc = make_tmap(&p->tm_h, p->h, rows, S * W, gemm::CTA_M, dt);
    if (!rc) rc = make_tmap(&p->tm_u, p->u, rows, S * 4 * W, gemm::CTA_M, dt);
    if (!rc) rc = make_tmap(&p->tm_conv, w->conv1_w, W, S * p->kpatch, gemm::CTA_NB, dt);
    // attention maps: 128-token boxes for the L = 197 kernel, 64-token boxes (loads and stores) for the L <= 64 kernel
    if (!rc) rc = make_tmap_tokens(&p->tm_qkv, p->qkv, max_batch, p->L, S * 3 * W, p->L <= 64 ? 64 : 128, dt);
    if (!rc && p->L <= 64) rc = make_tmap_tokens(&p->tm_ho128, p->h, max_batch, p->L, S * W, 64, dt);
    if (!rc && p->L == 197) rc = make_tmap_tokens(&p->tm_ho128, p->h, max_batch, p->L, S * W, 128, dt);
    if (!rc && p->L == 197) rc = make_tmap_tokens(&p->tm_ho72, p->h, max_batch, p->L, S * W, 72, dt);
    if (!rc) rc = make_tmap(&p->tm_hc, p->h_cls, max_batch, S * W, gemm::CTA_M, dt);
    if (!rc) rc = make_tmap(&p->tm_qkv_st, p->qkv, rows, S * 3 * W, 32, dt);
    if (!rc) rc = make_tmap(&p->tm_u_st, p->u, rows, S * 4 * W, 32, dt);
    if (!rc) rc = make_tmap(&p->tm_uc_st, p->u_cls, max_batch, S * 4 * W, 32, dt);
    if (!rc) rc = make_tmap(&p->tm_uc, p->u_cls, max_batch, S * 4 * W, gemm::CTA_M, dt);
    for (int i = 0; i < w->n_layers && !rc; ++i) {
        const eoe_vit_layer& l = p->layers[i];
        rc = make_tmap(&p->tm_in[i], l.in_proj_w, 3 * W, S * W, gemm::CTA_NB, dt);
        if (!rc) rc = make_tmap(&p->tm_out[i], l.out_proj_w, W, S * W, gemm::CTA_NB, dt);
        if (!rc) rc = make_tmap(&p->tm_fc[i], l.c_fc_w, 4 * W, S * W, gemm::CTA_NB, dt);
        if (!rc) rc = make_tmap(&p->tm_proj[i], l.c_proj_w, W, S * 4 * W, gemm::CTA_NB, dt);
        if (!rc && p->fused_ln) rc = make_tmap(&p->tm_inf[i], l.in_proj_wf, 3 * W, S * W, gemm::CTA_NB, dt);
        if (!rc && p->fused_ln) rc = make_tmap(&p->tm_fcf[i], l.c_fc_wf, 4 * W, S * W, gemm::CTA_NB, dt);
        if (!rc && p->fused_ln && l.c_proj_w_div1702) rc = make_tmap(&p->tm_projs[i], l.c_proj_w_div1702, W, S * 4 * W, gemm::CTA_NB, dt);
    }
    if (!rc && p->fused_ln) rc = make_tmap(&p->tm_xb, p->xb, rows, S * W, gemm::CTA_M, dt);
    if (!rc && p->fused_ln) {
        const eoe_vit_layer& ll = p->layers[w->n_layers - 1];
        rc = make_tmap(&p->tm_kv_w, (const uint16_t*)ll.in_proj_wf + (size_t)W * S * W, 2 * W, S * W, gemm::CTA_NB, dt);
        // K;V columns of qkv: [W, 3W) and, split, their lo halves 3W further on -- one window from column W to the row's end
        if (!rc) rc = make_tmap_ld(&p->tm_kv_st, p->qkv + W, rows, S == 2 ? 5 * W : 2 * W, S * 3 * W, 32, dt);
        if (!rc) rc = make_tmap(&p->tm_xbc, p->xb_cls, max_batch, S * W, gemm::CTA_M, dt);
        if (!rc) rc = make_tmap(&p->tm_qc_st, p->q_cls, max_batch, S * W, 32, dt);
    }
    if (rc) { eoe_vit_plan_destroy(p); return rc; }
    *plan_out = p;
    return EOE_OK;
}

extern "C" void eoe_vit_plan_destroy(eoe_vit_plan* p) {
    if (!p) return;
    for (auto& s : p->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    delete[] p->layers;
    delete[] p->tm_in;
    delete[] p->tm_out;
    delete[] p->tm_fc;
    delete[] p->tm_proj;
    delete[] p->tm_inf;
    delete[] p->tm_fcf;
    delete[] p->tm_projs;
    delete p;
}

// imgs_f32 != null: normalised fp32 NCHW input; else imgs_u8 (+ layout, norm): raw pixels, ToTensor + Normalize fused
static int vit_encode_impl(eoe_vit_plan* p, const float* imgs_f32, const uint8_t* imgs_u8, int layout, const NormParams& norm,
                           int64_t B, float* feats_out, const float* text, int64_t K, float scale, float* scores_out,
                           void* stream, const ResizeGeom* geom = nullptr) {
    cudaStream_t st = (cudaStream_t)stream;
    const eoe_vit_weights& w = p->w;
    const int W = w.width, dt = w.operand_dtype, L = p->L;
    const bool split = dt == EOE_F16X2;              // split fp16 operands: SPLIT kernel variants, fp16 everywhere else
    const int64_t M = B * L, Mp = B * p->g2;
    int rc;
    if (text && !clip_prompts_ok(K)) return EOE_ERR_SHAPE;      // before any work: not after the whole forward pass
    NvtxRange nvtx_all("eoe:vit_encode");
    // 1. patchify
    if (imgs_f32) {
        const int64_t total = B * 3 * (int64_t)w.resolution * w.resolution / 8;
        int grid = (int)((total + 255) / 256 < (int64_t)num_sms() * 16 ? (total + 255) / 256 : (int64_t)num_sms() * 16);
        if (dt == EOE_BF16) im2col_kernel<true><<<grid, 256, 0, st>>>(imgs_f32, p->patches, B, w.resolution, w.patch);
        else if (split) im2col_kernel<false, true><<<grid, 256, 0, st>>>(imgs_f32, p->patches, B, w.resolution, w.patch);
        else im2col_kernel<false><<<grid, 256, 0, st>>>(imgs_f32, p->patches, B, w.resolution, w.patch);
        if ((rc = check_launch("im2col_kernel"))) return rc;
    } else if (geom) {
        // raw [B, H, W, 3] pixels of any size: Resize + CenterCrop + ToTensor + Normalize + patchify in one kernel
        const int R = w.resolution, P = w.patch;
        const double fv = geom->scv < 1.0 ? 1.0 : geom->scv;
        const int max_rows = (int)(P * geom->scv + 4.0 * fv) + 4;               // source rows one row of patches can touch
        const size_t smem = (size_t)(2 * R + R * geom->ksh + P * geom->ksv) * sizeof(int) + (size_t)max_rows * R * 3;
        if (P > 32 || smem > 200 * 1024) return EOE_ERR_SHAPE;                   // down-scaling beyond ~12x per patch row
        auto kern = dt == EOE_BF16 ? resize_patchify_kernel<true> : (split ? resize_patchify_kernel<false, true> : resize_patchify_kernel<false>);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_cuda_error(e, "resize_patchify smem attr"); return EOE_ERR_CUDA; }
        kern<<<dim3((unsigned)(R / P), (unsigned)B), 256, smem, st>>>(imgs_u8, p->patches, R, P, *geom, norm, max_rows);
        if ((rc = check_launch("resize_patchify_kernel"))) return rc;
    } else {
        const int R = w.resolution;
        const int64_t total = (layout ? 1 : 3) * B * (int64_t)R * (R / 16);
        int grid = (int)((total + 255) / 256 < (int64_t)num_sms() * 8 ? (total + 255) / 256 : (int64_t)num_sms() * 8);
#define EOE_U8(BF, HWC) im2col_u8_kernel<BF, HWC><<<grid, 256, 0, st>>>(imgs_u8, p->patches, B, R, w.patch, norm)
        if (dt == EOE_BF16) { if (layout) EOE_U8(true, true); else EOE_U8(true, false); }
        else if (split) {
            if (layout) im2col_u8_kernel<false, true, true><<<grid, 256, 0, st>>>(imgs_u8, p->patches, B, R, w.patch, norm);
            else im2col_u8_kernel<false, false, true><<<grid, 256, 0, st>>>(imgs_u8, p->patches, B, R, w.patch, norm);
        }
        else { if (layout) EOE_U8(false, true); else EOE_U8(false, false); }
#undef EOE_U8
        if ((rc = check_launch("im2col_u8_kernel"))) return rc;
    }
    // 2. patch-embed GEMM, epilogue adds positional embedding and scatters to token rows 1..g2 of each image
    {
        gemm::Params gp{Mp, W, p->kpatch, nullptr, p->x, w.positional_embedding, p->g2, nullptr, nullptr, nullptr};
        if ((rc = timed_gemm(p, KIND_PATCH, p->tm_patches, p->tm_conv, gp, dt, EOE_EPI_PATCH_EMBED, st))) return rc;
    }
    // 3. class token + ln_pre (in place, fp32)
    if (!p->fused_ln) {
        if ((rc = layernorm_dispatch(p->x, w.ln_pre_w, w.ln_pre_b, p->x, EOE_F32, M, W, w.class_embedding,
                                     w.positional_embedding, L, st))) return rc;
    } else {
        const int grid = (int)((M + 7) / 8);
#define EOE_LN_PRE(BF, IT) ln_pre_stats_kernel<BF, IT><<<grid, 256, 0, st>>>( \
        p->x, w.ln_pre_w, w.ln_pre_b, p->xb, p->stats, p->shift, M, W, w.class_embedding, w.positional_embedding, L)
        if (dt == EOE_BF16) { if (W <= 768) EOE_LN_PRE(true, 6); else EOE_LN_PRE(true, 8); }
        else if (split)
            ln_pre_stats_kernel<false, 6, true><<<grid, 256, 0, st>>>(p->x, w.ln_pre_w, w.ln_pre_b, p->xb, p->stats, p->shift, M, W,
                                                                      w.class_embedding, w.positional_embedding, L);
        else { if (W <= 768) EOE_LN_PRE(false, 6); else EOE_LN_PRE(false, 8); }
#undef EOE_LN_PRE
        if ((rc = check_launch("ln_pre_stats_kernel"))) return rc;
    }
    // 4. transformer blocks
    const float* x_tail = p->x;          // rows that feed ln_post: token 0 of every image
    int64_t tail_stride_rows = L;
    const int epi_res = p->fused_ln ? EOE_EPI_RESIDUAL_STATS : EOE_EPI_BIAS_RESIDUAL_F32;
    float2* st_cur = p->stats;           // chunk sums of the residual stream as it is now (written by ln_pre)
    float2* st_nxt = p->stats2;
    for (int i = 0; i < w.n_layers; ++i) {
        char nvtx_name[32];
        snprintf(nvtx_name, sizeof(nvtx_name), "eoe:block %d", i);
        NvtxRange nvtx_block(nvtx_name);
        const eoe_vit_layer& l = p->layers[i];
        const bool last = (i == w.n_layers - 1);
        if (!p->fused_ln) {
            if ((rc = layernorm_dispatch(p->x, l.ln_1_w, l.ln_1_b, p->h, dt, M, W, nullptr, nullptr, L, st))) return rc;
            gemm::Params g1{M, 3 * W, W, l.in_proj_b, p->qkv, nullptr, 0, nullptr, nullptr, nullptr};
            if ((rc = timed_gemm(p, KIND_QKV, p->tm_h, p->tm_in[i], g1, dt, EOE_EPI_BIAS, st, &p->tm_qkv_st))) return rc;
        } else if (last && !(g_gemm_debug & 8)) {
            // Last block: downstream only the class-token rows matter (model.py:231), so K and V are needed for every
            // token but Q for B rows only.  K;V: the same folded GEMM over weight rows [W, 3W) into columns [W, 3W) of
            // qkv (N = 2W instead of 3W); Q: the same folded GEMM over the B compacted class-token rows.  Same operands,
            // same accumulation order as the full GEMM: the features do not change by a bit (diagnostics bit 3 = off).
            gemm::Params gkv{M, 2 * W, W, (const float*)l.in_proj_c2 + W, p->qkv + W, (const float*)l.in_proj_c1 + W, 0, st_cur,
                             nullptr, nullptr, p->shift, nullptr};
            gkv.lo_off = 3 * W;                      // split operands: lo halves of K;V sit 3W columns behind the hi halves
            if ((rc = timed_gemm(p, KIND_QKV, p->tm_xb, p->tm_kv_w, gkv, dt, EOE_EPI_LNFOLD_BIAS, st, &p->tm_kv_st))) return rc;
            gather_cls_rows_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(p->xb, st_cur, p->shift, p->xb_cls, p->stats_cls,
                                                                            p->shift_cls, B, L, W, split ? 2 * W : W);
            if ((rc = check_launch("gather_cls_rows_kernel"))) return rc;
            gemm::Params gq{B, W, W, l.in_proj_c2, p->q_cls, l.in_proj_c1, 0, p->stats_cls, nullptr, nullptr, p->shift_cls, nullptr};
            if ((rc = timed_gemm(p, KIND_QKV, p->tm_xbc, p->tm_inf[i], gq, dt, EOE_EPI_LNFOLD_BIAS, st, &p->tm_qc_st))) return rc;
        } else {
            // ln_1 folded into the QKV GEMM: A = 16-bit residual stream, epilogue rstd*(acc - mean*c1) + c2
            gemm::Params g1{M, 3 * W, W, l.in_proj_c2, p->qkv, l.in_proj_c1, 0, st_cur, nullptr, nullptr, p->shift, nullptr};
            if ((rc = timed_gemm(p, KIND_QKV, p->tm_xb, p->tm_inf[i], g1, dt, EOE_EPI_LNFOLD_BIAS, st, &p->tm_qkv_st))) return rc;
        }
        if (!last) {
            if ((rc = attention_dispatch(p->qkv, p->h, B, L, w.heads, dt, st, &p->tm_qkv, &p->tm_ho128, &p->tm_ho72))) return rc;
            gemm::Params g2{M, W, W, l.out_proj_b, p->x, nullptr, 0, st_cur, st_nxt, p->xb, nullptr, p->shift};
            if ((rc = timed_gemm(p, KIND_OUT, p->tm_h, p->tm_out[i], g2, dt, epi_res, st))) return rc;
            { float2* t = st_cur; st_cur = st_nxt; st_nxt = t; }
            if (!p->fused_ln) {
                if ((rc = layernorm_dispatch(p->x, l.ln_2_w, l.ln_2_b, p->h, dt, M, W, nullptr, nullptr, L, st))) return rc;
                gemm::Params g3{M, 4 * W, W, l.c_fc_b, p->u, nullptr, 0, nullptr, nullptr, nullptr};
                if ((rc = timed_gemm(p, KIND_FC, p->tm_h, p->tm_fc[i], g3, dt, EOE_EPI_BIAS_QUICKGELU, st, &p->tm_u_st))) return rc;
            } else {
                gemm::Params g3{M, 4 * W, W, l.c_fc_c2, p->u, l.c_fc_c1, 0, st_cur, nullptr, nullptr, p->shift, nullptr};
                const bool gelu_x = l.c_proj_w_div1702 && !(g_gemm_debug & 64);      // diagnostics bit 6: plain QuickGELU epilogue
                const int epi_fc = gelu_x ? EOE_EPI_LNFOLD_QUICKGELU_X1702 : EOE_EPI_LNFOLD_QUICKGELU;
                if ((rc = timed_gemm(p, KIND_FC, p->tm_xb, p->tm_fcf[i], g3, dt, epi_fc, st, &p->tm_u_st))) return rc;
            }
            gemm::Params g4{M, W, 4 * W, l.c_proj_b, p->x, nullptr, 0, st_cur, st_nxt, p->xb, nullptr, p->shift};
            const CUtensorMap& tm_pw = (p->fused_ln && l.c_proj_w_div1702 && !(g_gemm_debug & 64)) ? p->tm_projs[i] : p->tm_proj[i];
            if ((rc = timed_gemm(p, KIND_PROJ, p->tm_u, tm_pw, g4, dt, epi_res, st))) return rc;
            { float2* t = st_cur; st_cur = st_nxt; st_nxt = t; }
        } else {
            // last block: only the class-token rows are needed downstream (model.py:231)
            const unsigned grid = (unsigned)((B * w.heads + 3) / 4);
            const uint16_t* qc = (p->fused_ln && !(g_gemm_debug & 8)) ? p->q_cls : nullptr;
            if (dt == EOE_BF16) attention_cls_kernel<true><<<grid, 128, 0, st>>>(p->qkv, qc, p->x, p->h_cls, p->x_cls, B, L, w.heads);
            else if (split) attention_cls_kernel<false, true><<<grid, 128, 0, st>>>(p->qkv, qc, p->x, p->h_cls, p->x_cls, B, L, w.heads);
            else attention_cls_kernel<false><<<grid, 128, 0, st>>>(p->qkv, qc, p->x, p->h_cls, p->x_cls, B, L, w.heads);
            if ((rc = check_launch("attention_cls_kernel"))) return rc;
            gemm::Params g2{B, W, W, l.out_proj_b, p->x_cls, nullptr, 0, nullptr, nullptr, nullptr};
            if ((rc = timed_gemm(p, KIND_OUT, p->tm_hc, p->tm_out[i], g2, dt, EOE_EPI_BIAS_RESIDUAL_F32, st))) return rc;
            if ((rc = layernorm_dispatch(p->x_cls, l.ln_2_w, l.ln_2_b, p->h_cls, dt, B, W, nullptr, nullptr, 1, st))) return rc;
            gemm::Params g3{B, 4 * W, W, l.c_fc_b, p->u_cls, nullptr, 0, nullptr, nullptr, nullptr};
            if ((rc = timed_gemm(p, KIND_FC, p->tm_hc, p->tm_fc[i], g3, dt, EOE_EPI_BIAS_QUICKGELU, st, &p->tm_uc_st))) return rc;
            gemm::Params g4{B, W, 4 * W, l.c_proj_b, p->x_cls, nullptr, 0, nullptr, nullptr, nullptr};
            if ((rc = timed_gemm(p, KIND_PROJ, p->tm_uc, p->tm_proj[i], g4, dt, EOE_EPI_BIAS_RESIDUAL_F32, st))) return rc;
            x_tail = p->x_cls;
            tail_stride_rows = 1;
        }
    }
    // 5. ln_post + proj (+ zero-shot score head)
    NvtxRange nvtx_tail("eoe:tail + score head");
    float* feats = feats_out ? feats_out : p->feats;
    {
        const dim3 grid((unsigned)((B + kTailImgs - 1) / kTailImgs), (unsigned)((w.embed_dim + kTailCols - 1) / kTailCols));
        tail_kernel<<<grid, 256, (kTailImgs * W + kTailImgs * kTailCols) * sizeof(float), st>>>(
            x_tail, w.ln_post_w, w.ln_post_b, w.proj, feats, B, (int)tail_stride_rows, W, w.embed_dim);
        if ((rc = check_launch("tail_kernel"))) return rc;
    }
    if (text) {
        if ((rc = clip_score_f32(feats, text, B, w.embed_dim, K, scale, scores_out, st))) return rc;
    }
    return EOE_OK;
}

extern "C" int eoe_vit_encode(eoe_vit_plan* p, const float* imgs, int64_t B, float* feats_out, const float* text,
                              int64_t K, float scale, float* scores_out, void* stream) {
    if (!p || !imgs || B <= 0 || B > p->max_batch) return EOE_ERR_ARG;
    if ((text == nullptr) != (scores_out == nullptr)) return EOE_ERR_ARG;
    if ((uintptr_t)imgs % 16 != 0) return EOE_ERR_ALIGN;
    return vit_encode_impl(p, imgs, nullptr, 0, NormParams{}, B, feats_out, text, K, scale, scores_out, stream);
}

extern "C" int eoe_vit_encode_u8(eoe_vit_plan* p, const uint8_t* imgs, int layout, const float* mean_host,
                                 const float* std_host, int64_t B, float* feats_out, const float* text, int64_t K,
                                 float scale, float* scores_out, void* stream) {
    if (!p || !imgs || !mean_host || !std_host || B <= 0 || B > p->max_batch) return EOE_ERR_ARG;
    if (layout != EOE_LAYOUT_NCHW && layout != EOE_LAYOUT_NHWC) return EOE_ERR_ARG;
    if ((text == nullptr) != (scores_out == nullptr)) return EOE_ERR_ARG;
    if ((uintptr_t)imgs % 16 != 0) return EOE_ERR_ALIGN;
    if (p->w.patch % 16 != 0 || p->w.resolution % 16 != 0) return EOE_ERR_SHAPE;
    NormParams np;
    for (int c = 0; c < 3; ++c) {
        np.mean[c] = mean_host[c];
        np.stdv[c] = std_host[c];
        if (!(std_host[c] > 0.f)) return EOE_ERR_ARG;
    }
    return vit_encode_impl(p, nullptr, imgs, layout == EOE_LAYOUT_NHWC, np, B, feats_out, text, K, scale, scores_out, stream);
}

extern "C" int eoe_vit_encode_u8_resize(eoe_vit_plan* p, const uint8_t* imgs, int64_t H, int64_t W, const float* mean_host,
                                        const float* std_host, int64_t B, float* feats_out, const float* text, int64_t K,
                                        float scale, float* scores_out, void* stream) {
    if (!p || !imgs || !mean_host || !std_host || B <= 0 || B > p->max_batch) return EOE_ERR_ARG;
    if ((text == nullptr) != (scores_out == nullptr)) return EOE_ERR_ARG;
    if (p->w.patch % 8 != 0 || p->w.resolution % 8 != 0) return EOE_ERR_SHAPE;
    ResizeGeom gm;
    int rc = resize_geometry(H, W, p->w.resolution, &gm);
    if (rc) return rc;
    NormParams np;
    for (int c = 0; c < 3; ++c) {
        np.mean[c] = mean_host[c];
        np.stdv[c] = std_host[c];
        if (!(std_host[c] > 0.f)) return EOE_ERR_ARG;
    }
    return vit_encode_impl(p, nullptr, imgs, 1, np, B, feats_out, text, K, scale, scores_out, stream, &gm);
}

// Test hook: the transform alone.  out [B, R, R, 3] uint8 = CenterCrop(Resize(imgs [B, H, W, 3])) reconstructed from the
// patches the kernel wrote (mean 0 / std 1/255 make the LUT the identity on the pixel value).
extern "C" int eoe_resize_geometry(int64_t H, int64_t W, int n_px, int* out6_host) {
    ResizeGeom gm;
    int rc = resize_geometry(H, W, n_px, &gm);
    if (rc) return rc;
    out6_host[0] = gm.nh; out6_host[1] = gm.nw; out6_host[2] = gm.top; out6_host[3] = gm.left; out6_host[4] = gm.ksh; out6_host[5] = gm.ksv;
    return EOE_OK;
}

extern "C" int eoe_vit_profile_enable(eoe_vit_plan* p, int enable) {
    if (!p) return EOE_ERR_ARG;
    p->profile = enable != 0;
    p->spans_used = 0;
    return EOE_OK;
}

extern "C" int eoe_vit_profile_read(eoe_vit_plan* p, double* ms_out, int64_t* launches_out, double* flops_out) {
    if (!p || !ms_out || !launches_out || !flops_out) return EOE_ERR_ARG;
    for (int k = 0; k < 5; ++k) { ms_out[k] = 0.0; launches_out[k] = 0; flops_out[k] = 0.0; }
    for (size_t i = 0; i < p->spans_used; ++i) {
        eoe_vit_plan::Span& s = p->spans[i];
        cudaError_t e = cudaEventSynchronize(s.b);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, s.a, s.b);
        if (e != cudaSuccess) { set_cuda_error(e, "eoe_vit_profile_read"); return EOE_ERR_CUDA; }
        ms_out[s.kind] += ms;
        launches_out[s.kind] += 1;
        flops_out[s.kind] += s.flops;
    }
    p->spans_used = 0;
    return EOE_OK;
}

// ------------------------------------------------------------------------------------------ building blocks (tests)
extern "C" int eoe_gemm(const void* A, const void* Wt, const float* bias, void* out, int64_t M, int64_t N, int64_t K,
                        int operand_dtype, int epilogue, const float* aux, int64_t aux_i, void* stream) {
    if (!A || !Wt || !out) return EOE_ERR_ARG;
    int rc = gemm_check(M, N, K, operand_dtype);
    if (rc) return rc;
    if (epilogue == EOE_EPI_PATCH_EMBED && (!aux || aux_i <= 0)) return EOE_ERR_ARG;
    if ((uintptr_t)A % 16 != 0 || (uintptr_t)Wt % 16 != 0 || (uintptr_t)out % 16 != 0 || (uintptr_t)bias % 16 != 0)
        return EOE_ERR_ALIGN;
    CUtensorMap ta, tb;
    const int S = operand_dtype == EOE_F16X2 ? 2 : 1;       // split fp16: A [M, 2K], W [N, 2K], 16-bit out [M, 2N]
    if ((rc = make_tmap(&ta, A, M, S * K, gemm::CTA_M, operand_dtype))) return rc;
    if ((rc = make_tmap(&tb, Wt, N, S * K, gemm::CTA_NB, operand_dtype))) return rc;
    gemm::Params p{M, N, K, bias, out, aux, aux_i, nullptr, nullptr, nullptr};
    return gemm_launch(ta, tb, p, operand_dtype, epilogue, (cudaStream_t)stream);
}

extern "C" void eoe_debug_set(int flags) { g_gemm_debug = flags; }
extern "C" int eoe_debug_gemm_prof(unsigned long long* out4_host) {
    return cudaMemcpyFromSymbol(out4_host, gemm::g_gemm_prof, 4 * sizeof(unsigned long long)) == cudaSuccess ? 0 : EOE_ERR_CUDA;
}
extern "C" int eoe_debug_max_clusters(int clp) { return (clp == 1 || clp == 2) ? g_last_max_clusters[clp - 1] : -1; }


extern "C" int eoe_gemm_lnfold(const void* A, const void* Wf, const float* c1, const float* c2, const float* stats,
                               const float* shift, void* out, int64_t M, int64_t N, int64_t K, int operand_dtype,
                               int quick_gelu, void* stream) {
    if (!A || !Wf || !c1 || !c2 || !stats || !out) return EOE_ERR_ARG;
    int rc = gemm_check(M, N, K, operand_dtype);
    if (rc) return rc;
    if (K % 256 != 0 || K > 128 * gemm::MAX_NCH) return EOE_ERR_SHAPE;      // K in {256, 512, 768}
    if ((uintptr_t)A % 16 != 0 || (uintptr_t)Wf % 16 != 0 || (uintptr_t)out % 16 != 0 || (uintptr_t)stats % 16 != 0 ||
        (uintptr_t)c1 % 16 != 0 || (uintptr_t)c2 % 16 != 0 || (uintptr_t)shift % 16 != 0) return EOE_ERR_ALIGN;
    CUtensorMap ta, tb;
    const int S = operand_dtype == EOE_F16X2 ? 2 : 1;
    if ((rc = make_tmap(&ta, A, M, S * K, gemm::CTA_M, operand_dtype))) return rc;
    if ((rc = make_tmap(&tb, Wf, N, S * K, gemm::CTA_NB, operand_dtype))) return rc;
    gemm::Params p{M, N, K, c2, out, c1, 0, reinterpret_cast<const float2*>(stats), nullptr, nullptr, shift, nullptr};
    return gemm_launch(ta, tb, p, operand_dtype,
                       quick_gelu == 2 ? EOE_EPI_LNFOLD_QUICKGELU_X1702 : (quick_gelu ? EOE_EPI_LNFOLD_QUICKGELU : EOE_EPI_LNFOLD_BIAS),
                       (cudaStream_t)stream);
}

extern "C" int eoe_gemm_residual_stats(const void* A, const void* Wt, const float* bias, const float* stats_in, float* x,
                                       void* xb_out, float* stats_out, float* shift_out, int64_t M, int64_t N, int64_t K,
                                       int operand_dtype, void* stream) {
    if (!A || !Wt || !x || !xb_out || !stats_out) return EOE_ERR_ARG;
    if (stats_in == stats_out) return EOE_ERR_ARG;          // other CTAs still read the previous sums while these are written
    if (N > 128 * gemm::MAX_NCH && stats_in) return EOE_ERR_SHAPE;
    if ((uintptr_t)stats_in % 16 != 0) return EOE_ERR_ALIGN;
    int rc = gemm_check(M, N, K, operand_dtype);
    if (rc) return rc;
    if ((uintptr_t)A % 16 != 0 || (uintptr_t)Wt % 16 != 0 || (uintptr_t)x % 16 != 0 || (uintptr_t)xb_out % 8 != 0 || (uintptr_t)bias % 16 != 0 ||
        (uintptr_t)stats_out % 8 != 0) return EOE_ERR_ALIGN;
    CUtensorMap ta, tb;
    const int S = operand_dtype == EOE_F16X2 ? 2 : 1;
    if ((rc = make_tmap(&ta, A, M, S * K, gemm::CTA_M, operand_dtype))) return rc;
    if ((rc = make_tmap(&tb, Wt, N, S * K, gemm::CTA_NB, operand_dtype))) return rc;
    gemm::Params p{M, N, K, bias, x, nullptr, 0, reinterpret_cast<const float2*>(stats_in), reinterpret_cast<float2*>(stats_out),
                   reinterpret_cast<uint16_t*>(xb_out), nullptr, shift_out};
    return gemm_launch(ta, tb, p, operand_dtype, EOE_EPI_RESIDUAL_STATS, (cudaStream_t)stream);
}

extern "C" int eoe_vit_fold_layernorm(const float* w_f32, const float* ln_w, const float* ln_b, const float* bias,
                                      int64_t N, int64_t K, int operand_dtype, void* w_folded_out, float* c1_out,
                                      float* c2_out, void* stream) {
    if (!w_f32 || !ln_w || !ln_b || !w_folded_out || !c1_out || !c2_out || N <= 0 || K <= 0) return EOE_ERR_ARG;
    if (operand_dtype != EOE_BF16 && operand_dtype != EOE_F16 && operand_dtype != EOE_F16X2) return EOE_ERR_DTYPE;
    if (K % 4 != 0) return EOE_ERR_SHAPE;
    if ((uintptr_t)w_f32 % 16 != 0 || (uintptr_t)ln_w % 16 != 0 || (uintptr_t)ln_b % 16 != 0 || (uintptr_t)w_folded_out % 8 != 0)
        return EOE_ERR_ALIGN;
    const int grid = (int)((N + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (operand_dtype == EOE_BF16)
        fold_ln_kernel<true><<<grid, 256, 0, st>>>(w_f32, ln_w, ln_b, bias, N, K, (uint16_t*)w_folded_out, c1_out, c2_out);
    else if (operand_dtype == EOE_F16X2)
        fold_ln_kernel<false, true><<<grid, 256, 0, st>>>(w_f32, ln_w, ln_b, bias, N, K, (uint16_t*)w_folded_out, c1_out, c2_out);
    else
        fold_ln_kernel<false><<<grid, 256, 0, st>>>(w_f32, ln_w, ln_b, bias, N, K, (uint16_t*)w_folded_out, c1_out, c2_out);
    return check_launch("fold_ln_kernel");
}

extern "C" int eoe_layernorm(const float* x, const float* w, const float* b, void* y, int out_dtype, int64_t M,
                             int64_t width, void* stream) {
    if (!x || !w || !b || !y || M <= 0 || width <= 0) return EOE_ERR_ARG;
    return layernorm_dispatch(x, w, b, y, out_dtype, M, width, nullptr, nullptr, 1, (cudaStream_t)stream);
}

extern "C" int eoe_attention(const void* qkv, void* out, int64_t B, int64_t L, int64_t heads, int operand_dtype,
                             void* stream) {
    if (!qkv || !out) return EOE_ERR_ARG;
    return attention_dispatch(qkv, out, B, L, heads, operand_dtype, (cudaStream_t)stream);
}

extern "C" int eoe_attention_causal(const void* qkv, void* out, int64_t B, int64_t L, int64_t heads, int operand_dtype,
                                    void* stream) {
    if (!qkv || !out) return EOE_ERR_ARG;
    if ((uintptr_t)qkv % 16 != 0) return EOE_ERR_ALIGN;
    if (operand_dtype == EOE_F16X2) {                // split fp16 pairs: fp32 CUDA-core kernel (the text tower is 2 310 rows)
        if (B <= 0 || L <= 0 || heads <= 0) return EOE_ERR_ARG;
        if (L > 224 || B * heads * L > 0x7fffffff) return EOE_ERR_SHAPE;
        attention_causal_split_kernel<<<(unsigned)((B * heads * L + 3) / 4), 128, 0, (cudaStream_t)stream>>>(
            (const uint16_t*)qkv, (uint16_t*)out, B, (int)L, (int)heads);
        return check_launch("attention_causal_split_kernel");
    }
    return attention_dispatch(qkv, out, B, L, heads, operand_dtype, (cudaStream_t)stream, nullptr, nullptr, nullptr, true);
}

extern "C" int eoe_text_embed(const int64_t* tokens, const float* token_embedding, const float* positional_embedding,
                              float* x, int64_t n, int64_t ctx, int64_t width, int64_t vocab, void* stream) {
    if (!tokens || !token_embedding || !positional_embedding || !x || n <= 0 || ctx <= 0 || vocab <= 0) return EOE_ERR_ARG;
    if (width <= 0 || width % 4 != 0) return EOE_ERR_SHAPE;
    if ((uintptr_t)token_embedding % 16 != 0 || (uintptr_t)positional_embedding % 16 != 0 || (uintptr_t)x % 16 != 0) return EOE_ERR_ALIGN;
    const int64_t rows = n * ctx;
    text_embed_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        tokens, token_embedding, positional_embedding, x, rows, (int)ctx, (int)width, vocab);
    return check_launch("text_embed_kernel");
}

extern "C" int eoe_text_tail(const float* x, const int64_t* tokens, const float* ln_w, const float* ln_b, const float* proj,
                             float* feats, int64_t n, int64_t ctx, int64_t width, int64_t embed, void* stream) {
    if (!x || !tokens || !ln_w || !ln_b || !proj || !feats || n <= 0 || ctx <= 0) return EOE_ERR_ARG;
    if (width <= 0 || width % 2 != 0 || embed <= 0 || width > 2048) return EOE_ERR_SHAPE;     // 4 rows of fp32 in static-limit shared memory
    const dim3 grid((unsigned)((n + kTailImgs - 1) / kTailImgs), (unsigned)((embed + kTailCols - 1) / kTailCols));
    tail_kernel<<<grid, 256, (kTailImgs * width + kTailImgs * kTailCols) * sizeof(float), (cudaStream_t)stream>>>(
        x, ln_w, ln_b, proj, feats, n, (int)ctx, (int)width, (int)embed, tokens);
    return check_launch("tail_kernel(text)");
}
