// Fused loss / anomaly-score heads (HBM-bound row kernels, sm_100a).
//
//   hsc_*   : HSCTrainer.loss + autograd backward + compute_anomaly_score  (reference src/eoe/training/hsc.py:12-21)
//   bce_*   : BCETrainer.loss + backward + compute_anomaly_score           (reference src/eoe/training/bce.py:15-20)
//   clip_*  : ADClipTrainer.compute_anomaly_score / loss + backward         (reference src/eoe/training/clip.py:66-103)
//   dsad_*  : DSADTrainer.loss + backward + compute_anomaly_score           (reference src/eoe/training/dsad.py:13-22)
//   dsvdd_* : DSVDDTrainer.loss + backward + compute_anomaly_score          (reference src/eoe/training/dsvdd.py:23-27)
//   focal_* : FocalTrainer.loss (FocalLoss) + backward + score              (reference src/eoe/training/focal.py:11-39)
//
// Layout: features [n,d] row-major.  One warp owns ROWS rows at a time and keeps them in registers
// between the reduction pass (||z||^2, logits) and the gradient pass, so every feature byte is read from
// HBM exactly once and every gradient byte written exactly once (algorithmic bytes = traffic).
// Lane l holds elements {(it*32 + l)*4 .. +3}: 16-byte (fp32) / 8-byte (16-bit) fully coalesced accesses.
#include "common.cuh"

namespace eoe {

constexpr int kHeadBlock = 256;
constexpr int kHeadWarps = kHeadBlock / 32;

// ------------------------------------------------------------------------------------------ HSC
// Reference evaluation order (hsc.py:18-20): nrm = ||z||; dist = sqrt(nrm^2 + 1) - 1; score = 1 - exp(-dist);
// anomalous loss = -log(score + 1e-9).  Kept literally (no expm1/log1p) because the cancellation for small
// dist is part of the reference's answer.
struct HscRow {
    float dist, score, loss, coef;
};
__device__ __forceinline__ HscRow hsc_row_math(float sumsq, bool nominal, float inv_n) {
    HscRow r;
    float nrm = sqrtf(sumsq);
    float rad = sqrtf(nrm * nrm + 1.0f);
    r.dist = rad - 1.0f;
    float e = expf(-r.dist);
    r.score = 1.0f - e;
    float g;
    if (nominal) {
        r.loss = r.dist;
        g = 1.0f;
    } else {
        float s = r.score + 1e-9f;
        r.loss = -logf(s);
        g = -e / s;
    }
    r.coef = g * inv_n / rad;      // d loss / d z_ij = coef * z_ij
    return r;
}

// The three "row norm" objectives share one kernel: s = sum_j (z_j - c_j)^2 (c = 0 unless DSVDD), then per row
//   HSC   (hsc.py:17-21)    loss = dist | -log(score + 1e-9),            score = 1 - exp(-(sqrt(s+1)-1))
//   DSAD  (dsad.py:13-22)   loss = s | (s + 1e-9)^-1 (s via sqrt then square as in the reference), score as HSC
//   DSVDD (dsvdd.py:23-27)  loss = score = s (labels unused)
// and d loss / d z_j = coef * (z_j - c_j).
enum { ROW_HSC = 0, ROW_DSAD = 1, ROW_DSVDD = 2 };
template <int MODE>
__device__ __forceinline__ HscRow row_math(float sumsq, bool nominal, float inv_n) {
    if (MODE == ROW_HSC) return hsc_row_math(sumsq, nominal, inv_n);
    HscRow r;
    if (MODE == ROW_DSVDD) {
        r.dist = sumsq;
        r.score = sumsq;
        r.loss = sumsq;
        r.coef = 2.0f * inv_n;
        return r;
    }
    const float nrm = sqrtf(sumsq);
    const float d2 = nrm * nrm;                        // torch.norm(z, 2, dim=1) ** 2   (dsad.py:19)
    r.dist = sqrtf(d2 + 1.0f) - 1.0f;                  // dsad.py:14
    r.score = 1.0f - expf(-r.dist);
    if (nominal) {
        r.loss = d2;
        r.coef = 2.0f * inv_n;
    } else {
        const float t = d2 + 1e-9f;
        r.loss = 1.0f / t;                             // (dists + 1e-9) ** (-1)          (dsad.py:21)
        r.coef = -2.0f * inv_n / (t * t);
    }
    return r;
}

template <typename T, int VEC, int ITERS, int ROWS, int MODE>
__global__ void __launch_bounds__(kHeadBlock)
hsc_rows_kernel(const T* __restrict__ z, const int64_t* __restrict__ labels, int64_t n, int d,
                int64_t nominal_label, float* __restrict__ scores, T* __restrict__ grad,
                HeadWorkspace* ws, float* loss_out, float inv_n_f, double inv_n, const float* __restrict__ center) {
    const int lane = threadIdx.x & 31;
    const int nvec = d / VEC;
    const int64_t warp0 = (int64_t)blockIdx.x * kHeadWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kHeadWarps;
    float loss_acc = 0.f;

    for (int64_t base = warp0 * ROWS; base < n; base += nwarps * ROWS) {
        float v[ROWS][ITERS][VEC];
        float ss[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int64_t row = base + r;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const int vi = it * 32 + lane;
                if (row < n && vi < nvec) {
                    loadv_stream<T, VEC>(z + row * d + vi * VEC, v[r][it]);
                    if (MODE == ROW_DSVDD) {           // (features - center), dsvdd.py:24,27
#pragma unroll
                        for (int e = 0; e < VEC; e += 4) {
                            const float4 c4 = __ldg(reinterpret_cast<const float4*>(center + vi * VEC + e));
                            v[r][it][e] -= c4.x; v[r][it][e + 1] -= c4.y; v[r][it][e + 2] -= c4.z; v[r][it][e + 3] -= c4.w;
                        }
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) v[r][it][e] = 0.f;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            float s = 0.f;
#pragma unroll
            for (int it = 0; it < ITERS; ++it)
#pragma unroll
                for (int e = 0; e < VEC; ++e) s += v[r][it][e] * v[r][it][e];
            ss[r] = warp_sum(s);
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int64_t row = base + r;
            if (row >= n) break;
            const bool nominal = labels ? (labels[row] == nominal_label) : true;
            HscRow h = row_math<MODE>(ss[r], nominal, inv_n_f);
            if (lane == 0) {
                if (scores) scores[row] = h.score;
                loss_acc += h.loss;
            }
            if (grad) {
#pragma unroll
                for (int it = 0; it < ITERS; ++it) {
                    const int vi = it * 32 + lane;
                    if (vi < nvec) {
                        float g[VEC];
#pragma unroll
                        for (int e = 0; e < VEC; ++e) g[e] = h.coef * v[r][it][e];
                        storev_stream<T, VEC>(grad + row * d + vi * VEC, g);
                    }
                }
            }
        }
    }
    if (loss_out) grid_mean_finish<kHeadBlock>(loss_acc, ws, loss_out, inv_n);
}

// Any d / any alignment: scalar lane-strided loops, second pass re-reads the row (L1/L2 hit).
template <typename T, int MODE>
__global__ void __launch_bounds__(kHeadBlock)
hsc_rows_generic_kernel(const T* __restrict__ z, const int64_t* __restrict__ labels, int64_t n, int64_t d,
                        int64_t nominal_label, float* __restrict__ scores, T* __restrict__ grad,
                        HeadWorkspace* ws, float* loss_out, float inv_n_f, double inv_n, const float* __restrict__ center) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kHeadWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kHeadWarps;
    float loss_acc = 0.f;
    for (int64_t row = warp0; row < n; row += nwarps) {
        const T* zr = z + row * d;
        float s = 0.f;
        for (int64_t j = lane; j < d; j += 32) {
            float x = to_f32<T>(zr[j]) - (MODE == ROW_DSVDD ? center[j] : 0.f);
            s += x * x;
        }
        s = warp_sum(s);
        const bool nominal = labels ? (labels[row] == nominal_label) : true;
        HscRow h = row_math<MODE>(s, nominal, inv_n_f);
        if (lane == 0) {
            if (scores) scores[row] = h.score;
            loss_acc += h.loss;
        }
        if (grad)
            for (int64_t j = lane; j < d; j += 32)
                grad[row * d + j] = from_f32<T>(h.coef * (to_f32<T>(zr[j]) - (MODE == ROW_DSVDD ? center[j] : 0.f)));
    }
    if (loss_out) grid_mean_finish<kHeadBlock>(loss_acc, ws, loss_out, inv_n);
}

static inline int head_grid(int64_t units, int units_per_block) {
    int64_t blocks = (units + units_per_block - 1) / units_per_block;
    const int64_t cap = (int64_t)kNumSMs * 8;        // 8 resident 256-thread CTAs per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <typename T, int MODE = ROW_HSC>
static int hsc_launch(const void* z_, const int64_t* labels, int64_t n, int64_t d, int64_t nominal,
                      float* loss_out, float* scores, void* grad_, void* ws_, cudaStream_t st,
                      const float* center = nullptr) {
    const T* z = (const T*)z_;
    T* grad = (T*)grad_;
    HeadWorkspace* ws = (HeadWorkspace*)ws_;
    const float inv_n_f = 1.0f / (float)n;
    const double inv_n = 1.0 / (double)n;
    constexpr int VEC = sizeof(T) == 4 ? 4 : 8;            // 16 bytes per lane per access
    const bool vec_ok = (d % VEC == 0) && d <= 32 * VEC * 8 && ((uintptr_t)z % 16 == 0) && (!grad || (uintptr_t)grad % 16 == 0) &&
                        (uintptr_t)center % 16 == 0;
    if (!vec_ok) {
        hsc_rows_generic_kernel<T, MODE><<<head_grid(n, kHeadWarps), kHeadBlock, 0, st>>>(
            z, labels, n, d, nominal, scores, grad, ws, loss_out, inv_n_f, inv_n, center);
        return check_launch("hsc_rows_generic_kernel");
    }
    const int iters = (int)((d / VEC + 31) / 32);
#define EOE_HSC_CASE(IT, RW)                                                                           \
    hsc_rows_kernel<T, VEC, IT, RW, MODE><<<head_grid(n, kHeadWarps * RW), kHeadBlock, 0, st>>>(        \
        z, labels, n, (int)d, nominal, scores, grad, ws, loss_out, inv_n_f, inv_n, center)
    if (iters <= 1) EOE_HSC_CASE(1, 4);
    else if (iters <= 2) EOE_HSC_CASE(2, VEC == 4 ? 4 : 2);
    else if (iters <= 4) EOE_HSC_CASE(4, VEC == 4 ? 2 : 1);
    else EOE_HSC_CASE(8, 1);
#undef EOE_HSC_CASE
    return check_launch("hsc_rows_kernel");
}

// ------------------------------------------------------------------------------------------ BCE
// bce.py:19-20 -> torch binary_cross_entropy_with_logits (mean): max(x,0) - x*y + log1p(exp(-|x|)).
struct BceOut {
    float loss, sig;
};
__device__ __forceinline__ BceOut bce_math(float x, float y) {
    BceOut o;
    float e = expf(-fabsf(x));
    o.loss = fmaxf(x, 0.f) - x * y + log1pf(e);
    float inv = 1.0f / (1.0f + e);
    o.sig = x >= 0.f ? inv : e * inv;          // sigmoid(x) without overflow; NaN propagates
    if (x != x) o.sig = x;
    return o;
}

// FocalLoss (focal.py:19-24, gamma = 2, eps = 1e-7): bce as above, pt = clamp(exp(-bce), eps, 1 - eps),
// F = (1 - pt)^gamma * bce.  d F / d x = (sigmoid(x) - y) * [ (1-pt)^g + 1{eps <= exp(-bce) <= 1-eps} * g (1-pt)^(g-1) pt bce ]
// (torch's clamp passes the gradient on the closed interval).  FOCAL turns bce_kernel into that objective.
struct FocalParams { float gamma, eps; };
__device__ __forceinline__ void focal_math(const BceOut& o, float y, const FocalParams fp, float& loss, float& dldx) {
    const float pt_raw = expf(-o.loss);
    const float pt = fminf(fmaxf(pt_raw, fp.eps), 1.0f - fp.eps);
    const float q = 1.0f - pt;
    const float qg = fp.gamma == 2.0f ? q * q : powf(q, fp.gamma);
    const float qg1 = fp.gamma == 2.0f ? q : powf(q, fp.gamma - 1.0f);
    loss = qg * o.loss;
    const bool inside = pt_raw >= fp.eps && pt_raw <= 1.0f - fp.eps;
    dldx = (o.sig - y) * (qg + (inside ? fp.gamma * qg1 * pt * o.loss : 0.f));
    if (pt_raw != pt_raw) { loss = pt_raw; dldx = pt_raw; }      // NaN in -> NaN out (fminf/fmaxf would drop it)
}

template <typename T, bool VEC, bool FOCAL = false>
__global__ void __launch_bounds__(kHeadBlock)
bce_kernel(const T* __restrict__ x, const int64_t* __restrict__ labels, int64_t n, int flip_score,
           float* __restrict__ scores, T* __restrict__ grad, HeadWorkspace* ws, float* loss_out, float inv_n_f,
           double inv_n, FocalParams fp = FocalParams{2.0f, 1e-7f}) {
    const int64_t tid = (int64_t)blockIdx.x * kHeadBlock + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * kHeadBlock;
    float loss_acc = 0.f;
    int64_t done = 0;
    if (VEC) {
        const int64_t n4 = n >> 2;
        // two independent 4-sample groups per iteration: 6 x 16-byte loads in flight per thread before any store
        for (int64_t i0 = tid; i0 < n4; i0 += 2 * nthreads) {
            const int64_t idx[2] = {i0, i0 + nthreads};
            float xv[2][4];
            longlong2 lb[2][2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (idx[u] < n4) {
                    load4_stream<T>(x + idx[u] * 4, xv[u]);
                    if (labels) {
                        lb[u][0] = __ldg(reinterpret_cast<const longlong2*>(labels + idx[u] * 4));
                        lb[u][1] = __ldg(reinterpret_cast<const longlong2*>(labels + idx[u] * 4) + 1);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (idx[u] >= n4) continue;
                float y[4] = {0.f, 0.f, 0.f, 0.f};
                if (labels) { y[0] = (float)lb[u][0].x; y[1] = (float)lb[u][0].y; y[2] = (float)lb[u][1].x; y[3] = (float)lb[u][1].y; }
                float sc[4], g[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    BceOut o = bce_math(xv[u][j], y[j]);
                    sc[j] = flip_score ? 1.0f - o.sig : o.sig;
                    if (FOCAL) {
                        float fl, dl;
                        focal_math(o, y[j], fp, fl, dl);
                        loss_acc += fl;
                        g[j] = dl * inv_n_f;
                    } else {
                        loss_acc += o.loss;
                        g[j] = (o.sig - y[j]) * inv_n_f;
                    }
                }
                if (scores) store4_stream<float>(scores + idx[u] * 4, sc);
                if (grad) store4_stream<T>(grad + idx[u] * 4, g);
            }
        }
        done = n4 << 2;
    }
    for (int64_t i = done + tid; i < n; i += nthreads) {
        const float xv = to_f32<T>(x[i]);
        const float y = labels ? (float)labels[i] : 0.f;
        BceOut o = bce_math(xv, y);
        float gi;
        if (FOCAL) {
            float fl, dl;
            focal_math(o, y, fp, fl, dl);
            loss_acc += fl;
            gi = dl * inv_n_f;
        } else {
            loss_acc += o.loss;
            gi = (o.sig - y) * inv_n_f;
        }
        if (scores) scores[i] = flip_score ? 1.0f - o.sig : o.sig;
        if (grad) grad[i] = from_f32<T>(gi);
    }
    if (loss_out) grid_mean_finish<kHeadBlock>(loss_acc, ws, loss_out, inv_n);
}

template <typename T, bool FOCAL = false>
static int bce_launch(const void* x_, const int64_t* labels, int64_t n, int64_t nominal, float* loss_out,
                      float* scores, void* grad_, void* ws_, cudaStream_t st, FocalParams fp = FocalParams{2.0f, 1e-7f}) {
    const T* x = (const T*)x_;
    T* grad = (T*)grad_;
    const float inv_n_f = 1.0f / (float)n;
    const double inv_n = 1.0 / (double)n;
    const size_t vb = 4 * sizeof(T);
    const bool vec = ((uintptr_t)x % vb == 0) && (!grad || (uintptr_t)grad % vb == 0) &&
                     (!labels || (uintptr_t)labels % 16 == 0) && (!scores || (uintptr_t)scores % 16 == 0);
    const int grid = head_grid(n, kHeadBlock * 8);
    if (vec)
        bce_kernel<T, true, FOCAL><<<grid, kHeadBlock, 0, st>>>(x, labels, n, nominal != 0, scores, grad,
                                                                (HeadWorkspace*)ws_, loss_out, inv_n_f, inv_n, fp);
    else
        bce_kernel<T, false, FOCAL><<<grid, kHeadBlock, 0, st>>>(x, labels, n, nominal != 0, scores, grad,
                                                                 (HeadWorkspace*)ws_, loss_out, inv_n_f, inv_n, fp);
    return check_launch("bce_kernel");
}

// ------------------------------------------------------------------------------------------ CLIP heads
// Text rows live in shared memory as fp32 [K][ITERS*128] (zero padded), optionally re-normalised
// (clip.py:69 re-normalises `center` for the score, clip.py:82-86 does not for the loss).
template <int ITERS>
__device__ __forceinline__ void clip_stage_text(const float* __restrict__ text, int d, int K, bool renorm,
                                                float* s_text) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec = d >> 2;
    constexpr int DP = ITERS * 128;
    for (int k = warp; k < K; k += kHeadWarps) {
        float t[ITERS][4];
        float s = 0.f;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int vi = it * 32 + lane;
            if (vi < nvec) {
                float4 q = __ldg(reinterpret_cast<const float4*>(text + (int64_t)k * d) + vi);
                t[it][0] = q.x; t[it][1] = q.y; t[it][2] = q.z; t[it][3] = q.w;
            } else {
                t[it][0] = t[it][1] = t[it][2] = t[it][3] = 0.f;
            }
            s += t[it][0] * t[it][0] + t[it][1] * t[it][1] + t[it][2] * t[it][2] + t[it][3] * t[it][3];
        }
        float nrm = renorm ? sqrtf(warp_sum(s)) : 1.0f;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            float4 q = make_float4(t[it][0] / nrm, t[it][1] / nrm, t[it][2] / nrm, t[it][3] / nrm);
            reinterpret_cast<float4*>(s_text + k * DP)[it * 32 + lane] = q;
        }
    }
    __syncthreads();
}

template <typename T, int ITERS, int ROWS>
__global__ void __launch_bounds__(kHeadBlock)
clip_score_kernel(const T* __restrict__ z, const float* __restrict__ text, int64_t n, int d, int K, float scale,
                  float* __restrict__ scores) {
    extern __shared__ __align__(16) float s_text[];
    constexpr int DP = ITERS * 128;
    clip_stage_text<ITERS>(text, d, K, true, s_text);
    const int lane = threadIdx.x & 31;
    const int nvec = d >> 2;
    const int64_t warp0 = (int64_t)blockIdx.x * kHeadWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kHeadWarps;
    for (int64_t base = warp0 * ROWS; base < n; base += nwarps * ROWS) {
        float v[ROWS][ITERS][4];
        float inv_nrm[ROWS], mx[ROWS], se[ROWS], last[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int64_t row = base + r;
            float s = 0.f;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const int vi = it * 32 + lane;
                if (row < n && vi < nvec) load4_stream<T>(z + row * d + vi * 4, v[r][it]);
                else v[r][it][0] = v[r][it][1] = v[r][it][2] = v[r][it][3] = 0.f;
                s += v[r][it][0] * v[r][it][0] + v[r][it][1] * v[r][it][1] + v[r][it][2] * v[r][it][2] +
                     v[r][it][3] * v[r][it][3];
            }
            inv_nrm[r] = scale / sqrtf(warp_sum(s));
            mx[r] = -INFINITY; se[r] = 0.f; last[r] = 0.f;
        }
        for (int k = 0; k < K; ++k) {
            float dot[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) dot[r] = 0.f;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const float4 t = reinterpret_cast<const float4*>(s_text + k * DP)[it * 32 + lane];
#pragma unroll
                for (int r = 0; r < ROWS; ++r)
                    dot[r] += v[r][it][0] * t.x + v[r][it][1] * t.y + v[r][it][2] * t.z + v[r][it][3] * t.w;
            }
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const float l = warp_sum(dot[r]) * inv_nrm[r];       // logit_k = scale * z^ . T^_k
                if (l > mx[r]) { se[r] = se[r] * expf(mx[r] - l) + 1.0f; mx[r] = l; }
                else se[r] += expf(l - mx[r]);                         // NaN falls through here and sticks
                last[r] = l;
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < ROWS; ++r)
                if (base + r < n) scores[base + r] = expf(last[r] - mx[r]) / se[r];
        }
    }
}

// clip.py:81-103 + backward.  One warp per row; logit k lives in lane k%32, slot k/32 (K <= 64).
template <typename T, int ITERS>
__global__ void __launch_bounds__(kHeadBlock)
clip_oe_loss_kernel(const T* __restrict__ z, const float* __restrict__ text, const int64_t* __restrict__ labels,
                    int64_t n, int d, int K, float scale, int64_t nominal_label, int loo, T* __restrict__ grad,
                    HeadWorkspace* ws, float* loss_out, float inv_n_f, double inv_n) {
    extern __shared__ __align__(16) float s_text[];
    constexpr int DP = ITERS * 128;
    clip_stage_text<ITERS>(text, d, K, false, s_text);
    const int lane = threadIdx.x & 31;
    const int nvec = d >> 2;
    const int64_t warp0 = (int64_t)blockIdx.x * kHeadWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kHeadWarps;
    const int64_t anom_label = 1 - nominal_label;
    float loss_acc = 0.f;
    for (int64_t row = warp0; row < n; row += nwarps) {
        float v[ITERS][4];
        float s = 0.f;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int vi = it * 32 + lane;
            if (vi < nvec) load4_stream<T>(z + row * d + vi * 4, v[it]);
            else v[it][0] = v[it][1] = v[it][2] = v[it][3] = 0.f;
            s += v[it][0] * v[it][0] + v[it][1] * v[it][1] + v[it][2] * v[it][2] + v[it][3] * v[it][3];
        }
        const float inv_nrm = 1.0f / sqrtf(warp_sum(s));
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {           // v <- z^
            v[it][0] *= inv_nrm; v[it][1] *= inv_nrm; v[it][2] *= inv_nrm; v[it][3] *= inv_nrm;
        }
        float lg[2] = {-INFINITY, -INFINITY};
        for (int k = 0; k < K; ++k) {
            float dot = 0.f;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const float4 t = reinterpret_cast<const float4*>(s_text + k * DP)[it * 32 + lane];
                dot += v[it][0] * t.x + v[it][1] * t.y + v[it][2] * t.z + v[it][3] * t.w;
            }
            const float l = scale * warp_sum(dot);
            if ((k & 31) == lane) lg[k >> 5] = l;
        }
        const float m = warp_max(fmaxf(lg[0], lg[1]));
        float p0 = (lane < K) ? expf(lg[0] - m) : 0.f;
        float p1 = (lane + 32 < K) ? expf(lg[1] - m) : 0.f;
        const float sum = warp_sum(p0 + p1);
        const float lse = m + logf(sum);
        const int64_t lab = labels[row];
        int t = -1;
        if (lab == anom_label) t = K - 1;
        else if (lab == nominal_label) {
            t = 0;
            if (loo) {          // argmax over k < K-1, first maximal index (clip.py:95)
                float a0 = (lane < K - 1) ? lg[0] : -INFINITY;
                float a1 = (lane + 32 < K - 1) ? lg[1] : -INFINITY;
                const float am = warp_max(fmaxf(a0, a1));
                unsigned b0 = __ballot_sync(kFullMask, a0 == am), b1 = __ballot_sync(kFullMask, a1 == am);
                t = b0 ? (__ffs(b0) - 1) : (b1 ? 32 + __ffs(b1) - 1 : 0);
            }
        }
        float lt = __shfl_sync(kFullMask, (t >= 32) ? lg[1] : lg[0], t < 0 ? 0 : (t & 31));
        if (t >= 0 && lane == 0) loss_acc += lse - lt;
        if (grad) {
            // G_k = (softmax_k - [k==t]) / n ; g = scale * sum_k G_k c_k ; dz = (g - (g.z^) z^) / ||z||
            float G0 = (t >= 0 && lane < K) ? (p0 / sum - (lane == t ? 1.f : 0.f)) * inv_n_f : 0.f;
            float G1 = (t >= 0 && lane + 32 < K) ? (p1 / sum - (lane + 32 == t ? 1.f : 0.f)) * inv_n_f : 0.f;
            float g[ITERS][4];
#pragma unroll
            for (int it = 0; it < ITERS; ++it) g[it][0] = g[it][1] = g[it][2] = g[it][3] = 0.f;
            for (int k = 0; k < K; ++k) {
                const float Gk = scale * __shfl_sync(kFullMask, (k >= 32) ? G1 : G0, k & 31);
#pragma unroll
                for (int it = 0; it < ITERS; ++it) {
                    const float4 c = reinterpret_cast<const float4*>(s_text + k * DP)[it * 32 + lane];
                    g[it][0] += Gk * c.x; g[it][1] += Gk * c.y; g[it][2] += Gk * c.z; g[it][3] += Gk * c.w;
                }
            }
            float gd = 0.f;
#pragma unroll
            for (int it = 0; it < ITERS; ++it)
                gd += g[it][0] * v[it][0] + g[it][1] * v[it][1] + g[it][2] * v[it][2] + g[it][3] * v[it][3];
            gd = warp_sum(gd);
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const int vi = it * 32 + lane;
                if (vi < nvec) {
                    float o[4] = {(g[it][0] - gd * v[it][0]) * inv_nrm, (g[it][1] - gd * v[it][1]) * inv_nrm,
                                  (g[it][2] - gd * v[it][2]) * inv_nrm, (g[it][3] - gd * v[it][3]) * inv_nrm};
                    store4_stream<T>(grad + row * d + vi * 4, o);
                }
            }
        }
    }
    grid_mean_finish<kHeadBlock>(loss_acc, ws, loss_out, inv_n);
}

template <typename T>
static int clip_score_launch(const void* z, const float* text, int64_t n, int64_t d, int64_t K, float scale,
                             float* scores, cudaStream_t st) {
    const int iters = (int)((d / 4 + 31) / 32);
    const int it_pad = iters <= 2 ? 2 : (iters <= 4 ? 4 : 8);
    const size_t smem = (size_t)K * it_pad * 128 * sizeof(float);
#define EOE_CLIP_CASE(IT, RW)                                                                            \
    {                                                                                                    \
        auto kern = clip_score_kernel<T, IT, RW>;                                                        \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) { set_cuda_error(e, "clip_score smem attr"); return EOE_ERR_CUDA; }        \
        int grid = head_grid(n, kHeadWarps * RW);                                                        \
        if (grid > kNumSMs * 2) grid = kNumSMs * 2;  /* text staging is per block: keep blocks few */    \
        kern<<<grid, kHeadBlock, smem, st>>>((const T*)z, text, n, (int)d, (int)K, scale, scores);       \
    }
    if (it_pad == 2) EOE_CLIP_CASE(2, 4)
    else if (it_pad == 4) EOE_CLIP_CASE(4, 4)
    else EOE_CLIP_CASE(8, 2)
#undef EOE_CLIP_CASE
    return check_launch("clip_score_kernel");
}

template <typename T>
static int clip_loss_launch(const void* z, const float* text, const int64_t* labels, int64_t n, int64_t d,
                            int64_t K, float scale, int64_t nominal, int loo, float* loss_out, void* grad,
                            void* ws, cudaStream_t st) {
    const int iters = (int)((d / 4 + 31) / 32);
    const int it_pad = iters <= 2 ? 2 : (iters <= 4 ? 4 : 8);
    const size_t smem = (size_t)K * it_pad * 128 * sizeof(float);
    const float inv_n_f = 1.0f / (float)n;
    const double inv_n = 1.0 / (double)n;
#define EOE_CLIPL_CASE(IT)                                                                               \
    {                                                                                                    \
        auto kern = clip_oe_loss_kernel<T, IT>;                                                          \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) { set_cuda_error(e, "clip_oe_loss smem attr"); return EOE_ERR_CUDA; }      \
        int grid = head_grid(n, kHeadWarps);                                                             \
        if (grid > kNumSMs * 2) grid = kNumSMs * 2;                                                      \
        kern<<<grid, kHeadBlock, smem, st>>>((const T*)z, text, labels, n, (int)d, (int)K, scale, nominal, loo, \
                                             (T*)grad, (HeadWorkspace*)ws, loss_out, inv_n_f, inv_n);    \
    }
    if (it_pad == 2) EOE_CLIPL_CASE(2)
    else if (it_pad == 4) EOE_CLIPL_CASE(4)
    else EOE_CLIPL_CASE(8)
#undef EOE_CLIPL_CASE
    return check_launch("clip_oe_loss_kernel");
}

// called by the fused encoder tail as well
int clip_score_f32(const float* z, const float* text, int64_t n, int64_t d, int64_t K, float scale, float* scores,
                   cudaStream_t st) {
    return clip_score_launch<float>(z, text, n, d, K, scale, scores, st);
}

}  // namespace eoe

using namespace eoe;

#define EOE_DISPATCH_DTYPE(dt, CALL)                       \
    switch (dt) {                                          \
        case EOE_F32: { using T = float; return CALL; }    \
        case EOE_F16: { using T = __half; return CALL; }   \
        case EOE_BF16: { using T = __nv_bfloat16; return CALL; } \
        default: return EOE_ERR_DTYPE;                     \
    }

extern "C" int eoe_hsc_fwd_bwd(const void* z, int z_dtype, const int64_t* labels, int64_t n, int64_t d,
                               int64_t nominal_label, float* loss_out, float* scores_out, void* grad_z_out,
                               void* head_ws, void* stream) {
    if (!z || !labels || !loss_out || !head_ws || n <= 0 || d <= 0) return EOE_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    EOE_DISPATCH_DTYPE(z_dtype, (hsc_launch<T>(z, labels, n, d, nominal_label, loss_out, scores_out, grad_z_out, head_ws, st)))
}

extern "C" int eoe_hsc_score(const void* z, int z_dtype, int64_t n, int64_t d, float* scores_out, void* stream) {
    if (!z || !scores_out || n <= 0 || d <= 0) return EOE_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    EOE_DISPATCH_DTYPE(z_dtype, (hsc_launch<T>(z, nullptr, n, d, 0, nullptr, scores_out, nullptr, nullptr, st)))
}

extern "C" int eoe_bce_fwd_bwd(const void* x, int x_dtype, const int64_t* labels, int64_t n, int64_t nominal_label,
                               float* loss_out, float* scores_out, void* grad_x_out, void* head_ws, void* stream) {
    if (!x || !labels || !loss_out || !head_ws || n <= 0) return EOE_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    EOE_DISPATCH_DTYPE(x_dtype, (bce_launch<T>(x, labels, n, nominal_label, loss_out, scores_out, grad_x_out, head_ws, st)))
}

extern "C" int eoe_bce_score(const void* x, int x_dtype, int64_t n, int64_t nominal_label, float* scores_out,
                             void* stream) {
    if (!x || !scores_out || n <= 0) return EOE_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    EOE_DISPATCH_DTYPE(x_dtype, (bce_launch<T>(x, nullptr, n, nominal_label, nullptr, scores_out, nullptr, nullptr, st)))
}

extern "C" int eoe_dsad_fwd_bwd(const void* z, int z_dtype, const int64_t* labels, int64_t n, int64_t d,
                                int64_t nominal_label, float* loss_out, float* scores_out, void* grad_z_out,
                                void* head_ws, void* stream) {
    if (!z || !labels || !loss_out || !head_ws || n <= 0 || d <= 0) return EOE_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    EOE_DISPATCH_DTYPE(z_dtype, (hsc_launch<T, ROW_DSAD>(z, labels, n, d, nominal_label, loss_out, scores_out, grad_z_out, head_ws, st)))
}

extern "C" int eoe_dsvdd_fwd_bwd(const void* z, int z_dtype, const float* center, int64_t n, int64_t d,
                                 float* loss_out, float* scores_out, void* grad_z_out, void* head_ws, void* stream) {
    if (!z || !center || n <= 0 || d <= 0) return EOE_ERR_ARG;
    if (loss_out && !head_ws) return EOE_ERR_ARG;
    if (!loss_out && !scores_out && !grad_z_out) return EOE_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    EOE_DISPATCH_DTYPE(z_dtype, (hsc_launch<T, ROW_DSVDD>(z, nullptr, n, d, 0, loss_out, scores_out, grad_z_out, head_ws, st, center)))
}

extern "C" int eoe_focal_fwd_bwd(const void* x, int x_dtype, const int64_t* labels, int64_t n, int64_t nominal_label,
                                 float gamma, float eps, float* loss_out, float* scores_out, void* grad_x_out,
                                 void* head_ws, void* stream) {
    if (!x || !labels || !loss_out || !head_ws || n <= 0) return EOE_ERR_ARG;
    if (!(gamma >= 0.f) || !(eps >= 0.f && eps < 0.5f)) return EOE_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const FocalParams fp{gamma, eps};
    EOE_DISPATCH_DTYPE(x_dtype, (bce_launch<T, true>(x, labels, n, nominal_label, loss_out, scores_out, grad_x_out, head_ws, st, fp)))
}

static int clip_check(const void* z, const float* text, int64_t n, int64_t d, int64_t K) {
    if (!z || !text || n <= 0 || d <= 0 || K <= 0) return EOE_ERR_ARG;
    if (d % 4 != 0 || d > 1024 || K > 64) return EOE_ERR_SHAPE;
    const int iters = (int)((d / 4 + 31) / 32);
    const int it_pad = iters <= 2 ? 2 : (iters <= 4 ? 4 : 8);
    if ((size_t)K * it_pad * 128 * 4 > 192 * 1024) return EOE_ERR_SHAPE;
    if ((uintptr_t)z % 16 != 0 || (uintptr_t)text % 16 != 0) return EOE_ERR_ALIGN;
    return EOE_OK;
}

extern "C" int eoe_clip_score(const void* z, int z_dtype, const float* text, int64_t n, int64_t d, int64_t K,
                              float scale, float* scores_out, void* stream) {
    int rc = clip_check(z, text, n, d, K);
    if (rc) return rc;
    if (!scores_out) return EOE_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    EOE_DISPATCH_DTYPE(z_dtype, (clip_score_launch<T>(z, text, n, d, K, scale, scores_out, st)))
}

extern "C" int eoe_clip_oe_loss_fwd_bwd(const void* z, int z_dtype, const float* text, const int64_t* labels,
                                        int64_t n, int64_t d, int64_t K, float scale, int64_t nominal_label,
                                        int leave_one_out, float* loss_out, void* grad_z_out, void* head_ws,
                                        void* stream) {
    int rc = clip_check(z, text, n, d, K);
    if (rc) return rc;
    if (!labels || !loss_out || !head_ws) return EOE_ERR_ARG;
    if (grad_z_out && (uintptr_t)grad_z_out % 16 != 0) return EOE_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    EOE_DISPATCH_DTYPE(z_dtype, (clip_loss_launch<T>(z, text, labels, n, d, K, scale, nominal_label, leave_one_out, loss_out, grad_z_out, head_ws, st)))
}
