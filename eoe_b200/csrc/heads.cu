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

// GRAD = false (score only, grad == nullptr): nothing outlives the reduction pass, so the rows need not stay in registers.
template <typename T, int VEC, int ITERS, int ROWS, int MODE, bool GRAD = true>
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
            if (GRAD && grad) {
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
    if (grad) hsc_rows_kernel<T, VEC, IT, RW, MODE, true><<<head_grid(n, kHeadWarps * RW), kHeadBlock, 0, st>>>(     \
        z, labels, n, (int)d, nominal, scores, grad, ws, loss_out, inv_n_f, inv_n, center);            \
    else hsc_rows_kernel<T, VEC, IT, RW, MODE, false><<<head_grid(n, kHeadWarps * RW), kHeadBlock, 0, st>>>(         \
        z, labels, n, (int)d, nominal, scores, grad, ws, loss_out, inv_n_f, inv_n, center)
    if (iters <= 1) { EOE_HSC_CASE(1, 4); }
    else if (iters <= 2) { EOE_HSC_CASE(2, VEC == 4 ? 4 : 2); }
    else if (iters <= 4) { EOE_HSC_CASE(4, VEC == 4 ? 2 : 1); }
    else { EOE_HSC_CASE(8, 1); }
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

// Large prompt sets (leave-one-out on cifar100 / cub / dtd builds 100 / 200 / 47 prompts, clip.py:53-54) do not fit the
// shared-memory text tile: GTEXT reads the text rows through the read-only path from L2 / L1 instead (K * d * 4 <= 1 MB)
// and keeps only 1 / ||t_k|| (score head: clip.py:69 re-normalises) per prompt in shared memory.
template <bool RENORM>
__device__ __forceinline__ void clip_stage_inv_norms(const float* __restrict__ text, int d, int K, float* s_inv) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = warp; k < K; k += kHeadWarps) {
        float s = 0.f;
        if (RENORM)
            for (int i = lane; i < d; i += 32) { const float v = __ldg(text + (int64_t)k * d + i); s += v * v; }
        s = warp_sum(s);
        if (lane == 0) s_inv[k] = RENORM ? 1.0f / sqrtf(s) : 1.0f;
    }
    __syncthreads();
}

template <typename T, int ITERS, int ROWS, bool GTEXT = false>
__global__ void __launch_bounds__(kHeadBlock)
clip_score_kernel(const T* __restrict__ z, const float* __restrict__ text, int64_t n, int d, int K, float scale,
                  float* __restrict__ scores) {
    extern __shared__ __align__(16) float s_text[];
    constexpr int DP = ITERS * 128;
    if (GTEXT) clip_stage_inv_norms<true>(text, d, K, s_text);
    else clip_stage_text<ITERS>(text, d, K, true, s_text);
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
                float4 t;
                if (GTEXT) {
                    const int vi = it * 32 + lane;
                    t = (vi < nvec) ? __ldg(reinterpret_cast<const float4*>(text + (int64_t)k * d) + vi) : make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
                    t = reinterpret_cast<const float4*>(s_text + k * DP)[it * 32 + lane];
                }
#pragma unroll
                for (int r = 0; r < ROWS; ++r)
                    dot[r] += v[r][it][0] * t.x + v[r][it][1] * t.y + v[r][it][2] * t.z + v[r][it][3] * t.w;
            }
            const float tk = GTEXT ? s_text[k] : 1.0f;               // 1 / ||t_k|| when the rows are not pre-normalised
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const float l = warp_sum(dot[r]) * inv_nrm[r] * tk;  // logit_k = scale * z^ . T^_k
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

// clip.py:81-103 + backward.  One warp per row; logit k lives in lane k%32, slot k/32 (K <= 32 * SLOTS).
// GTEXT: text rows come from global memory (L2 / L1) instead of the shared-memory tile (large prompt sets).
template <typename T, int ITERS, int SLOTS = 2, bool GTEXT = false>
__global__ void __launch_bounds__(kHeadBlock)
clip_oe_loss_kernel(const T* __restrict__ z, const float* __restrict__ text, const int64_t* __restrict__ labels,
                    int64_t n, int d, int K, float scale, int64_t nominal_label, int loo, T* __restrict__ grad,
                    HeadWorkspace* ws, float* loss_out, float inv_n_f, double inv_n) {
    extern __shared__ __align__(16) float s_text[];
    constexpr int DP = ITERS * 128;
    if (!GTEXT) clip_stage_text<ITERS>(text, d, K, false, s_text);
    const int lane = threadIdx.x & 31;
    const int nvec = d >> 2;
    const int64_t warp0 = (int64_t)blockIdx.x * kHeadWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kHeadWarps;
    const int64_t anom_label = 1 - nominal_label;
    auto text4 = [&](int k, int it) {
        if (GTEXT) {
            const int vi = it * 32 + lane;
            return (vi < nvec) ? __ldg(reinterpret_cast<const float4*>(text + (int64_t)k * d) + vi) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return reinterpret_cast<const float4*>(s_text + k * DP)[it * 32 + lane];
    };
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
        float lg[SLOTS];
#pragma unroll
        for (int q = 0; q < SLOTS; ++q) lg[q] = -INFINITY;
        for (int k = 0; k < K; ++k) {
            float dot = 0.f;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const float4 t = text4(k, it);
                dot += v[it][0] * t.x + v[it][1] * t.y + v[it][2] * t.z + v[it][3] * t.w;
            }
            const float l = scale * warp_sum(dot);
#pragma unroll
            for (int q = 0; q < SLOTS; ++q)
                if ((k >> 5) == q && (k & 31) == lane) lg[q] = l;
        }
        float m = lg[0];
#pragma unroll
        for (int q = 1; q < SLOTS; ++q) m = fmaxf(m, lg[q]);
        m = warp_max(m);
        float p[SLOTS], psum = 0.f;
#pragma unroll
        for (int q = 0; q < SLOTS; ++q) {
            p[q] = (lane + 32 * q < K) ? expf(lg[q] - m) : 0.f;
            psum += p[q];
        }
        const float sum = warp_sum(psum);
        const float lse = m + logf(sum);
        const int64_t lab = labels[row];
        int t = -1;
        if (lab == anom_label) t = K - 1;
        else if (lab == nominal_label) {
            t = 0;
            if (loo) {          // argmax over k < K-1, first maximal index (clip.py:95)
                float a[SLOTS], am = -INFINITY;
#pragma unroll
                for (int q = 0; q < SLOTS; ++q) {
                    a[q] = (lane + 32 * q < K - 1) ? lg[q] : -INFINITY;
                    am = fmaxf(am, a[q]);
                }
                am = warp_max(am);
                bool found = false;
#pragma unroll
                for (int q = 0; q < SLOTS; ++q) {
                    const unsigned bq = __ballot_sync(kFullMask, a[q] == am);
                    if (!found && bq) { t = 32 * q + __ffs(bq) - 1; found = true; }
                }
            }
        }
        float lt_src = lg[0];
#pragma unroll
        for (int q = 1; q < SLOTS; ++q)
            if ((t >> 5) == q) lt_src = lg[q];
        const float lt = __shfl_sync(kFullMask, lt_src, t < 0 ? 0 : (t & 31));
        if (t >= 0 && lane == 0) loss_acc += lse - lt;
        if (grad) {
            // G_k = (softmax_k - [k==t]) / n ; g = scale * sum_k G_k c_k ; dz = (g - (g.z^) z^) / ||z||
            float G[SLOTS];
#pragma unroll
            for (int q = 0; q < SLOTS; ++q)
                G[q] = (t >= 0 && lane + 32 * q < K) ? (p[q] / sum - (lane + 32 * q == t ? 1.f : 0.f)) * inv_n_f : 0.f;
            float g[ITERS][4];
#pragma unroll
            for (int it = 0; it < ITERS; ++it) g[it][0] = g[it][1] = g[it][2] = g[it][3] = 0.f;
            for (int k = 0; k < K; ++k) {
                float gsrc = G[0];
#pragma unroll
                for (int q = 1; q < SLOTS; ++q)
                    if ((k >> 5) == q) gsrc = G[q];
                const float Gk = scale * __shfl_sync(kFullMask, gsrc, k & 31);
#pragma unroll
                for (int it = 0; it < ITERS; ++it) {
                    const float4 c = text4(k, it);
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

// ------------------------------------------------------------------------------------------ CLIP score head on tensor cores
// logits = z [n, d] @ T^T [d, K] is a dense contraction: at K = 10 / 30 the FP32-FMA kernel above needs 5 / 15 kFMA per row
// and runs at 28 % / 12 % of HBM bandwidth.  Here a warp owns 16 rows and feeds them to mma.sync.m16n8k16 straight from
// its coalesced 16-byte global loads -- no shared-memory staging of z at all: the k dimension of a dot product may be
// permuted freely, so the four consecutive features a lane loads ARE its four k-slots (2t, 2t+1, 2t+8, 2t+9) of one
// k-step, and the text rows are laid out once per CTA in the matching fragment order (one conflict-free LDS.128 per
// (k-step, 8-prompt tile) yields the hi and lo B fragments).
// Precision: products must carry ~1e-6 (logits are 100 * cos): fp32 features are split z = hi + lo (two bf16), unit text
// rows t = hi + lo likewise, and acc += hi*hi + lo*hi + hi*lo in fp32 (the dropped lo*lo term is 2^-18 relative);
// 16-bit features are exact operands and only the text is split.  Row norms are fp32 sums of the original values.
template <typename T>
struct ClipMma {
    static constexpr bool kBF16 = !std::is_same<T, __half>::value;          // operand type: bf16 for fp32 / bf16 features
    static constexpr bool kSplitA = std::is_same<T, float>::value;
};

template <bool BF16>
__device__ __forceinline__ uint32_t clip_pack2(float a, float b) {
    if (BF16) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    } else {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
}
template <bool BF16>
__device__ __forceinline__ float2 clip_unpack2(uint32_t w) {
    if (BF16) return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
    return __half22float2(*reinterpret_cast<__half2*>(&w));
}
template <bool BF16>
__device__ __forceinline__ void clip_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if (BF16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// feature column of k-step j, lane quad position t, slot s (0..3): what the A-side loads below put into that k-slot
template <typename T>
__device__ __forceinline__ int clip_frag_col(int j, int t) {
    if (sizeof(T) == 4) return 16 * j + 4 * t;                        // one 16-byte load = 4 fp32 = one k-step
    return 32 * (j >> 1) + 8 * t + 4 * (j & 1);                      // one 16-byte load = 8 x 16 bit = two k-steps
}

// s_frag [d/16][NT][32] uint4 = {hi(s0,s1), hi(s2,s3), lo(s0,s1), lo(s2,s3)} of prompt nt*8 + lane/4 (zero rows past K)
template <typename T, int NT>
__device__ __forceinline__ void clip_stage_fragments(const float* __restrict__ text, int d, int K, bool renorm,
                                                     uint4* s_frag, float* s_inv) {
    constexpr bool BF = ClipMma<T>::kBF16;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = warp; k < NT * 8; k += (int)(blockDim.x >> 5)) {
        float s = 0.f;
        if (k < K && renorm)
            for (int i = lane; i < d; i += 32) { const float v = __ldg(text + (int64_t)k * d + i); s += v * v; }
        s = warp_sum(s);
        if (lane == 0) s_inv[k] = (k < K) ? (renorm ? 1.0f / sqrtf(s) : 1.0f) : 0.f;
    }
    __syncthreads();
    const int total = (d >> 4) * NT * 32;
    for (int i = threadIdx.x; i < total; i += (int)blockDim.x) {
        const int ln = i & 31, nt = (i >> 5) % NT, j = (i >> 5) / NT;
        const int k = nt * 8 + (ln >> 2);
        uint4 f = make_uint4(0u, 0u, 0u, 0u);
        if (k < K) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(text + (int64_t)k * d + clip_frag_col<T>(j, ln & 3)));
            const float inv = s_inv[k];
            const float e[4] = {v.x * inv, v.y * inv, v.z * inv, v.w * inv};
            f.x = clip_pack2<BF>(e[0], e[1]);
            f.y = clip_pack2<BF>(e[2], e[3]);
            const float2 h0 = clip_unpack2<BF>(f.x), h1 = clip_unpack2<BF>(f.y);
            f.z = clip_pack2<BF>(e[0] - h0.x, e[1] - h0.y);
            f.w = clip_pack2<BF>(e[2] - h1.x, e[3] - h1.y);
        }
        s_frag[i] = f;
    }
    __syncthreads();
}

// acc[nt][2r + e] += z[row r] . T[prompt nt*8 + 2t + e] over all d features (this lane's quarter of each row is summed by
// the tensor core), ss[r] = this lane's share of sum z^2.  Rows g and g + 8 of the warp's 16-row tile.
// KEEP: the rows are read again by the caller (loss backward): ask L2 to keep them (evict_last policy)
template <typename T, int NT, bool KEEP = false>
__device__ __forceinline__ void clip_forward_tile(const T* __restrict__ z, const int64_t (&row)[2], const bool (&ok)[2], int d,
                                                  const uint4* s_frag, int lane, float (&acc)[NT][4], float (&ss)[2],
                                                  uint64_t keep_policy = 0) {
    constexpr bool BF = ClipMma<T>::kBF16;
    constexpr bool SPLIT = ClipMma<T>::kSplitA;
    const int t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    ss[0] = ss[1] = 0.f;
    for (int c0 = 0; c0 < d; c0 += 128) {                       // 8 k-steps per batch: 8 KB (fp32) of loads in flight per warp
        constexpr int LOADS = sizeof(T) == 4 ? 8 : 4;           // 16-byte loads per row per batch
        uint4 raw[2][LOADS];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < LOADS; ++q) {
                const int col = c0 + (sizeof(T) == 4 ? 16 * q + 4 * t : 32 * q + 8 * t);
                if (!ok[r]) raw[r][q] = make_uint4(0u, 0u, 0u, 0u);
                else if (KEEP)
                    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                                 : "=r"(raw[r][q].x), "=r"(raw[r][q].y), "=r"(raw[r][q].z), "=r"(raw[r][q].w)
                                 : "l"(z + row[r] * d + col), "l"(keep_policy));
                else
                    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(raw[r][q].x), "=r"(raw[r][q].y), "=r"(raw[r][q].z), "=r"(raw[r][q].w)
                                 : "l"(z + row[r] * d + col));
            }
#pragma unroll
        for (int js = 0; js < 8; ++js) {
            uint32_t a_hi[4], a_lo[4];
            if (SPLIT) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const uint4 w = raw[r][js];
                    const float e[4] = {__uint_as_float(w.x), __uint_as_float(w.y), __uint_as_float(w.z), __uint_as_float(w.w)};
                    ss[r] += (e[0] * e[0] + e[1] * e[1]) + (e[2] * e[2] + e[3] * e[3]);
                    a_hi[r] = clip_pack2<true>(e[0], e[1]);
                    a_hi[2 + r] = clip_pack2<true>(e[2], e[3]);
                    const float2 h0 = clip_unpack2<true>(a_hi[r]), h1 = clip_unpack2<true>(a_hi[2 + r]);
                    a_lo[r] = clip_pack2<true>(e[0] - h0.x, e[1] - h0.y);
                    a_lo[2 + r] = clip_pack2<true>(e[2] - h1.x, e[3] - h1.y);
                }
            } else {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const uint4 w = raw[r][js >> 1];
                    a_hi[r] = (js & 1) ? w.z : w.x;
                    a_hi[2 + r] = (js & 1) ? w.w : w.y;
                    const float2 f0 = clip_unpack2<BF>(a_hi[r]), f1 = clip_unpack2<BF>(a_hi[2 + r]);
                    ss[r] += (f0.x * f0.x + f0.y * f0.y) + (f1.x * f1.x + f1.y * f1.y);
                }
            }
            const uint4* fr = s_frag + ((size_t)((c0 >> 4) + js) * NT) * 32 + lane;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const uint4 f = fr[nt * 32];
                clip_mma<BF>(acc[nt], a_hi, f.z, f.w);          // small terms first
                if (SPLIT) clip_mma<BF>(acc[nt], a_lo, f.x, f.y);
                clip_mma<BF>(acc[nt], a_hi, f.x, f.y);
            }
        }
    }
}

// The same tile, software pipelined: the 16-byte loads of the NEXT 64 feature columns are issued before the MMAs of the
// current 64 (two register buffers, each holding a 64-column slab of both rows), so a warp's loads and tensor work overlap
// instead of alternating -- with 16 resident warps per SM the non-pipelined form leaves both the tensor pipe and HBM at
// ~45 % (ncu, profiles/r1_v10_ncu_full_heads.md).  Same k order per accumulator: bit-identical results.  Used by the score
// kernel (+2 ... +4 % of HBM bandwidth); in the loss kernel the second buffer spills (128 registers at 512 threads) and it
// measured 1-2 % slower, so the loss keeps the form above.
template <typename T, int NT, bool KEEP = false>
__device__ __forceinline__ void clip_forward_tile_pipe(const T* __restrict__ z, const int64_t (&row)[2], const bool (&ok)[2], int d,
                                                       const uint4* s_frag, int lane, float (&acc)[NT][4], float (&ss)[2],
                                                       uint64_t keep_policy = 0) {
    constexpr bool BF = ClipMma<T>::kBF16;
    constexpr bool SPLIT = ClipMma<T>::kSplitA;
    constexpr int HL = sizeof(T) == 4 ? 4 : 2;                    // 16-byte loads per row per 64-column slab
    const int t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    ss[0] = ss[1] = 0.f;
    uint4 raw[2][2][HL];
    auto issue = [&](uint4 (&buf)[2][HL], int c0) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < HL; ++q) {
                const int col = c0 + (sizeof(T) == 4 ? 16 * q + 4 * t : 32 * q + 8 * t);
                if (!ok[r]) buf[r][q] = make_uint4(0u, 0u, 0u, 0u);
                else if (KEEP)
                    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                                 : "=r"(buf[r][q].x), "=r"(buf[r][q].y), "=r"(buf[r][q].z), "=r"(buf[r][q].w)
                                 : "l"(z + row[r] * d + col), "l"(keep_policy));
                else
                    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(buf[r][q].x), "=r"(buf[r][q].y), "=r"(buf[r][q].z), "=r"(buf[r][q].w)
                                 : "l"(z + row[r] * d + col));
            }
    };
    auto compute = [&](const uint4 (&buf)[2][HL], int c0) {
#pragma unroll
        for (int js = 0; js < 4; ++js) {                          // 4 k-steps of 16 columns per slab
            uint32_t a_hi[4], a_lo[4];
            if (SPLIT) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const uint4 w = buf[r][js];
                    const float e[4] = {__uint_as_float(w.x), __uint_as_float(w.y), __uint_as_float(w.z), __uint_as_float(w.w)};
                    ss[r] += (e[0] * e[0] + e[1] * e[1]) + (e[2] * e[2] + e[3] * e[3]);
                    a_hi[r] = clip_pack2<true>(e[0], e[1]);
                    a_hi[2 + r] = clip_pack2<true>(e[2], e[3]);
                    const float2 h0 = clip_unpack2<true>(a_hi[r]), h1 = clip_unpack2<true>(a_hi[2 + r]);
                    a_lo[r] = clip_pack2<true>(e[0] - h0.x, e[1] - h0.y);
                    a_lo[2 + r] = clip_pack2<true>(e[2] - h1.x, e[3] - h1.y);
                }
            } else {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const uint4 w = buf[r][js >> 1];
                    a_hi[r] = (js & 1) ? w.z : w.x;
                    a_hi[2 + r] = (js & 1) ? w.w : w.y;
                    const float2 f0 = clip_unpack2<BF>(a_hi[r]), f1 = clip_unpack2<BF>(a_hi[2 + r]);
                    ss[r] += (f0.x * f0.x + f0.y * f0.y) + (f1.x * f1.x + f1.y * f1.y);
                }
            }
            const uint4* fr = s_frag + ((size_t)((c0 >> 4) + js) * NT) * 32 + lane;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const uint4 f = fr[nt * 32];
                clip_mma<BF>(acc[nt], a_hi, f.z, f.w);          // small terms first
                if (SPLIT) clip_mma<BF>(acc[nt], a_lo, f.x, f.y);
                clip_mma<BF>(acc[nt], a_hi, f.x, f.y);
            }
        }
    };
    issue(raw[0], 0);
    for (int c0 = 0; c0 < d; c0 += 128) {                         // d % 128 == 0
        issue(raw[1], c0 + 64);
        compute(raw[0], c0);
        if (c0 + 128 < d) issue(raw[0], c0 + 128);
        compute(raw[1], c0 + 64);
    }
}

template <typename T, int NT>
__global__ void __launch_bounds__(kHeadBlock, 2)
clip_score_mma_kernel(const T* __restrict__ z, const float* __restrict__ text, int64_t n, int d, int K, float scale,
                      float* __restrict__ scores) {
    extern __shared__ __align__(16) uint8_t s_clip_raw[];
    uint4* s_frag = reinterpret_cast<uint4*>(s_clip_raw);
    float* s_inv = reinterpret_cast<float*>(s_clip_raw + (size_t)(d >> 4) * NT * 32 * sizeof(uint4));
    clip_stage_fragments<T, NT>(text, d, K, true, s_frag, s_inv);
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int64_t warp0 = (int64_t)blockIdx.x * kHeadWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kHeadWarps;
    const int64_t tiles = (n + 15) >> 4;
    const int k_last = K - 1;
    for (int64_t tile = warp0; tile < tiles; tile += nwarps) {
        const int64_t row[2] = {tile * 16 + g, tile * 16 + g + 8};
        const bool ok[2] = {row[0] < n, row[1] < n};
        float acc[NT][4];
        float ss[2];
        clip_forward_tile_pipe<T, NT>(z, row, ok, d, s_frag, lane, acc, ss);
        // rows g (accumulator slots 0, 1) and g + 8 (slots 2, 3): prompt nt*8 + 2t + e lives in this lane
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float s = ss[r];
            s += __shfl_xor_sync(kFullMask, s, 1);
            s += __shfl_xor_sync(kFullMask, s, 2);
            const float inv = scale / sqrtf(s);
            float mx = -INFINITY, last = 0.f;
            float l[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = nt * 8 + 2 * t + e;
                    l[nt][e] = (k < K) ? acc[nt][2 * r + e] * inv : -INFINITY;     // logit_k = scale * z^ . T^_k
                    mx = fmaxf(mx, l[nt][e]);
                    if (k == k_last) last = l[nt][e];
                }
            mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, 2));
            float se = 0.f;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) se += expf(l[nt][0] - mx) + expf(l[nt][1] - mx);   // NaN logits stick here
            se += __shfl_xor_sync(kFullMask, se, 1);
            se += __shfl_xor_sync(kFullMask, se, 2);
            last = __shfl_sync(kFullMask, last, (lane & ~3) | ((k_last & 7) >> 1));
            if (t == 0 && ok[r]) scores[row[r]] = expf(last - mx) / se;
        }
    }
}

// ---- CLIP OE loss + backward on tensor cores (clip.py:81-103).  Forward as above (text rows used as given).  Backward:
// g = scale * G @ C is the second small GEMM; the forward accumulator layout (row g / g+8, prompts nt*8 + 2t + e) IS the
// A-fragment layout of a k-step over prompts, so G goes from registers straight back into mma.sync.  The output tile's
// n index is mapped to feature columns such that every lane ends up with the same 16-byte run of columns it loads z in
// (clip_out_col), and g . z^ = sum_k G_k logit_k needs no pass over g:  dz = (g - (sum_k G_k l_k) z^) / ||z||.
constexpr int kClipLossBlock = 512;

template <typename T>
__device__ __forceinline__ int clip_out_col(int ct, int nn) {       // feature column of output tile ct, tile column nn
    if (sizeof(T) == 4) return 16 * (ct >> 1) + 4 * (nn >> 1) + 2 * (ct & 1) + (nn & 1);
    return 32 * (ct >> 2) + 8 * (nn >> 1) + 2 * (ct & 3) + (nn & 1);
}

// s_fragT [d/8][KS][32] uint4 = B fragments (k = prompt, n = feature) {hi b0, hi b1, lo b0, lo b1}
template <typename T, int KS>
__device__ __forceinline__ void clip_stage_fragments_bwd(const float* __restrict__ text, int d, int K, uint4* s_fragT) {
    constexpr bool BF = ClipMma<T>::kBF16;
    const int total = (d >> 3) * KS * 32;
    for (int i = threadIdx.x; i < total; i += (int)blockDim.x) {
        const int ln = i & 31, ks = (i >> 5) % KS, ct = (i >> 5) / KS;
        const int col = clip_out_col<T>(ct, ln >> 2);
        const int k0 = 16 * ks + 2 * (ln & 3);
        float e[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + (q & 1) + (q >> 1) * 8;
            e[q] = (k < K) ? __ldg(text + (int64_t)k * d + col) : 0.f;
        }
        uint4 f;
        f.x = clip_pack2<BF>(e[0], e[1]);
        f.y = clip_pack2<BF>(e[2], e[3]);
        const float2 h0 = clip_unpack2<BF>(f.x), h1 = clip_unpack2<BF>(f.y);
        f.z = clip_pack2<BF>(e[0] - h0.x, e[1] - h0.y);
        f.w = clip_pack2<BF>(e[2] - h1.x, e[3] - h1.y);
        s_fragT[i] = f;
    }
}

template <typename T, int NT>
__global__ void __launch_bounds__(kClipLossBlock, 1)
clip_oe_loss_mma_kernel(const T* __restrict__ z, const float* __restrict__ text, const int64_t* __restrict__ labels,
                        int64_t n, int d, int K, float scale, int64_t nominal_label, int loo, T* __restrict__ grad,
                        HeadWorkspace* ws, float* loss_out, float inv_n_f, double inv_n) {
    constexpr bool BF = ClipMma<T>::kBF16;
    constexpr int KS = (NT + 1) / 2;                                // k-steps of 16 prompts in the backward product
    constexpr int GROUP = sizeof(T) == 4 ? 2 : 4;                    // output tiles (8 columns each) per 16-byte run of a lane
    extern __shared__ __align__(16) uint8_t s_clip_raw[];
    uint4* s_frag = reinterpret_cast<uint4*>(s_clip_raw);
    uint4* s_fragT = s_frag + (size_t)(d >> 4) * NT * 32;
    float* s_inv = reinterpret_cast<float*>(s_fragT + (size_t)(grad ? (d >> 3) * KS * 32 : 0));
    clip_stage_fragments<T, NT>(text, d, K, false, s_frag, s_inv);
    if (grad) clip_stage_fragments_bwd<T, KS>(text, d, K, s_fragT);
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int wpb = kClipLossBlock / 32;
    const int64_t warp0 = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * wpb;
    const int64_t tiles = (n + 15) >> 4;
    const int64_t anom_label = 1 - nominal_label;
    float loss_acc = 0.f;
    // L2 policies: the forward pass asks L2 to keep the tile (it is read again ~10 us later by the backward pass), the
    // second read and the gradient stores are marked evict_first so that they do not push waiting tiles out
    uint64_t pol_keep, pol_stream;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
    for (int64_t tile = warp0; tile < tiles; tile += nwarps) {
        const int64_t row[2] = {tile * 16 + g, tile * 16 + g + 8};
        const bool ok[2] = {row[0] < n, row[1] < n};
        float acc[NT][4];
        float ss[2];
        if (grad) clip_forward_tile<T, NT, true>(z, row, ok, d, s_frag, lane, acc, ss, pol_keep);
        else clip_forward_tile<T, NT>(z, row, ok, d, s_frag, lane, acc, ss);
        uint32_t a_hi[KS][4], a_lo[KS][4];
        float gdz[2], osc[2];                                        // (g . z^) / ||z||  and  1 / (n ||z||) per row
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float s = ss[r];
            s += __shfl_xor_sync(kFullMask, s, 1);
            s += __shfl_xor_sync(kFullMask, s, 2);
            const float inv_nrm = 1.0f / sqrtf(s);
            const float sc = scale * inv_nrm;
            float l[NT][2];
            float mx = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    l[nt][e] = (nt * 8 + 2 * t + e < K) ? acc[nt][2 * r + e] * sc : -INFINITY;    // logit_k = scale * z^ . c_k
                    mx = fmaxf(mx, l[nt][e]);
                }
            mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, 2));
            float pr[NT][2];
            float sum = 0.f;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                pr[nt][0] = expf(l[nt][0] - mx);
                pr[nt][1] = expf(l[nt][1] - mx);
                sum += pr[nt][0] + pr[nt][1];
            }
            sum += __shfl_xor_sync(kFullMask, sum, 1);
            sum += __shfl_xor_sync(kFullMask, sum, 2);
            const float lse = mx + logf(sum);
            const int64_t lab = ok[r] ? labels[row[r]] : 0;
            // argmax over k < K-1, first maximal index (clip.py:95) -- evaluated by every lane: the shuffles below must
            // not sit inside the per-row label branch (rows of one warp carry different labels)
            int loo_t = 0;
            if (loo) {
                float best = -INFINITY;
                int bi = 0x7fffffff;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int k = nt * 8 + 2 * t + e;
                        if (k < K - 1 && (l[nt][e] > best || bi == 0x7fffffff)) { best = l[nt][e]; bi = k; }
                    }
#pragma unroll
                for (int o = 1; o <= 2; o <<= 1) {
                    const float ob = __shfl_xor_sync(kFullMask, best, o);
                    const int oi = __shfl_xor_sync(kFullMask, bi, o);
                    if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi))) { best = ob; bi = oi; }
                }
                loo_t = bi == 0x7fffffff ? 0 : bi;
            }
            int tg = -1;
            if (ok[r] && lab == anom_label) tg = K - 1;
            else if (ok[r] && lab == nominal_label) tg = loo_t;
            float lt = 0.f;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (nt * 8 + 2 * t + e == tg) lt = l[nt][e];
            lt = __shfl_sync(kFullMask, lt, (lane & ~3) | (((tg < 0 ? 0 : tg) & 7) >> 1));
            if (tg >= 0 && t == 0) loss_acc += lse - lt;
            // G_k = softmax_k - [k == t] (x scale as the operand; 1/n is applied to the result), gd = sum_k G_k l_k
            float gd = 0.f;
            float G[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = nt * 8 + 2 * t + e;
                    const float gk = (tg >= 0 && k < K) ? (pr[nt][e] / sum - (k == tg ? 1.f : 0.f)) : 0.f;
                    if (tg >= 0 && k < K) gd += gk * l[nt][e];
                    G[nt][e] = gk * scale;
                }
            gd += __shfl_xor_sync(kFullMask, gd, 1);
            gd += __shfl_xor_sync(kFullMask, gd, 2);
            gdz[r] = gd * inv_nrm;
            osc[r] = inv_nrm * inv_n_f;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {                     // prompts 16ks + 2t + e (a0 / a1) and + 8 (a2 / a3)
                    const int nt = 2 * ks + hf;
                    const float g0 = nt < NT ? G[nt < NT ? nt : 0][0] : 0.f, g1 = nt < NT ? G[nt < NT ? nt : 0][1] : 0.f;
                    const uint32_t hi = clip_pack2<BF>(g0, g1);
                    const float2 hv = clip_unpack2<BF>(hi);
                    a_hi[ks][2 * hf + r] = hi;
                    a_lo[ks][2 * hf + r] = clip_pack2<BF>(g0 - hv.x, g1 - hv.y);
                }
            }
        }
        if (!grad) continue;
        // pass 2: per 16-byte run of columns, g = G @ C on the tensor core, z again (L2), dz out
        const int runs = d / (8 * GROUP);                          // d % 128 == 0: a multiple of QB
        constexpr int QB = 4;                                        // runs whose z loads are in flight together
        for (int q0 = 0; q0 < runs; q0 += QB) {
            uint4 zr[QB][2];
#pragma unroll
            for (int qq = 0; qq < QB; ++qq)
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int col = (q0 + qq) * 8 * GROUP + (sizeof(T) == 4 ? 4 : 8) * t;
                    if (ok[r])
                        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                                     : "=r"(zr[qq][r].x), "=r"(zr[qq][r].y), "=r"(zr[qq][r].z), "=r"(zr[qq][r].w)
                                     : "l"(z + row[r] * d + col), "l"(pol_stream));
                    else zr[qq][r] = make_uint4(0u, 0u, 0u, 0u);
                }
#pragma unroll
            for (int qq = 0; qq < QB; ++qq) {
                const int q = q0 + qq;
                float c[GROUP][4];
#pragma unroll
                for (int i = 0; i < GROUP; ++i) {
                    c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
                    const uint4* fr = s_fragT + ((size_t)(q * GROUP + i) * KS) * 32 + lane;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        const uint4 f = fr[ks * 32];
                        clip_mma<BF>(c[i], a_hi[ks], f.z, f.w);
                        clip_mma<BF>(c[i], a_lo[ks], f.x, f.y);
                        clip_mma<BF>(c[i], a_hi[ks], f.x, f.y);
                    }
                }
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    if (!ok[r]) continue;
                    const int col = q * 8 * GROUP + (sizeof(T) == 4 ? 4 : 8) * t;
                    const uint4 zq = zr[qq][r];
                    if (sizeof(T) == 4) {
                        const float zv[4] = {__uint_as_float(zq.x), __uint_as_float(zq.y), __uint_as_float(zq.z), __uint_as_float(zq.w)};
                        float o[4];
#pragma unroll
                        for (int i = 0; i < 2; ++i)
#pragma unroll
                            for (int e = 0; e < 2; ++e) o[2 * i + e] = (c[i][2 * r + e] - gdz[r] * zv[2 * i + e]) * osc[r];
                        asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                                     :: "l"(grad + row[r] * d + col), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]), "l"(pol_stream) : "memory");
                    } else {
                        const uint32_t zw[4] = {zq.x, zq.y, zq.z, zq.w};
                        uint32_t ow[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float2 zv = clip_unpack2<BF>(zw[i]);
                            ow[i] = clip_pack2<BF>((c[i & (GROUP - 1)][2 * r] - gdz[r] * zv.x) * osc[r],
                                                   (c[i & (GROUP - 1)][2 * r + 1] - gdz[r] * zv.y) * osc[r]);
                        }
                        asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;"
                                     :: "l"(grad + row[r] * d + col), "r"(ow[0]), "r"(ow[1]), "r"(ow[2]), "r"(ow[3]), "l"(pol_stream) : "memory");
                    }
                }
            }
        }
    }
    grid_mean_finish<kClipLossBlock>(loss_acc, ws, loss_out, inv_n);
}

template <typename T>
static int clip_loss_mma_launch(const void* z, const float* text, const int64_t* labels, int64_t n, int64_t d, int64_t K,
                                float scale, int64_t nominal, int loo, float* loss_out, void* grad, void* ws,
                                cudaStream_t st) {
    const int nt = (int)((K + 7) / 8), ks = (nt + 1) / 2;
    const size_t smem = (size_t)(d / 16) * nt * 512 + (grad ? (size_t)(d / 8) * ks * 512 : 0) + 32 * sizeof(float);
    const int64_t tiles = (n + 15) / 16;
    const int wpb = kClipLossBlock / 32;
    int grid = (int)((tiles + wpb - 1) / wpb);
    if (grid > kNumSMs) grid = kNumSMs;
    const float inv_n_f = 1.0f / (float)n;
    const double inv_n = 1.0 / (double)n;
#define EOE_CLIPL_MMA_CASE(NT)                                                                           \
    {                                                                                                    \
        auto kern = clip_oe_loss_mma_kernel<T, NT>;                                                      \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) { set_cuda_error(e, "clip_oe_loss_mma smem attr"); return EOE_ERR_CUDA; }  \
        kern<<<grid, kClipLossBlock, smem, st>>>((const T*)z, text, labels, n, (int)d, (int)K, scale, nominal, loo, \
                                                 (T*)grad, (HeadWorkspace*)ws, loss_out, inv_n_f, inv_n); \
    }
    if (nt == 1) EOE_CLIPL_MMA_CASE(1)
    else if (nt == 2) EOE_CLIPL_MMA_CASE(2)
    else if (nt == 3) EOE_CLIPL_MMA_CASE(3)
    else EOE_CLIPL_MMA_CASE(4)
#undef EOE_CLIPL_MMA_CASE
    return check_launch("clip_oe_loss_mma_kernel");
}

// K <= 64 prompts whose fp32 rows fit the shared-memory tile use it; larger prompt sets (up to kClipMaxPrompts) read the
// text rows from global memory
constexpr int64_t kClipMaxPrompts = 256;
static bool clip_text_from_global(int64_t d, int64_t K) {
    const int iters = (int)((d / 4 + 31) / 32);
    const int it_pad = iters <= 2 ? 2 : (iters <= 4 ? 4 : 8);
    return K > 64 || (size_t)K * it_pad * 128 * 4 > 192 * 1024;
}

template <typename T>
static bool clip_mma_ok(const void* z, const float* text, int64_t n, int64_t d, int64_t K) {
    return n >= 2048 && d % 128 == 0 && d <= 1024 && K >= 1 && K <= 32 && (uintptr_t)z % 16 == 0 && (uintptr_t)text % 16 == 0;
}

template <typename T>
static int clip_score_mma_launch(const void* z, const float* text, int64_t n, int64_t d, int64_t K, float scale,
                                 float* scores, cudaStream_t st) {
    const int nt = (int)((K + 7) / 8);
    const size_t smem = (size_t)(d / 16) * nt * 32 * sizeof(uint4) + 32 * sizeof(float);
    const int64_t tiles = (n + 15) / 16;
    int grid = (int)((tiles + kHeadWarps - 1) / kHeadWarps);
    if (grid > kNumSMs * 2) grid = kNumSMs * 2;
#define EOE_CLIP_MMA_CASE(NT)                                                                            \
    {                                                                                                    \
        auto kern = clip_score_mma_kernel<T, NT>;                                                        \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) { set_cuda_error(e, "clip_score_mma smem attr"); return EOE_ERR_CUDA; }    \
        kern<<<grid, kHeadBlock, smem, st>>>((const T*)z, text, n, (int)d, (int)K, scale, scores);       \
    }
    if (nt == 1) EOE_CLIP_MMA_CASE(1)
    else if (nt == 2) EOE_CLIP_MMA_CASE(2)
    else if (nt == 3) EOE_CLIP_MMA_CASE(3)
    else EOE_CLIP_MMA_CASE(4)
#undef EOE_CLIP_MMA_CASE
    return check_launch("clip_score_mma_kernel");
}

// vit.cu: the tcgen05 / TMEM / TMA score head for many 16-bit rows (clip_head_sm100.cuh)
int clip_score_tc16(const void* z, int dtype, const float* text, int64_t n, int64_t d, int64_t K, float scale, float* scores,
                    cudaStream_t st);
int clip_loss_tc16(const void* z, int dtype, const float* text, const int64_t* labels, int64_t n, int64_t d, int64_t K,
                   float scale, int64_t nominal, int loo, float* loss_out, void* grad, void* ws, cudaStream_t st);
static int g_clip_tc_min_rows = 16384;          // below: the warp-level kernel (fewer than one 128-row tile per SM otherwise)

template <typename T>
static int clip_score_launch(const void* z, const float* text, int64_t n, int64_t d, int64_t K, float scale,
                             float* scores, cudaStream_t st) {
    // 16-bit rows, many of them: logits on tcgen05 with the rows used as UMMA operands where TMA puts them
    if (sizeof(T) == 2 && n >= g_clip_tc_min_rows && d % 64 == 0 && d <= 512 && K >= 1 && K <= 32 &&
        (uintptr_t)z % 16 == 0 && (uintptr_t)text % 16 == 0)
        return clip_score_tc16(z, std::is_same<T, __half>::value ? EOE_F16 : EOE_BF16, text, n, d, K, scale, scores, st);
    if (clip_mma_ok<T>(z, text, n, d, K)) return clip_score_mma_launch<T>(z, text, n, d, K, scale, scores, st);
    const int iters = (int)((d / 4 + 31) / 32);
    const int it_pad = iters <= 2 ? 2 : (iters <= 4 ? 4 : 8);
    const bool gtext = clip_text_from_global(d, K);
    const size_t smem = gtext ? (size_t)K * sizeof(float) : (size_t)K * it_pad * 128 * sizeof(float);
#define EOE_CLIP_CASE(IT, RW)                                                                            \
    if (gtext) {                                                                                         \
        int grid = head_grid(n, kHeadWarps * RW);                                                        \
        if (grid > kNumSMs * 4) grid = kNumSMs * 4;                                                      \
        clip_score_kernel<T, IT, RW, true><<<grid, kHeadBlock, smem, st>>>((const T*)z, text, n, (int)d, (int)K, scale, scores); \
    } else {                                                                                             \
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
    if (sizeof(T) == 2 && grad && n >= g_clip_tc_min_rows && d % 64 == 0 && d <= 512 && K >= 1 && K <= 32 &&
        (uintptr_t)z % 16 == 0 && (uintptr_t)text % 16 == 0 && (uintptr_t)grad % 16 == 0)
        return clip_loss_tc16(z, std::is_same<T, __half>::value ? EOE_F16 : EOE_BF16, text, labels, n, d, K, scale, nominal, loo,
                              loss_out, grad, ws, st);
    if (clip_mma_ok<T>(z, text, n, d, K) && (uintptr_t)grad % 16 == 0 && d <= 512)
        return clip_loss_mma_launch<T>(z, text, labels, n, d, K, scale, nominal, loo, loss_out, grad, ws, st);
    const int iters = (int)((d / 4 + 31) / 32);
    const int it_pad = iters <= 2 ? 2 : (iters <= 4 ? 4 : 8);
    const bool gtext = clip_text_from_global(d, K);
    const size_t smem = gtext ? 0 : (size_t)K * it_pad * 128 * sizeof(float);
    const float inv_n_f = 1.0f / (float)n;
    const double inv_n = 1.0 / (double)n;
#define EOE_CLIPL_CASE(IT)                                                                               \
    if (gtext) {                                                                                         \
        int grid = head_grid(n, kHeadWarps);                                                             \
        if (grid > kNumSMs * 4) grid = kNumSMs * 4;                                                      \
        clip_oe_loss_kernel<T, IT, 8, true><<<grid, kHeadBlock, 0, st>>>((const T*)z, text, labels, n, (int)d, (int)K, scale, \
            nominal, loo, (T*)grad, (HeadWorkspace*)ws, loss_out, inv_n_f, inv_n);                       \
    } else {                                                                                             \
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

int clip_prompts_ok(int64_t K) { return K >= 1 && K <= kClipMaxPrompts; }      // the fused encoder validates K up front

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
    NvtxRange nvtx("eoe:hsc fwd+bwd+score");
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
    NvtxRange nvtx("eoe:bce fwd+bwd+score");
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
    if (d % 4 != 0 || d > 1024 || K > kClipMaxPrompts) return EOE_ERR_SHAPE;
    if ((uintptr_t)z % 16 != 0 || (uintptr_t)text % 16 != 0) return EOE_ERR_ALIGN;
    return EOE_OK;
}

extern "C" int eoe_clip_score(const void* z, int z_dtype, const float* text, int64_t n, int64_t d, int64_t K,
                              float scale, float* scores_out, void* stream) {
    int rc = clip_check(z, text, n, d, K);
    if (rc) return rc;
    if (!scores_out) return EOE_ERR_ARG;
    NvtxRange nvtx("eoe:clip score");
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
    NvtxRange nvtx("eoe:clip oe loss fwd+bwd");
    cudaStream_t st = (cudaStream_t)stream;
    EOE_DISPATCH_DTYPE(z_dtype, (clip_loss_launch<T>(z, text, labels, n, d, K, scale, nominal_label, leave_one_out, loss_out, grad_z_out, head_ws, st)))
}
