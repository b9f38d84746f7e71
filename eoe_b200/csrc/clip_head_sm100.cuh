// tcgen05 / TMEM / TMA version of the CLIP zero-shot score head for 16-bit feature rows and many rows
// (ADClipTrainer.compute_anomaly_score, reference src/eoe/training/clip.py:66-79):
//     score[i] = softmax_k( scale * z^_i . t^_k )[K - 1]
// The warp-level kernel (heads.cu, mma.sync fed from registers) is bound by the legacy-HMMA issue rate and the fragment
// loads once the rows are 16 bit and K = 30 (53 % of HBM).  Here the logits of a 128-row tile are ONE accumulator tile:
//   A   z tile, 128 rows x 64 columns per chunk, brought by TMA (SWIZZLE_128B) into an 8-stage ring -- the feature rows are
//       already K-major 16-bit, i.e. the UMMA operand layout, and are used as they are (exact operands)
//   B   the unit text rows, staged once per CTA as [hi (32 prompts) ; lo (32 prompts)] x 64 columns per chunk: N = 64
//   D   [128 x 64] fp32 in TMEM (two accumulator stages): columns k and 32 + k hold z . t_hi_k and z . t_lo_k
// Warp roles (192 threads, one CTA per SM, persistent over tiles):
//   warps 0-3  row owners: thread t owns row t of the tile -- while the chunks stream by it reads its own 128 bytes of each
//              chunk from shared memory for the row norm (conflict-free: 8 neighbouring threads start at 8 different
//              16-byte columns), then reads its accumulator row with tcgen05.ld and finishes softmax + score in registers
//   warp 4     TMA producer          warp 5     TMEM allocator + MMA issuer
// Same arithmetic as the warp-level kernel (text split hi + lo in the feature dtype, fp32 accumulation, fp32 row norms,
// expf softmax), so the same parity tests hold it to 1e-3 on the scores.
#pragma once
#include "common.cuh"
#include "gemm_sm100.cuh"
#include "sm100_ptx.cuh"

namespace eoe {
namespace cliptc {

constexpr int THREADS = 192;
constexpr int NST = 8;                               // z chunks in flight: 128 KB per SM
constexpr uint32_t ZCH = 128 * 128;                  // one z chunk: 128 rows x 64 x 16 bit
constexpr uint32_t TCH = 64 * 128;                   // one text chunk: 64 rows (hi ; lo) x 64 x 16 bit
constexpr int MAX_CH = 8;                            // d <= 512
constexpr uint32_t SMEM_BYTES = MAX_CH * TCH + NST * ZCH + 512 + 1024;
constexpr uint32_t TMEM_COLS = 128;

template <bool BF16>
__device__ __forceinline__ float2 unpack2(uint32_t w) {
    if (BF16) return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
    return __half22float2(*reinterpret_cast<__half2*>(&w));
}

template <bool BF16>
__global__ void __launch_bounds__(THREADS, 1)
clip_score_tc_kernel(const __grid_constant__ CUtensorMap tm_z, const float* __restrict__ text, int64_t n, int d, int K,
                     float scale, float* __restrict__ scores) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sT = smem;                                             // [nch][64 rows][128 B]
    uint8_t* sZ = smem + MAX_CH * TCH;                              // [NST][128 rows][128 B]
    uint64_t* full = reinterpret_cast<uint64_t*>(sZ + NST * ZCH);   // [NST] TMA -> MMA and row owners
    uint64_t* empty = full + NST;                                   // [NST] MMA commit + 4 row-owner warps -> TMA
    uint64_t* tfull = empty + NST;                                  // [2]   MMA -> row owners
    uint64_t* tempty = tfull + 2;                                   // [2]   row owners -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* s_inv = reinterpret_cast<float*>(tmem_slot + 2);         // [32] 1 / ||t_k||

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(kFullMask, tid >> 5, 0);
    const int nch = d >> 6;
    const int64_t tiles = (n + 127) >> 7;

    if (tid == 0) {
        ptx::prefetch_tensormap(&tm_z);
        for (int s = 0; s < NST; ++s) {
            ptx::mbar_init(ptx::smem_u32(&full[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&empty[s]), 5);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(ptx::smem_u32(&tfull[a]), 1);
            ptx::mbar_init(ptx::smem_u32(&tempty[a]), 4);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 5) {
        ptx::tmem_alloc<1>(ptx::smem_u32(tmem_slot), TMEM_COLS);
        ptx::tmem_relinquish<1>();
    }
    // ---- text rows -> unit norm -> (hi, lo) in the feature dtype -> UMMA B tiles (128-byte swizzle, written by hand)
    for (int k = tid >> 5; k < 32; k += THREADS / 32) {
        float s = 0.f;
        if (k < K)
            for (int i = lane; i < d; i += 32) { const float v = __ldg(text + (int64_t)k * d + i); s += v * v; }
        s = warp_sum(s);
        if (lane == 0) s_inv[k] = k < K ? 1.0f / sqrtf(s) : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < 32 * (d >> 3); i += THREADS) {
        const int k = i / (d >> 3), c8 = i % (d >> 3);               // prompt, group of 8 columns
        uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
        if (k < K) {
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(text + (int64_t)k * d + c8 * 8));
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(text + (int64_t)k * d + c8 * 8) + 1);
            const float inv = s_inv[k];
            const float e[8] = {v0.x * inv, v0.y * inv, v0.z * inv, v0.w * inv, v1.x * inv, v1.y * inv, v1.z * inv, v1.w * inv};
            uint32_t h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                h[j] = gemm::pack2<BF16>(e[2 * j], e[2 * j + 1]);
                const float2 f = unpack2<BF16>(h[j]);
                l[j] = gemm::pack2<BF16>(e[2 * j] - f.x, e[2 * j + 1] - f.y);
            }
            hi = make_uint4(h[0], h[1], h[2], h[3]);
            lo = make_uint4(l[0], l[1], l[2], l[3]);
        }
        const int ch = c8 >> 3, q = c8 & 7;
        uint8_t* base = sT + ch * TCH;
        *reinterpret_cast<uint4*>(base + k * 128 + ((q ^ (k & 7)) << 4)) = hi;
        *reinterpret_cast<uint4*>(base + (32 + k) * 128 + ((q ^ (k & 7)) << 4)) = lo;      // (32 + k) & 7 == k & 7
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            // ------------------------------------------------------------------ TMA producer
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                for (int ch = 0; ch < nch; ++ch) {
                    ptx::mbar_wait(ptx::smem_u32(&empty[stage]), phase ^ 1);
                    ptx::mbar_arrive_expect_tx(ptx::smem_u32(&full[stage]), ZCH);
                    ptx::tma_load_2d(ptx::smem_u32(sZ + stage * ZCH), &tm_z, ptx::smem_u32(&full[stage]), ch * 64, (int)(tile * 128));
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            // ------------------------------------------------------------------ MMA issuer
            constexpr uint32_t idesc = ptx::make_idesc_f16(BF16 ? 1u : 0u, 128, 64);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                ptx::mbar_wait(ptx::smem_u32(&tempty[acc]), acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem + (uint32_t)acc * 64;
                for (int ch = 0; ch < nch; ++ch) {
                    ptx::mbar_wait(ptx::smem_u32(&full[stage]), phase);
                    ptx::tc_fence_after();
                    const uint64_t a_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sZ + stage * ZCH));
                    const uint64_t b_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sT + ch * TCH));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_f16<1>(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (ch | k) != 0 ? 1u : 0u);
                    ptx::umma_commit(ptx::smem_u32(&empty[stage]));
                    if (++stage == NST) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(ptx::smem_u32(&tfull[acc]));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ---------------------------------------------------------------------- row owners (warps 0-3)
        const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        const int k_last = K - 1;
        for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            float ss = 0.f;
            for (int ch = 0; ch < nch; ++ch) {
                ptx::mbar_wait(ptx::smem_u32(&full[stage]), phase);
                const uint8_t* rowp = sZ + stage * ZCH + tid * 128;
#pragma unroll
                for (int p = 0; p < 8; ++p) {       // the sum does not care which 16-byte column comes first: stagger them
                    const uint4 v = *reinterpret_cast<const uint4*>(rowp + ((p ^ (tid & 7)) << 4));
                    const float2 a = unpack2<BF16>(v.x), b = unpack2<BF16>(v.y), c = unpack2<BF16>(v.z), e = unpack2<BF16>(v.w);
                    ss += (a.x * a.x + a.y * a.y) + (b.x * b.x + b.y * b.y) + (c.x * c.x + c.y * c.y) + (e.x * e.x + e.y * e.y);
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&empty[stage]));
                if (++stage == NST) { stage = 0; phase ^= 1; }
            }
            ptx::mbar_wait(ptx::smem_u32(&tfull[acc]), acc_phase);
            ptx::tc_fence_after();
            uint32_t r0[32], r1[32];
            ptx::tmem_ld_32x32b_x32(t_lane + (uint32_t)acc * 64, r0);
            ptx::tmem_ld_32x32b_x32(t_lane + (uint32_t)acc * 64 + 32, r1);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&tempty[acc]));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            const float inv = scale / sqrtf(ss);
            float mx = -INFINITY, last = 0.f;
            float l[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                l[k] = k < K ? (__uint_as_float(r1[k]) + __uint_as_float(r0[k])) * inv : -INFINITY;    // small term first
                mx = fmaxf(mx, l[k]);
                if (k == k_last) last = l[k];
            }
            float se = 0.f;
#pragma unroll
            for (int k = 0; k < 32; ++k) se += expf(l[k] - mx);          // padded prompts: exp(-inf) = 0; NaN logits stick
            const int64_t row = tile * 128 + tid;
            if (row < n) scores[row] = expf(last - mx) / se;
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem, TMEM_COLS);
    }
}

}  // namespace cliptc
}  // namespace eoe
