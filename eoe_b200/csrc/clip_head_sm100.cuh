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

// ------------------------------------------------------------------------------------------------------------------------
// CLIP outlier-exposure loss + backward w.r.t. the features (ADClipTrainer.loss, clip.py:81-103) for many 16-bit rows.
// Forward as above.  Backward: with G = softmax - onehot(target) (x scale) per row,
//     dz = (G @ C - (sum_k G_k logit_k) z^) / (n ||z||),        C = the text rows as given (the loss does not renormalise)
// G @ C is the second small GEMM: A = G (hi, lo as 16-bit pairs) written by a row owner into TMEM -- a thread owns a row =
// a TMEM lane, so it goes from registers straight to where tcgen05.mma reads its A operand --, B = the forward's text
// tiles read as MN-major operands (features contiguous, prompts = k): six MMAs per 64-column chunk (G_hi T_hi, G_hi T_lo,
// G_lo T_hi over two 16-prompt k-steps), D_bwd [128 x 64] fp32, double buffered.
// The two halves of a tile's work are done by DIFFERENT warps, one tile apart, so that neither HBM nor the tensor pipe
// waits for a softmax:
//   warps 0-3  "F": norm pass over the forward chunks, logits from TMEM, softmax / loss / G of tile j+1 ...
//   warps 4-11 "B": ... while they turn tile j's accumulator chunks into dz: z is streamed a second time (TMA, L2 hits:
//              148 x 128 KB in flight is far below the 126 MB L2), dz goes through a per-warp staging block and out by TMA.
//              A chunk costs a warp ~1 500 clocks of mostly fixed latencies (three mbarrier waits, TMEM load, proxy
//              fence, the serial TMA-store tail), so two sets share the work: B0 = warps 4-7 takes the even chunks
//              (accumulator 0), B1 = warps 8-11 the odd ones (accumulator 1)
//   warp 12    TMA producer of the forward ring (3 stages; consumers: F and the forward issuer, in lockstep on every use)
//   warp 13    TMEM allocator + issuer of the forward MMAs;  warp 14  issuer of the backward MMAs.  One thread issuing
//              all 80 MMAs + 24 commits of a tile was busy for 10 000 of the tile's 14 000 clocks and starved B (clock64
//              probes, profiles/r2_clip_loss_tc_clocks.json): issuing a small tcgen05.mma costs its thread ~100 clocks,
//              three times what an M 128 x N 64 x K 16 product occupies the tensor pipe.
// Each consumer group has its OWN ring -- an mbarrier carries one phase bit, so a waiter may never skip a use of a stage:
// the forward ring above, and a private 2-stage ring per backward set that the set feeds itself (when its four warps have
// passed a named barrier behind a chunk, one thread issues the TMA load of the set's chunk after next into that stage).
// F hands a tile to B and to the MMA issuer through `gready` (G in TMEM, the two row scalars in shared memory, both double
// buffered) and gets the buffers back through `gfree`.
// TMEM (512 columns, one CTA per SM): D_fwd 2 x 64 | G 2 x 32 | D_bwd 2 x 64.
constexpr int LOSS_THREADS = 480;                    // warps 0-3 F, 4-7 B0 (even chunks), 8-11 B1 (odd chunks), 12 TMA, 13 / 14 MMA
constexpr uint32_t LOSS_TMEM_COLS = 512, G_COL = 128, DB_COL = 192;
constexpr int LNST = 3;                              // forward ring stages (text 64 KB + 3 x 16 KB + 2 x 2 x 16 KB + 8 x 4 KB staging)
constexpr uint32_t LOSS_SMEM_BYTES = MAX_CH * TCH + (LNST + 4) * ZCH + 8 * 4096 + 512 + 4 * 512 + 1024;

template <bool BF16>
__global__ void __launch_bounds__(LOSS_THREADS, 1)
clip_oe_loss_tc_kernel(const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_g,
                       const float* __restrict__ text, const int64_t* __restrict__ labels, int64_t n, int d, int K, float scale,
                       int64_t nominal_label, int loo, HeadWorkspace* ws, float* loss_out, float inv_n_f, double inv_n, int dbg) {
    constexpr int NST = LNST;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sT = smem;
    uint8_t* sZ = smem + MAX_CH * TCH;
    uint8_t* sZb = sZ + NST * ZCH;                                  // [2 sets][2 stages][128 rows][128 B] the backward sets' own rings
    uint8_t* sOut = sZb + 4 * ZCH;                                  // [8 B warps][32 rows][128 B] dz staging
    uint64_t* full = reinterpret_cast<uint64_t*>(sOut + 8 * 4096);  // [NST] forward ring: TMA -> F, forward issuer
    uint64_t* empty = full + NST;                                   // [NST] forward issuer's commit + the 4 F warps -> TMA
    uint64_t* fullb = empty + NST;                                  // [2][2] backward rings: TMA -> the set
    uint64_t* tfull = fullb + 4;                                    // [2]   forward accumulator: MMA -> F
    uint64_t* tempty = tfull + 2;                                   // [2]   F -> MMA
    uint64_t* bfull = tempty + 2;                                   // [2]   backward accumulator: MMA -> B
    uint64_t* bempty = bfull + 2;                                   // [2]   B -> MMA
    uint64_t* gready = bempty + 2;                                  // [2]   G + row scalars of a tile are published: F -> MMA, B
    uint64_t* gfree = gready + 2;                                   // [2]   ... and consumed: MMA commit + the 8 B warps -> F
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gfree + 2);
    float* s_gdz = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + 512);     // [2][128] (sum_k G_k l_k) / ||z||
    float* s_osc = s_gdz + 256;                                     // [2][128] 1 / (n ||z||)

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(kFullMask, tid >> 5, 0);
    const int nch = d >> 6;
    const int64_t tiles = (n + 127) >> 7;
    const int64_t m = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;      // tiles of this CTA
    auto fwd_use = [&](int64_t j) { return j * (int64_t)nch; };      // forward ring: use index of chunk 0 of the CTA's j-th tile

    if (tid == 0) {
        ptx::prefetch_tensormap(&tm_z);
        ptx::prefetch_tensormap(&tm_g);
        for (int s = 0; s < NST; ++s) {
            ptx::mbar_init(ptx::smem_u32(&full[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&empty[s]), 5);
        }
        for (int a = 0; a < 4; ++a) ptx::mbar_init(ptx::smem_u32(&fullb[a]), 1);
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(ptx::smem_u32(&tfull[a]), 1);
            ptx::mbar_init(ptx::smem_u32(&tempty[a]), 4);
            ptx::mbar_init(ptx::smem_u32(&bfull[a]), 1);
            ptx::mbar_init(ptx::smem_u32(&bempty[a]), 4);
            ptx::mbar_init(ptx::smem_u32(&gready[a]), 4);
            ptx::mbar_init(ptx::smem_u32(&gfree[a]), 9);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 13) {
        ptx::tmem_alloc<1>(ptx::smem_u32(tmem_slot), LOSS_TMEM_COLS);
        ptx::tmem_relinquish<1>();
    }
    // ---- text rows AS GIVEN (clip.py:81-103 works on the centre it is handed) -> (hi, lo) -> UMMA tiles
    for (int i = tid; i < 32 * (d >> 3); i += LOSS_THREADS) {
        const int k = i / (d >> 3), c8 = i % (d >> 3);
        uint4 hi = make_uint4(0u, 0u, 0u, 0u), lo = hi;
        if (k < K) {
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(text + (int64_t)k * d + c8 * 8));
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(text + (int64_t)k * d + c8 * 8) + 1);
            const float e[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
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
        *reinterpret_cast<uint4*>(base + (32 + k) * 128 + ((q ^ (k & 7)) << 4)) = lo;
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    float loss_acc = 0.f;
    auto tile_of = [&](int64_t j) { return (int64_t)blockIdx.x + j * gridDim.x; };

    if (warp == 12) {
        if (lane == 0) {
            // ------------------------------------------------------------------ TMA producer
            int64_t u = 0;
            for (int64_t j = 0; j < m; ++j) {
                const int64_t tile = tile_of(j);
                for (int ch = 0; ch < nch; ++ch, ++u) {
                    const int s = (int)(u % NST);
                    ptx::mbar_wait(ptx::smem_u32(&empty[s]), (uint32_t)(((u / NST) & 1) ^ 1));
                    ptx::mbar_arrive_expect_tx(ptx::smem_u32(&full[s]), ZCH);
                    ptx::tma_load_2d(ptx::smem_u32(sZ + s * ZCH), &tm_z, ptx::smem_u32(&full[s]), ch * 64, (int)(tile * 128));
                }
            }
        }
    } else if (warp == 13) {
        if (lane == 0) {
            // ------------------------------------------------------------------ issuer of the forward MMAs
            constexpr uint32_t idesc_f = ptx::make_idesc_f16(BF16 ? 1u : 0u, 128, 64);
            for (int64_t j = 0; j < m; ++j) {
                const int acc = (int)(j & 1);
                ptx::mbar_wait(ptx::smem_u32(&tempty[acc]), (uint32_t)(((j >> 1) & 1) ^ 1));
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem + (uint32_t)acc * 64;
                int64_t u = fwd_use(j);
                for (int ch = 0; ch < nch; ++ch, ++u) {
                    const int s = (int)(u % NST);
                    ptx::mbar_wait(ptx::smem_u32(&full[s]), (uint32_t)((u / NST) & 1));
                    ptx::tc_fence_after();
                    const uint64_t a_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sZ + s * ZCH));
                    const uint64_t b_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sT + ch * TCH));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::umma_f16<1>(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc_f, (ch | k) != 0 ? 1u : 0u);
                    ptx::umma_commit(ptx::smem_u32(&empty[s]));
                }
                ptx::umma_commit(ptx::smem_u32(&tfull[acc]));
            }
        }
    } else if (warp == 14) {
        if (lane == 0) {
            // ------------------------------------------------------------------ issuer of the backward MMAs
            constexpr uint32_t idesc_b = ptx::make_idesc_f16(BF16 ? 1u : 0u, 128, 64, 0, 1);      // B = text chunk, MN-major
            uint32_t b_phase[2] = {0, 0};
            for (int64_t j = 0; j < m; ++j) {
                const int g = (int)(j & 1);
                ptx::mbar_wait(ptx::smem_u32(&gready[g]), (uint32_t)((j >> 1) & 1));
                ptx::tc_fence_after();
                const uint32_t gcol = tmem + G_COL + (uint32_t)g * 32;
                for (int ch = 0; ch < nch; ++ch) {
                    const int bb = ch & 1;
                    ptx::mbar_wait(ptx::smem_u32(&bempty[bb]), b_phase[bb] ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t db = tmem + DB_COL + (uint32_t)bb * 64;
                    const uint64_t t_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sT + ch * TCH));
                    // A: packed G columns (8 per 16 prompts): hi [0, 16), lo [16, 32); B: 16-prompt row blocks of 2 KB: hi 0, 1, lo 2, 3
                    ptx::umma_f16_ts(db, gcol + 0, t_desc + (uint64_t)(0 * 2048 >> 4), idesc_b, 0u);      // G_hi T_hi
                    ptx::umma_f16_ts(db, gcol + 8, t_desc + (uint64_t)(1 * 2048 >> 4), idesc_b, 1u);
                    ptx::umma_f16_ts(db, gcol + 0, t_desc + (uint64_t)(2 * 2048 >> 4), idesc_b, 1u);      // G_hi T_lo
                    ptx::umma_f16_ts(db, gcol + 8, t_desc + (uint64_t)(3 * 2048 >> 4), idesc_b, 1u);
                    ptx::umma_f16_ts(db, gcol + 16, t_desc + (uint64_t)(0 * 2048 >> 4), idesc_b, 1u);     // G_lo T_hi
                    ptx::umma_f16_ts(db, gcol + 24, t_desc + (uint64_t)(1 * 2048 >> 4), idesc_b, 1u);
                    ptx::umma_commit(ptx::smem_u32(&bfull[bb]));
                    b_phase[bb] ^= 1;
                }
                ptx::umma_commit(ptx::smem_u32(&gfree[g]));             // every MMA that reads this G buffer has completed
            }
        }
    } else if (warp < 4) {
        // ---------------------------------------------------------------------- F: norm, logits, softmax, loss, G
        const int rt = tid;                                              // row of the tile = TMEM lane
        const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
        const int64_t anom_label = 1 - nominal_label;
        for (int64_t j = 0; j < m; ++j) {
            const int64_t tile = tile_of(j);
            float ss = 0.f;
            int64_t u = fwd_use(j);
            for (int ch = 0; ch < nch; ++ch, ++u) {
                const int s = (int)(u % NST);
                ptx::mbar_wait(ptx::smem_u32(&full[s]), (uint32_t)((u / NST) & 1));
                const uint8_t* rowp = sZ + s * ZCH + rt * 128;
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const uint4 v = *reinterpret_cast<const uint4*>(rowp + ((p ^ (rt & 7)) << 4));
                    const float2 a = unpack2<BF16>(v.x), b = unpack2<BF16>(v.y), c = unpack2<BF16>(v.z), e = unpack2<BF16>(v.w);
                    ss += (a.x * a.x + a.y * a.y) + (b.x * b.x + b.y * b.y) + (c.x * c.x + c.y * c.y) + (e.x * e.x + e.y * e.y);
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&empty[s]));
            }
            const int acc = (int)(j & 1);
            ptx::mbar_wait(ptx::smem_u32(&tfull[acc]), (uint32_t)((j >> 1) & 1));
            ptx::tc_fence_after();
            uint32_t r0[32], r1[32];
            ptx::tmem_ld_32x32b_x32(t_lane + (uint32_t)acc * 64, r0);
            ptx::tmem_ld_32x32b_x32(t_lane + (uint32_t)acc * 64 + 32, r1);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&tempty[acc]));
            const int64_t row = tile * 128 + rt;
            const bool ok = row < n;
            const float inv_nrm = 1.0f / sqrtf(ss);
            const float sc = scale * inv_nrm;
            float l[32];
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                l[k] = k < K ? (__uint_as_float(r1[k]) + __uint_as_float(r0[k])) * sc : -INFINITY;
                mx = fmaxf(mx, l[k]);
            }
            float pr[32];
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < 32; ++k) { pr[k] = expf(l[k] - mx); sum += pr[k]; }
            const float lse = mx + logf(sum);
            const int64_t lab = ok ? labels[row] : 0;
            int loo_t = 0;                                            // argmax over k < K - 1, first maximal index (clip.py:95)
            if (loo) {
                float best = -INFINITY;
                bool have = false;
#pragma unroll
                for (int k = 0; k < 32; ++k)
                    if (k < K - 1 && (!have || l[k] > best)) { best = l[k]; loo_t = k; have = true; }
            }
            int tg = -1;
            if (ok && lab == anom_label) tg = K - 1;
            else if (ok && lab == nominal_label) tg = loo_t;
            float lt = 0.f;
#pragma unroll
            for (int k = 0; k < 32; ++k) if (k == tg) lt = l[k];
            if (tg >= 0) loss_acc += lse - lt;
            float gd = 0.f;
            uint32_t gh[16], gl[16];
            const float inv_sum = 1.0f / sum;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                float g2[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = 2 * c + e;
                    const float gk = (tg >= 0 && k < K) ? (pr[k] * inv_sum - (k == tg ? 1.f : 0.f)) : 0.f;
                    if (tg >= 0 && k < K) gd += gk * l[k];
                    g2[e] = gk * scale;
                }
                gh[c] = gemm::pack2<BF16>(g2[0], g2[1]);
                const float2 f = unpack2<BF16>(gh[c]);
                gl[c] = gemm::pack2<BF16>(g2[0] - f.x, g2[1] - f.y);
            }
            // publish: the G buffer and the scalar slots of tile j-2 must have been consumed
            const int g = (int)(j & 1);
            ptx::mbar_wait(ptx::smem_u32(&gfree[g]), (uint32_t)(((j >> 1) & 1) ^ 1));
            ptx::tc_fence_after();
            s_gdz[g * 128 + rt] = gd * inv_nrm;
            s_osc[g * 128 + rt] = inv_nrm * inv_n_f;
            ptx::tmem_st_32x32b_x16(t_lane + G_COL + (uint32_t)g * 32, gh);
            ptx::tmem_st_32x32b_x16(t_lane + G_COL + (uint32_t)g * 32 + 16, gl);
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            __threadfence_block();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&gready[g]));
        }
    } else if (warp < 12) {
        // ---------------------------------------------------------------------- B0 / B1: dz = (G @ C - gdz z) osc per chunk
        const int set = (warp - 4) >> 2;                                 // 0: even chunks / accumulator 0, 1: odd / 1
        const int bw = (warp - 4) & 3;                                   // TMEM lane quarter
        const int rt = (tid - 128) & 127;
        const uint32_t t_lane = tmem + ((uint32_t)(bw * 32) << 16);
        uint8_t* blk = sOut + (warp - 4) * 4096;                         // this warp's staging block, 32 rows x 128 B
        uint8_t* outp = blk + (rt & 31) * 128;
        uint8_t* ring = sZb + set * 2 * ZCH;                             // this set's own 2-stage ring
        uint64_t* rfull = fullb + set * 2;
        const bool loader = bw == 0 && lane == 0;
        uint32_t b_phase = 0;
        // the set's chunk sequence: (tile j, chunk ch) for ch = set, set + 2, ...; `seq` numbers them over all tiles
        int64_t seq = 0, lseq = 0, lj = 0;
        int lch = set;
        auto load_next = [&]() {                                         // loader thread: the next chunk of the sequence -> stage lseq & 1
            while (lj < m && lch >= nch) { ++lj; lch = set; }
            if (lj >= m) return;
            const int st = (int)(lseq & 1);
            ptx::mbar_arrive_expect_tx(ptx::smem_u32(&rfull[st]), ZCH);
            ptx::tma_load_2d(ptx::smem_u32(ring + st * ZCH), &tm_z, ptx::smem_u32(&rfull[st]), lch * 64, (int)(tile_of(lj) * 128));
            ++lseq;
            lch += 2;
        };
        if (loader) { load_next(); load_next(); }
        for (int64_t j = 0; j < m; ++j) {
            const int64_t tile = tile_of(j);
            const int g = (int)(j & 1);
            ptx::mbar_wait(ptx::smem_u32(&gready[g]), (uint32_t)((j >> 1) & 1));
            const float gdz = s_gdz[g * 128 + rt], osc = s_osc[g * 128 + rt];
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&gfree[g]));
            for (int ch = set; ch < nch; ch += 2, ++seq) {
                const int st = (int)(seq & 1);
                ptx::mbar_wait(ptx::smem_u32(&bfull[set]), b_phase);
                b_phase ^= 1;
                ptx::tc_fence_after();
                uint32_t g0[32], g1[32];
                ptx::tmem_ld_32x32b_x32(t_lane + DB_COL + (uint32_t)set * 64, g0);
                ptx::tmem_ld_32x32b_x32(t_lane + DB_COL + (uint32_t)set * 64 + 32, g1);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive(ptx::smem_u32(&bempty[set]));
                    ptx::bulk_wait_group_read0();                       // this warp's previous dz block has left its staging
                }
                ptx::mbar_wait(ptx::smem_u32(&rfull[st]), (uint32_t)((seq >> 1) & 1));
                __syncwarp();
                const uint8_t* rowp = ring + st * ZCH + rt * 128;
#pragma unroll
                for (int q = 0; q < 8; ++q) {                          // logical 16-byte column q = features 8q .. 8q+7 of the chunk
                    const uint32_t off = (uint32_t)((q ^ (rt & 7)) << 4);
                    const uint4 v = *reinterpret_cast<const uint4*>(rowp + off);
                    const uint32_t zw[4] = {v.x, v.y, v.z, v.w};
                    uint32_t ow[4];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const float2 zv = unpack2<BF16>(zw[jj]);
                        const int c = 8 * q + 2 * jj;
                        const float ga = __uint_as_float(c < 32 ? g0[c & 31] : g1[c & 31]);
                        const float gb = __uint_as_float(c < 32 ? g0[(c + 1) & 31] : g1[(c + 1) & 31]);
                        ow[jj] = gemm::pack2<BF16>((ga - gdz * zv.x) * osc, (gb - gdz * zv.y) * osc);
                    }
                    *reinterpret_cast<uint4*>(outp + off) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                }
                ptx::fence_proxy_async_smem();
                // all four warps of the set are done with the ring stage: it takes the set's chunk after next
                asm volatile("bar.sync %0, 128;" ::"r"(2 + set) : "memory");
                if (loader) load_next();
                if (lane == 0 && !(dbg & 1)) {                          // (diagnostics bit 0: no dz stores)
                    ptx::tma_store_2d(&tm_g, ptx::smem_u32(blk), ch * 64, (int)(tile * 128 + bw * 32));
                    ptx::bulk_commit_group();
                }
            }
        }
        if (lane == 0) ptx::bulk_wait_group_read0();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 13) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem, LOSS_TMEM_COLS);
    }
    grid_mean_finish<LOSS_THREADS>(loss_acc, ws, loss_out, inv_n);
}

}  // namespace cliptc
}  // namespace eoe
