// tcgen05 attention for the ViT-B/16 token count (L = 197): softmax(Q K^T / 8) V per (image, head)
// (reference: nn.MultiheadAttention inside ResidualAttentionBlock.attention, clip_official/clip/model.py:181-183).
//
// One CTA of 128 threads per (image, head) item, persistent over items, 2 CTAs per SM (their TMA / MMA / softmax phases
// interleave).  Per item: TMA brings the head's Q (2 tiles of 128 query rows), K and V slices ([rows][64] bf16, 128-byte
// rows, SWIZZLE_128B) straight out of the qkv GEMM output through a 3-D [image][token][column] map (tokens past the image
// are zero-filled for free).  Per query tile:
//   S[128 x 208] = Q K^T        tcgen05.mma, both operands K-major from smem, fp32 in TMEM columns [0, 208)
//   softmax                      thread t owns TMEM lane t = query row t: row max and exp2 need no cross-thread traffic;
//                                un-normalised P is written back to TMEM as packed 16-bit pairs over columns [0, 104)
//   O[128 x 64] = P V           tcgen05.mma with A = P from TMEM and B = V as an MN-major smem operand (no transpose),
//                                fp32 in TMEM columns [128, 192)
//   epilogue                     O * (1 / row sum) -> 16-bit -> staged in the dead Q tile -> one TMA tile store (clipped at L)
// L = 197 rows do not fill two 128-lane tiles; the second tile alternates between rows [128, 256) and rows [69, 197)
// from item to item so that the valid rows load all four TMEM lane quarters (= the four SM sub-partitions whose MUFU
// units do the exp2 work) evenly.
#pragma once
#include "common.cuh"
#include "gemm_sm100.cuh"
#include "sm100_ptx.cuh"

namespace eoe {
namespace attn {

constexpr int THREADS = 128;
constexpr uint32_t TILE_BYTES = 128 * 128;                    // 128 rows x 64 x 16 bit
constexpr uint32_t SMEM_BYTES = 6 * TILE_BYTES + 64 + 1024;   // Q0, Q1, K (2 boxes), V (2 boxes), barriers, alignment
// SPLIT (operand dtype EOE_F16X2): qkv rows are [q k v | q_lo k_lo v_lo] fp16 pairs and the output rows [hi | lo].  A second
// set of six tiles holds the lo halves; S = Qhi Khi^T + Qlo Khi^T + Qhi Klo^T and O = P Vhi + P Vlo accumulate in the same
// TMEM columns; the probabilities are an fp16 pair as well (P_lo in its own TMEM columns, O = P_hi V_hi + P_hi V_lo +
// P_lo V_hi).  One CTA per SM (193 KB of shared memory, all 512 TMEM columns).
constexpr uint32_t SMEM_BYTES_SPLIT = 12 * TILE_BYTES + 64 + 1024;
// hi / lo fp16 halves of two values scaled by `s` -> one word each
__device__ __forceinline__ void split_scaled(uint32_t a, uint32_t b, float s, uint32_t& hi, uint32_t& lo) {
    gemm::split2(__uint_as_float(a) * s, __uint_as_float(b) * s, hi, lo);
}
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t O_COL = 128;
// SPLIT: the whole TMEM of the SM (one CTA per SM): S / P_hi [0, 208), P_lo [256, 360), O [384, 448)
constexpr uint32_t TMEM_COLS_SPLIT = 512, PLO_COL_SPLIT = 256, O_COL_SPLIT = 384;

__device__ __forceinline__ float fast_exp2(float x) {      // x <= 0 here; MUFU.EX2, flushes denormal results to 0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}


// Softmax of one query row held in TMEM lane `t_lane` (S in fp32 over columns [0, LP)): un-normalised P is written back
// over columns [0, LP/2) as packed 16-bit pairs, the fp32 row sum's reciprocal is returned.
//
// TMEM reads are the attention kernel's bottleneck (64 B/clk/SM: two passes over S cost 3 us per (image, head)), so S is
// read ONCE: the exponent reference starts as (a function of) the maximum of the first 32 keys, held in registers,
// instead of the row maximum -- softmax is invariant to the reference.
//   bf16 probabilities: reference = that maximum + 32 (log2 units).  bf16 / fp32 keep full relative precision at any
//     magnitude; a key may exceed the reference by 2^142 before the (NaN-preserving) clamp at 2^110 engages, i.e. scores
//     more than ~100 above the first 32 keys' maximum -- far beyond what the reference's own fp16 GPU path can represent.
//   fp16 probabilities (largest finite value 65504): reference = that maximum, so the first chunk's P <= 1; a later chunk
//     whose fp32 partial sum reaches 2^15 (which every element >= 2^15 forces) raises the reference to that chunk's maximum,
//     rescales the P already stored in TMEM and the running sum by 2^(old - new) and recomputes the chunk -- the
//     flash-attention rescale, taken only when needed and warp-uniformly (tcgen05.ld / .st are warp-collective; lanes
//     that did not overflow rescale by exactly 1).  Exact for any finite scores; never taken when the row maximum is within
//     ~10 nats of the first 32 keys' (always, at random initialisation).  EOE_F16_ATTN_TWOPASS=1 restores the two-pass form
//     (exact row maximum first) for A/B timing.
#ifndef EOE_F16_ATTN_TWOPASS
#define EOE_F16_ATTN_TWOPASS 0
#endif

// rare path of the fp16 single-pass softmax: P of chunks < c (columns [16 cc, 16 cc + 16)) times f (warp-collective)
// (plo_col != 0: the lo halves of split probabilities, stored plo_col columns further on, are rescaled too; after a rescale
// the stored pairs are accurate to one fp16 rounding only, like the single-fp16 path)
__device__ __noinline__ void softmax_rescale_stored_p(uint32_t t_lane, int c, float f, uint32_t plo_col = 0) {
    const __half2 f2 = __float2half2_rn(f);
#pragma unroll 1
    for (int part = 0; part < (plo_col ? 2 : 1); ++part) {
#pragma unroll 1
        for (int cc = 0; cc < c; ++cc) {
            uint32_t pk[16];
            ptx::tmem_ld_32x32b_x16(t_lane + part * plo_col + cc * 16, pk);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                __half2 h = __hmul2(*reinterpret_cast<__half2*>(&pk[j]), f2);
                pk[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            ptx::tmem_st_32x32b_x16(t_lane + part * plo_col + cc * 16, pk);
        }
    }
    ptx::tmem_st_wait();
}

// PLO != 0 (split operands): P is stored as an fp16 (hi, lo) pair, the lo halves PLO columns behind the hi halves.
template <bool BF16, int L, uint32_t PLO = 0>
__device__ __forceinline__ float softmax_row_tmem(uint32_t t_lane, float sl2) {
    constexpr int LP = (L + 15) / 16 * 16;
    constexpr int NC = LP / 32;                 // full 32-column chunks; LP % 32 == 16 leaves one half chunk
    constexpr bool ONEPASS = BF16 || !EOE_F16_ATTN_TWOPASS;
    static_assert(LP % 32 == 0 || LP % 32 == 16, "chunking");
    static_assert(L >= 32, "first chunk fully valid");
    uint32_t r[2][32];
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
    float ms;
    if (ONEPASS) {
        ptx::tmem_ld_32x32b_x32(t_lane, r[0]);
        ptx::tmem_ld_wait();
        if (NC > 1) ptx::tmem_ld_32x32b_x32(t_lane + 32, r[1]);
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(r[0][j]));
        ms = fmaf(fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])), sl2, BF16 ? 32.0f : 0.0f);
    } else {
        // pass 1: exact row maximum (TMEM loads software pipelined, four independent accumulators)
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        ptx::tmem_ld_32x32b_x32(t_lane, r[0]);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            ptx::tmem_ld_wait();
            if (c + 1 < NC) ptx::tmem_ld_32x32b_x32(t_lane + (c + 1) * 32, r[(c + 1) & 1]);
            else if (LP % 32) ptx::tmem_ld_32x32b_x16(t_lane + NC * 32, reinterpret_cast<uint32_t(&)[16]>(r[(c + 1) & 1]));
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c * 32 + j < L) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(r[c & 1][j]));
        }
        if (LP % 32) {
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (NC * 32 + j < L) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(r[NC & 1][j]));
        }
        ms = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * sl2;
        ptx::tmem_ld_32x32b_x32(t_lane, r[0]);
    }
    // p = exp2(s / 8 * log2 e - ms), fp32 row sum, P -> TMEM as packed 16-bit pairs
    auto prob = [&](uint32_t bits) {
        float e = fmaf(__uint_as_float(bits), sl2, -ms);
        if (BF16) e = e > 110.0f ? 110.0f : e;          // NaN stays NaN (fminf would drop it)
        return fast_exp2(e);
    };
    constexpr bool RESCALE = !BF16 && ONEPASS;          // fp16, single pass: the reference may have to rise
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        uint32_t pk[16], pl[PLO ? 16 : 1];
        if (!(ONEPASS && c == 0)) {                     // single pass: chunk 0 is already in registers, chunk 1 in flight
            ptx::tmem_ld_wait();
            if (c + 1 < NC) ptx::tmem_ld_32x32b_x32(t_lane + (c + 1) * 32, r[(c + 1) & 1]);
            else if (LP % 32) ptx::tmem_ld_32x32b_x16(t_lane + NC * 32, reinterpret_cast<uint32_t(&)[16]>(r[(c + 1) & 1]));
        }
        float c4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            const float p0 = (c * 32 + j < L) ? prob(r[c & 1][j]) : 0.f;
            const float p1 = (c * 32 + j + 1 < L) ? prob(r[c & 1][j + 1]) : 0.f;
            c4[(j >> 1) & 3] += p0 + p1;
            if (PLO) gemm::split2(p0, p1, pk[j >> 1], pl[j >> 1]);
            else pk[j >> 1] = gemm::pack2<BF16>(p0, p1);
        }
        if (RESCALE && c > 0) {
            const bool ovf = (c4[0] + c4[1]) + (c4[2] + c4[3]) >= 32768.0f;
            if (__any_sync(0xffffffffu, ovf)) {
                float m = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c * 32 + j < L) m = fmaxf(m, __uint_as_float(r[c & 1][j]));
                const float new_ms = ovf ? fmaxf(ms, m * sl2) : ms;
                const float f = fast_exp2(ms - new_ms);          // exactly 1 for lanes that keep their reference
                ms = new_ms;
#pragma unroll
                for (int i = 0; i < 4; ++i) s4[i] *= f;
                softmax_rescale_stored_p(t_lane, c, f, PLO);
#pragma unroll
                for (int i = 0; i < 4; ++i) c4[i] = 0.f;
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const float p0 = (c * 32 + j < L) ? prob(r[c & 1][j]) : 0.f;
                    const float p1 = (c * 32 + j + 1 < L) ? prob(r[c & 1][j + 1]) : 0.f;
                    c4[(j >> 1) & 3] += p0 + p1;
                    if (PLO) gemm::split2(p0, p1, pk[j >> 1], pl[j >> 1]);
                    else pk[j >> 1] = gemm::pack2<BF16>(p0, p1);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) s4[i] += c4[i];
        // columns [16c, 16c+16) hold scores consumed in rounds <= c; the in-flight load of round c+1 reads
        // columns >= 32(c+1) > 16c+16, so the store cannot clobber unread scores
        ptx::tmem_st_32x32b_x16(t_lane + c * 16, pk);
        if (PLO) ptx::tmem_st_32x32b_x16(t_lane + PLO + c * 16, reinterpret_cast<const uint32_t(&)[16]>(pl));
    }
    if (LP % 32) {
        uint32_t pk[8], pl[PLO ? 8 : 1];
        ptx::tmem_ld_wait();
        float c4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            const float p0 = (NC * 32 + j < L) ? prob(r[NC & 1][j]) : 0.f;
            const float p1 = (NC * 32 + j + 1 < L) ? prob(r[NC & 1][j + 1]) : 0.f;
            c4[(j >> 1) & 3] += p0 + p1;
            if (PLO) gemm::split2(p0, p1, pk[j >> 1], pl[j >> 1]);
            else pk[j >> 1] = gemm::pack2<BF16>(p0, p1);
        }
        if (RESCALE) {
            const bool ovf = (c4[0] + c4[1]) + (c4[2] + c4[3]) >= 32768.0f;
            if (__any_sync(0xffffffffu, ovf)) {
                float m = -INFINITY;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (NC * 32 + j < L) m = fmaxf(m, __uint_as_float(r[NC & 1][j]));
                const float new_ms = ovf ? fmaxf(ms, m * sl2) : ms;
                const float f = fast_exp2(ms - new_ms);
                ms = new_ms;
#pragma unroll
                for (int i = 0; i < 4; ++i) s4[i] *= f;
                softmax_rescale_stored_p(t_lane, NC, f, PLO);
#pragma unroll
                for (int i = 0; i < 4; ++i) c4[i] = 0.f;
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    const float p0 = (NC * 32 + j < L) ? prob(r[NC & 1][j]) : 0.f;
                    const float p1 = (NC * 32 + j + 1 < L) ? prob(r[NC & 1][j + 1]) : 0.f;
                    c4[(j >> 1) & 3] += p0 + p1;
                    if (PLO) gemm::split2(p0, p1, pk[j >> 1], pl[j >> 1]);
                    else pk[j >> 1] = gemm::pack2<BF16>(p0, p1);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) s4[i] += c4[i];
        ptx::tmem_st_32x32b_x8(t_lane + NC * 16, pk);
        if (PLO) ptx::tmem_st_32x32b_x8(t_lane + PLO + NC * 16, reinterpret_cast<const uint32_t(&)[8]>(pl));
    }
    ptx::tmem_st_wait();
    return 1.0f / ((s4[0] + s4[1]) + (s4[2] + s4[3]));
}

template <bool BF16, int L, bool SPLIT = false>
__global__ void __launch_bounds__(THREADS, SPLIT ? 1 : 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_out128,
                    const __grid_constant__ CUtensorMap tm_out72, int num_items, int heads) {
    constexpr int LP = (L + 15) / 16 * 16;          // keys padded to the UMMA N granularity (208)
    constexpr int KSTEPS = LP / 16;                  // k-steps of the P*V product
    static_assert(L >= 193 && L <= 200, "two 128-row query tiles; the second tile's 72-row store window [L-72, L) must be covered by active warps");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                         // [2][128][128 B]
    uint8_t* sK = smem + 2 * TILE_BYTES;        // [256][128 B]
    uint8_t* sV = smem + 4 * TILE_BYTES;        // [256][128 B]
    // One load barrier per buffer, so that every buffer of the NEXT item is refilled as soon as this item is done with it
    // (K after the last S product, Q0 once its staged output has left, V after the last P.V product, Q1 at the item
    // boundary): the first S product of an item never waits for HBM.  Only the MMA-issuing thread waits on them.
    constexpr uint32_t LO = 6 * TILE_BYTES;     // SPLIT: the lo half of every tile sits one 6-tile set further on
    constexpr uint32_t NH = SPLIT ? 2 : 1;
    uint64_t* bar_q0 = reinterpret_cast<uint64_t*>(smem + NH * 6 * TILE_BYTES);
    uint64_t* bar_q1 = bar_q0 + 1;
    uint64_t* bar_k = bar_q0 + 2;
    uint64_t* bar_v = bar_q0 + 3;
    uint64_t* bar_mma = bar_q0 + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q0 + 5);
    constexpr uint32_t o_col = SPLIT ? O_COL_SPLIT : O_COL;
    constexpr uint32_t tmem_cols = SPLIT ? TMEM_COLS_SPLIT : TMEM_COLS;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int width = heads * 64;
    if (tid == 0) {
        ptx::prefetch_tensormap(&tm_qkv);
        ptx::mbar_init(ptx::smem_u32(bar_q0), 1);
        ptx::mbar_init(ptx::smem_u32(bar_q1), 1);
        ptx::mbar_init(ptx::smem_u32(bar_k), 1);
        ptx::mbar_init(ptx::smem_u32(bar_v), 1);
        ptx::mbar_init(ptx::smem_u32(bar_mma), 1);
        ptx::fence_barrier_init();
    }
    if (warp == 0) {
        ptx::tmem_alloc<1>(ptx::smem_u32(tmem_slot), tmem_cols);
        ptx::tmem_relinquish<1>();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);     // this warp's TMEM lane quarter
    ptx::pdl_wait();                 // the prologue above overlapped the QKV GEMM's tail; qkv is complete from here on
    ptx::pdl_launch_dependents();

    constexpr uint32_t idesc_s = ptx::make_idesc_f16(BF16 ? 1u : 0u, 128, LP);
    constexpr uint32_t idesc_o = ptx::make_idesc_f16(BF16 ? 1u : 0u, 128, 64, 0, 1);     // B = V is MN-major
    const float sl2 = 0.125f * 1.4426950408889634f;                   // 1/sqrt(64) * log2(e)

    uint32_t load_phase = 0, mma_phase = 0;
    // 3-D map [image][token][3*width]: tokens >= L of a box are zero-filled without touching memory, so K / V rows
    // L..255 cost no traffic and meet P == 0 with finite values
    auto load_q = [&](int tile, int bb, int hh, int row0) {
        const uint32_t lb = ptx::smem_u32(tile ? bar_q1 : bar_q0);
        ptx::mbar_arrive_expect_tx(lb, NH * TILE_BYTES);
        ptx::tma_load_3d(ptx::smem_u32(sQ + tile * TILE_BYTES), &tm_qkv, lb, hh * 64, row0, bb);
        if (SPLIT) ptx::tma_load_3d(ptx::smem_u32(sQ + LO + tile * TILE_BYTES), &tm_qkv, lb, 3 * width + hh * 64, row0, bb);
    };
    auto load_kv = [&](int which, int bb, int hh) {         // which: 0 = K, 1 = V
        const uint32_t lb = ptx::smem_u32(which ? bar_v : bar_k);
        uint8_t* dst = which ? sV : sK;
        ptx::mbar_arrive_expect_tx(lb, NH * 2 * TILE_BYTES);
        ptx::tma_load_3d(ptx::smem_u32(dst), &tm_qkv, lb, (1 + which) * width + hh * 64, 0, bb);
        ptx::tma_load_3d(ptx::smem_u32(dst + TILE_BYTES), &tm_qkv, lb, (1 + which) * width + hh * 64, 128, bb);
        if (SPLIT) {
            ptx::tma_load_3d(ptx::smem_u32(dst + LO), &tm_qkv, lb, (4 + which) * width + hh * 64, 0, bb);
            ptx::tma_load_3d(ptx::smem_u32(dst + LO + TILE_BYTES), &tm_qkv, lb, (4 + which) * width + hh * 64, 128, bb);
        }
    };
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int b = item / heads, h = item % heads;
        const int t1_start = (it & 1) ? (L - 128) : 128;              // first query row of the second tile
        const int next = item + gridDim.x;
        const bool has_next = next < num_items;
        const int nb = next / heads, nh = next % heads;
        if (tid == 0) {
            ptx::bulk_wait_group_read0();        // the previous item's output tiles (staged in the Q buffers) have left
            if (it == 0) {                       // later items find Q0, K and V already requested (see below)
                load_q(0, b, h, 0);
                load_kv(0, b, h);
                load_kv(1, b, h);
            }
            load_q(1, b, h, t1_start);
        }
#pragma unroll 1
        for (int tile = 0; tile < 2; ++tile) {
            if (tid == 0) {
                ptx::mbar_wait(ptx::smem_u32(tile ? bar_q1 : bar_q0), load_phase);
                if (tile == 0) ptx::mbar_wait(ptx::smem_u32(bar_k), load_phase);
                ptx::tc_fence_after();
                const uint64_t a_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sQ + tile * TILE_BYTES));
                const uint64_t b_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sK));
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::umma_f16<1>(tmem, a_desc + 2 * k, b_desc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                if (SPLIT) {
                    const uint64_t a_lo = ptx::make_smem_desc_sw128(ptx::smem_u32(sQ + LO + tile * TILE_BYTES));
                    const uint64_t b_lo = ptx::make_smem_desc_sw128(ptx::smem_u32(sK + LO));
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::umma_f16<1>(tmem, a_lo + 2 * k, b_desc + 2 * k, idesc_s, 1u);
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::umma_f16<1>(tmem, a_desc + 2 * k, b_lo + 2 * k, idesc_s, 1u);
                }
                ptx::umma_commit(ptx::smem_u32(bar_mma));
            }
            // query row owned by this thread, and whether its warp has any row to compute
            const int wfirst = (tile == 0 ? 0 : t1_start) + warp * 32;
            const bool warp_active = (tile == 0) ? true : (wfirst + 31 >= 128 && wfirst < L);
            ptx::mbar_wait(ptx::smem_u32(bar_mma), mma_phase);
            mma_phase ^= 1;
            ptx::tc_fence_after();
            if (tile == 1 && tid == 0 && has_next) {
                // the last S product of this item is done: K is free, and tile 0's staged output has long left Q0
                ptx::bulk_wait_group_read0();
                load_q(0, nb, nh, 0);
                load_kv(0, nb, nh);
            }
            float inv_sum = 0.f;
            if (warp_active) {
                inv_sum = softmax_row_tmem<BF16, L, SPLIT ? PLO_COL_SPLIT : 0>(t_lane, sl2);
            }
            ptx::tc_fence_before();
            __syncthreads();                       // P of all rows is in TMEM (and the V tail is zeroed)
            if (tid == 0) {
                if (tile == 0) ptx::mbar_wait(ptx::smem_u32(bar_v), load_phase);
                ptx::tc_fence_after();
                const uint64_t v_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sV));
#pragma unroll
                for (int kk = 0; kk < KSTEPS; ++kk)    // 16 keys per step: 8 packed columns of P, 16 rows (2 KB) of V
                    ptx::umma_f16_ts(tmem + o_col, tmem + kk * 8, v_desc + (uint64_t)(kk * 2048 >> 4), idesc_o, kk != 0 ? 1u : 0u);
                if (SPLIT) {                           // + P_hi V_lo + P_lo V_hi
                    const uint64_t v_lo = ptx::make_smem_desc_sw128(ptx::smem_u32(sV + LO));
#pragma unroll
                    for (int kk = 0; kk < KSTEPS; ++kk)
                        ptx::umma_f16_ts(tmem + o_col, tmem + kk * 8, v_lo + (uint64_t)(kk * 2048 >> 4), idesc_o, 1u);
#pragma unroll
                    for (int kk = 0; kk < KSTEPS; ++kk)
                        ptx::umma_f16_ts(tmem + o_col, tmem + PLO_COL_SPLIT + kk * 8, v_desc + (uint64_t)(kk * 2048 >> 4), idesc_o, 1u);
                }
                ptx::umma_commit(ptx::smem_u32(bar_mma));
            }
            ptx::mbar_wait(ptx::smem_u32(bar_mma), mma_phase);
            mma_phase ^= 1;
            ptx::tc_fence_after();
            if (tile == 1 && tid == 0 && has_next) load_kv(1, nb, nh);      // the last P.V product is done: V is free
            // Epilogue: O * (1 / row sum) -> 16 bit, staged row-major in the (now dead) Q tile with the 128-byte swizzle
            // and written by ONE TMA tile store: full 128-byte lines instead of 32 half-sector writes per instruction.
            // The output map is 3-D [image][token][width], so rows past the image's last token are clipped by the TMA unit.
            uint8_t* stg = sQ + tile * TILE_BYTES;
            if (warp_active) {
                uint32_t o0[32], o1[32];
                ptx::tmem_ld_32x32b_x32(t_lane + o_col, o0);
                ptx::tmem_ld_32x32b_x32(t_lane + o_col + 32, o1);
                ptx::tmem_ld_wait();
                const uint32_t srow = ptx::smem_u32(stg + tid * 128);
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const uint32_t (&oo)[32] = hh ? o1 : o0;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        const uint32_t off = ((hh * 4 + (j >> 3)) ^ (tid & 7)) << 4;
                        if (SPLIT) {
                            uint32_t qh[4], ql[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) split_scaled(oo[j + 2 * e], oo[j + 2 * e + 1], inv_sum, qh[e], ql[e]);
                            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(srow + off), "r"(qh[0]), "r"(qh[1]), "r"(qh[2]), "r"(qh[3]) : "memory");
                            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(srow + LO + off), "r"(ql[0]), "r"(ql[1]), "r"(ql[2]), "r"(ql[3]) : "memory");
                        } else {
                            const uint32_t q0 = gemm::pack2<BF16>(__uint_as_float(oo[j]) * inv_sum, __uint_as_float(oo[j + 1]) * inv_sum);
                            const uint32_t q1 = gemm::pack2<BF16>(__uint_as_float(oo[j + 2]) * inv_sum, __uint_as_float(oo[j + 3]) * inv_sum);
                            const uint32_t q2 = gemm::pack2<BF16>(__uint_as_float(oo[j + 4]) * inv_sum, __uint_as_float(oo[j + 5]) * inv_sum);
                            const uint32_t q3 = gemm::pack2<BF16>(__uint_as_float(oo[j + 6]) * inv_sum, __uint_as_float(oo[j + 7]) * inv_sum);
                            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(srow + off), "r"(q0), "r"(q1), "r"(q2), "r"(q3) : "memory");
                        }
                    }
                }
            }
            ptx::fence_proxy_async_smem();
            ptx::tc_fence_before();
            __syncthreads();                       // O has been read and staged: TMEM is free
            if (tid == 0) {
#pragma unroll
                for (uint32_t part = 0; part < NH; ++part) {     // SPLIT: the lo tile goes to columns [width, 2 width) of the output rows
                    const uint32_t src = ptx::smem_u32(stg + part * LO);
                    const int col = (int)part * width + h * 64;
                    if (tile == 0) {
                        ptx::tma_store_3d(&tm_out128, src, col, 0, b);
                    } else if (t1_start == 128) {      // rows [128, 200): rows >= L are clipped
                        ptx::tma_store_3d(&tm_out72, src, col, 128, b);
                    } else {                           // tile rows [L-128, L): store from staging row 56 = token L-72 (8-row aligned),
                        ptx::tma_store_3d(&tm_out72, src + 56 * 128, col, L - 72, b);   // tokens < 128 repeat tile 0's values
                    }
                }
                ptx::bulk_commit_group();
            }
        }
        load_phase ^= 1;                           // each of the four load barriers completed exactly one phase for this item
    }
    if (tid == 0) ptx::bulk_wait_group_read0();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem, tmem_cols);
    }
}


// ------------------------------------------------------------------------------------------------------------------------
// tcgen05 attention for short sequences (L <= 64: ViT-B/32 has 50 tokens).  One 128-row UMMA tile holds TWO (image, head)
// items of the same image (heads h and h + 1: rows 0..63 and 64..127); their keys are concatenated to 128 "keys", so one
//   S[128 x 128] = [Q_h; Q_h+1] [K_h; K_h+1]^T
// produces both score blocks on the diagonal (the off-diagonal blocks are computed and ignored: tensor work is not what
// bounds this kernel), and one
//   O[128 x 64] = P[128 x 128] [V_h; V_h+1]
// with P zero off the diagonal gives both outputs.  A thread owns one query row and only ever reads its own block's 64
// score columns (two tcgen05.ld), so the whole row lives in registers: exact row maximum, one pass.  TMEM plan (128 columns,
// four CTAs per SM): S [0, 128) -> P packed over [0, 64) (own block's 32 columns, zeros in the other block's 32),
// O over [64, 128) once every row has been read.  Loads and stores go through 3-D [image][token][column] maps with
// 64-token boxes: tokens >= L are zero-filled on load and clipped on store.
namespace tc64 {
constexpr int THREADS = 128;
constexpr uint32_t SMEM_BYTES = 3 * TILE_BYTES + 64 + 1024;     // Q, K, V tiles of 128 rows x 128 B, barriers, alignment
constexpr uint32_t SMEM_BYTES_SPLIT = 6 * TILE_BYTES + 64 + 1024;   // + the lo halves (attn::SMEM_BYTES_SPLIT); two CTAs per SM
constexpr uint32_t TMEM_COLS_SPLIT = 256, PLO_COL_SPLIT = 128;       // SPLIT: P_lo packed over [128, 192), same block structure
constexpr uint32_t TMEM_COLS = 128;
constexpr uint32_t O_COL = 64;
}  // namespace tc64

template <bool BF16, bool SPLIT = false>
__global__ void __launch_bounds__(tc64::THREADS, SPLIT ? 2 : 4)
attention_tc64_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_out, int num_pairs,
                      int heads, int L) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                         // [128][128 B]: rows 0..63 head h, rows 64..127 head h + 1
    uint8_t* sK = smem + TILE_BYTES;
    uint8_t* sV = smem + 2 * TILE_BYTES;
    constexpr uint32_t LO = 3 * TILE_BYTES, NH = SPLIT ? 2 : 1;
    uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + NH * 3 * TILE_BYTES);
    uint64_t* bar_mma = bar_load + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 2);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int width = heads * 64;
    const int hp = heads >> 1;                  // head pairs per image
    if (tid == 0) {
        ptx::prefetch_tensormap(&tm_qkv);
        ptx::prefetch_tensormap(&tm_out);
        ptx::mbar_init(ptx::smem_u32(bar_load), 1);
        ptx::mbar_init(ptx::smem_u32(bar_mma), 1);
        ptx::fence_barrier_init();
    }
    if (warp == 0) {
        ptx::tmem_alloc<1>(ptx::smem_u32(tmem_slot), SPLIT ? tc64::TMEM_COLS_SPLIT : tc64::TMEM_COLS);
        ptx::tmem_relinquish<1>();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t own = (uint32_t)(warp >> 1) * 64;          // first score column of this row's own block
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();

    constexpr uint32_t idesc_s = ptx::make_idesc_f16(BF16 ? 1u : 0u, 128, 128);
    constexpr uint32_t idesc_o = ptx::make_idesc_f16(BF16 ? 1u : 0u, 128, 64, 0, 1);     // B = V is MN-major
    const float sl2 = 0.125f * 1.4426950408889634f;                   // 1/sqrt(64) * log2(e)

    uint32_t load_phase = 0, mma_phase = 0;
    for (int pair = blockIdx.x; pair < num_pairs; pair += gridDim.x) {
        const int b = pair / hp, h = 2 * (pair % hp);
        if (tid == 0) {
            ptx::bulk_wait_group_read0();        // the previous pair's output (staged in the Q tile) has left
            const uint32_t lb = ptx::smem_u32(bar_load);
            ptx::mbar_arrive_expect_tx(lb, NH * 3 * TILE_BYTES);
#pragma unroll
            for (int i = 0; i < 2; ++i) {        // 64-token boxes: tokens >= L arrive as zeros
                ptx::tma_load_3d(ptx::smem_u32(sQ + i * (TILE_BYTES / 2)), &tm_qkv, lb, (h + i) * 64, 0, b);
                ptx::tma_load_3d(ptx::smem_u32(sK + i * (TILE_BYTES / 2)), &tm_qkv, lb, width + (h + i) * 64, 0, b);
                ptx::tma_load_3d(ptx::smem_u32(sV + i * (TILE_BYTES / 2)), &tm_qkv, lb, 2 * width + (h + i) * 64, 0, b);
                if (SPLIT) {
                    ptx::tma_load_3d(ptx::smem_u32(sQ + LO + i * (TILE_BYTES / 2)), &tm_qkv, lb, 3 * width + (h + i) * 64, 0, b);
                    ptx::tma_load_3d(ptx::smem_u32(sK + LO + i * (TILE_BYTES / 2)), &tm_qkv, lb, 4 * width + (h + i) * 64, 0, b);
                    ptx::tma_load_3d(ptx::smem_u32(sV + LO + i * (TILE_BYTES / 2)), &tm_qkv, lb, 5 * width + (h + i) * 64, 0, b);
                }
            }
        }
        ptx::mbar_wait(ptx::smem_u32(bar_load), load_phase);
        load_phase ^= 1;
        if (tid == 0) {
            ptx::tc_fence_after();
            const uint64_t a_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sQ));
            const uint64_t b_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sK));
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::umma_f16<1>(tmem, a_desc + 2 * k, b_desc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
            if (SPLIT) {
                const uint64_t a_lo = ptx::make_smem_desc_sw128(ptx::smem_u32(sQ + LO));
                const uint64_t b_lo = ptx::make_smem_desc_sw128(ptx::smem_u32(sK + LO));
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::umma_f16<1>(tmem, a_lo + 2 * k, b_desc + 2 * k, idesc_s, 1u);
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::umma_f16<1>(tmem, a_desc + 2 * k, b_lo + 2 * k, idesc_s, 1u);
            }
            ptx::umma_commit(ptx::smem_u32(bar_mma));
        }
        ptx::mbar_wait(ptx::smem_u32(bar_mma), mma_phase);
        mma_phase ^= 1;
        ptx::tc_fence_after();
        // softmax of this thread's row over its own block's keys: the 64 scores fit in registers
        float inv_sum;
        {
            uint32_t r0[32], r1[32];
            ptx::tmem_ld_32x32b_x32(t_lane + own, r0);
            ptx::tmem_ld_32x32b_x32(t_lane + own + 32, r1);
            ptx::tmem_ld_wait();
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (j < L) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(r0[j]));
                if (32 + j < L) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(r1[j]));
            }
            const float ms = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * sl2;
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t pk[32], pl[SPLIT ? 32 : 1];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float p0 = (j < L) ? fast_exp2(fmaf(__uint_as_float(r0[j]), sl2, -ms)) : 0.f;
                const float p1 = (j + 1 < L) ? fast_exp2(fmaf(__uint_as_float(r0[j + 1]), sl2, -ms)) : 0.f;
                const float p2 = (32 + j < L) ? fast_exp2(fmaf(__uint_as_float(r1[j]), sl2, -ms)) : 0.f;
                const float p3 = (33 + j < L) ? fast_exp2(fmaf(__uint_as_float(r1[j + 1]), sl2, -ms)) : 0.f;
                s4[(j >> 1) & 3] += (p0 + p1) + (p2 + p3);
                if (SPLIT) {
                    gemm::split2(p0, p1, pk[j >> 1], pl[j >> 1]);
                    gemm::split2(p2, p3, pk[16 + (j >> 1)], pl[(SPLIT ? 16 : 0) + (j >> 1)]);
                } else {
                    pk[j >> 1] = gemm::pack2<BF16>(p0, p1);
                    pk[16 + (j >> 1)] = gemm::pack2<BF16>(p2, p3);
                }
            }
            inv_sum = 1.0f / ((s4[0] + s4[1]) + (s4[2] + s4[3]));
            // P over columns [0, 64): keys 2c, 2c+1 in column c; this row's block at [own / 2, +32), zeros in the other block
            uint32_t zero[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) zero[j] = 0u;
            const uint32_t pcol = own >> 1, zcol = 32 - pcol;
            ptx::tmem_st_32x32b_x16(t_lane + pcol, reinterpret_cast<const uint32_t(&)[16]>(pk[0]));
            ptx::tmem_st_32x32b_x16(t_lane + pcol + 16, reinterpret_cast<const uint32_t(&)[16]>(pk[16]));
            ptx::tmem_st_32x32b_x16(t_lane + zcol, zero);
            ptx::tmem_st_32x32b_x16(t_lane + zcol + 16, zero);
            if (SPLIT) {                           // the lo halves of P: same block structure, PLO_COL_SPLIT columns further on
                constexpr uint32_t PL = tc64::PLO_COL_SPLIT;
                ptx::tmem_st_32x32b_x16(t_lane + PL + pcol, reinterpret_cast<const uint32_t(&)[16]>(pl[0]));
                ptx::tmem_st_32x32b_x16(t_lane + PL + pcol + 16, reinterpret_cast<const uint32_t(&)[16]>(pl[SPLIT ? 16 : 0]));
                ptx::tmem_st_32x32b_x16(t_lane + PL + zcol, zero);
                ptx::tmem_st_32x32b_x16(t_lane + PL + zcol + 16, zero);
            }
            ptx::tmem_st_wait();
        }
        ptx::tc_fence_before();
        __syncthreads();                           // P of all rows is in TMEM, every S column has been read
        if (tid == 0) {
            ptx::tc_fence_after();
            const uint64_t v_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sV));
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)         // 16 keys per step: 8 packed columns of P, 16 rows (2 KB) of V
                ptx::umma_f16_ts(tmem + tc64::O_COL, tmem + kk * 8, v_desc + (uint64_t)(kk * 2048 >> 4), idesc_o, kk != 0 ? 1u : 0u);
            if (SPLIT) {
                const uint64_t v_lo = ptx::make_smem_desc_sw128(ptx::smem_u32(sV + LO));
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                    ptx::umma_f16_ts(tmem + tc64::O_COL, tmem + kk * 8, v_lo + (uint64_t)(kk * 2048 >> 4), idesc_o, 1u);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)     // + P_lo V_hi
                    ptx::umma_f16_ts(tmem + tc64::O_COL, tmem + tc64::PLO_COL_SPLIT + kk * 8, v_desc + (uint64_t)(kk * 2048 >> 4), idesc_o, 1u);
            }
            ptx::umma_commit(ptx::smem_u32(bar_mma));
        }
        ptx::mbar_wait(ptx::smem_u32(bar_mma), mma_phase);
        mma_phase ^= 1;
        ptx::tc_fence_after();
        // O * (1 / row sum) -> 16 bit -> the dead Q tile (128-byte swizzle) -> two TMA tile stores (rows >= L clipped)
        {
            uint32_t o0[32], o1[32];
            ptx::tmem_ld_32x32b_x32(t_lane + tc64::O_COL, o0);
            ptx::tmem_ld_32x32b_x32(t_lane + tc64::O_COL + 32, o1);
            ptx::tmem_ld_wait();
            const uint32_t srow = ptx::smem_u32(sQ + tid * 128);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const uint32_t (&oo)[32] = hh ? o1 : o0;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    const uint32_t off = ((hh * 4 + (j >> 3)) ^ (tid & 7)) << 4;
                    if (SPLIT) {
                        uint32_t qh[4], ql[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) split_scaled(oo[j + 2 * e], oo[j + 2 * e + 1], inv_sum, qh[e], ql[e]);
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(srow + off), "r"(qh[0]), "r"(qh[1]), "r"(qh[2]), "r"(qh[3]) : "memory");
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(srow + LO + off), "r"(ql[0]), "r"(ql[1]), "r"(ql[2]), "r"(ql[3]) : "memory");
                    } else {
                        const uint32_t q0 = gemm::pack2<BF16>(__uint_as_float(oo[j]) * inv_sum, __uint_as_float(oo[j + 1]) * inv_sum);
                        const uint32_t q1 = gemm::pack2<BF16>(__uint_as_float(oo[j + 2]) * inv_sum, __uint_as_float(oo[j + 3]) * inv_sum);
                        const uint32_t q2 = gemm::pack2<BF16>(__uint_as_float(oo[j + 4]) * inv_sum, __uint_as_float(oo[j + 5]) * inv_sum);
                        const uint32_t q3 = gemm::pack2<BF16>(__uint_as_float(oo[j + 6]) * inv_sum, __uint_as_float(oo[j + 7]) * inv_sum);
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(srow + off), "r"(q0), "r"(q1), "r"(q2), "r"(q3) : "memory");
                    }
                }
            }
        }
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        __syncthreads();                           // O has been read and staged: TMEM and the K / V tiles are free
        if (tid == 0) {
            ptx::tma_store_3d(&tm_out, ptx::smem_u32(sQ), h * 64, 0, b);
            ptx::tma_store_3d(&tm_out, ptx::smem_u32(sQ + TILE_BYTES / 2), (h + 1) * 64, 0, b);
            if (SPLIT) {
                ptx::tma_store_3d(&tm_out, ptx::smem_u32(sQ + LO), width + h * 64, 0, b);
                ptx::tma_store_3d(&tm_out, ptx::smem_u32(sQ + LO + TILE_BYTES / 2), width + (h + 1) * 64, 0, b);
            }
            ptx::bulk_commit_group();
        }
    }
    if (tid == 0) ptx::bulk_wait_group_read0();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem, SPLIT ? tc64::TMEM_COLS_SPLIT : tc64::TMEM_COLS);
    }
}

}  // namespace attn
}  // namespace eoe
