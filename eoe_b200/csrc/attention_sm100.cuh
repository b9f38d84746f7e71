// tcgen05 attention for the ViT-B/16 token count (L = 197): softmax(Q K^T / 8) V per (image, head)
// (reference: nn.MultiheadAttention inside ResidualAttentionBlock.attention, clip_official/clip/model.py:181-183).
//
// One CTA of 128 threads per (image, head) item, persistent over items, 2 CTAs per SM (their TMA / MMA / softmax phases
// interleave).  Per item: TMA brings the head's Q (2 tiles of 128 query rows), K and V slices ([rows][64] bf16, 128-byte
// rows, SWIZZLE_128B) straight out of the qkv GEMM output.  Per query tile:
//   S[128 x 208] = Q K^T        tcgen05.mma, both operands K-major from smem, fp32 in TMEM columns [0, 208)
//   softmax                      thread t owns TMEM lane t = query row t: row max and exp2 need no cross-thread traffic;
//                                un-normalised P is written back to TMEM as packed 16-bit pairs over columns [0, 104)
//   O[128 x 64] = P V           tcgen05.mma with A = P from TMEM and B = V as an MN-major smem operand (no transpose),
//                                fp32 in TMEM columns [128, 192)
//   epilogue                     O * (1 / row sum) -> 16-bit -> global
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
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t O_COL = 128;

__device__ __forceinline__ float fast_exp2(float x) {      // x <= 0 here; MUFU.EX2, flushes denormal results to 0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool BF16, int L>
__global__ void __launch_bounds__(THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, uint16_t* __restrict__ out, int num_items, int heads) {
    constexpr int LP = (L + 15) / 16 * 16;          // keys padded to the UMMA N granularity (208)
    constexpr int KSTEPS = LP / 16;                  // k-steps of the P*V product
    static_assert(L > 128 && L <= 208, "two 128-row query tiles");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;                         // [2][128][128 B]
    uint8_t* sK = smem + 2 * TILE_BYTES;        // [256][128 B]
    uint8_t* sV = smem + 4 * TILE_BYTES;        // [256][128 B]
    uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + 6 * TILE_BYTES);
    uint64_t* bar_mma = bar_load + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 2);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int width = heads * 64;
    if (tid == 0) {
        ptx::prefetch_tensormap(&tm_qkv);
        ptx::mbar_init(ptx::smem_u32(bar_load), 1);
        ptx::mbar_init(ptx::smem_u32(bar_mma), 1);
        ptx::fence_barrier_init();
    }
    if (warp == 0) {
        ptx::tmem_alloc<1>(ptx::smem_u32(tmem_slot), TMEM_COLS);
        ptx::tmem_relinquish<1>();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);     // this warp's TMEM lane quarter

    constexpr uint32_t idesc_s = ptx::make_idesc_f16(BF16 ? 1u : 0u, 128, LP);
    constexpr uint32_t idesc_o = ptx::make_idesc_f16(BF16 ? 1u : 0u, 128, 64, 0, 1);     // B = V is MN-major
    const float sl2 = 0.125f * 1.4426950408889634f;                   // 1/sqrt(64) * log2(e)

    uint32_t load_phase = 0, mma_phase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int b = item / heads, h = item % heads;
        const int t1_start = (it & 1) ? (L - 128) : 128;              // first query row of the second tile
        if (tid == 0) {
            const uint32_t lb = ptx::smem_u32(bar_load);
            ptx::mbar_arrive_expect_tx(lb, 6 * TILE_BYTES);
            const int r0 = b * L;
            ptx::tma_load_2d(ptx::smem_u32(sQ), &tm_qkv, lb, h * 64, r0);
            ptx::tma_load_2d(ptx::smem_u32(sQ + TILE_BYTES), &tm_qkv, lb, h * 64, r0 + t1_start);
            ptx::tma_load_2d(ptx::smem_u32(sK), &tm_qkv, lb, width + h * 64, r0);
            ptx::tma_load_2d(ptx::smem_u32(sK + TILE_BYTES), &tm_qkv, lb, width + h * 64, r0 + 128);
            ptx::tma_load_2d(ptx::smem_u32(sV), &tm_qkv, lb, 2 * width + h * 64, r0);
            ptx::tma_load_2d(ptx::smem_u32(sV + TILE_BYTES), &tm_qkv, lb, 2 * width + h * 64, r0 + 128);
        }
        ptx::mbar_wait(ptx::smem_u32(bar_load), load_phase);
        load_phase ^= 1;
        // rows L..LP-1 of V belong to the next image (or to slack): they meet P == 0, but 0 * NaN would poison O
        if (tid < (LP - L) * 8) {
            const int r = L + (tid >> 3), c = tid & 7;
            asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(ptx::smem_u32(sV + r * 128 + ((c ^ (r & 7)) << 4))), "r"(0u) : "memory");
        }
        ptx::fence_proxy_async_smem();

#pragma unroll 1
        for (int tile = 0; tile < 2; ++tile) {
            if (tid == 0) {
                ptx::tc_fence_after();
                const uint64_t a_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sQ + tile * TILE_BYTES));
                const uint64_t b_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sK));
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::umma_f16<1>(tmem, a_desc + 2 * k, b_desc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                ptx::umma_commit(ptx::smem_u32(bar_mma));
            }
            // query row owned by this thread, and whether its warp has any row to compute
            const int qrow = (tile == 0 ? 0 : t1_start) + tid;
            const bool own = (tile == 0) ? true : (qrow >= 128 && qrow < L);
            const int wfirst = (tile == 0 ? 0 : t1_start) + warp * 32;
            const bool warp_active = (tile == 0) ? true : (wfirst + 31 >= 128 && wfirst < L);
            ptx::mbar_wait(ptx::smem_u32(bar_mma), mma_phase);
            mma_phase ^= 1;
            ptx::tc_fence_after();
            float inv_sum = 0.f;
            if (warp_active) {
                constexpr int NC = LP / 32;                 // full 32-column chunks; LP % 32 == 16 leaves one half chunk
                static_assert(LP % 32 == 0 || LP % 32 == 16, "chunking");
                // pass 1: row maximum over the L valid keys.  TMEM loads are software pipelined (chunk c+1 is in
                // flight while chunk c is reduced) and four independent accumulators break the FMNMX dependency chain.
                float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                uint32_t r[2][32];
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
                const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
                // pass 2: p = exp2((s - m) / 8 * log2 e), fp32 row sum, P -> TMEM as packed 16-bit pairs
                const float ms = m * sl2;
                float s4[4] = {0.f, 0.f, 0.f, 0.f};
                ptx::tmem_ld_32x32b_x32(t_lane, r[0]);
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    uint32_t pk[16];
                    ptx::tmem_ld_wait();
                    if (c + 1 < NC) ptx::tmem_ld_32x32b_x32(t_lane + (c + 1) * 32, r[(c + 1) & 1]);
                    else if (LP % 32) ptx::tmem_ld_32x32b_x16(t_lane + NC * 32, reinterpret_cast<uint32_t(&)[16]>(r[(c + 1) & 1]));
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const float p0 = (c * 32 + j < L) ? fast_exp2(fmaf(__uint_as_float(r[c & 1][j]), sl2, -ms)) : 0.f;
                        const float p1 = (c * 32 + j + 1 < L) ? fast_exp2(fmaf(__uint_as_float(r[c & 1][j + 1]), sl2, -ms)) : 0.f;
                        s4[(j >> 1) & 3] += p0 + p1;
                        pk[j >> 1] = gemm::pack2<BF16>(p0, p1);
                    }
                    // columns [16c, 16c+16) hold scores consumed in rounds <= c; the in-flight load of round c+1 reads
                    // columns >= 32(c+1) > 16c+16, so the store cannot clobber unread scores
                    ptx::tmem_st_32x32b_x16(t_lane + c * 16, pk);
                }
                if (LP % 32) {
                    uint32_t pk[8];
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        const float p0 = (NC * 32 + j < L) ? fast_exp2(fmaf(__uint_as_float(r[NC & 1][j]), sl2, -ms)) : 0.f;
                        const float p1 = (NC * 32 + j + 1 < L) ? fast_exp2(fmaf(__uint_as_float(r[NC & 1][j + 1]), sl2, -ms)) : 0.f;
                        s4[(j >> 1) & 3] += p0 + p1;
                        pk[j >> 1] = gemm::pack2<BF16>(p0, p1);
                    }
                    ptx::tmem_st_32x32b_x8(t_lane + NC * 16, pk);
                }
                ptx::tmem_st_wait();
                inv_sum = 1.0f / ((s4[0] + s4[1]) + (s4[2] + s4[3]));
            }
            ptx::tc_fence_before();
            __syncthreads();                       // P of all rows is in TMEM (and the V tail is zeroed)
            if (tid == 0) {
                ptx::tc_fence_after();
                const uint64_t v_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sV));
#pragma unroll
                for (int kk = 0; kk < KSTEPS; ++kk)    // 16 keys per step: 8 packed columns of P, 16 rows (2 KB) of V
                    ptx::umma_f16_ts(tmem + O_COL, tmem + kk * 8, v_desc + (uint64_t)(kk * 2048 >> 4), idesc_o, kk != 0 ? 1u : 0u);
                ptx::umma_commit(ptx::smem_u32(bar_mma));
            }
            ptx::mbar_wait(ptx::smem_u32(bar_mma), mma_phase);
            mma_phase ^= 1;
            ptx::tc_fence_after();
            if (warp_active) {
                uint32_t o0[32], o1[32];
                ptx::tmem_ld_32x32b_x32(t_lane + O_COL, o0);
                ptx::tmem_ld_32x32b_x32(t_lane + O_COL + 32, o1);
                ptx::tmem_ld_wait();
                if (own) {
                    uint16_t* dst = out + ((int64_t)b * L + qrow) * width + h * 64;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint4 q;
                        q.x = gemm::pack2<BF16>(__uint_as_float(o0[j]) * inv_sum, __uint_as_float(o0[j + 1]) * inv_sum);
                        q.y = gemm::pack2<BF16>(__uint_as_float(o0[j + 2]) * inv_sum, __uint_as_float(o0[j + 3]) * inv_sum);
                        q.z = gemm::pack2<BF16>(__uint_as_float(o0[j + 4]) * inv_sum, __uint_as_float(o0[j + 5]) * inv_sum);
                        q.w = gemm::pack2<BF16>(__uint_as_float(o0[j + 6]) * inv_sum, __uint_as_float(o0[j + 7]) * inv_sum);
                        *reinterpret_cast<uint4*>(dst + j) = q;
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint4 q;
                        q.x = gemm::pack2<BF16>(__uint_as_float(o1[j]) * inv_sum, __uint_as_float(o1[j + 1]) * inv_sum);
                        q.y = gemm::pack2<BF16>(__uint_as_float(o1[j + 2]) * inv_sum, __uint_as_float(o1[j + 3]) * inv_sum);
                        q.z = gemm::pack2<BF16>(__uint_as_float(o1[j + 4]) * inv_sum, __uint_as_float(o1[j + 5]) * inv_sum);
                        q.w = gemm::pack2<BF16>(__uint_as_float(o1[j + 6]) * inv_sum, __uint_as_float(o1[j + 7]) * inv_sum);
                        *reinterpret_cast<uint4*>(dst + 32 + j) = q;
                    }
                }
            }
            ptx::tc_fence_before();
            __syncthreads();                       // O has been read: TMEM and (after tile 1) the smem tiles are free
        }
    }
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem, TMEM_COLS);
    }
}

}  // namespace attn
}  // namespace eoe
