// tcgen05 / TMEM / TMA GEMM for the CLIP ViT encoder (sm_100a):   out = epilogue( A[M,K] * W[N,K]^T )
//
// Replaces the 48 cuBLAS addmm + conv1 of VisualTransformer.forward (reference
// src/eoe/models/clip_official/clip/model.py:171,174-176,220).  Both operands are K-major 16-bit (activations
// [tokens, K] and nn.Linear weights [N, K] are already in that layout), accumulation is fp32 in TMEM.
//
// A CTA PAIR (cluster of 2, `cta_group::2`) owns a 256 x 256 output tile: CTA r holds rows [r*128, +128) of A and rows
// [r*128, +128) of the W tile, so every operand byte is fetched from L2 once per pair and the tensor cores of the two
// SMs read each other's B half.  Persistent over tiles (n fastest: the pairs running concurrently share A rows in L2,
// W <= 4.7 MB stays L2 resident).  Warp roles per CTA:
//   warp 0     TMA producer: 5-stage ring of {A 128x64, B 128x64} tiles (SWIZZLE_128B); completion bytes of both CTAs
//              are signalled on the LEADER CTA's `full` barrier
//   warp 1     MMA issuer (leader CTA only): one thread issues tcgen05.mma.cta_group::2 256x256x16; tcgen05.commit
//              multicasts to both CTAs' `empty` / `tmem_full` barriers
//   warp 2     TMEM allocator (512 columns = 2 accumulator stages of 256 fp32 columns)
//   warps 4-11 epilogue: warp (q, half) reads TMEM lanes [32q, 32q+32) x columns [128*half, +128) with tcgen05.ld,
//              applies bias / QuickGELU in the row-owner layout, transposes 32x32 blocks through padded shared memory and
//              issues fully coalesced 128-byte global accesses (stores, or the fp32 residual read-modify-write);
//              overlapped with the next tile's main loop through the second accumulator stage
#pragma once
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace eoe {
namespace gemm {

constexpr int BM = 256, BN = 256, BK = 64, UMMA_K = 16;   // tile of a CTA pair
constexpr int CTA_M = 128, CTA_NB = 128;                   // rows of A / rows of W each CTA loads
constexpr int STAGES = 5;
constexpr uint32_t A_BYTES = CTA_M * BK * 2, B_BYTES = CTA_NB * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;   // per CTA
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 128 + EPI_WARPS * 32;
constexpr uint32_t TMEM_COLS = 512;
constexpr int STAGE_LD = 36;                                // row stride (words) of the transpose buffers: 16-byte aligned rows,
                                                            // conflict-free for 128-bit accesses by row owners and by 8-lane row readers
constexpr size_t EPI_STAGE_BYTES = (size_t)EPI_WARPS * 32 * STAGE_LD * 4;
constexpr size_t EPI_BIAS_BYTES = (size_t)EPI_WARPS * 128 * 4;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + EPI_STAGE_BYTES + EPI_BIAS_BYTES + 256 /*barriers*/ +
                              1024 /*alignment slack*/;

struct Params {
    int64_t M, N, K;
    const float* bias;      // [N] or null
    void* out;              // [M,N] operand dtype (BIAS, QUICKGELU) or fp32 (RESIDUAL, PATCH_EMBED)
    const float* aux;       // PATCH_EMBED: positional embedding [g2+1, N]
    int64_t aux_i;          // PATCH_EMBED: g2 (patches per image)
};

template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    if (BF16) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    } else {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
}

// x * sigmoid(1.702 x).  bf16 output (2^-9 rounding): sigmoid(y) = 0.5 + 0.5 tanh(y/2) with the single-MUFU
// tanh.approx (abs error ~5e-4 on the sigmoid, below the output rounding); fp16 output keeps exp + divide.
template <bool BF16>
__device__ __forceinline__ float quick_gelu(float x) {
    if (BF16) {
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
        const float hx = 0.5f * x;
        return fmaf(hx, t, hx);
    }
    return __fdividef(x, 1.0f + __expf(-1.702f * x));
}

template <int EPI, bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;                                   // [STAGES][128 rows][128 B]
    uint8_t* smem_b = smem + STAGES * A_BYTES;                // [STAGES][128 rows][128 B]
    float* epi_stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
    float* epi_bias = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + EPI_STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_STAGE_BYTES + EPI_BIAS_BYTES);
    uint64_t* full = bars;                         // [STAGES]  TMA (both CTAs) -> MMA (leader)
    uint64_t* empty = bars + STAGES;               // [STAGES]  MMA -> TMA, one copy per CTA
    uint64_t* tmem_full = bars + 2 * STAGES;       // [2] MMA -> epilogue, one copy per CTA
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2] epilogue (both CTAs) -> MMA (leader)
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = __shfl_sync(kFullMask, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = ptx::cluster_ctarank();
    const bool leader = cta_rank == 0;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int num_n = (int)(p.N / BN);
    const int num_m = (int)((p.M + BM - 1) / BM);
    const int num_tiles = num_m * num_n;
    const int num_k = (int)(p.K / BK);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tma_a);
        ptx::prefetch_tensormap(&tma_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(ptx::smem_u32(&full[s]), 1);          // the leader's producer arrive.expect_tx
            ptx::mbar_init(ptx::smem_u32(&empty[s]), 1);         // one tcgen05.commit arrival
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(ptx::smem_u32(&tmem_full[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&tmem_empty[s]), 2 * EPI_WARPS);   // every epilogue warp of both CTAs
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc<2>(ptx::smem_u32(tmem_base_slot), TMEM_COLS);
        ptx::tmem_relinquish<2>();
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();            // barriers of both CTAs are initialised before anyone signals them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------------ TMA producer (both CTAs)
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                const int m_blk = tile / num_n, n_blk = tile % num_n;
                const int a_row = m_blk * BM + (int)cta_rank * CTA_M;
                const int b_row = n_blk * BN + (int)cta_rank * CTA_NB;
                for (int kb = 0; kb < num_k; ++kb) {
                    ptx::mbar_wait(ptx::smem_u32(&empty[stage]), phase ^ 1);
                    const uint32_t fb_leader = ptx::mapa(ptx::smem_u32(&full[stage]), 0);
                    if (leader) ptx::mbar_arrive_expect_tx(ptx::smem_u32(&full[stage]), 2 * STAGE_BYTES);
                    ptx::tma_load_2d_cg2(ptx::smem_u32(smem_a + stage * A_BYTES), &tma_a, fb_leader, kb * BK, a_row);
                    ptx::tma_load_2d_cg2(ptx::smem_u32(smem_b + stage * B_BYTES), &tma_b, fb_leader, kb * BK, b_row);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            // ------------------------------------------------------------------ MMA issuer (leader CTA)
            constexpr uint32_t idesc = ptx::make_idesc_f16(BF16 ? 1u : 0u, BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                ptx::mbar_wait(ptx::smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    ptx::mbar_wait(ptx::smem_u32(&full[stage]), phase);
                    ptx::tc_fence_after();
                    const uint64_t a_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_a + stage * A_BYTES));
                    const uint64_t b_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_b + stage * B_BYTES));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advance 16 elements = 32 bytes along K inside the 128-byte swizzle row: +2 in the >>4 field
                        ptx::umma_f16<2>(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit_cg2(ptx::smem_u32(&empty[stage]), 3);      // frees the stage in both CTAs
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit_cg2(ptx::smem_u32(&tmem_full[acc]), 3);        // accumulator complete -> both epilogues
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------------- epilogue (warps 4..11)
        const int ew = warp - 4;
        const int wq = warp & 3;                       // TMEM lane quarter this warp may access (warp id % 4)
        const int half = ew >> 2;                      // which 128 of the 256 accumulator columns
        float* st = epi_stage + ew * (32 * STAGE_LD);
        float* sbias = epi_bias + ew * 128;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs) {
            const int m_blk = tile / num_n, n_blk = tile % num_n;
            const int64_t row0 = (int64_t)m_blk * BM + (int64_t)cta_rank * CTA_M + wq * 32;   // first of this warp's 32 rows
            const int64_t nb = (int64_t)n_blk * BN + half * 128;
            int pe_orow = 0, pe_prow = 0;
            if (EPI != EOE_EPI_PATCH_EMBED) {
#pragma unroll
                for (int i = 0; i < 4; ++i) sbias[i * 32 + lane] = p.bias ? __ldg(p.bias + nb + i * 32 + lane) : 0.f;
            } else {
                const int64_t row = row0 + lane, img = row / p.aux_i, pi = row % p.aux_i;
                pe_orow = (int)(img * (p.aux_i + 1) + 1 + pi);      // token 0 of every image is the class token
                pe_prow = (int)(1 + pi);
            }
            __syncwarp();
            ptx::mbar_wait(ptx::smem_u32(&tmem_full[acc]), acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)acc * BN + (uint32_t)half * 128;

            // Transposed access pattern shared by all variants: after the row owners (lane = row) have written their
            // 32 words, lane l reads the 16-byte group (l & 7) of row (it*4 + (l >> 3)): 8 lanes cover one 128-byte row
            // segment, one instruction covers 4 rows.
            const int sub = lane >> 3, grp = lane & 7;
            if (EPI == EOE_EPI_BIAS || EPI == EOE_EPI_BIAS_QUICKGELU) {
                // 16-bit output: 64 columns per round = 32 packed words per row
                uint16_t* out16 = reinterpret_cast<uint16_t*>(p.out);
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    uint32_t r0[32], r1[32];
                    ptx::tmem_ld_32x32b_x32(taddr + c * 64, r0);
                    ptx::tmem_ld_32x32b_x32(taddr + c * 64 + 32, r1);
                    ptx::tmem_ld_wait();
                    if (c == 1) {                      // all TMEM reads of this tile are done: release the accumulator
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive_cluster(ptx::smem_u32(&tmem_empty[acc]), 0);
                    }
                    __syncwarp();                      // previous round's readers are done with `st`
                    uint32_t* strow = reinterpret_cast<uint32_t*>(st) + lane * STAGE_LD;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        float a[8], b[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            a[e] = __uint_as_float(r0[j + e]) + sbias[c * 64 + j + e];
                            b[e] = __uint_as_float(r1[j + e]) + sbias[c * 64 + 32 + j + e];
                            if (EPI == EOE_EPI_BIAS_QUICKGELU) {      // QuickGELU (model.py:162-164): x * sigmoid(1.702 x)
                                a[e] = quick_gelu<BF16>(a[e]);
                                b[e] = quick_gelu<BF16>(b[e]);
                            }
                        }
                        *reinterpret_cast<uint4*>(strow + (j >> 1)) =
                            make_uint4(pack2<BF16>(a[0], a[1]), pack2<BF16>(a[2], a[3]), pack2<BF16>(a[4], a[5]), pack2<BF16>(a[6], a[7]));
                        *reinterpret_cast<uint4*>(strow + 16 + (j >> 1)) =
                            make_uint4(pack2<BF16>(b[0], b[1]), pack2<BF16>(b[2], b[3]), pack2<BF16>(b[4], b[5]), pack2<BF16>(b[6], b[7]));
                    }
                    __syncwarp();
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int r = it * 4 + sub;
                        const int64_t row = row0 + r;
                        const uint4 q = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(st) + r * STAGE_LD + grp * 4);
                        if (row < p.M) *reinterpret_cast<uint4*>(out16 + row * p.N + nb + c * 64 + grp * 8) = q;
                    }
                }
            } else {
                // fp32 output (residual += or patch-embed scatter): 32 columns per round
                float* out32 = reinterpret_cast<float*>(p.out);
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    const int64_t col = nb + c * 32 + grp * 4;
                    float4 pos[8];
                    if (EPI == EOE_EPI_PATCH_EMBED) {          // positional embedding rows: independent of the accumulator
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int prow = __shfl_sync(kFullMask, pe_prow, it * 4 + sub);
                            pos[it] = __ldg(reinterpret_cast<const float4*>(p.aux + (int64_t)prow * p.N + col));
                        }
                    }
                    uint32_t r0[32];
                    ptx::tmem_ld_32x32b_x32(taddr + c * 32, r0);
                    ptx::tmem_ld_wait();
                    if (c == 3) {
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive_cluster(ptx::smem_u32(&tmem_empty[acc]), 0);
                    }
                    __syncwarp();
                    float* strow = st + lane * STAGE_LD;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 v = make_float4(__uint_as_float(r0[j]), __uint_as_float(r0[j + 1]), __uint_as_float(r0[j + 2]),
                                               __uint_as_float(r0[j + 3]));
                        if (EPI == EOE_EPI_BIAS_RESIDUAL_F32) {
                            v.x += sbias[c * 32 + j]; v.y += sbias[c * 32 + j + 1];
                            v.z += sbias[c * 32 + j + 2]; v.w += sbias[c * 32 + j + 3];
                        }
                        *reinterpret_cast<float4*>(strow + j) = v;
                    }
                    __syncwarp();
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int r = it * 4 + sub;
                        const float4 v = *reinterpret_cast<const float4*>(st + r * STAGE_LD + grp * 4);
                        if (EPI == EOE_EPI_BIAS_RESIDUAL_F32) {
                            // x += attn/mlp branch (model.py:186-187).  Each element receives exactly one addend per
                            // launch, so the fire-and-forget L2 reduction is deterministic and keeps the read of x off the SM.
                            if (row0 + r < p.M)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out32 + (row0 + r) * p.N + col),
                                             "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                        } else {
                            const int orow = __shfl_sync(kFullMask, pe_orow, r);      // token row of patch row row0 + r
                            if (row0 + r < p.M)
                                *reinterpret_cast<float4*>(out32 + (int64_t)orow * p.N + col) =
                                    make_float4(v.x + pos[it].x, v.y + pos[it].y, v.z + pos[it].z, v.w + pos[it].w);
                        }
                    }
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();            // the peer may still be reading this CTA's smem / signalling its barriers
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<2>(tmem_base, TMEM_COLS);
    }
}

}  // namespace gemm
}  // namespace eoe
