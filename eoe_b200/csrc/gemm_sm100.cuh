// tcgen05 / TMEM / TMA GEMM for the CLIP ViT encoder (sm_100a):   out = epilogue( A[M,K] * W[N,K]^T )
//
// Replaces the 48 cuBLAS addmm + conv1 of VisualTransformer.forward (reference
// src/eoe/models/clip_official/clip/model.py:171,174-176,220).  Both operands are K-major 16-bit (activations
// [tokens, K] and nn.Linear weights [N, K] are already in that layout), accumulation is fp32 in TMEM.
//
// CTA = 128 x 256 output tile, persistent over tiles (n fastest so that the CTAs running concurrently share the
// same A rows in L2 while W, <= 4.7 MB, stays L2 resident).  Warp roles:
//   warp 0   TMA producer: 4-stage ring of {A 128x64, B 256x64} bf16 tiles (SWIZZLE_128B), mbarrier full/empty
//   warp 1   MMA issuer: one thread issues tcgen05.mma 128x256x16, tcgen05.commit frees the smem stage
//   warp 2   TMEM allocator (512 columns = 2 accumulator stages of 256 columns)
//   warps 4-7 epilogue: tcgen05.ld 32 rows x 32 columns per warp, fused bias / QuickGELU / residual / pos-emb,
//            overlapped with the next tile's main loop through the second accumulator stage
#pragma once
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace eoe {
namespace gemm {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4, UMMA_K = 16;
constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int THREADS = 256;
constexpr uint32_t TMEM_COLS = 512;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 256 /*barriers*/ + 1024 /*alignment slack*/;

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

template <int EPI, bool BF16>
__global__ void __launch_bounds__(THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;                                   // [STAGES][128 rows][128 B]
    uint8_t* smem_b = smem + STAGES * A_BYTES;                // [STAGES][256 rows][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;                 // [STAGES]  TMA -> MMA
    uint64_t* empty = bars + STAGES;       // [STAGES]  MMA -> TMA
    uint64_t* tmem_full = bars + 2 * STAGES;       // [2] MMA -> epilogue
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2] epilogue -> MMA
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = __shfl_sync(kFullMask, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
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
            ptx::mbar_init(ptx::smem_u32(&full[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&empty[s]), 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(ptx::smem_u32(&tmem_full[s]), 1);
            ptx::mbar_init(ptx::smem_u32(&tmem_empty[s]), 4);   // one arrival per epilogue warp
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc<1>(ptx::smem_u32(tmem_base_slot), TMEM_COLS);
        ptx::tmem_relinquish<1>();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------------ TMA producer
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile / num_n, n_blk = tile % num_n;
                for (int kb = 0; kb < num_k; ++kb) {
                    ptx::mbar_wait(ptx::smem_u32(&empty[stage]), phase ^ 1);
                    const uint32_t fb = ptx::smem_u32(&full[stage]);
                    ptx::mbar_arrive_expect_tx(fb, STAGE_BYTES);
                    ptx::tma_load_2d(ptx::smem_u32(smem_a + stage * A_BYTES), &tma_a, fb, kb * BK, m_blk * BM);
                    ptx::tma_load_2d(ptx::smem_u32(smem_b + stage * B_BYTES), &tma_b, fb, kb * BK, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------------------------ MMA issuer
            constexpr uint32_t idesc = ptx::make_idesc_f16(BF16 ? 1u : 0u, BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
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
                        ptx::umma_f16<1>(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(ptx::smem_u32(&empty[stage]));      // frees the smem stage when the MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(ptx::smem_u32(&tmem_full[acc]));        // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------------- epilogue (warps 4..7)
        const int wq = warp & 3;                       // TMEM lane quarter this warp may access
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / num_n, n_blk = tile % num_n;
            ptx::mbar_wait(ptx::smem_u32(&tmem_full[acc]), acc_phase);
            ptx::tc_fence_after();
            const int64_t row = (int64_t)m_blk * BM + wq * 32 + lane;
            const bool row_ok = row < p.M;
            const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)acc * BN;
            int64_t orow = row;
            const float* pos = nullptr;
            if (EPI == EOE_EPI_PATCH_EMBED) {
                const int64_t img = row / p.aux_i, pi = row % p.aux_i;
                orow = img * (p.aux_i + 1) + 1 + pi;
                pos = p.aux + (1 + pi) * p.N;
            }
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32];
                ptx::tmem_ld_32x32b_x32(taddr + c * 32, r);
                ptx::tmem_ld_wait();
                const int64_t n0 = (int64_t)n_blk * BN + c * 32;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (EPI != EOE_EPI_PATCH_EMBED && p.bias) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                    }
                }
                if (EPI == EOE_EPI_BIAS_QUICKGELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)            // QuickGELU (model.py:162-164): x * sigmoid(1.702 x)
                        v[j] = __fdividef(v[j], 1.0f + __expf(-1.702f * v[j]));
                }
                if (!row_ok) continue;
                if (EPI == EOE_EPI_BIAS || EPI == EOE_EPI_BIAS_QUICKGELU) {
                    uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + orow * p.N + n0;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint4 q;
                        q.x = pack2<BF16>(v[j], v[j + 1]); q.y = pack2<BF16>(v[j + 2], v[j + 3]);
                        q.z = pack2<BF16>(v[j + 4], v[j + 5]); q.w = pack2<BF16>(v[j + 6], v[j + 7]);
                        *reinterpret_cast<uint4*>(o + j) = q;
                    }
                } else if (EPI == EOE_EPI_BIAS_RESIDUAL_F32) {
                    float* o = reinterpret_cast<float*>(p.out) + orow * p.N + n0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 x = *reinterpret_cast<const float4*>(o + j);
                        x.x += v[j]; x.y += v[j + 1]; x.z += v[j + 2]; x.w += v[j + 3];
                        *reinterpret_cast<float4*>(o + j) = x;
                    }
                } else {   // PATCH_EMBED
                    float* o = reinterpret_cast<float*>(p.out) + orow * p.N + n0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 e = __ldg(reinterpret_cast<const float4*>(pos + n0 + j));
                        *reinterpret_cast<float4*>(o + j) = make_float4(v[j] + e.x, v[j + 1] + e.y, v[j + 2] + e.z, v[j + 3] + e.w);
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&tmem_empty[acc]));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem_base, TMEM_COLS);
    }
}

}  // namespace gemm
}  // namespace eoe
