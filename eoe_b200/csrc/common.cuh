// Shared device/host helpers for the eoe_b200 kernels (sm_100a only).
#pragma once
#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges cost a branch unless a profiler is attached
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/eoe_b200.h"

namespace eoe {

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs; grids are sized in multiples of this
constexpr unsigned kFullMask = 0xffffffffu;

void set_cuda_error(cudaError_t e, const char* where);
int check_launch(const char* where, int n_launched = 1);   // also feeds eoe_launch_count()

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 4 consecutive elements of T <-> 4 floats; one 16-byte (fp32) or 8-byte (16-bit) transaction.
// Streaming variants bypass L1 allocation: every head tensor is touched exactly once.
template <typename T>
__device__ __forceinline__ void load4_stream(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4_stream<float>(const float* p, float (&v)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
}
template <>
__device__ __forceinline__ void load4_stream<__half>(const __half* p, float (&v)[4]) {
    uint32_t a, b;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
    __half2 h0 = *reinterpret_cast<__half2*>(&a), h1 = *reinterpret_cast<__half2*>(&b);
    float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
}
template <>
__device__ __forceinline__ void load4_stream<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
    uint32_t a, b;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
    v[0] = __uint_as_float(a << 16); v[1] = __uint_as_float(a & 0xffff0000u);
    v[2] = __uint_as_float(b << 16); v[3] = __uint_as_float(b & 0xffff0000u);
}

template <typename T>
__device__ __forceinline__ void store4_stream(T* p, const float (&v)[4]);
template <>
__device__ __forceinline__ void store4_stream<float>(float* p, const float (&v)[4]) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}
template <>
__device__ __forceinline__ void store4_stream<__half>(__half* p, const float (&v)[4]) {
    __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};"
                 :: "l"(p), "r"(*reinterpret_cast<uint32_t*>(&h0)), "r"(*reinterpret_cast<uint32_t*>(&h1)) : "memory");
}
template <>
__device__ __forceinline__ void store4_stream<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};"
                 :: "l"(p), "r"(*reinterpret_cast<uint32_t*>(&h0)), "r"(*reinterpret_cast<uint32_t*>(&h1)) : "memory");
}

// 8 consecutive 16-bit elements <-> 8 floats in one 16-byte transaction (HSC rows in fp16 / bf16)
template <typename T>
__device__ __forceinline__ void load8_stream(const T* p, float (&v)[8]) {
    uint32_t w[4];
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (sizeof(T) == 2 && std::is_same<T, __nv_bfloat16>::value) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        } else {
            const float2 f = __half22float2(*reinterpret_cast<__half2*>(&w[i]));
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    }
}
template <typename T>
__device__ __forceinline__ void store8_stream(T* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (std::is_same<T, __nv_bfloat16>::value) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        } else {
            __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
    }
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
}
// VEC elements per lane per iteration: 4 (fp32, 16 B) or 8 (16-bit, 16 B)
template <typename T, int VEC>
__device__ __forceinline__ void loadv_stream(const T* p, float (&v)[VEC]) {
    if constexpr (VEC == 4) load4_stream<T>(p, v);
    else load8_stream<T>(p, v);
}
template <typename T, int VEC>
__device__ __forceinline__ void storev_stream(T* p, const float (&v)[VEC]) {
    if constexpr (VEC == 4) store4_stream<T>(p, v);
    else store8_stream<T>(p, v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}

// Deterministic mean over the grid: every block deposits one partial, the last block to arrive
// (threadfence + ticket) adds the partials in index order in fp64 and writes sum/n.  The ticket is
// reset so the (zero-initialised) workspace can be reused by the next call.
// RAII NVTX range around a host-side launch sequence (visible as named spans in Nsight Systems / Compute timelines:
// "eoe:vit_encode" > "eoe:block 3", "eoe:auc", ...).  SURVEY section 5 ("tracing").
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

struct HeadWorkspace {
    unsigned int ticket;
    unsigned int pad[31];
    float partial[(EOE_HEAD_WS_BYTES - 128) / 4];
};
constexpr int kMaxHeadBlocks = (EOE_HEAD_WS_BYTES - 128) / 4;

template <int BLOCK>
__device__ __forceinline__ void grid_mean_finish(float thread_val, HeadWorkspace* ws, float* out, double inv_n) {
    __shared__ float s_warp[BLOCK / 32];
    __shared__ bool s_last;
    float w = warp_sum(thread_val);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
        float b = 0.f;
#pragma unroll
        for (int i = 0; i < BLOCK / 32; ++i) b += s_warp[i];
        ws->partial[blockIdx.x] = b;
        __threadfence();
        unsigned int t = atomicAdd(&ws->ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x < 32) {
        __threadfence();
        double acc = 0.0;
        // fixed order: lane l adds partial[l], partial[l+32], ...; then a fixed shuffle tree
        for (unsigned i = threadIdx.x; i < gridDim.x; i += 32) acc += (double)__ldcg(&ws->partial[i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFullMask, acc, o);
        if (threadIdx.x == 0) {
            *out = (float)(acc * inv_n);
            ws->ticket = 0;
        }
    }
}

}  // namespace eoe
