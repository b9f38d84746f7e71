// Error plumbing shared by all entry points of libeoe_b200.so (see include/eoe_b200.h).
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace eoe {

static thread_local char g_cuda_err[512] = "";

void set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}

static std::atomic<long long> g_launches{0};

int check_launch(const char* where, int n_launched) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_cuda_error(e, where);
        return EOE_ERR_CUDA;
    }
    g_launches.fetch_add(n_launched, std::memory_order_relaxed);
    return EOE_OK;
}

}  // namespace eoe

extern "C" int eoe_abi_version(void) { return EOE_ABI_VERSION; }

#ifndef EOE_BUILD_ID
#define EOE_BUILD_ID "unknown"
#endif
extern "C" const char* eoe_build_id(void) { return EOE_BUILD_ID; }

extern "C" long long eoe_launch_count(void) { return eoe::g_launches.load(std::memory_order_relaxed); }

extern "C" const char* eoe_last_cuda_error(void) { return eoe::g_cuda_err; }

extern "C" const char* eoe_strerror(int code) {
    switch (code) {
        case EOE_OK: return "ok";
        case EOE_ERR_ARG: return "invalid argument (null pointer, non-positive size or bad enum)";
        case EOE_ERR_DTYPE: return "unsupported dtype";
        case EOE_ERR_SHAPE: return "unsupported shape";
        case EOE_ERR_ALIGN: return "pointer not sufficiently aligned";
        case EOE_ERR_WORKSPACE: return "workspace missing or too small";
        case EOE_ERR_CUDA: return "CUDA error (see eoe_last_cuda_error)";
        case EOE_ERR_ARCH: return "device is not an sm_100 (B200) GPU";
        default: return "unknown error";
    }
}
