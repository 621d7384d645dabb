// Error plumbing shared by the C-ABI entry points (no torch, no pybind headers).
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

namespace glenet {

// thread-local last-error text, returned by glenet_last_error()
inline char* last_error_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

inline int fail(int code, const char* fmt, const char* what = "") {
    snprintf(last_error_buf(), 512, fmt, what);
    return code;
}

// negative return = -(cudaError_t); -1000.. = argument errors
enum { GLENET_OK = 0, GLENET_EINVAL = -1000, GLENET_EWORKSPACE = -1001, GLENET_EALIGN = -1002 };

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(last_error_buf(), 512, "%s: %s", what, cudaGetErrorString(e));
        return -(int)e;
    }
    return GLENET_OK;
}

template <typename K>
inline int set_smem(K kernel, size_t bytes, const char* what, bool max_carveout = false) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        snprintf(last_error_buf(), 512, "%s: cudaFuncSetAttribute(%zu B): %s", what, bytes, cudaGetErrorString(e));
        return -(int)e;
    }
    // a hint (failure is harmless): kernels sized to fill the SM's shared memory with resident CTAs ask for the largest carveout;
    // the others keep the driver's choice, i.e. more L1 for the clip code's local-memory vertex lists
    if (max_carveout) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return GLENET_OK;
}

// per-device caches of "attribute already set" flags are indexed by the CUDA device ordinal
#define GLENET_MAX_DEVICES 64
#define GLENET_MAX_PEERS 8      // GPUs of one NVLink / NVSwitch box that can take part in an exchange

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace glenet
