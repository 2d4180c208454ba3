// Shared helpers for the hybdrt_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>

#include "hybdrt_b200.h"

namespace hdrt {

void set_error(const char* fmt, ...);

#define HDRT_CUDA_CHECK(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            hdrt::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return HDRT_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

}  // namespace hdrt

struct hdrt_handle {
    int device;
    int sm_count;
    int* work_counter;  // device
};
