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

// Per-launch resources of the fit kernel.  A handle owns a small ring of them so that fits launched on different
// streams of one handle never share a work counter: launch i takes slot i % kLaunchSlots and, before reusing it, makes
// its stream wait for the event recorded behind the previous launch that used the slot.
constexpr int kLaunchSlots = 4;
struct hdrt_launch_slot {
    int* work_counter;      // device: the persistent CTAs pull spectrum indices from it
    cudaEvent_t done;       // recorded behind the last launch that used the slot
    bool used;
};

struct hdrt_handle {
    int device;
    int sm_count;
    size_t smem_per_sm, smem_reserved_per_cta;   // device limits the fit kernel sizes its grid with
    int regs_per_sm, threads_per_sm;
    hdrt_launch_slot slots[kLaunchSlots];
    unsigned long long launches;        // guarded by mu
    void* mu;                           // std::mutex*
};

