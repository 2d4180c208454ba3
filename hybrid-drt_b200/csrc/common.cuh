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

// 64-bit shuffles, low word first: the toolkit's double overloads move the high word first, which lands the halves in
// the wrong registers of the destination pair and costs three LOP3 (an xor swap) per shuffle.
__device__ __forceinline__ double shfl_xor_d(double v, int m) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(kFull, lo, m);
    hi = __shfl_xor_sync(kFull, hi, m);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_idx_d(double v, int src) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(kFull, lo, src);
    hi = __shfl_sync(kFull, hi, src);
    return __hiloint2double(hi, lo);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_d(v, o);
    return v;
}

// Eight per-lane partial sums -> eight warp totals in nine shuffles: lane l returns the total of v[l / 4].
__device__ __forceinline__ double warp_reduce8(const double (&v)[8], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    double k[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) k[i] = (b4 ? v[4 + i] : v[i]) + shfl_xor_d(b4 ? v[i] : v[4 + i], 16);
    const double m0 = (b3 ? k[2] : k[0]) + shfl_xor_d(b3 ? k[0] : k[2], 8);
    const double m1 = (b3 ? k[3] : k[1]) + shfl_xor_d(b3 ? k[1] : k[3], 8);
    double t = (b2 ? m1 : m0) + shfl_xor_d(b2 ? m0 : m1, 4);
    t += shfl_xor_d(t, 2);
    t += shfl_xor_d(t, 1);
    return t;
}

// max(t, v) for a running maximum t that is never NaN; a NaN v is skipped, as fmax would.  One compare and two selects
// where fmax on doubles is eight instructions on this target.
__device__ __forceinline__ double dmax_run(double t, double v) { return (v > t) ? v : t; }

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, shfl_xor_d(v, o));
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

