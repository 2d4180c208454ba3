// Development probe: latency of the 8x8 diagonal-tile factor (diag_factor of qphb_kernel.cu) and of its pieces.
#include <cstdio>
#include <cuda_runtime.h>
constexpr unsigned kFull = 0xffffffffu;
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = fma(y, fma(-hx * y, y, 0.5), y);
    y = fma(y, fma(-hx * y, y, 0.5), y);
    return y;
}
template <int VAR>
__device__ __forceinline__ bool diag_factor(const double2 s, double* binv, double2& ykk, int lane) {
    const int g = lane >> 2, q = lane & 3;
    double m0 = -s.x, m1 = -s.y;
    double m2 = (g == 2 * q) ? 1.0 : 0.0, m3 = (g == 2 * q + 1) ? 1.0 : 0.0;
    bool ok = true;
    double piv = __shfl_sync(kFull, m0, 0);
#pragma unroll 1
    for (int cc = 0; cc < 8; ++cc) {
        const int h = cc >> 1, cn = min(cc + 1, 7);
        const double colv = (cc & 1) ? m1 : m0;
        const double agc = __shfl_sync(kFull, colv, 4 * g + h);
        const double nl = __shfl_sync(kFull, colv, 4 * cn + h);
        const double nd = __shfl_sync(kFull, (cn & 1) ? m1 : m0, 4 * cn + (cn >> 1));
        const int src = 4 * cc + q;
        double r0 = __shfl_sync(kFull, m0, src), r1 = __shfl_sync(kFull, m1, src);
        double r2 = __shfl_sync(kFull, m2, src), r3 = __shfl_sync(kFull, m3, src);
        ok = ok && (piv > 0.0) && (piv < INFINITY);
        double rinv;
        if (VAR == 1) rinv = rsqrt(piv);
        else if (VAR == 2) { asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rinv) : "d"(piv)); }
        else rinv = fast_rsqrt(piv);
        const double t = nl * rinv;
        piv = fma(-t, t, nd);
        const double lr = agc * rinv;
        r0 *= rinv; r1 *= rinv; r2 *= rinv; r3 *= rinv;
        if (VAR == 3) {
            const bool prow = g == cc, act = g >= cc;
            const double f = act ? (prow ? 0.0 : -lr) : 0.0, bm = prow ? 0.0 : 1.0;
            m0 = act ? fma(f, r0, prow ? r0 : m0 * bm) : m0;
            m1 = act ? fma(f, r1, prow ? r1 : m1 * bm) : m1;
            m2 = act ? fma(f, r2, prow ? r2 : m2 * bm) : m2;
            m3 = act ? fma(f, r3, prow ? r3 : m3 * bm) : m3;
        } else if (g >= cc) {
            const bool prow = g == cc;
            m0 = prow ? r0 : fma(-lr, r0, m0);
            m1 = prow ? r1 : fma(-lr, r1, m1);
            m2 = prow ? r2 : fma(-lr, r2, m2);
            m3 = prow ? r3 : fma(-lr, r3, m3);
        }
    }
    *reinterpret_cast<double2*>(binv + 2 * lane) = make_double2(-m2, -m3);
    const int sx = 8 * q + (g >> 1);
    const double a2 = __shfl_sync(kFull, m2, sx), a3 = __shfl_sync(kFull, m3, sx);
    const double b2 = __shfl_sync(kFull, m2, sx + 4), b3 = __shfl_sync(kFull, m3, sx + 4);
    ykk = make_double2((g & 1) ? a3 : a2, (g & 1) ? b3 : b2);
    return ok;
}
template <int VAR>
__global__ void k(double* out, long long* cyc, int iters) {
    __shared__ double binv[64];
    const int lane = threadIdx.x;
    const int g = lane >> 2, q = lane & 3;
    double2 s = make_double2(-((g == 2 * q) ? 10.0 : 1.0 / (1 + g + 2 * q)), -((g == 2 * q + 1) ? 10.0 : 1.0 / (2 + g + 2 * q)));
    double2 y = make_double2(0, 0);
    bool ok = true;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        ok &= diag_factor<VAR>(s, binv, y, lane);
        s.x -= 1e-9 * y.x; s.y -= 1e-9 * y.y;
    }
    long long t1 = clock64();
    out[lane] = y.x + y.y + (ok ? 0 : 1);
    if (lane == 0) *cyc = t1 - t0;
}
__global__ void k_shfl(double* out, long long* cyc, int iters) {
    double v = threadIdx.x; int src = threadIdx.x ^ 5;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) v = __shfl_sync(kFull, v, src) + 1.0;
    long long t1 = clock64(); out[threadIdx.x] = v; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_rsq(double* out, long long* cyc, int iters) {
    double v = 1.0 + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v)); v = y + 1.5; }
    long long t1 = clock64(); out[threadIdx.x] = v; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_frsq(double* out, long long* cyc, int iters) {
    double v = 1.0 + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) v = fast_rsqrt(v) + 1.5;
    long long t1 = clock64(); out[threadIdx.x] = v; if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* out; cudaMalloc(&out, 8 * 64); long long* cyc; cudaMallocManaged(&cyc, 8);
    const int it = 2000;
    k<0><<<1, 32>>>(out, cyc, it); cudaDeviceSynchronize(); printf("diag_factor fast_rsqrt        %.1f cycles\n", *cyc / (double)it);
    k<1><<<1, 32>>>(out, cyc, it); cudaDeviceSynchronize(); printf("diag_factor rsqrt()           %.1f cycles\n", *cyc / (double)it);
    k<2><<<1, 32>>>(out, cyc, it); cudaDeviceSynchronize(); printf("diag_factor approx only       %.1f cycles\n", *cyc / (double)it);
    k<3><<<1, 32>>>(out, cyc, it); cudaDeviceSynchronize(); printf("diag_factor branch-free       %.1f cycles\n", *cyc / (double)it);
    k_shfl<<<1, 32>>>(out, cyc, 10000); cudaDeviceSynchronize(); printf("DSHFL+DADD dependent          %.1f cycles\n", *cyc / 10000.0);
    k_rsq<<<1, 32>>>(out, cyc, 10000); cudaDeviceSynchronize(); printf("rsqrt.approx.f64+DADD dep     %.1f cycles\n", *cyc / 10000.0);
    k_frsq<<<1, 32>>>(out, cyc, 10000); cudaDeviceSynchronize(); printf("fast_rsqrt+DADD dependent     %.1f cycles\n", *cyc / 10000.0);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
