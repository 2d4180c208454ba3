// Development probe: isolated latency of the two diagonal-tile variants of qphb_kernel.cu (reciprocal per pivot vs
// division-free), one warp alone on an SM, and of their building blocks.
#include <cstdio>
#include <cuda_runtime.h>
constexpr unsigned kFull = 0xffffffffu;
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x; y = fma(y, fma(-hx * y, y, 0.5), y); y = fma(y, fma(-hx * y, y, 0.5), y); return y;
}
__device__ __forceinline__ double fast_rcp(double x) {
    double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(y, fma(-x, y, 1.0), y); y = fma(y, fma(-x, y, 1.0), y); return y;
}
template <int VAR>
__device__ __forceinline__ bool diag_factor(const double2 s, double* binv, double2& ykk, int lane) {
    const int g = lane >> 2, q = lane & 3;
    double m0 = -s.x, m1 = -s.y;
    double m2 = (g == 2 * q) ? 1.0 : 0.0, m3 = (g == 2 * q + 1) ? 1.0 : 0.0;
    double cg = 1.0;
#pragma unroll 1
    for (int cc = 0; cc < 7; ++cc) {
        const int h = cc >> 1;
        const double colv = (cc & 1) ? m1 : m0;
        const double agc = __shfl_sync(kFull, colv, 4 * g + h);
        const double piv = __shfl_sync(kFull, colv, 4 * cc + h);
        const int src = 4 * cc + q;
        const double r0 = __shfl_sync(kFull, m0, src), r1 = __shfl_sync(kFull, m1, src);
        const double r2 = __shfl_sync(kFull, m2, src), r3 = __shfl_sync(kFull, m3, src);
        if (VAR == 0) {
            const double f = (g > cc) ? -agc * fast_rcp(piv) : 0.0;
            m0 = fma(f, r0, m0); m1 = fma(f, r1, m1); m2 = fma(f, r2, m2); m3 = fma(f, r3, m3);
        } else if (VAR == 1) {
            const int e = (__double2hiint(piv) >> 20) & 0x7ff;
            const double sc = __hiloint2double((2046 - e) << 20, 0);
            const bool below = g > cc;
            const double pn = below ? piv * sc : 1.0, an = below ? -(agc * sc) : 0.0;
            cg *= pn;
            m0 = fma(an, r0, pn * m0); m1 = fma(an, r1, pn * m1); m2 = fma(an, r2, pn * m2); m3 = fma(an, r3, pn * m3);
        } else {   // shuffles only (no arithmetic chain): what the data movement alone costs
            m0 += r0 * 1e-30 + agc * 1e-30; m1 += r1 * 1e-30 + piv * 1e-30; m2 += r2 * 1e-30; m3 += r3 * 1e-30;
        }
    }
    const double dg = __shfl_sync(kFull, (g & 1) ? m1 : m0, 4 * g + (g >> 1));
    const bool ok = __all_sync(kFull, (dg > 0.0) && (dg < INFINITY));
    const double rinv = fast_rsqrt(cg * dg);
    m2 *= rinv; m3 *= rinv;
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
    const int lane = threadIdx.x, g = lane >> 2, q = lane & 3;
    double2 s = make_double2(-((g == 2 * q) ? 10.0 : 1.0 / (1 + g + 2 * q)), -((g == 2 * q + 1) ? 10.0 : 1.0 / (2 + g + 2 * q)));
    double2 y = make_double2(0, 0);
    bool ok = true;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { ok &= diag_factor<VAR>(s, binv, y, lane); s.x -= 1e-9 * y.x; s.y -= 1e-9 * y.y; }
    long long t1 = clock64();
    out[lane] = y.x + y.y + (ok ? 0 : 1);
    if (lane == 0) *cyc = t1 - t0;
}
__global__ void k_dfma(double* out, long long* cyc, int iters) {
    double v = 1.0 + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) v = fma(v, 1.0000001, 1e-9);
    long long t1 = clock64(); out[threadIdx.x] = v; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_dmma(double* out, long long* cyc, int iters) {
    double d0 = 0, d1 = 0, a = 1e-3 * threadIdx.x, b = 1e-3;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
    long long t1 = clock64(); out[threadIdx.x] = d0 + d1; if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* out; cudaMalloc(&out, 8 * 64); long long* cyc; cudaMallocManaged(&cyc, 8);
    const int it = 2000;
    k<0><<<1, 32>>>(out, cyc, it); cudaDeviceSynchronize(); printf("diag_factor reciprocal/pivot   %.1f cycles\n", *cyc / (double)it);
    k<1><<<1, 32>>>(out, cyc, it); cudaDeviceSynchronize(); printf("diag_factor division-free      %.1f cycles\n", *cyc / (double)it);
    k<2><<<1, 32>>>(out, cyc, it); cudaDeviceSynchronize(); printf("diag_factor shuffles only      %.1f cycles\n", *cyc / (double)it);
    k_dfma<<<1, 32>>>(out, cyc, 10000); cudaDeviceSynchronize(); printf("DFMA dependent                 %.1f cycles\n", *cyc / 10000.0);
    k_dmma<<<1, 32>>>(out, cyc, 10000); cudaDeviceSynchronize(); printf("DMMA m8n8k4 dependent          %.1f cycles\n", *cyc / 10000.0);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
