// Development probe: FP64 throughput of DFMA vs DMMA (mma.sync m8n8k4 / m16n8k8 / m16n8k16 f64) on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void k_dfma(double* out, int iters) {
    double acc[ILP];
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    const double m = 1.0000001, b = 1e-9;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], m, b);
    double s = 0; for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_884(double* out, int iters) {
    double c[ILP][2];
    for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; }
    double a = threadIdx.x * 1e-9 + 1e-3, b = 1e-3;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(c[i], a, b);
    double s = 0; for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_1688(double* out, int iters) {
    double c[ILP][4];
    for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    double a[4] = {threadIdx.x * 1e-9 + 1e-3, 1e-3, 2e-3, 3e-3}, b[2] = {1e-3, 2e-3};
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma1688(c[i], a, b);
    double s = 0; for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_16816(double* out, int iters) {
    double c[ILP][4];
    for (int i = 0; i < ILP; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-9 + 1e-3 * i;
    for (int i = 0; i < 4; ++i) b[i] = 1e-3 * i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma16816(c[i], a, b);
    double s = 0; for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// latency probes: a single dependent chain per warp, one warp per SM sub-partition
__global__ void k_lat_dfma(double* out, int iters, long long* cyc) {
    double a = threadIdx.x * 1e-9; const double m = 1.0000001, b = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) a = fma(a, m, b);
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_lat_884(double* out, int iters, long long* cyc) {
    double c[2] = {0, 1}; double a = threadIdx.x * 1e-9 + 1e-3, b = 1e-3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) dmma884(c, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = c[0] + c[1]; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_lat_bar(double* out, int iters, long long* cyc) {
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) __syncthreads();
    long long t1 = clock64();
    out[threadIdx.x] = 0; if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <typename F>
static double time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 16 * 1024);
    long long* cyc; cudaMallocManaged(&cyc, 8);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        const int threads = 32 * warps, blocks = sms;
        double ms;
        ms = time_ms([&] { k_dfma<8><<<blocks, threads>>>(out, iters); });
        printf("warps/SM %2d  DFMA ilp8        %7.2f TFLOP/s\n", warps, 2.0 * 8 * iters * blocks * threads / ms / 1e9);
        ms = time_ms([&] { k_884<8><<<blocks, threads>>>(out, iters); });
        printf("warps/SM %2d  DMMA m8n8k4 ilp8  %7.2f TFLOP/s\n", warps, 2.0 * 256 * 8 * iters * blocks * warps / ms / 1e9);
        ms = time_ms([&] { k_884<2><<<blocks, threads>>>(out, iters); });
        printf("warps/SM %2d  DMMA m8n8k4 ilp2  %7.2f TFLOP/s\n", warps, 2.0 * 256 * 2 * iters * blocks * warps / ms / 1e9);
        ms = time_ms([&] { k_1688<4><<<blocks, threads>>>(out, iters); });
        printf("warps/SM %2d  DMMA m16n8k8 ilp4 %7.2f TFLOP/s\n", warps, 2.0 * 1024 * 4 * iters * blocks * warps / ms / 1e9);
        ms = time_ms([&] { k_16816<4><<<blocks, threads>>>(out, iters); });
        printf("warps/SM %2d  DMMA m16n8k16 ilp4 %6.2f TFLOP/s\n", warps, 2.0 * 2048 * 4 * iters * blocks * warps / ms / 1e9);
    }
    k_lat_dfma<<<1, 32>>>(out, 10000, cyc); cudaDeviceSynchronize(); printf("DFMA dependent latency  %.2f cycles\n", *cyc / 10000.0);
    k_lat_884<<<1, 32>>>(out, 10000, cyc); cudaDeviceSynchronize(); printf("DMMA m8n8k4 dependent latency %.2f cycles\n", *cyc / 10000.0);
    for (int t : {64, 128, 256, 512}) { k_lat_bar<<<1, t>>>(out, 10000, cyc); cudaDeviceSynchronize(); printf("__syncthreads %d threads %.2f cycles\n", t, *cyc / 10000.0); }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
