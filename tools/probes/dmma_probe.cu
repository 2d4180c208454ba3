// Development probe: issue behaviour of mma.sync.m8n8k4.f64 (DMMA.8x8x4) on sm_100a.
// For K independent accumulators per warp and W warps per SM sub-partition: cycles per DMMA per sub-partition,
//   mode 0: K independent chains, one DMMA of each per round ("sweep")
//   mode 1: K tiles, two dependent DMMAs back to back per tile ("pairs", the Gram loop as written)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe tools/probes/dmma_probe.cu ; run on the GPU box
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double2& d, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d.x), "+d"(d.y) : "d"(a), "d"(b));
}

template <int K, int MODE>
__global__ void probe(double* out, long long* cyc, int iters) {
    double2 acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = make_double2(0.0, 0.0);
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) dmma(acc[k], a, b);
#pragma unroll
            for (int k = 0; k < K; ++k) dmma(acc[k], b, a);
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) { dmma(acc[k], a, b); dmma(acc[k], b, a); }
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += acc[k].x + acc[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int K, int MODE>
void run(int warps_per_smsp, double* out, long long* cyc) {
    const int iters = 2000;
    probe<K, MODE><<<148, 128 * warps_per_smsp>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    probe<K, MODE><<<148, 128 * warps_per_smsp>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double per_warp = (double)c / (2.0 * K * iters);
    printf("K=%2d mode=%s warps/SMSP=%d: %.1f cycles per DMMA per warp, %.1f per sub-partition\n", K, MODE ? "pairs" : "sweep", warps_per_smsp,
           per_warp, per_warp / warps_per_smsp);
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 148 * 1024);
    cudaMalloc(&cyc, sizeof(long long));
    for (int w = 1; w <= 3; ++w) {
        run<1, 0>(w, out, cyc); run<2, 0>(w, out, cyc); run<4, 0>(w, out, cyc); run<8, 0>(w, out, cyc); run<28, 0>(w, out, cyc);
        run<1, 1>(w, out, cyc); run<4, 1>(w, out, cyc); run<28, 1>(w, out, cyc);
    }
    return 0;
}
