// Development probe: can several CTAs per SM hold tensor-memory allocations at the same time?
// Each CTA allocates COLS columns, spins ~200 us, frees.  Prints, per configuration, the largest number of CTAs whose
// [start, end) intervals overlap on one SM.
#include <cstdio>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

template <int COLS>
__global__ void k(unsigned long long* out, int spin_us, int smem_dummy) {
    extern __shared__ char dyn[];
    __shared__ unsigned s_t;
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    unsigned long long t0, t1, t2;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_t)), "r"(COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2)); } while (t2 - t1 < (unsigned long long)spin_us * 1000ull);
    if (smem_dummy < 0) dyn[threadIdx.x] = 1;
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_t), "r"(COLS) : "memory");
    if (threadIdx.x == 0) { out[4 * blockIdx.x] = smid; out[4 * blockIdx.x + 1] = t0; out[4 * blockIdx.x + 2] = t1; out[4 * blockIdx.x + 3] = t2; }
}

template <int COLS>
static void run(int per_sm, int smem_bytes) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * per_sm;
    unsigned long long* d; cudaMalloc(&d, 32 * grid);
    cudaFuncSetAttribute(k<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<COLS>, 128, smem_bytes);
    k<COLS><<<grid, 128, smem_bytes>>>(d, 200, 0);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<unsigned long long> h(4 * grid);
    cudaMemcpy(h.data(), d, 32 * grid, cudaMemcpyDeviceToHost);
    int best = 0; double wait_max = 0;
    for (int s = 0; s < sms; ++s) {
        std::vector<std::pair<unsigned long long, int>> ev;
        for (int b = 0; b < grid; ++b) if ((int)h[4 * b] == s) { ev.push_back({h[4 * b + 2], 1}); ev.push_back({h[4 * b + 3], -1}); wait_max = std::max(wait_max, (double)(h[4 * b + 2] - h[4 * b + 1])); }
        std::sort(ev.begin(), ev.end());
        int cur = 0; for (auto& x : ev) { cur += x.second; best = std::max(best, cur); }
    }
    printf("cols %3d  launched %d CTAs/SM (occupancy API %d, smem %d KB): max concurrently holding TMEM on one SM = %d, longest alloc wait %.0f us  [%s]\n",
           COLS, per_sm, occ, smem_bytes >> 10, best, wait_max / 1000.0, cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run<128>(3, 69 * 1024);
    run<128>(3, 8 * 1024);
    run<128>(4, 8 * 1024);
    run<256>(2, 8 * 1024);
    run<64>(4, 8 * 1024);
    run<64>(8, 8 * 1024);
    run<32>(8, 8 * 1024);
    run<32>(16, 8 * 1024);
    return 0;
}
