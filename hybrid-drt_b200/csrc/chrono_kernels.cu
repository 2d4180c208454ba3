// Chrono conditioning ahead of fit_chrono / fit_hybrid (reference: hybdrt/preprocessing.py downsample_data
// :335-468 -> filter_chrono_signal :507-574 -> filters/_filters.py nonuniform_gaussian_filter1d :261-341).
//
// The reference filters every raw sample of a trace with a position-dependent Gaussian (a blend of
// scipy.ndimage.gaussian_filter1d outputs at log-spaced widths, mirrored at the step boundaries) and then keeps
// the decimation index.  Only the kept samples are needed, and their filter taps depend on the time grid alone,
// not on the signal: the host lays the blended, normalised taps of every kept sample out once per grid
// (hybdrt_b200/preprocessing.py), and this kernel evaluates them for a whole batch of traces -- a gather-FMA
// over windows that tile each trace about twice, so a trace is read from HBM once (bound: HBM read bandwidth).
#include "common.cuh"

namespace hdrt {

constexpr int kFThreads = 256;

// out[s][m] = sum_{k=-lw..lw} taps[woff[m] + lw + k] * y[s][seg_lo[m] + reflect(idx[m] - seg_lo[m] + k, seg_len[m])]
// reflect = scipy.ndimage 'reflect' (d c b a | a b c d | d c b a), periodic in 2 * seg_len.
__global__ void filter_gather_kernel(const double* __restrict__ y, int n_sig, int nt, const int* __restrict__ idx,
                                     const int* __restrict__ seg_lo, const int* __restrict__ seg_len,
                                     const long long* __restrict__ woff, const int* __restrict__ lw,
                                     const double* __restrict__ taps, int m_total, int m_per_cta,
                                     double* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m0 = blockIdx.y * m_per_cta, m1 = min(m0 + m_per_cta, m_total);
    for (int s = blockIdx.x; s < n_sig; s += gridDim.x) {
        const double* __restrict__ ys = y + (size_t)s * nt;
        for (int m = m0 + warp; m < m1; m += kFThreads / 32) {
            const int i = idx[m], lo = seg_lo[m], len = seg_len[m], r = lw[m];
            const double* __restrict__ w = taps + woff[m] + r;
            const int c = i - lo;
            double acc = 0.0;
            if (c - r >= 0 && c + r < len) {          // window inside the segment: no mirroring
                for (int k = -r + lane; k <= r; k += 32) acc = fma(w[k], ys[i + k], acc);
            } else {
                const int p2 = 2 * len;
                for (int k = -r + lane; k <= r; k += 32) {
                    int p = (c + k) % p2;
                    if (p < 0) p += p2;
                    if (p >= len) p = p2 - 1 - p;
                    acc = fma(w[k], ys[lo + p], acc);
                }
            }
            acc = warp_sum(acc);
            if (lane == 0) out[(size_t)s * m_total + m] = acc;
        }
    }
}

}  // namespace hdrt

using namespace hdrt;

extern "C" int hdrt_filter_gather(const double* y, int n_sig, int nt, const int* idx, const int* seg_lo,
                                  const int* seg_len, const long long* woff, const int* lw, const double* taps, int m,
                                  double* out, void* stream) {
    if (n_sig == 0) return HDRT_OK;        // an empty batch is a no-op (its buffers may be null)
    if (!y || !idx || !seg_lo || !seg_len || !woff || !lw || !taps || !out || n_sig < 0 || nt <= 0 || m <= 0) {
        set_error("hdrt_filter_gather: invalid argument");
        return HDRT_ERR_ARG;
    }
    int dev = 0, sms = 148;
    HDRT_CUDA_CHECK(cudaGetDevice(&dev));
    HDRT_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // one CTA per trace when there are enough traces to fill the GPU; otherwise the kept samples are split too
    int slices = 1;
    while ((long long)n_sig * slices < 4LL * sms && m / (slices * 2) >= kFThreads / 32) slices *= 2;
    const int m_per_cta = (m + slices - 1) / slices;
    dim3 grid(n_sig < 8 * sms ? n_sig : 8 * sms, (m + m_per_cta - 1) / m_per_cta);
    filter_gather_kernel<<<grid, kFThreads, 0, (cudaStream_t)stream>>>(y, n_sig, nt, idx, seg_lo, seg_len, woff, lw, taps, m,
                                                                       m_per_cta, out);
    HDRT_CUDA_CHECK(cudaGetLastError());
    return HDRT_OK;
}
