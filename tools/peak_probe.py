"""Development aid: write-only and read-only HBM bandwidth of this GPU (torch fill_ / sum over 1 GiB), next to the
copy figure of MEASURED_PEAKS.json -- the denominators for write-bound (matrix builders) and read-bound (chrono filter) kernels."""
import torch
n = 1 << 27
a = torch.empty(n, dtype=torch.float64, device='cuda')
b = torch.empty(n, dtype=torch.float64, device='cuda')
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e-3
print('fill_  (write only) %.0f GB/s' % (8 * n / t(lambda: a.fill_(1.5)) / 1e9))
print('sum    (read only)  %.0f GB/s' % (8 * n / t(lambda: a.sum()) / 1e9))
print('copy_  (read+write) %.0f GB/s' % (16 * n / t(lambda: b.copy_(a)) / 1e9))
