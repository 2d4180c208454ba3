"""Development aid: achieved HBM write bandwidth of impedance_interp_kernel on G per-spectrum grids (C2 shape)."""
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybdrt_b200 import engine as E, synth  # noqa: E402

eng = E.get_engine(0)
eps = 1 / np.log(10 ** 0.1)
tab = eng.build_lookup(eps)
tau = np.logspace(-7, 3, 101)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for g, nf in ((10000, 70), (1, 70), (200, 70), (4096, 30)):
    freq = synth.C2_FREQ[:nf]
    f_dev = eng.dev(np.repeat(freq[None], g, 0) * (1 + 1e-3 * np.arange(g)[:, None] / g))
    t_dev = eng.dev(np.repeat(tau[None], g, 0))
    for _ in range(3):
        eng.build_impedance(f_dev, t_dev, eps, E.MODE_INTERP, tab)
    ms = []
    for k in range(5):
        flush.fill_(k)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.build_impedance(f_dev, t_dev, eps, E.MODE_INTERP, tab)
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    nbytes = 16.0 * g * nf * tau.size
    print(f'grids {g:6d} nf {nf}: {np.mean(ms) * 1e3:9.1f} us  {nbytes / (np.mean(ms) * 1e-3) / 1e9:8.1f} GB/s', flush=True)
