"""Development aid: time of the trapz-mode matrix builders (1000-point quadrature per entry)."""
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybdrt_b200 import engine as E, synth  # noqa: E402
eng = E.get_engine(0)
eps = 1 / np.log(10 ** 0.1)
tau = np.logspace(-7, 3, 101)
for g in (1, 64):
    f = np.repeat(synth.C2_FREQ[None], g, 0)
    t = np.repeat(tau[None], g, 0)
    for _ in range(2):
        eng.build_impedance(f, t, eps, E.MODE_TRAPZ)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.build_impedance(f, t, eps, E.MODE_TRAPZ); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    evals = 2.0 * g * 70 * 101 * 1000
    print(f'trapz A_re + A_im, {g:3d} grid(s) 70 x 101: {ms:8.3f} ms  {evals / ms / 1e6:8.1f} G integrand evaluations/s', flush=True)
