"""Development aid: host-side profile of DRTMD.fit_all on the C5 map (where the time outside the fit kernel goes)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybdrt_b200 import synth  # noqa: E402
from hybdrt_b200.mapping import DRTMD  # noqa: E402

rows = cols = 256
freq, z = synth.make_map_batch(rows, cols, seed=3)
psi = np.array([(r, c) for r in range(rows) for c in range(cols)], dtype=float)


def run():
    md = DRTMD(tau_supergrid=np.logspace(-8, 3, 111), psi_dim_names=['row', 'col'], print_progress=False)
    t0 = time.perf_counter()
    md.add_observations(psi, freq, z)
    t1 = time.perf_counter()
    md.fit_all(ignore_errors=True)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


run()
print('add_observations %.3f s, fit_all %.3f s' % run())
pr = cProfile.Profile()
pr.enable()
run()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
