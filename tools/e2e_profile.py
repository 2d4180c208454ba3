"""Development aid: where the end-to-end time of DRT.fit_eis_batch goes (host stages vs the fit kernel)."""
import os
import sys
import time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybdrt_b200 import synth  # noqa: E402
from hybdrt_b200.models import DRT  # noqa: E402
import cProfile, pstats  # noqa: E402

freq, z = synth.make_eis_batch(10000, seed=0)
drt = DRT()
for _ in range(2):
    drt.fit_eis_batch(freq, z).fit_parameters()
torch.cuda.synchronize()
t0 = time.perf_counter()
res = drt.fit_eis_batch(freq, z)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
fp = res.fit_parameters()
t3 = time.perf_counter()
print(f'fit_eis_batch returns after {1e3 * (t1 - t0):.1f} ms; kernel done after {1e3 * (t2 - t0):.1f} ms; '
      f'fit_parameters (D2H + unscale) {1e3 * (t3 - t2):.1f} ms; total {1e3 * (t3 - t0):.1f} ms')
pr = cProfile.Profile()
pr.enable()
res = drt.fit_eis_batch(freq, z)
fp = res.fit_parameters()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
