"""Development aid: achieved HBM bandwidth of filter_gather_kernel on a batch of raw traces."""
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybdrt_b200 import engine as E, synth, preprocessing as pp  # noqa: E402
eng = E.get_engine(0)
rt, ri, rv = synth.make_raw_chrono_batch(1, seed=5)
dec = pp.get_decimation_index(rt, rt[pp.identify_steps(ri, True)], np.min(np.diff(rt)), 25, 8, 2, None)
plan = pp.filter_plan(rt, pp.identify_steps(ri, allow_consecutive=False), dec)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for n_tr in (4096, 64, 1):
    raw = eng.dev(np.repeat(rv, n_tr, 0))
    for _ in range(3):
        eng.filter_gather(raw, plan)
    ms = []
    for k in range(5):
        flush.fill_(k)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.filter_gather(raw, plan); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    nbytes = 8.0 * n_tr * (len(rt) + len(dec))
    print(f'traces {n_tr:5d}: {np.mean(ms) * 1e3:8.1f} us  {nbytes / (np.mean(ms) * 1e-3) / 1e9:8.1f} GB/s', flush=True)
