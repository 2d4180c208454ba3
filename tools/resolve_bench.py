"""Development aid: throughput of the batched resolve QP (hdrt_resolve_qp_batch) on a synthetic hybrid map:
G groups of 32 observations each, resolved in windows of 7 with 2 shared neighbours."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybdrt_b200 import synth  # noqa: E402
from hybdrt_b200.mapping import DRTMD  # noqa: E402

n_groups, per_group = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 32
times = np.concatenate([np.linspace(-0.01, -1e-4, 25), np.logspace(-4, 0, 220)])
t, i_sig, v, freq, z = synth.make_hybrid_batch(n_groups * per_group, times=times, seed=17)
md = DRTMD(tau_supergrid=np.logspace(-8, 3, 111), psi_dim_names=['g', 'k'], print_progress=False, keep_pq=True)
for b in range(len(z)):
    md.add_observation([float(b // per_group), float(b % per_group)], (t, i_sig, v[b]), (freq, z[b]), group_id=f'g{b // per_group}')
t0 = time.perf_counter()
md.fit_all(ignore_errors=True)
torch.cuda.synchronize()
t1 = time.perf_counter()
print(f'fit_all: {len(z)} hybrid observations in {t1 - t0:.2f} s')
nwin = 0
for rep in range(2):
    t0 = time.perf_counter()
    nwin = 0
    for gi in range(n_groups):
        md.resolve_group(f'g{gi}', batch_size=7, overlap=2, psi_sort_dims=['k'])
        nwin += len(md.last_resolve['iters'])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(f'resolve_group x {n_groups}: {nwin} windows of 7 x {md.obs_tau_indices[0][1] - md.obs_tau_indices[0][0] + 2} unknowns in {dt:.2f} s '
      f'({nwin / dt:.1f} windows/s, {np.mean(md.last_resolve["iters"]):.1f} interior-point iterations per window)')
# kernel alone: all windows of all groups in one launch
wins = []
for gi in range(n_groups):
    idx = md.get_group_index(f'g{gi}')
    for a in range(0, per_group - 6, 5):
        wins.append(idx[a:a + 7])
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    md._resolve_windows(wins, False, 1, 1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(f'one launch, {len(wins)} windows: {dt:.2f} s ({len(wins) / dt:.1f} windows/s incl. host assembly)')
# kernel alone (CUDA events around the launch inside _resolve_windows)
eng = md.drt1d.engine
orig = eng.resolve_qp_batch
times_ms = []


def timed(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = orig(*a, **k)
    e1.record()
    torch.cuda.synchronize()
    times_ms.append(e0.elapsed_time(e1))
    return out


eng.resolve_qp_batch = timed
md._resolve_windows(wins, False, 1, 1)
md._resolve_windows(wins, False, 1, 1)
it = md.last_resolve['iters']
print(f'resolve_qp_kernel alone: {times_ms[-1]:.1f} ms for {len(wins)} windows ({len(wins) / times_ms[-1] * 1e3:.0f} windows/s, '
      f'{np.mean(it):.1f} interior-point iterations per window, {times_ms[-1] / np.mean(it):.2f} ms per iteration of a wave)')
