"""Development aid: wall-clock throughput of the BASELINE.json configurations C3 (hybrid), C4 (DOP) and C5 (map)
through the public API (host buffers in, fit parameters out), plus size-independent sanity properties.
usage (GPU box): python tools/config_bench.py [c3] [c4] [c5]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybdrt_b200 import synth  # noqa: E402
from hybdrt_b200.models import DRT  # noqa: E402
from hybdrt_b200.mapping import DRTMD  # noqa: E402

which = [a for a in sys.argv[1:] if a in ('c2r', 'c3', 'c4', 'c5')] or ['c2r', 'c3', 'c4', 'c5']
out = {}


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, r


if 'c2r' in which:
    B = 10000
    freq, z = synth.make_eis_batch(B, seed=0)
    freqs = freq[None, :] * (1 + 1e-3 * np.arange(B)[:, None] / B)      # every spectrum on its own grid
    drt = DRT()
    dt, res = timed(lambda: drt.fit_eis_batch(freqs, z).fit_parameters() and drt.last_batch)
    h = res.host(['status', 'n_outer', 'x'])
    out['C2 ragged (10000 spectra, one frequency grid each: matrices built and read per spectrum)'] = dict(
        fits_per_s=B / dt, seconds=dt, mean_outer=float(h['n_outer'].mean()), finite=bool(np.isfinite(h['x']).all()),
        n=res.plan['n'], matrix_bytes=int(res.plan['rm'].numel() + res.plan['pen'].numel() + res.plan['vmm_eis'].numel()) * 8)
if 'c3' in which:
    B = 4096
    times, i_sig, v, freq, z = synth.make_hybrid_batch(B, seed=1)
    drt = DRT()
    dt, res = timed(lambda: drt.fit_hybrid_batch(times, i_sig, v, freq, z).fit_parameters() and drt.last_batch)
    h = res.host(['status', 'n_outer', 'x'])
    out['C3 hybrid (4096 x [2000 samples + 30 freqs])'] = dict(fits_per_s=B / dt, seconds=dt, mean_outer=float(h['n_outer'].mean()),
                                                            finite=bool(np.isfinite(h['x']).all()), n_rows=res.plan['n_rows'], n=res.plan['n'])
if 'c4' in which:
    B = 10000
    freq, z = synth.make_dop_batch(B, seed=2)
    drt = DRT(fit_dop=True)
    dt, res = timed(lambda: drt.fit_eis_batch(freq, z).fit_parameters() and drt.last_batch)
    h = res.host(['status', 'n_outer', 'x'])
    out['C4 DOP (10000 spectra, n = %d)' % res.plan['n']] = dict(fits_per_s=B / dt, seconds=dt, mean_outer=float(h['n_outer'].mean()),
                                                                finite=bool(np.isfinite(h['x']).all()))
if 'c5' in which:
    rows = cols = 256
    freq, z = synth.make_map_batch(rows, cols, seed=3)
    psi = np.array([(r, c) for r in range(rows) for c in range(cols)], dtype=float)

    def run():
        md = DRTMD(tau_supergrid=np.logspace(-8, 3, 111), psi_dim_names=['row', 'col'], print_progress=False)
        md.add_observations(psi, freq, z)
        md.fit_all(ignore_errors=True)
        return md
    dt, md = timed(run, reps=1)
    out['C5 map 256x256 through DRTMD (fit + drt_var + llh + rss)'] = dict(
        fits_per_s=rows * cols / dt, seconds=dt, fitted=int(md.obs_fit_status.sum()), mean_outer=float(md.obs_outer_iterations.mean()),
        finite=bool(np.isfinite(md.obs_x).all() and np.isfinite(md.obs_drt_var[md.obs_fit_status]).all()))
print(json.dumps(out, indent=1))
