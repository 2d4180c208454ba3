"""Reduced set of fits for compute-sanitizer (memcheck / racecheck): one spectrum of the reference's golden test, a
16-spectrum C2 batch (warp-per-spectrum kernel: Toeplitz path, and the general path through the raw engine call), one
hybrid fit (CTA kernel, N = 2060 rows cut to 260), one DRT + DOP fit (16-warp configuration), one optional path
(outlier error structure), and the matrix builders those fits need.
usage (GPU box): compute-sanitizer --tool memcheck python tools/sanitize_cases.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybdrt_b200 import synth  # noqa: E402
from hybdrt_b200.models import DRT  # noqa: E402

freq, z = synth.make_eis_batch(16, seed=0)
drt = DRT()
r = drt.fit_eis_batch(freq, z)
x_t = r.host(['x'])['x']
plan = r.plan
zs = z / r.scales['coefficient_scale'][:, None]
raw = drt.engine.qphb_fit_batch(plan['rm'], drt.engine.dev(np.concatenate([zs.real, zs.imag], axis=1)), plan['pen'], plan['h'],
                                plan['l1'], plan['n_special'], vmm_eis=plan['vmm_eis'], hypers=drt._c_hypers(plan['opts']))
torch.cuda.synchronize()
print('C2 x16: Toeplitz vs general path max rel', float(np.max(np.abs(raw['x'].cpu().numpy() - x_t)) / np.max(np.abs(x_t))))
drt.fit_eis(freq, z[0])
print('single fit_eis outer', drt.qphb_params['n_outer'])
t = np.concatenate([np.linspace(-0.01, -1e-4, 25), np.logspace(-4, 0, 220)])
tt, i_sig, v, f3, z3 = synth.make_hybrid_batch(2, times=t, seed=1)
drt.fit_hybrid(tt, i_sig, v[0], f3, z3[0])
print('hybrid outer', drt.qphb_params['n_outer'])
f4, z4 = synth.make_dop_batch(2, seed=2)
dd = DRT(fit_dop=True)
dd.fit_eis(f4, z4[0])
print('dop outer', dd.qphb_params['n_outer'])
drt.fit_eis(freq, z[1], outlier_p=0.05)
print('outlier outer', drt.qphb_params['n_outer'])
torch.cuda.synchronize()
print('done')
