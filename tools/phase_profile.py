"""Development aid: per-phase clock64 breakdown of the QPHB kernel (warp 0 of block 0).

Builds a -DHDRT_PROFILE copy of the library next to the product one, runs a C2 batch and prints cycles per fit.
usage (GPU box): python tools/phase_profile.py [c2|c3|c4] [batch]
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
csrc = os.path.join(ROOT, 'hybrid-drt_b200', 'csrc')
out = os.path.join(ROOT, 'hybrid-drt_b200', '_lib', 'libhybdrt_b200_prof.so')
if '--nobuild' not in sys.argv:
    subprocess.check_call(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-DHDRT_PROFILE',
                           '-Xcompiler', '-fPIC', '-shared', '-I', os.path.join(ROOT, 'include'), '-I', csrc, '-o', out] +
                          [os.path.join(csrc, f) for f in ('capi.cu', 'matrix_kernels.cu', 'chrono_kernels.cu', 'qphb_kernel.cu')] + ['-lcudart'])
if not torch.cuda.is_available():
    sys.exit(0)
from hybdrt_b200 import engine as E, synth  # noqa: E402
E.LIB_PATH = out
from hybdrt_b200.models import DRT  # noqa: E402

batch = int([a for a in sys.argv[1:] if a.isdigit()][0]) if any(a.isdigit() for a in sys.argv[1:]) else 444
which = ([a for a in sys.argv[1:] if a in ('c2', 'c3', 'c4')] or ['c2'])[0]
if which == 'c2':
    freq, z = synth.make_eis_batch(batch, seed=0)
    drt = DRT()
    fit = lambda: drt.fit_eis_batch(freq, z)
elif which == 'c3':
    t_, i_, v_, f_, z_ = synth.make_hybrid_batch(batch, seed=1)
    drt = DRT()
    fit = lambda: drt.fit_hybrid_batch(t_, i_, v_, f_, z_)
else:
    f_, z_ = synth.make_dop_batch(batch, seed=2)
    drt = DRT(fit_dop=True)
    fit = lambda: drt.fit_eis_batch(f_, z_)
lib = E.load_library()
lib.hdrt_debug_profile.argtypes = [C.c_void_p, C.c_int]
fit()
torch.cuda.synchronize()
lib.hdrt_debug_profile(None, 1)
res = fit()
torch.cuda.synchronize()
print('config', which, 'batch', batch, 'mean outer', float(res.host(['n_outer'])['n_outer'].mean()), 'mean ipm', float(res.host(['n_ipm'])['n_ipm'].mean()))
buf = (C.c_ulonglong * 32)()
lib.hdrt_debug_profile(buf, 0)
v = np.array(list(buf), dtype=np.float64)
grid = min(batch, 148 * 3)
fits_block0 = int(np.ceil(batch / grid))  # approximately; the work queue decides
names = {0: 'other(outer)', 1: 'gram', 2: 'qp total', 3: 'hyper', 4: 'weights', 8: 'qp: matvec+residual', 9: 'qp: factor_chol',
         10: 'qp: solves+step', 16: 'fc: load column 0', 17: 'fc: catch-up (updates)', 18: 'fc: diag tile (owner)', 19: 'fc: team barrier + scaling + next column', 20: 'fc: block barrier', 21: 'fc: inverse row (shadow team)',
         5: 'gram: chunks', 6: 'gram: L2 add', 11: 'w: residual', 12: 'w: vmm', 13: 'hyper: s loop', 14: 'hyper: rho loop'}
tot = v[0] + v[1] + v[2] + v[3] + v[4]
print(f'block 0 total cycles {tot:.3e}  fits {v[25]:.0f}  diag calls (warp 0) {v[24]:.0f} -> {v[18] / max(v[24], 1):.0f} cycles each; per fit {tot / max(v[25], 1):.3e}')
for k in sorted(names):
    print(f'{names[k]:24s} {v[k]:.3e}  {100 * v[k] / tot:5.1f}%')
