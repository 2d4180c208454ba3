"""Development aid: one batched fit of a BASELINE configuration (for ncu captures).  usage: python tools/config_run.py c2|c3|c4 [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybdrt_b200 import synth  # noqa: E402
from hybdrt_b200.models import DRT  # noqa: E402

which = sys.argv[1]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 592
if which == 'c2':
    f, z = synth.make_eis_batch(batch, seed=0)
    drt = DRT()
    fit = lambda: drt.fit_eis_batch(f, z)
elif which == 'c3':
    t, i_sig, v, f, z = synth.make_hybrid_batch(batch, seed=1)
    drt = DRT()
    fit = lambda: drt.fit_hybrid_batch(t, i_sig, v, f, z)
else:
    f, z = synth.make_dop_batch(batch, seed=2)
    drt = DRT(fit_dop=True)
    fit = lambda: drt.fit_eis_batch(f, z)
for _ in range(3):
    r = fit()
    torch.cuda.synchronize()
print(which, batch, 'mean outer', float(r.host(['n_outer'])['n_outer'].mean()))
