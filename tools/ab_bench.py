"""Development aid: A/B the resident-input QPHB throughput of several builds of the library in one process.
usage: python tools/ab_bench.py libA.so libB.so ...  (paths relative to hybrid-drt_b200/_lib)"""
import ctypes as C
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybdrt_b200 import engine as E, synth  # noqa: E402
from hybdrt_b200.models import DRT  # noqa: E402
CONFIG = os.environ.get('AB_CONFIG', 'c2')          # c2 (default), c3 (hybrid) or c4 (DRT + DOP, n = 153)
B = 10000 if CONFIG == 'c2' else 2960
freq, z = synth.make_eis_batch(B, seed=0) if CONFIG == 'c2' else synth.make_dop_batch(B, seed=2)
libs = sys.argv[1:]
res = {}
if CONFIG == 'c3':
    B = int(os.environ.get('AB_BATCH', '1184'))
    hyb = synth.make_hybrid_batch(B, seed=1)
for rep in range(2):
    for name in libs:
        E._lib = None
        E._engines.clear()
        E.LIB_PATH = os.path.join(ROOT, 'hybrid-drt_b200', '_lib', name)
        if CONFIG == 'c3':     # hybrid fits: the launch the public API made, repeated on the resident inputs
            r0 = DRT().fit_hybrid_batch(*hyb)
            torch.cuda.synchronize()
            relaunch, out = r0.extra['relaunch'], {}
            relaunch(out); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(2):
                relaunch(out)
            b.record(); torch.cuda.synchronize()
            fps = 2 * B / (a.elapsed_time(b) * 1e-3)
            res.setdefault(name, []).append(fps)
            print(f'{name:36s} {fps:10.0f} fits/s   x checksum {float(out["x"].sum()):.12e}', flush=True)
            continue
        drt = DRT() if CONFIG == 'c2' else DRT(fit_dop=True)
        r0 = drt.fit_eis_batch(freq, z)
        plan = r0.plan
        zs = z / r0.scales['coefficient_scale'][:, None]
        eng = drt.engine
        rv = eng.dev(np.concatenate([zs.real, zs.imag], axis=1))
        hyp = drt._c_hypers(plan['opts'])
        out = {}
        def step():
            eng.qphb_fit_batch(plan['rm'], rv, plan['pen'], plan['h'], plan['l1'], plan['n_special'], vmm_eis=plan['vmm_eis'], hypers=hyp, out=out, pen_hint=plan.get('pen_hint'),
                               dop_range=(drt.dop_indices if CONFIG != 'c2' else None))
        step(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            step()
        b.record(); torch.cuda.synchronize()
        fps = 3 * B / (a.elapsed_time(b) * 1e-3)
        res.setdefault(name, []).append(fps)
        print(f'{name:36s} {fps:10.0f} fits/s   x checksum {float(out["x"].sum()):.12e}', flush=True)
