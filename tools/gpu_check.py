"""Development aid: run every kernel against the oracle / fixtures on the GPU box and print errors + timings."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybdrt_b200 import engine as E, synth  # noqa: E402
from oracle import drt_oracle as orc  # noqa: E402

G = os.path.join(ROOT, 'tests', 'golden')


def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def main():
    eng = E.get_engine(0)
    print('SMs', eng.sm_count, 'fp64 probe TFLOP/s', eng.probe_fp64())
    lk = dict(np.load(os.path.join(G, 'lookup_eps_ppd10.npz')))
    eps = float(lk['eps'])
    tab = eng.build_lookup(eps)
    torch.cuda.synchronize()
    for k in ('re_x', 're_v', 'im_x', 'im_v', 'resp_x', 'resp_v'):
        print('lookup', k, rel(tab[k].cpu().numpy(), lk[k]))
    m = dict(np.load(os.path.join(G, 'matrices.npz')))
    for mode, name in ((E.MODE_INTERP, 'interp'), (E.MODE_TRAPZ, 'trapz')):
        a_re, a_im = eng.build_impedance(m['f_irreg'][None], m['tau_irreg'][None], eps, mode)
        print('imp', name, rel(a_re[0].cpu().numpy(), m[f'irreg_{name}_real']), rel(a_im[0].cpu().numpy(), m[f'irreg_{name}_imag']))
        rm = eng.build_response(m['t_resp'][None], m['tau_resp'][None], m['step_times'][None], m['step_sizes'][None], eps, mode)
        print('resp', name, rel(rm[0].cpu().numpy(), m[f'resp_{name}']))
    a_re, a_im = eng.build_impedance(synth.C2_FREQ[None], m['tau_c2'][None], eps, E.MODE_TRAPZ)
    print('imp c2 trapz', rel(a_re[0].cpu().numpy(), m['c2_trapz_real']), rel(a_im[0].cpu().numpy(), m['c2_trapz_imag']))
    pen = eng.build_penalty(np.log(m['tau_irreg'])[None], eps, False)[0].cpu().numpy()
    pen2 = eng.build_penalty(np.log(m['tau_c2'])[None], eps, True)[0].cpu().numpy()
    for k in range(3):
        print('pen', k, rel(pen[k], m[f'pen_irreg_{k}']), rel(pen2[k], m[f'pen_c2_{k}']))
    print('vmm', rel(eng.build_eis_vmm(m['f_irreg'][None])[0].cpu().numpy(), m['vmm_irreg']),
          rel(eng.build_eis_vmm(m['f_irreg'][None], uniform=True)[0].cpu().numpy(), m['vmm_uniform']))
    d = dict(np.load(os.path.join(G, 'dop.npz')))
    zd = eng.build_dop_z(d['freq'][None], d['basis_nu'], float(d['nu_epsilon']))[0].cpu().numpy()
    print('dop_z', rel(zd, d['zm_dop']))

    # ---- QPHB on C1 + C2 fixtures
    c1 = dict(np.load(os.path.join(G, 'c1_golden.npz')))
    prep = orc.EisPrep(c1['freq'], tables=lk)
    prob, scale = prep.problem(c1['z'])
    dev = eng.dev
    out = eng.qphb_fit_batch(dev(prep.rm), dev(prob['rv'][None]), dev(np.array(prep.pen)), dev(prep.h), dev(prep.l1), 2,
                             vmm_eis=dev(prep.vmm), want_pq=True)
    torch.cuda.synchronize()
    print('c1 status', out['status'].item(), 'n_outer', out['n_outer'].item(), int(c1['n_outer']), 'n_ipm', out['n_ipm'].item(), int(c1['qp_log'].sum()))
    print('c1 x', rel(out['x'][0].cpu().numpy(), c1['cvx_x']), 'w', rel(out['weights'][0].cpu().numpy(), c1['true_weights']),
          'p', rel(out['p_matrix'][0].cpu().numpy(), c1['p_matrix']), 'q', rel(out['q_vector'][0].cpu().numpy(), c1['q_vector']),
          'x_overfit', rel(out['x_overfit'][0].cpu().numpy(), c1['x_overfit_eis']), 'est_w', rel(out['est_weights'][0].cpu().numpy(), c1['est_weights']))

    c2 = dict(np.load(os.path.join(G, 'c2_eis.npz')))
    prep2 = orc.EisPrep(c2['freq'], tables=lk)
    rvs = np.array([prep2.problem(z)[0]['rv'] for z in c2['z']])
    out = eng.qphb_fit_batch(dev(prep2.rm), dev(rvs), dev(np.array(prep2.pen)), dev(prep2.h), dev(prep2.l1), 2, vmm_eis=dev(prep2.vmm))
    torch.cuda.synchronize()
    print('c2 status', out['status'].cpu().numpy())
    print('c2 n_outer', out['n_outer'].cpu().numpy(), c2['n_outer'])
    print('c2 n_ipm', out['n_ipm'].cpu().numpy(), c2['qp_log_total'])
    print('c2 x err', [f"{rel(out['x'][b].cpu().numpy(), c2['cvx_x'][b]):.1e}" for b in range(12)])

    # ---- throughput sweep
    for B in (148, 296, 2048, 10000):
        f, z = synth.make_eis_batch(B, seed=0)
        scale = (z.real.max(axis=1) - z.real.min(axis=1)) / 14.0
        zs = z / scale[:, None]
        rv = dev(np.concatenate([zs.real, zs.imag], axis=1))
        args = (dev(prep2.rm), rv, dev(np.array(prep2.pen)), dev(prep2.h), dev(prep2.l1), 2)
        o = {}
        eng.qphb_fit_batch(*args, vmm_eis=dev(prep2.vmm), out=o)
        torch.cuda.synchronize()
        t = time.time()
        eng.qphb_fit_batch(*args, vmm_eis=dev(prep2.vmm), out=o)
        torch.cuda.synchronize()
        dt = time.time() - t
        no = o['n_outer'].cpu().numpy(); ni = o['n_ipm'].cpu().numpy(); st = o['status'].cpu().numpy()
        print(f'B={B} time {dt*1e3:.1f} ms -> {B/dt:.0f} fits/s; outer mean {no.mean():.1f} ipm mean {ni.mean():.1f}; status counts', np.unique(st, return_counts=True))


if __name__ == '__main__':
    main()
