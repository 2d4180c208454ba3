"""CPU ORACLE support (test infrastructure, not product code).

Second fixture set, produced like oracle/make_golden.py by the UNMODIFIED reference on top of
oracle/refshim.py (authoring container only):

    python -m oracle.make_golden_ext [name ...]

* outlier_eis.npz     fit_eis with outlier_p (qphb.py:1497-1538, 1629-1655) on spectra with injected outliers,
                      and the remove_outliers two-pass flow (drt1d.py:217-303)
* chrono_flex.npz     construct_chrono_var_matrix with error_structure=None (mat1d.py:455-490) and the
                      fit_chrono / fit_hybrid results with chrono_error_structure=None (+ outlier_p: the
                      tutorial's flags)
* rescale.npz         solve_rp (drt1d.py:573-607, qphb.py:1684-1717) and update_scale (drt1d.py:914-936) fits
* dop_chrono.npz      phasance.construct_phasor_v_matrix (phasance.py:121-144) and DRT + DOP fits of time-domain data
* pfrt.npz            pfrt_fit_eis / pfrt_fit_hybrid (drt1d.py:2558-2716): per-factor coefficients, marginal
                      log-likelihood and P matrices of the continuation path
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402

refshim.install()
warnings.filterwarnings('ignore')

from hybdrt.models import DRT  # noqa: E402  (the reference)
from hybdrt.matrices import mat1d  # noqa: E402
from hybdrt_b200 import synth  # noqa: E402
from oracle.make_golden import fit_outputs  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def outlier_spectra():
    """Three C2 spectra with gross outliers at fixed positions (sizes relative to Rp)."""
    freq, z = synth.make_eis_batch(3, seed=4)
    z = z.copy()
    rp = z.real.max(axis=1) - z.real.min(axis=1)
    z[0, 20] += 0.08 * rp[0]
    z[0, 45] -= 0.05j * rp[0]
    z[1, 33] += (0.06 - 0.06j) * rp[1]
    z[2, 10] -= 0.1 * rp[2]
    z[2, 11] += 0.07j * rp[2]
    z[2, 60] += 0.05 * rp[2]
    return freq, z


def small_hybrid():
    th, ih, vh, fh, zh = synth.make_hybrid_batch(2, seed=1)
    keep = np.concatenate([np.arange(0, 12), np.arange(12, th.size, 5)])
    return th[keep], ih[keep], vh[:, keep], fh, zh


def gen_outlier_eis():
    freq, z = outlier_spectra()
    drt = DRT()
    rows = []
    for b in range(z.shape[0]):
        refshim.QP_LOG.clear()
        drt.fit_eis(freq, z[b], outlier_p=0.01)
        o = fit_outputs(drt, freq)
        o['qp_log_total'] = int(np.sum(refshim.QP_LOG))
        o['outlier_t'] = np.array(drt.qphb_history[-1]['outlier_t'])
        rows.append(o)
        print('outlier eis', b, 'outer', o['n_outer'], 'ipm', o['qp_log_total'],
              'min t', o['outlier_t'].min())
    keys = ['cvx_x', 'x', 'R_inf', 'inductance', 'est_weights', 'init_weights', 'weights', 'rho_vector', 's_vectors',
            'n_outer', 'z_pred', 'x_overfit_eis', 'qp_log_total', 'outlier_t', 'coefficient_scale', 'rv']
    d = {k: np.array([r[k] for r in rows]) for k in keys}
    d.update(freq=freq, z=z, outlier_p=0.01)
    # remove_outliers: first pass flags points from the initialisation, second pass refits without them
    ro = []
    for b in range(z.shape[0]):
        drt.fit_eis(freq, z[b], outlier_p=0.01, remove_outliers=True, outlier_thresh=0.75)
        o = fit_outputs(drt)
        idx = np.zeros(freq.size, dtype=bool) if drt.eis_outlier_index is None else np.asarray(drt.eis_outlier_index)
        ro.append(dict(index=idx, x=o['x'], R_inf=o['R_inf'], inductance=o['inductance'], n_outer=o['n_outer'],
                       z_pred=drt.predict_z(freq)))
        print('remove_outliers', b, 'flagged', np.where(idx)[0].tolist(), 'outer', o['n_outer'])
    d['ro_index'] = np.array([r['index'] for r in ro])
    for k in ('x', 'R_inf', 'inductance', 'n_outer', 'z_pred'):
        d[f'ro_{k}'] = np.array([r[k] for r in ro])
    np.savez_compressed(os.path.join(OUT, 'outlier_eis.npz'), **d)


def gen_chrono_flex():
    ts, is_, vs, fh, zh = small_hybrid()
    d = dict(times=ts, i_signal=is_, v_signal=vs, freq=fh, z=zh)
    # the matrix alone, on a two-step protocol as well
    t2 = np.concatenate([np.linspace(-0.004, -0.001, 4), np.linspace(0, 0.0495, 100), np.linspace(0.05, 0.3, 120)])
    st2 = np.array([0.0, 0.05])
    d.update(t2=t2, st2=st2, vmm_two_step_rows=mat1d.construct_chrono_var_matrix(t2, st2, 4, None)[::8])   # every 8th row
    drt = DRT()
    refshim.QP_LOG.clear()
    drt.fit_chrono(ts, is_, vs[1], error_structure=None)
    o = fit_outputs(drt)
    d['vmm_chrono_rows'] = drt.qphb_params['vmm'][::8]
    d['step_times'] = drt.step_times
    for k in ('cvx_x', 'x', 'R_inf', 'inductance', 'v_baseline', 'weights', 'est_weights', 'n_outer', 'rho_vector'):
        d[f'chrono_{k}'] = np.asarray(o[k])
    d['chrono_v_pred'] = drt.predict_response(ts)
    d['chrono_ipm'] = int(np.sum(refshim.QP_LOG))
    print('chrono flex: outer', o['n_outer'], 'ipm', d['chrono_ipm'])
    refshim.QP_LOG.clear()
    drt.fit_hybrid(ts, is_, vs[0], fh, zh[0], chrono_error_structure=None)
    o = fit_outputs(drt, fh)
    for k in ('cvx_x', 'x', 'R_inf', 'inductance', 'v_baseline', 'vz_offset', 'weights', 'est_weights', 'n_outer',
              'rho_vector', 'z_pred'):
        d[f'hybrid_{k}'] = np.asarray(o[k])
    d['hybrid_ipm'] = int(np.sum(refshim.QP_LOG))
    print('hybrid flex: outer', o['n_outer'], 'ipm', d['hybrid_ipm'])
    # the tutorial's flags (Probabilistic_DRT_fitting.ipynb): flexible chrono errors + outlier_p, with a spike
    v_sp = vs[0].copy()
    v_sp[40] += 400e-6
    v_sp[90] -= 300e-6
    refshim.QP_LOG.clear()
    drt.fit_hybrid(ts, is_, v_sp, fh, zh[0], chrono_error_structure=None, outlier_p=0.01)
    o = fit_outputs(drt, fh)
    d['tut_v_signal'] = v_sp
    for k in ('cvx_x', 'x', 'R_inf', 'inductance', 'v_baseline', 'vz_offset', 'weights', 'est_weights', 'n_outer',
              'rho_vector', 'z_pred'):
        d[f'tut_{k}'] = np.asarray(o[k])
    d['tut_outlier_t'] = np.array(drt.qphb_history[-1]['outlier_t'])
    d['tut_ipm'] = int(np.sum(refshim.QP_LOG))
    print('tutorial flags: outer', o['n_outer'], 'ipm', d['tut_ipm'], 'min t', d['tut_outlier_t'].min())
    np.savez_compressed(os.path.join(OUT, 'chrono_flex.npz'), **d)


def gen_pfrt():
    """DRT.pfrt_fit_eis / pfrt_fit_hybrid (drt1d.py:2558-2716): per-factor x, llh, P."""
    freq, z = synth.make_eis_batch(2, seed=0)
    drt = DRT()
    rows = []
    for b in range(2):
        refshim.QP_LOG.clear()
        drt.pfrt_fit_eis(freq, z[b])
        r = drt.pfrt_result
        rows.append(dict(step_x=np.array(r['step_x']), step_llh=np.array(r['step_llh']),
                         step_p_diag=np.array([np.diag(p) for p in r['step_p_mat']]),
                         step_p_last=np.array(r['step_p_mat'][-1]),
                         n_hist=len(drt.pfrt_history), ipm=int(np.sum(refshim.QP_LOG)),
                         init_n_outer=len(drt.qphb_history), coefficient_scale=drt.coefficient_scale))
        print('pfrt eis', b, 'history', rows[-1]['n_hist'], 'init outer', rows[-1]['init_n_outer'], 'ipm', rows[-1]['ipm'])
    d = {k: np.array([r[k] for r in rows]) for k in rows[0]}
    d.update(freq=freq, z=z, factors=np.array(drt.pfrt_result['factors']))
    ts, is_, vs, fh, zh = small_hybrid()
    refshim.QP_LOG.clear()
    drt.pfrt_fit_hybrid(ts, is_, vs[0], fh, zh[0], factors=np.logspace(-0.6, 0.6, 5))
    r = drt.pfrt_result
    d.update(h_times=ts, h_i=is_, h_v=vs[0], h_freq=fh, h_z=zh[0], h_factors=np.array(r['factors']),
             h_step_x=np.array(r['step_x']), h_step_llh=np.array(r['step_llh']),
             h_step_p_diag=np.array([np.diag(p) for p in r['step_p_mat']]), h_n_hist=len(drt.pfrt_history),
             h_ipm=int(np.sum(refshim.QP_LOG)))
    print('pfrt hybrid: history', d['h_n_hist'], 'ipm', d['h_ipm'])
    np.savez_compressed(os.path.join(OUT, 'pfrt.npz'), **d)


def gen_rescale():
    """solve_rp (drt1d.py:573-607) and update_scale (:914-936): EIS, DRT+DOP and hybrid."""
    d = {}
    freq, z = synth.make_eis_batch(2, seed=6)
    z = z * np.array([3.0, 0.2])[:, None]            # Rp estimates that start away from the target scale
    drt = DRT()
    cases = (('us', dict(update_scale=True)), ('rp', dict(solve_rp=True)), ('both', dict(solve_rp=True, update_scale=True)))
    for tag, kw in cases:
        rows = []
        for b in range(2):
            refshim.QP_LOG.clear()
            drt.fit_eis(freq, z[b], **kw)
            o = fit_outputs(drt, freq)
            o['ipm'] = int(np.sum(refshim.QP_LOG))
            rows.append(o)
            print('rescale eis', tag, b, 'outer', o['n_outer'], 'scale', o['coefficient_scale'])
        for k in ('cvx_x', 'x', 'R_inf', 'inductance', 'weights', 'est_weights', 'init_weights', 'n_outer', 'z_pred',
                  'coefficient_scale', 'rv', 'x_overfit_eis', 'xmx_norms', 'ipm', 'q_vector'):
            d[f'{tag}_{k}'] = np.array([r[k] for r in rows])
    d.update(freq=freq, z=z)
    fd, zd = synth.make_dop_batch(2, seed=2)
    drt_d = DRT(fit_dop=True)
    rows = []
    for b in range(2):
        refshim.QP_LOG.clear()
        drt_d.fit_eis(fd, zd[b], solve_rp=True)
        o = fit_outputs(drt_d, fd)
        o['ipm'] = int(np.sum(refshim.QP_LOG))
        o['dop_scale_vector'] = drt_d.dop_scale_vector.copy()
        o['rm'] = drt_d.qphb_params['rm'].copy()
        rows.append(o)
        print('rescale dop', b, 'outer', o['n_outer'], 'scale', o['coefficient_scale'])
    for k in ('cvx_x', 'x', 'x_dop', 'R_inf', 'inductance', 'weights', 'n_outer', 'z_pred', 'coefficient_scale',
              'dop_scale_vector', 'ipm', 'dop_rho_vector'):
        d[f'dop_{k}'] = np.array([r[k] for r in rows])
    d['dop_rm0'] = rows[0]['rm']
    d.update(dop_freq=fd, dop_z=zd)
    ts, is_, vs, fh, zh = small_hybrid()
    refshim.QP_LOG.clear()
    drt.fit_hybrid(ts, is_, vs[0], fh, zh[0], solve_rp=True, update_scale=True)
    o = fit_outputs(drt, fh)
    for k in ('cvx_x', 'x', 'R_inf', 'v_baseline', 'vz_offset', 'n_outer', 'z_pred', 'coefficient_scale', 'weights'):
        d[f'hyb_{k}'] = np.asarray(o[k])
    d.update(hyb_ipm=int(np.sum(refshim.QP_LOG)), hyb_response_signal_scale=drt.response_signal_scale,
             hyb_scaled_response_offset=drt.scaled_response_offset, hyb_v_pred=drt.predict_response(ts))
    print('rescale hybrid: outer', o['n_outer'], 'scale', o['coefficient_scale'])
    np.savez_compressed(os.path.join(OUT, 'rescale.npz'), **d)


def gen_dop_chrono():
    """DRT + DOP with time-domain data: phasance.construct_phasor_v_matrix and fit_chrono / fit_hybrid."""
    from hybdrt.matrices import phasance
    ts, is_, vs, fh, zh = small_hybrid()
    drt = DRT(fit_dop=True)
    refshim.QP_LOG.clear()
    drt.fit_hybrid(ts, is_, vs[0], fh, zh[0])
    o = fit_outputs(drt, fh)
    d = dict(times=ts, i_signal=is_, v_signal=vs, freq=fh, z=zh, basis_nu=drt.basis_nu, nu_epsilon=drt.nu_epsilon,
             step_times=drt.step_times, step_sizes=drt.step_sizes,
             rm_dop=phasance.construct_phasor_v_matrix(ts, drt.basis_nu, 'gaussian', drt.nu_epsilon, 'ideal',
                                                       drt.step_times, drt.step_sizes)[0],
             hyb_rm=drt.qphb_params['rm'], hyb_ipm=int(np.sum(refshim.QP_LOG)), hyb_v_pred=drt.predict_response(ts))
    for k in ('cvx_x', 'x', 'x_dop', 'R_inf', 'vz_offset', 'n_outer', 'z_pred', 'weights', 'dop_rho_vector'):
        d[f'hyb_{k}'] = np.asarray(o[k])
    print('dop hybrid: outer', o['n_outer'], 'ipm', d['hyb_ipm'])
    refshim.QP_LOG.clear()
    drt.fit_chrono(ts, is_, vs[1])
    o = fit_outputs(drt)
    for k in ('cvx_x', 'x', 'x_dop', 'R_inf', 'n_outer', 'weights'):
        d[f'chr_{k}'] = np.asarray(o[k])
    d.update(chr_ipm=int(np.sum(refshim.QP_LOG)), chr_v_pred=drt.predict_response(ts))
    print('dop chrono: outer', o['n_outer'], 'ipm', d['chr_ipm'])
    np.savez_compressed(os.path.join(OUT, 'dop_chrono.npz'), **d)


GENERATORS = dict(dop_chrono=gen_dop_chrono, outlier_eis=gen_outlier_eis, chrono_flex=gen_chrono_flex, pfrt=gen_pfrt, rescale=gen_rescale)


def main():
    names = sys.argv[1:] or list(GENERATORS)
    for name in names:
        GENERATORS[name]()
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)) // 1024, 'KiB')


if __name__ == '__main__':
    main()
