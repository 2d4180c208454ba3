"""CPU ORACLE support (test infrastructure, not product code).

Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference/hybdrt) on top of oracle/refshim.py.  Run in the authoring container only:

    python -m oracle.make_golden

The reference cannot travel to the GPU box, so the vectors it produces are committed.  Before
anything is written the reference's own golden test (tests/test_drt_fit.py) is executed under the
shim; if it fails, no fixture is produced.
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
from hybdrt.matrices import mat1d, basis, phasance  # noqa: E402
from hybdrt_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def run_reference_golden_test():
    src = open('/root/reference/tests/test_drt_fit.py').read()
    ns = {}
    exec(compile(src, 'test_drt_fit.py', 'exec'), ns)
    ns['test_drt_fit_eis']()
    # pull the inputs and expectations out of the test module source
    body = src[src.index('    z_noisy = np.array(['):src.index('    drt = DRT(')]
    env = {'np': np}
    exec('import numpy as np\n' + '\n'.join(line[4:] for line in body.splitlines()), env)
    return env['z_noisy'], env['expected_result']


def history_arrays(drt):
    h = drt.qphb_history
    return dict(
        hist_x=np.array([e['x'] for e in h]),
        hist_s=np.array([np.array(e['s_vectors']) for e in h]),
        hist_rho=np.array([e['rho_vector'] for e in h]),
        hist_w=np.array([e['weights'] for e in h]),
        hist_fun=np.array([e['fun'] for e in h]),
        hist_ipm=np.array([e['cvx_result']['iterations'] for e in h]),
    )


def fit_outputs(drt, freq=None):
    fp = drt.fit_parameters
    qp = drt.qphb_params
    out = dict(
        cvx_x=np.array(drt.cvx_result['x']),
        x=fp['x'], R_inf=fp['R_inf'], inductance=fp['inductance'],
        q_vector=fp['q_vector'],
        est_weights=qp['est_weights'], init_weights=qp['init_weights'],
        weights=qp['weights'], true_weights=qp['true_weights'],
        xmx_norms=qp['xmx_norms'], rho_vector=qp['rho_vector'],
        s_vectors=np.array(qp['s_vectors']),
        rv=qp['rv'],
        coefficient_scale=drt.coefficient_scale,
        n_outer=len(drt.qphb_history),
        basis_tau=drt.basis_tau,
    )
    if fp.get('z_sigma_tot') is not None:
        out['z_sigma_tot'] = fp['z_sigma_tot']
    if fp.get('v_sigma_tot') is not None:
        out['v_sigma_tot'] = fp['v_sigma_tot']
    for key in ('v_baseline', 'vz_offset', 'x_dop'):
        if key in fp:
            out[key] = np.asarray(fp[key])
    if qp.get('dop_rho_vector') is not None:
        out['dop_rho_vector'] = qp['dop_rho_vector']
        out['dop_xmx_norms'] = qp['dop_xmx_norms']
    if qp['x_overfit_eis'] is not None:
        out['x_overfit_eis'] = qp['x_overfit_eis']
    if qp['x_overfit_chrono'] is not None:
        out['x_overfit_chrono'] = qp['x_overfit_chrono']
    if freq is not None:
        out['z_pred'] = drt.predict_z(freq)
    out.update(history_arrays(drt))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    z_noisy, expected = run_reference_golden_test()
    print('reference golden test passes under the shim')

    # ---- C1: the reference's golden spectrum, with intermediates -------------------------------
    freq = np.logspace(6, -1, 71)
    drt = DRT(fit_inductance=True, fit_ohmic=True)
    lk = drt.interpolate_lookups
    np.savez_compressed(
        os.path.join(OUT, 'lookup_eps_ppd10.npz'),
        eps=drt.tau_epsilon,
        re_x=lk['z_real'][0], re_v=lk['z_real'][1],
        im_x=lk['z_imag'][0], im_v=lk['z_imag'][1],
        resp_x=lk['response'][0], resp_v=lk['response'][1])
    refshim.QP_LOG.clear()
    drt.fit_eis(freq, z_noisy)
    qp = drt.qphb_params
    c1 = fit_outputs(drt, freq)
    c1.update(
        freq=freq, z=z_noisy,
        expected_x=expected['x'], expected_R_inf=expected['R_inf'],
        expected_inductance=expected['inductance'], expected_z_sigma_tot=expected['z_sigma_tot'],
        expected_q_vector=expected['q_vector'],
        zm=drt.fit_matrices['impedance'],
        m0=qp['penalty_matrices']['m0'], m1=qp['penalty_matrices']['m1'], m2=qp['penalty_matrices']['m2'],
        vmm=qp['vmm'], rm=qp['rm'], p_matrix=qp['p_matrix'],
        tau_epsilon=drt.tau_epsilon,
        qp_log=np.array(refshim.QP_LOG),
    )
    np.savez_compressed(os.path.join(OUT, 'c1_golden.npz'), **c1)
    print('c1: outer', c1['n_outer'], 'ipm', refshim.QP_LOG)

    # ---- C2: a few spectra of the benchmark generator ------------------------------------------
    f2, z2 = synth.make_eis_batch(12, seed=0)
    drt2 = DRT()
    rows = []
    for b in range(z2.shape[0]):
        refshim.QP_LOG.clear()
        drt2.fit_eis(f2, z2[b])
        o = fit_outputs(drt2, f2)
        o['qp_log_total'] = int(np.sum(refshim.QP_LOG))
        rows.append(o)
        print('c2', b, 'outer', o['n_outer'], 'ipm total', o['qp_log_total'])
    keys = ['cvx_x', 'x', 'R_inf', 'inductance', 'q_vector', 'est_weights', 'init_weights', 'weights',
            'xmx_norms', 'rho_vector', 's_vectors', 'rv', 'coefficient_scale', 'n_outer', 'z_sigma_tot',
            'x_overfit_eis', 'z_pred', 'qp_log_total']
    c2 = {k: np.array([r[k] for r in rows]) for k in keys}
    c2['hist_fun_last'] = np.array([r['hist_fun'][-1] for r in rows])
    c2['hist_ipm_flat'] = np.concatenate([r['hist_ipm'] for r in rows])
    c2.update(freq=f2, z=z2, basis_tau=rows[0]['basis_tau'],
              rm=drt2.qphb_params['rm'], vmm=drt2.qphb_params['vmm'])
    np.savez_compressed(os.path.join(OUT, 'c2_eis.npz'), **c2)

    # nonneg=False and a non-default-hyper fit on one spectrum
    drt2.fit_eis(f2, z2[0], nonneg=False)
    o = fit_outputs(drt2, f2)
    np.savez_compressed(os.path.join(OUT, 'c2_free.npz'), freq=f2, z=z2[0],
                        **{k: o[k] for k in ['cvx_x', 'x', 'R_inf', 'inductance', 'weights', 'n_outer', 'z_pred',
                                             'rho_vector', 's_vectors']})

    # ---- matrix builders: trapz mode + non-Toeplitz grids --------------------------------------
    eps = drt.tau_epsilon
    f_irreg = np.array([9.3e5, 2.1e5, 4.4e4, 7.7e3, 1.3e3, 310.0, 55.0, 9.1, 2.2, 0.37, 0.081, 0.013])
    tau_irreg = np.logspace(-7, 2, 15) * (1 + 0.05 * np.sin(np.arange(15)))
    tz = dict(eps=eps, f_irreg=f_irreg, tau_irreg=tau_irreg)
    for part in ('real', 'imag'):
        tz[f'irreg_trapz_{part}'] = mat1d.construct_impedance_matrix(
            f_irreg, part, tau=tau_irreg, epsilon=eps, integrate_method='trapz')
        tz[f'irreg_interp_{part}'] = mat1d.construct_impedance_matrix(
            f_irreg, part, tau=tau_irreg, epsilon=eps, integrate_method='interp',
            interpolate_grids=lk['z_real' if part == 'real' else 'z_imag'])
    f_c2 = synth.C2_FREQ
    tau_c2 = rows[0]['basis_tau']
    tz['tau_c2'] = tau_c2
    for part in ('real', 'imag'):
        tz[f'c2_trapz_{part}'] = mat1d.construct_impedance_matrix(
            f_c2, part, tau=tau_c2, epsilon=eps, integrate_method='trapz')
    t_resp = np.concatenate([np.linspace(-0.004, 0.0, 5), np.logspace(-4, 0.5, 24)])
    step_times = np.array([-1e-6, 0.05])
    step_sizes = np.array([0.01, -0.004])
    tau_resp = np.logspace(-6, 1.5, 17)
    tz.update(t_resp=t_resp, step_times=step_times, step_sizes=step_sizes, tau_resp=tau_resp)
    tz['resp_trapz'], _ = mat1d.construct_response_matrix(
        tau_resp, t_resp, 'ideal', step_times, step_sizes, epsilon=eps, integrate_method='trapz')
    tz['resp_interp'], _ = mat1d.construct_response_matrix(
        tau_resp, t_resp, 'ideal', step_times, step_sizes, epsilon=eps, integrate_method='interp',
        interpolate_grids=lk['response'])
    for k in range(3):
        tz[f'pen_irreg_{k}'] = mat1d.construct_integrated_derivative_matrix(
            np.log(tau_irreg), order=k, epsilon=eps)
        tz[f'pen_c2_{k}'] = mat1d.construct_integrated_derivative_matrix(
            np.log(tau_c2), order=k, epsilon=eps)
    tz['vmm_irreg'] = mat1d.construct_eis_var_matrix(f_irreg, 0.25, 0.25, None)
    tz['vmm_uniform'] = mat1d.construct_eis_var_matrix(f_irreg, 0.25, 0.25, 'uniform')
    np.savez_compressed(os.path.join(OUT, 'matrices.npz'), **tz)

    # trapz-mode fit (interpolate_integrals=False) on one C2 spectrum
    drt_t = DRT(interpolate_integrals=False)
    drt_t.fit_eis(f2, z2[1])
    o = fit_outputs(drt_t, f2)
    np.savez_compressed(os.path.join(OUT, 'c2_trapz_fit.npz'), freq=f2, z=z2[1],
                        **{k: o[k] for k in ['cvx_x', 'x', 'R_inf', 'inductance', 'weights', 'n_outer', 'z_pred']})

    # ---- DOP -------------------------------------------------------------------------------------
    fd, zd = synth.make_dop_batch(3, seed=2)
    drt_d = DRT(fit_dop=True)
    rows = []
    for b in range(zd.shape[0]):
        refshim.QP_LOG.clear()
        drt_d.fit_eis(fd, zd[b])
        o = fit_outputs(drt_d, fd)
        o['qp_log_total'] = int(np.sum(refshim.QP_LOG))
        rows.append(o)
        print('dop', b, 'outer', o['n_outer'], 'ipm total', o['qp_log_total'])
    keys = ['cvx_x', 'x', 'x_dop', 'R_inf', 'inductance', 'weights', 'est_weights', 'xmx_norms', 'dop_xmx_norms',
            'rho_vector', 'dop_rho_vector', 's_vectors', 'rv', 'coefficient_scale', 'n_outer', 'z_pred',
            'x_overfit_eis', 'qp_log_total']
    dd = {k: np.array([r[k] for r in rows]) for k in keys}
    qpd = drt_d.qphb_params
    dd.update(freq=fd, z=zd, basis_tau=rows[0]['basis_tau'], basis_nu=drt_d.basis_nu,
              nu_epsilon=drt_d.nu_epsilon, dop_scale_vector=drt_d.dop_scale_vector,
              zm_dop=drt_d.fit_matrices['zm_dop'], rm=qpd['rm'],
              m0=qpd['penalty_matrices']['m0'], m1=qpd['penalty_matrices']['m1'],
              m2=qpd['penalty_matrices']['m2'])
    np.savez_compressed(os.path.join(OUT, 'dop.npz'), **dd)

    # ---- hybrid + chrono ----------------------------------------------------------------------
    th, ih, vh, fh, zh = synth.make_hybrid_batch(2, seed=1)
    # small case (every 5th sample after the step keeps the fixture small) with intermediates
    keep = np.concatenate([np.arange(0, 12), np.arange(12, th.size, 5)])
    ts, is_, vs = th[keep], ih[keep], vh[:, keep]
    drt_h = DRT()
    refshim.QP_LOG.clear()
    drt_h.fit_hybrid(ts, is_, vs[0], fh, zh[0])
    o = fit_outputs(drt_h, fh)
    qph = drt_h.qphb_params
    hs = {k: o[k] for k in ['cvx_x', 'x', 'R_inf', 'inductance', 'v_baseline', 'vz_offset', 'weights', 'est_weights',
                            'xmx_norms', 'rho_vector', 's_vectors', 'rv', 'coefficient_scale', 'n_outer', 'z_pred',
                            'z_sigma_tot', 'v_sigma_tot', 'basis_tau', 'hist_x', 'hist_ipm']}
    hs.update(times=ts, i_signal=is_, v_signal=vs[0], freq=fh, z=zh[0],
              rm=qph['rm'], vz_strength_vec=qph['vz_strength_vec'],
              step_times=drt_h.step_times, step_sizes=drt_h.step_sizes,
              response_signal_scale=drt_h.response_signal_scale, input_signal_scale=drt_h.input_signal_scale,
              scaled_response_offset=drt_h.scaled_response_offset, qp_log=np.array(refshim.QP_LOG),
              v_pred=drt_h.predict_response(ts))
    np.savez_compressed(os.path.join(OUT, 'hybrid_small.npz'), **hs)
    print('hybrid small: outer', o['n_outer'], 'ipm', int(np.sum(refshim.QP_LOG)))

    # full-size C3 case, outputs only
    rows = []
    for b in range(2):
        refshim.QP_LOG.clear()
        drt_h.fit_hybrid(th, ih, vh[b], fh, zh[b])
        o = fit_outputs(drt_h, fh)
        o['qp_log_total'] = int(np.sum(refshim.QP_LOG))
        rows.append(o)
        print('hybrid full', b, 'outer', o['n_outer'], 'ipm total', o['qp_log_total'])
    keys = ['cvx_x', 'x', 'R_inf', 'inductance', 'v_baseline', 'vz_offset', 'xmx_norms', 'rho_vector',
            'coefficient_scale', 'n_outer', 'z_pred', 'qp_log_total']
    hf = {k: np.array([r[k] for r in rows]) for k in keys}
    hf.update(basis_tau=rows[0]['basis_tau'], seed=1, batch=2)
    np.savez_compressed(os.path.join(OUT, 'hybrid_full.npz'), **hf)

    # chrono-only fit on the small trace
    refshim.QP_LOG.clear()
    drt_h.fit_chrono(ts, is_, vs[1])
    o = fit_outputs(drt_h)
    hc = {k: o[k] for k in ['cvx_x', 'x', 'R_inf', 'inductance', 'v_baseline', 'weights', 'est_weights',
                            'xmx_norms', 'rho_vector', 's_vectors', 'rv', 'coefficient_scale', 'n_outer',
                            'v_sigma_tot', 'basis_tau']}
    hc.update(times=ts, i_signal=is_, v_signal=vs[1], rm=drt_h.qphb_params['rm'],
              v_pred=drt_h.predict_response(ts), qp_log=np.array(refshim.QP_LOG))
    np.savez_compressed(os.path.join(OUT, 'chrono_small.npz'), **hc)
    print('chrono: outer', o['n_outer'])

    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)) // 1024, 'KiB')


if __name__ == '__main__':
    main()
