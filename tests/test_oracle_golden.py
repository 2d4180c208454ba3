"""The CPU oracle (oracle/) against fixtures produced by the UNMODIFIED reference.

Fixtures come from oracle/make_golden.py (reference run under oracle/refshim.py) and include the
reference's own golden vector (reference tests/test_drt_fit.py:55-141).  No GPU needed.
"""
import numpy as np
import pytest

from conftest import load_golden, rel_err
from oracle import drt_oracle as orc
from oracle.coneqp import coneqp_orthant


@pytest.fixture(scope='module')
def tables(lookup_golden):
    return lookup_golden


def test_lookup_tables_match_reference(lookup_golden):
    tab = orc.lookup_tables(float(lookup_golden['eps']))
    for key in ('re_x', 're_v', 'im_x', 'im_v', 'resp_x', 'resp_v'):
        assert rel_err(tab[key], lookup_golden[key]) < 1e-13, key


def test_matrix_builders_match_reference(tables):
    m = load_golden('matrices.npz')
    eps = float(m['eps'])
    for part in ('real', 'imag'):
        got = orc.impedance_matrix(m['f_irreg'], m['tau_irreg'], eps, part, 'interp', tables)
        assert rel_err(got, m[f'irreg_interp_{part}']) < 1e-13
        got = orc.impedance_matrix(m['f_irreg'], m['tau_irreg'], eps, part, 'trapz')
        assert rel_err(got, m[f'irreg_trapz_{part}']) < 1e-13
    args = (m['tau_resp'], m['t_resp'], m['step_times'], m['step_sizes'], eps)
    assert rel_err(orc.response_matrix(*args, 'interp', tables), m['resp_interp']) < 1e-13
    assert rel_err(orc.response_matrix(*args, 'trapz'), m['resp_trapz']) < 1e-13
    for k in range(3):
        assert rel_err(orc.penalty_matrix(np.log(m['tau_irreg']), k, eps), m[f'pen_irreg_{k}']) < 1e-13
        assert rel_err(orc.penalty_matrix(np.log(m['tau_c2']), k, eps), m[f'pen_c2_{k}']) < 1e-13
    assert rel_err(orc.eis_vmm(m['f_irreg']), m['vmm_irreg']) < 1e-13
    assert rel_err(orc.eis_vmm(m['f_irreg'], structure='uniform'), m['vmm_uniform']) < 1e-13


def test_reference_golden_vector(tables):
    """The hard-coded expectations of the reference's tests/test_drt_fit.py, same tolerance."""
    c1 = load_golden('c1_golden.npz')
    prep = orc.EisPrep(c1['freq'], tables=tables)
    assert rel_err(prep.zm, c1['zm']) < 1e-13
    assert rel_err(prep.rm, c1['rm']) < 1e-13
    assert rel_err(prep.vmm, c1['vmm']) < 1e-13
    for k in range(3):
        assert rel_err(prep.pen[k], c1[f'm{k}']) < 1e-13
    res = prep.fit(c1['z'])
    p = res['params']
    assert np.allclose(c1['expected_x'], p['x'])                      # np.allclose as in the reference test
    assert np.allclose(c1['expected_R_inf'], p['R_inf'])
    assert np.allclose(c1['expected_inductance'], p['inductance'])
    assert np.allclose(c1['expected_z_sigma_tot'], p['z_sigma_tot'])
    assert np.allclose(c1['expected_q_vector'], res['q_vector'])
    # and against the reference run here, to far tighter tolerance
    assert res['n_outer'] == int(c1['n_outer'])
    assert list(res['ipm_iters']) == list(c1['qp_log'])
    assert rel_err(res['x'], c1['cvx_x']) < 1e-9
    assert rel_err(res['weights'], c1['weights']) < 1e-9
    assert rel_err(res['p_matrix'], c1['p_matrix']) < 1e-9
    assert rel_err(prep.predict_z(p), c1['z_pred']) < 1e-9


def test_c2_spectra_match_reference(tables):
    c2 = load_golden('c2_eis.npz')
    prep = orc.EisPrep(c2['freq'], tables=tables)
    for b in range(c2['z'].shape[0]):
        res = prep.fit(c2['z'][b])
        assert res['n_outer'] == int(c2['n_outer'][b])
        assert int(res['ipm_iters'].sum()) == int(c2['qp_log_total'][b])
        assert rel_err(res['x'], c2['cvx_x'][b]) < 1e-9
        assert rel_err(res['weights'], c2['weights'][b]) < 1e-9
        assert abs(res['fun'] - c2['hist_fun_last'][b]) <= 1e-9 * abs(c2['hist_fun_last'][b])


def test_free_sign_fit_matches_reference(tables):
    g = load_golden('c2_free.npz')
    prep = orc.EisPrep(g['freq'], tables=tables, nonneg=False)
    res = prep.fit(g['z'])
    assert res['n_outer'] == int(g['n_outer'])
    assert rel_err(res['x'], g['cvx_x']) < 1e-9


def test_trapz_mode_fit_matches_reference():
    g = load_golden('c2_trapz_fit.npz')
    prep = orc.EisPrep(g['freq'], mode='trapz')
    res = prep.fit(g['z'])
    assert res['n_outer'] == int(g['n_outer'])
    assert rel_err(res['x'], g['cvx_x']) < 1e-9


def test_dop_matches_reference():
    d = load_golden('dop.npz')
    nu_eps = float(d['nu_epsilon'])
    assert rel_err(orc.dop_z_matrix(d['freq'], d['basis_nu'], nu_eps), d['zm_dop']) < 1e-13
    scale = orc.dop_scale_vector(d['basis_nu'], d['basis_tau']) / (np.sqrt(np.pi) / nu_eps)
    assert rel_err(scale, d['dop_scale_vector']) < 1e-13
    n = d['rm'].shape[1]
    b = 1
    prob = dict(rm=d['rm'], rv=d['rv'][b], vmm=orc.eis_vmm(d['freq']), pen=[d['m0'], d['m1'], d['m2']],
                h=np.zeros(n), l1=np.zeros(n), n_special=52, dop_range=(2, 52))
    res = orc.qphb_fit(prob)
    assert res['n_outer'] == int(d['n_outer'][b])
    assert int(res['ipm_iters'].sum()) == int(d['qp_log_total'][b])
    assert rel_err(res['x'], d['cvx_x'][b]) < 1e-8
    assert rel_err(res['dop_rho'], d['dop_rho_vector'][b]) < 1e-8


def _layout_pen(tau, eps, diag):
    ns = len(diag)
    n = ns + tau.size
    pen = []
    for k in range(3):
        mk = np.zeros((n, n))
        mk[np.arange(ns), np.arange(ns)] = diag
        mk[ns:, ns:] = orc.penalty_matrix(np.log(tau), k, eps)
        pen.append(mk)
    return pen


def test_hybrid_matches_reference(tables):
    hs = load_golden('hybrid_small.npz')
    eps = float(tables['eps'])
    rm0 = hs['rm'].copy()
    rm0[:, 1] = 0                                   # vz_offset column starts at zero (drt1d.py:5809)
    n = rm0.shape[1]
    nc = hs['times'].size
    pen = _layout_pen(hs['basis_tau'], eps, [1e-6, 1.0, 1e-6, 1e-6])
    h = np.zeros(n)
    h[:2] = 1000                                    # v_baseline, vz_offset unbounded (qphb.py:530-533)
    prob = dict(rm=rm0, rv=hs['rv'], vmm=dict(n_chrono=nc, chrono=None, eis=orc.eis_vmm(hs['freq'])),
                pen=pen, h=h, l1=np.zeros(n), n_special=4, vz_index=1, vb_range=(0, 1),
                vz_strength=hs['vz_strength_vec'], n_chrono=nc)
    res = orc.qphb_fit(prob)
    assert res['n_outer'] == int(hs['n_outer'])
    assert list(res['ipm_iters']) == list(hs['qp_log'])
    assert rel_err(res['x'], hs['cvx_x']) < 1e-8
    assert rel_err(res['weights'], hs['weights']) < 1e-8
    assert rel_err(res['rm_final'], hs['rm']) < 1e-10


def test_chrono_matches_reference(tables):
    cs = load_golden('chrono_small.npz')
    eps = float(tables['eps'])
    rm = cs['rm']
    n_rows, n = rm.shape
    pen = _layout_pen(cs['basis_tau'], eps, [1e-6, 1e-6, 1e-6])
    h = np.zeros(n)
    h[0] = 1000
    prob = dict(rm=rm, rv=cs['rv'], vmm=dict(n_chrono=n_rows, chrono=None, eis=None), pen=pen, h=h,
                l1=np.zeros(n), n_special=3, n_chrono=n_rows)
    res = orc.qphb_fit(prob)
    assert res['n_outer'] == int(cs['n_outer'])
    assert rel_err(res['x'], cs['cvx_x']) < 1e-8


def test_outlier_error_structure_matches_reference(tables):
    """outlier_p: Bernoulli-mixture t vector, T^1/2 vmm T^1/2 + (I - T), two-pass initialisation."""
    g = load_golden('outlier_eis.npz')
    prep = orc.EisPrep(g['freq'], tables=tables)
    for b in range(g['z'].shape[0]):
        res = prep.fit(g['z'][b], hypers=dict(outlier_p=float(g['outlier_p'])))
        assert res['n_outer'] == int(g['n_outer'][b])
        assert int(res['ipm_iters'].sum()) == int(g['qp_log_total'][b])
        assert rel_err(res['x'], g['cvx_x'][b]) < 1e-9
        assert rel_err(res['est_weights'], g['est_weights'][b]) < 1e-9
        assert rel_err(res['outlier_t'], g['outlier_t'][b]) < 1e-9
        flagged = (1 - res['init_outlier_t']) > 0.75                  # drt1d.py:820
        nf = g['freq'].size
        assert np.array_equal(flagged[:nf] | flagged[nf:], g['ro_index'][b])


def test_flexible_chrono_error_structure_matches_reference(tables):
    g = load_golden('chrono_flex.npz')
    assert rel_err(orc.chrono_vmm(g['t2'], g['st2'], 4.0)[::8], g['vmm_two_step_rows']) < 1e-13
    vmm = orc.chrono_vmm(g['times'], g['step_times'], 4.0)
    assert rel_err(vmm[::8], g['vmm_chrono_rows']) < 1e-13
    cs = load_golden('chrono_small.npz')                              # same trace: rm / rv are shared
    rm = cs['rm']
    n_rows, n = rm.shape
    pen = _layout_pen(cs['basis_tau'], float(tables['eps']), [1e-6, 1e-6, 1e-6])
    h = np.zeros(n)
    h[0] = 1000
    prob = dict(rm=rm, rv=cs['rv'], vmm=dict(n_chrono=n_rows, chrono=vmm, eis=None), pen=pen, h=h,
                l1=np.zeros(n), n_special=3, n_chrono=n_rows)
    res = orc.qphb_fit(prob)
    assert res['n_outer'] == int(g['chrono_n_outer'])
    assert int(res['ipm_iters'].sum()) == int(g['chrono_ipm'])
    assert rel_err(res['x'], g['chrono_cvx_x']) < 1e-8


def test_pfrt_matches_reference(tables):
    """DRT.pfrt_fit_eis / pfrt_fit_hybrid: per-factor x, marginal llh and P of the continuation path."""
    g = load_golden('pfrt.npz')
    prep = orc.EisPrep(g['freq'], tables=tables)
    for b in range(2):
        prob, _ = prep.problem(g['z'][b])
        r = orc.pfrt_fit(prob)
        assert sum(st['n_iter'] for st in r['steps']) == int(g['n_hist'][b])
        assert r['init']['n_outer'] == int(g['init_n_outer'][b])
        assert rel_err(np.array([st['x'] for st in r['steps']]), g['step_x'][b]) < 1e-8
        llh = np.array([st['llh'] for st in r['steps']])
        assert np.max(np.abs(llh - g['step_llh'][b]) / np.abs(g['step_llh'][b])) < 1e-8
        assert rel_err(np.array([np.diag(st['p_matrix']) for st in r['steps']]), g['step_p_diag'][b]) < 1e-8
        assert rel_err(r['steps'][-1]['p_matrix'], g['step_p_last'][b]) < 1e-8
    hs = load_golden('hybrid_small.npz')                  # same trace and spectrum as the hybrid PFRT fixture
    rm0 = hs['rm'].copy()
    rm0[:, 1] = 0
    n, nc = rm0.shape[1], hs['times'].size
    h = np.zeros(n)
    h[:2] = 1000
    prob = dict(rm=rm0, rv=hs['rv'], vmm=dict(n_chrono=nc, chrono=None, eis=orc.eis_vmm(hs['freq'])),
                pen=_layout_pen(hs['basis_tau'], float(tables['eps']), [1e-6, 1.0, 1e-6, 1e-6]), h=h, l1=np.zeros(n),
                n_special=4, vz_index=1, vb_range=(0, 1), vz_strength=hs['vz_strength_vec'], n_chrono=nc)
    r = orc.pfrt_fit(prob, factors=g['h_factors'])
    assert sum(st['n_iter'] for st in r['steps']) == int(g['h_n_hist'])
    assert rel_err(np.array([st['x'] for st in r['steps']]), g['h_step_x']) < 1e-7
    llh = np.array([st['llh'] for st in r['steps']])
    assert np.max(np.abs(llh - g['h_step_llh']) / np.abs(g['h_step_llh'])) < 1e-7


def test_solve_rp_and_update_scale_match_reference(tables):
    """solve_rp (drt1d.py:573-607) and update_scale (:914-936) on EIS spectra whose first Rp estimate is off."""
    g = load_golden('rescale.npz')
    prep = orc.EisPrep(g['freq'], tables=tables)
    area = np.sqrt(np.pi) / prep.eps
    for tag, kw in (('us', dict(update_scale=True)), ('rp', dict(solve_rp=True)),
                    ('both', dict(solve_rp=True, update_scale=True))):
        for b in range(2):
            r = prep.fit(g['z'][b], basis_area=area, **kw)
            sf = r['scale_factors']
            assert r['n_outer'] == int(g[f'{tag}_n_outer'][b]) and int(r['ipm_iters'].sum()) == int(g[f'{tag}_ipm'][b])
            assert rel_err(r['x'], g[f'{tag}_cvx_x'][b]) < 1e-9
            assert abs(r['coefficient_scale'] / (sf[0] * sf[1]) / g[f'{tag}_coefficient_scale'][b] - 1) < 1e-9
            assert rel_err(r['est_weights'], g[f'{tag}_est_weights'][b]) < 1e-9
            assert rel_err(r['init_weights'], g[f'{tag}_init_weights'][b]) < 1e-9
            assert rel_err(r['xmx_norms'], g[f'{tag}_xmx_norms'][b]) < 1e-9


def test_dop_voltage_matrix_matches_reference():
    d = load_golden('dop_chrono.npz')
    m = orc.dop_v_matrix(d['times'], d['basis_nu'], float(d['nu_epsilon']), d['step_times'], d['step_sizes'])
    assert rel_err(m, d['rm_dop']) < 1e-13


DS_CASES = dict(tut=dict(decimation_interval=8, decimation_factor=2, prestep_samples=25),
                size=dict(target_size=300, decimation_factor=2, prestep_samples=10, decimation_max_period=0.05),
                raw=dict(decimation_interval=20, decimation_factor=3, prestep_samples=5, antialiased=False))


def test_downsample_matches_reference():
    """preprocessing.downsample_data (decimation index + antialiasing filter): the oracle against the reference,
    and the host-side tap layout of the product (evaluated here with numpy) against the oracle."""
    from oracle import chrono_oracle as co
    from hybdrt_b200 import synth, preprocessing as pp
    g = load_golden('downsample.npz')
    times, i_sig, v = synth.make_raw_chrono_batch(3, seed=5)
    for tag, kw in DS_CASES.items():
        st, si, sv, idx = co.downsample_data(times, i_sig, v[1], **kw)
        assert np.array_equal(idx, g[f'{tag}_index']) and np.array_equal(st, g[f'{tag}_times'])
        assert rel_err(si, g[f'{tag}_i']) < 1e-13 and rel_err(sv, g[f'{tag}_v'][1]) < 1e-13
    # scipy's own filter, where the explicit restatement stands in for it
    from scipy import ndimage
    y = np.random.default_rng(0).normal(size=57)
    for sigma in (0.3, 1.7, 9.0, 40.0):
        assert rel_err(co.gaussian_filter1d_reflect(y, sigma), ndimage.gaussian_filter1d(y, sigma, mode='reflect')) < 1e-13
    # product host logic: index selection and the tap layout (no GPU: the taps are applied with numpy here)
    step_times = times[co.identify_steps(i_sig, True)]
    t_sample = np.min(np.diff(times))
    idx = pp.get_decimation_index(times, step_times, t_sample, 25, 8, 2, None)
    assert np.array_equal(idx, g['tut_index'])
    iv = pp.select_decimation_interval(times, step_times, t_sample, 10, 2, 0.05, 300)
    assert np.array_equal(pp.get_decimation_index(times, step_times, t_sample, 10, iv, 2, 0.05), g['size_index'])
    plan = pp.filter_plan(times, pp.identify_steps(i_sig, allow_consecutive=False), idx)
    out = np.empty(len(idx))
    for j in range(len(idx)):
        lw, lo, ln = plan['lw'][j], plan['seg_lo'][j], plan['seg_len'][j]
        pos = (plan['idx'][j] - lo + np.arange(-lw, lw + 1)) % (2 * ln)
        pos = np.where(pos >= ln, 2 * ln - 1 - pos, pos)
        out[j] = plan['taps'][plan['woff'][j]:plan['woff'][j] + 2 * lw + 1] @ v[2][lo + pos]
    assert rel_err(out, g['tut_v'][2]) < 1e-13
    assert abs(plan['taps'].sum() - len(idx)) < 1e-9                     # every kept sample's taps sum to one


def test_hybrid_options_match_reference(tables):
    """hybrid_weight_factor_method='weight' (drt1d.py:749-759) and init_weights_separately (:647-669)."""
    g = load_golden('hybrid_options.npz')
    hs = load_golden('hybrid_small.npz')
    rm0 = hs['rm'].copy()
    rm0[:, 1] = 0
    n, nc = rm0.shape[1], hs['times'].size
    h = np.zeros(n)
    h[:2] = 1000
    base = dict(rm=rm0, rv=hs['rv'], vmm=dict(n_chrono=nc, chrono=None, eis=orc.eis_vmm(hs['freq'])),
                pen=_layout_pen(hs['basis_tau'], float(tables['eps']), [1e-6, 1.0, 1e-6, 1e-6]), h=h, l1=np.zeros(n),
                n_special=4, vz_index=1, vb_range=(0, 1), vz_strength=hs['vz_strength_vec'], n_chrono=nc)
    res = orc.qphb_fit(dict(base, hybrid_weight_factor_method='weight'))
    assert res['n_outer'] == int(g['weight_n_outer']) and int(res['ipm_iters'].sum()) == int(g['weight_ipm'])
    assert rel_err(np.array([res['chrono_weight_factor'], res['eis_weight_factor']]), g['weight_factors']) < 1e-10
    assert rel_err(res['x'], g['weight_cvx_x']) < 1e-8
    res = orc.qphb_fit(dict(base, init_weights_separately=True))
    assert res['n_outer'] == int(g['sep_n_outer']) and int(res['ipm_iters'].sum()) == int(g['sep_ipm'])
    assert rel_err(res['est_weights'], g['sep_est_weights']) < 1e-8
    assert rel_err(res['x'], g['sep_cvx_x']) < 1e-8
    cf, ef = g['rp_factors']                        # 'rp': the factors come from the data, the loop is the plain one
    res = orc.qphb_fit(dict(base, chrono_weight_factor=float(cf), eis_weight_factor=float(ef)))
    assert res['n_outer'] == int(g['rp_n_outer']) and rel_err(res['x'], g['rp_cvx_x']) < 1e-8


def test_coneqp_small_kat():
    """Known answer: min 1/2 x'x - c'x, x >= 0 has x = max(c, 0)."""
    c = np.array([1.0, -2.0, 0.5, -0.1])
    res = coneqp_orthant(np.eye(4), -c, np.zeros(4))
    assert res['status'] == 'optimal'
    assert np.allclose(res['x'], np.maximum(c, 0), atol=1e-6)


def test_resolve_oracle_against_the_reference_fixture():
    """mapping/resolve.py through DRTMD.resolve_observations (one window of seven hybrid observations)."""
    from oracle import resolve_oracle as ro
    g = load_golden('drtmd_resolve.npz')
    names = [str(s) for s in g['special_names']]
    assert names == ['v_baseline', 'vz_offset', 'R_inf', 'inductance']
    p_list, q_list = [], []
    match = (int(g['obs_tau_indices'][:7, 0].min()), int(g['obs_tau_indices'][:7, 1].max()))
    for i in range(7):
        xr = np.concatenate([ro.scaled_v_baseline(g['fit_v_baseline'][i], g['fit_response_signal_scale'][i],
                                                  g['fit_scaled_response_offset'][i], g['fit_v_baseline_scale'][i]),
                             [g['fit_vz_offset'][i]]])
        p, q = ro.offset_pq(g['fit_p'][i], g['fit_q'][i], xr)
        p, q = ro.resize_pq(p, q, 2, tuple(g['obs_tau_indices'][i]), match)
        p_list.append(p)
        q_list.append(q)
    scale = g['fit_coefficient_scale'][:7]
    x, iters, _ = ro.resolve_window(p_list, q_list, scale, g['fit_R_inf'][:7] / scale, 0, True, [0, 1])
    assert iters == int(g['win_ipm'][0])
    x_drt = x[:, 2:] * scale[:, None]
    assert rel_err(x_drt, g['win_x_resolved'][:, match[0]:match[1]]) < 1e-9
    assert rel_err(x[:, 0] * scale, g['win_special_R_inf']) < 1e-9
    assert rel_err(x[:, 1] * scale * g['fit_inductance_scale'][:7], g['win_special_inductance']) < 1e-9
