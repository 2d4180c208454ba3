"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on seeded inputs and
against the fixtures produced by the unmodified reference.  Run on the B200 box: pytest -m gpu.

Tolerances (BASELINE.json north_star): response matrices 1e-10 relative; fitted coefficients, predicted
impedance and the objective 1e-6 relative.  'Relative' is norm-wise, max|a-b|/max|b| (coefficients
pinned at the non-negativity bound sit at ~1e-7 of the peak)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu

MAT_TOL = 1e-10
FIT_TOL = 1e-6


@pytest.fixture(scope='module')
def eng():
    import __graft_entry__ as g
    g.build()
    from hybdrt_b200 import engine
    return engine.get_engine(0)


@pytest.fixture(scope='module')
def orc():
    from oracle import drt_oracle
    return drt_oracle


def _np(t):
    return t.detach().cpu().numpy()


# ---------------------------------------------------------------------------------------------------
# L1: matrix builders
# ---------------------------------------------------------------------------------------------------
def test_lookup_tables(eng, lookup_golden):
    tab = eng.build_lookup(float(lookup_golden['eps']))
    for key in ('re_x', 're_v', 'im_x', 'im_v', 'resp_x', 'resp_v'):
        assert rel_err(_np(tab[key]), lookup_golden[key]) < MAT_TOL, key


def test_matrix_fixtures(eng, lookup_golden):
    from hybdrt_b200 import engine as E, synth
    m = load_golden('matrices.npz')
    eps = float(m['eps'])
    for mode, name in ((E.MODE_INTERP, 'interp'), (E.MODE_TRAPZ, 'trapz')):
        a_re, a_im = eng.build_impedance(m['f_irreg'][None], m['tau_irreg'][None], eps, mode)
        assert rel_err(_np(a_re[0]), m[f'irreg_{name}_real']) < MAT_TOL
        assert rel_err(_np(a_im[0]), m[f'irreg_{name}_imag']) < MAT_TOL
        rm = eng.build_response(m['t_resp'][None], m['tau_resp'][None], m['step_times'][None],
                                m['step_sizes'][None], eps, mode)
        assert rel_err(_np(rm[0]), m[f'resp_{name}']) < MAT_TOL
    a_re, a_im = eng.build_impedance(synth.C2_FREQ[None], m['tau_c2'][None], eps, E.MODE_TRAPZ)
    assert rel_err(_np(a_re[0]), m['c2_trapz_real']) < MAT_TOL
    assert rel_err(_np(a_im[0]), m['c2_trapz_imag']) < MAT_TOL
    pen_i = _np(eng.build_penalty(np.log(m['tau_irreg'])[None], eps, False)[0])
    pen_u = _np(eng.build_penalty(np.log(m['tau_c2'])[None], eps, True)[0])
    for k in range(3):
        assert rel_err(pen_i[k], m[f'pen_irreg_{k}']) < MAT_TOL
        assert rel_err(pen_u[k], m[f'pen_c2_{k}']) < MAT_TOL
    assert rel_err(_np(eng.build_eis_vmm(m['f_irreg'][None])[0]), m['vmm_irreg']) < MAT_TOL
    assert rel_err(_np(eng.build_eis_vmm(m['f_irreg'][None], uniform=True)[0]), m['vmm_uniform']) < MAT_TOL
    d = load_golden('dop.npz')
    zd = _np(eng.build_dop_z(d['freq'][None], d['basis_nu'], float(d['nu_epsilon']))[0])
    assert rel_err(zd, d['zm_dop']) < MAT_TOL


def test_matrices_many_grids_vs_oracle(eng, orc, lookup_golden):
    """Ragged/edge grids: per-spectrum grids, values far outside the table span (edge clamping), one row."""
    from hybdrt_b200 import engine as E
    eps = float(lookup_golden['eps'])
    rng = np.random.default_rng(11)
    g, nf, nb = 5, 9, 13
    freq = 10 ** rng.uniform(-4, 7.5, (g, nf))
    tau = np.sort(10 ** rng.uniform(-9, 4, (g, nb)), axis=1)
    a_re, a_im = eng.build_impedance(freq, tau, eps, E.MODE_INTERP)
    for i in range(g):
        assert rel_err(_np(a_re[i]), orc.impedance_matrix(freq[i], tau[i], eps, 'real', 'interp', lookup_golden)) < MAT_TOL
        assert rel_err(_np(a_im[i]), orc.impedance_matrix(freq[i], tau[i], eps, 'imag', 'interp', lookup_golden)) < MAT_TOL
    a_re, a_im = eng.build_impedance(freq[:1, :1], tau[:1], eps, E.MODE_TRAPZ)
    assert rel_err(_np(a_re[0]), orc.impedance_matrix(freq[0, :1], tau[0], eps, 'real', 'trapz')) < MAT_TOL
    times = np.sort(rng.uniform(-0.01, 3.0, (g, 40)), axis=1)
    st = np.stack([np.array([0.0, 1.0 + 0.1 * i]) for i in range(g)])
    sa = rng.uniform(-1, 1, (g, 2))
    rm = eng.build_response(times, tau, st, sa, eps, E.MODE_INTERP)
    for i in range(g):
        assert rel_err(_np(rm[i]), orc.response_matrix(tau[i], times[i], st[i], sa[i], eps, 'interp', lookup_golden)) < MAT_TOL
    pen = eng.build_penalty(np.log(tau), eps, False)
    for i in range(g):
        for k in range(3):
            assert rel_err(_np(pen[i, k]), orc.penalty_matrix(np.log(tau[i]), k, eps)) < MAT_TOL
    vmm = eng.build_eis_vmm(freq)
    for i in range(g):
        assert rel_err(_np(vmm[i]), orc.eis_vmm(freq[i])) < MAT_TOL
        assert np.allclose(_np(vmm[i]).sum(axis=1), 1.0, atol=1e-14)      # rows are averaging weights


def test_interp_matrix_properties_at_scale(eng):
    """Size-independent properties on a large batch of grids: A_re in (0, area], monotone in omega, A_im < 0."""
    from hybdrt_b200 import engine as E
    rng = np.random.default_rng(5)
    g = 2048
    freq = np.sort(10 ** rng.uniform(-2, 6, (g, 70)), axis=1)[:, ::-1].copy()
    tau = np.logspace(-7.2, 2.2, 101)[None].repeat(g, axis=0)
    eps = 1 / np.log(10 ** 0.1)
    a_re, a_im = eng.build_impedance(freq, tau, eps, E.MODE_INTERP)
    area = np.sqrt(np.pi) / eps
    assert float(a_re.max()) <= area * (1 + 1e-12) and float(a_re.min()) > 0
    assert float(a_im.max()) < 0
    assert bool((a_re[:, 1:, :] >= a_re[:, :-1, :] - 1e-15).all())        # rows ordered by decreasing frequency


# ---------------------------------------------------------------------------------------------------
# L2: batched QPHB solver vs oracle and reference fixtures
# ---------------------------------------------------------------------------------------------------
def _eis_launch(eng, prep, rvs, **kw):
    d = eng.dev
    out = eng.qphb_fit_batch(d(prep.rm), d(rvs), d(np.array(prep.pen)), d(prep.h), d(prep.l1), 2,
                             vmm_eis=d(prep.vmm), **kw)
    torch.cuda.synchronize()
    return {k: _np(v) for k, v in out.items()}


def test_reference_golden_vector_through_the_kernel(eng, orc, lookup_golden):
    c1 = load_golden('c1_golden.npz')
    prep = orc.EisPrep(c1['freq'], tables=lookup_golden)
    prob, scale = prep.problem(c1['z'])
    out = _eis_launch(eng, prep, prob['rv'][None], want_pq=True)
    assert int(out['n_outer'][0]) == int(c1['n_outer'])
    assert int(out['n_ipm'][0]) == int(c1['qp_log'].sum())
    assert rel_err(out['x'][0], c1['cvx_x']) < FIT_TOL
    assert rel_err(out['weights'][0], c1['true_weights']) < FIT_TOL
    assert rel_err(out['p_matrix'][0], c1['p_matrix']) < FIT_TOL
    assert rel_err(out['q_vector'][0], c1['q_vector']) < FIT_TOL
    assert abs(out['fun'][0] - c1['hist_fun'][-1]) <= FIT_TOL * abs(c1['hist_fun'][-1])
    # the reference's own hard-coded expectations (reference tests/test_drt_fit.py:164-170)
    assert np.allclose(c1['expected_x'], out['x'][0, 2:] * scale)
    assert np.allclose(c1['expected_R_inf'], out['x'][0, 0] * scale)
    assert np.allclose(c1['expected_inductance'], out['x'][0, 1] * scale * 1e-5)
    assert np.allclose(c1['expected_q_vector'], out['q_vector'][0])


def test_qphb_batch_vs_oracle_seeded(eng, orc, lookup_golden):
    from hybdrt_b200 import synth
    freq, z = synth.make_eis_batch(40, seed=123)
    prep = orc.EisPrep(freq, tables=lookup_golden)
    rvs = np.array([prep.problem(zz)[0]['rv'] for zz in z])
    out = _eis_launch(eng, prep, rvs)
    n_same = 0
    for b in range(len(z)):
        ref = prep.fit(z[b])
        if int(out['n_outer'][b]) == ref['n_outer'] and int(out['n_ipm'][b]) == int(ref['ipm_iters'].sum()):
            n_same += 1
        assert rel_err(out['x'][b], ref['x']) < FIT_TOL, b
        assert rel_err(out['weights'][b], ref['weights']) < FIT_TOL, b
        assert abs(out['fun'][b] - ref['fun']) <= FIT_TOL * abs(ref['fun']), b
        assert rel_err(out['rho'][b], ref['rho']) < FIT_TOL, b
        # s = u^2 with u the root of a quadratic whose discriminant cancels (qphb.py:346-360): 1e-6 on x leaves ~1e-5 on s
        assert rel_err(out['s_vectors'][b], ref['s_vectors']) < 1e-5, b
        assert rel_err(out['xmx_norms'][b], ref['xmx_norms']) < FIT_TOL, b
    assert n_same == len(z)       # iteration counts are discrete outcomes; all should agree


def test_batch_invariances(eng, orc, lookup_golden):
    """Size-independent properties: a spectrum's result does not depend on its batch neighbours or slot;
    per-spectrum matrices (stride != 0) give the same answer as shared ones."""
    from hybdrt_b200 import synth
    freq, z = synth.make_eis_batch(600, seed=9)
    prep = orc.EisPrep(freq, tables=lookup_golden)
    scale = (z.real.max(axis=1) - z.real.min(axis=1)) / 14.0
    zs = z / scale[:, None]
    rvs = np.concatenate([zs.real, zs.imag], axis=1)
    full = _eis_launch(eng, prep, rvs)
    perm = np.random.default_rng(0).permutation(len(z))[:64]
    sub = _eis_launch(eng, prep, rvs[perm])
    assert np.array_equal(sub['x'], full['x'][perm])                      # bitwise: same code path, same data
    assert np.array_equal(sub['n_ipm'], full['n_ipm'][perm])
    d = eng.dev
    k = 5
    out = eng.qphb_fit_batch(d(np.repeat(prep.rm[None], k, 0)), d(rvs[:k]), d(np.repeat(np.array(prep.pen)[None], k, 0)),
                             d(prep.h), d(prep.l1), 2, vmm_eis=d(np.repeat(prep.vmm[None], k, 0)))
    torch.cuda.synchronize()
    assert np.array_equal(_np(out['x']), full['x'][:k])
    st = full['status']
    assert np.all((st & 3) != 0) and np.all((st & 3) != 3)               # exactly one of converged / max_iter
    assert np.all(full['x'][:, 0] >= 0) and np.all(np.isfinite(full['x']))
    assert np.all(full['n_outer'] >= 1) and np.all(full['n_outer'] <= 50)


def test_problem_size_edges(eng, orc, lookup_golden):
    """Shapes around the kernel's internal boundaries: a handful of columns (one or two 8 x 8 tiles), the last size of
    the small configuration (n = 104) and the first of the large one (n = 105), the largest supported n (160), few
    and many data rows -- each against the oracle."""
    from hybdrt_b200 import synth
    cases = [np.logspace(3.0, 2.2, 3),           # n = 32 (four tiles), N = 6: fewer rows than columns
             np.logspace(4.0, 3.0, 41),          # n = 34 (one column into the fifth tile), N = 82
             np.logspace(6.0, -2.05, 70),        # n = 104: last CfgS size
             np.logspace(6.0, -2.15, 70),        # n = 105: first CfgL size
             np.logspace(7.0, -6.65, 400)]       # n = 160: the largest, N = 800
    for k, freq in enumerate(cases):
        _, z = synth.make_eis_batch(2, freq=freq, seed=20 + k)
        prep = orc.EisPrep(freq, tables=lookup_golden)
        n = prep.rm.shape[1]
        assert n == (32, 34, 104, 105, 160)[k]
        rvs = np.array([prep.problem(zz)[0]['rv'] for zz in z])
        out = _eis_launch(eng, prep, rvs, want_pq=True)
        for b in range(2):
            ref = prep.fit(z[b])
            assert int(out['n_outer'][b]) == ref['n_outer'], (k, b)
            assert int(out['n_ipm'][b]) == int(ref['ipm_iters'].sum()), (k, b)
            assert rel_err(out['x'][b], ref['x']) < FIT_TOL, (k, b)
            assert rel_err(out['weights'][b], ref['weights']) < FIT_TOL, (k, b)
            assert rel_err(out['p_matrix'][b], ref['p_matrix']) < FIT_TOL, (k, b)
    with pytest.raises(Exception):               # n = 161 does not fit
        freq = np.logspace(7.0, -6.75, 50)
        prep = orc.EisPrep(freq, tables=lookup_golden)
        assert prep.rm.shape[1] > 160
        _eis_launch(eng, prep, np.zeros((1, 2 * len(freq))))


def test_empty_batch_and_bad_arguments(eng, orc, lookup_golden):
    from hybdrt_b200 import engine as E
    c1 = load_golden('c1_golden.npz')
    prep = orc.EisPrep(c1['freq'], tables=lookup_golden)
    d = eng.dev
    out = eng.qphb_fit_batch(d(prep.rm), eng.empty(0, prep.rm.shape[0]), d(np.array(prep.pen)), d(prep.h), d(prep.l1), 2,
                             vmm_eis=d(prep.vmm))
    assert out['x'].shape[0] == 0
    with pytest.raises(E.EngineError):      # EIS rows without a vmm block
        eng.qphb_fit_batch(d(prep.rm), d(np.zeros((1, prep.rm.shape[0]))), d(np.array(prep.pen)), d(prep.h), d(prep.l1), 2)
    with pytest.raises(E.EngineError):      # nu_epsilon too small for the complex-erf kernel
        eng.build_dop_z(c1['freq'][None], np.linspace(-1, 1, 5), 1.0)


# ---------------------------------------------------------------------------------------------------
# the drop-in model class against outputs of the unmodified reference
# ---------------------------------------------------------------------------------------------------
def test_model_fit_eis_golden():
    """The reference's tests/test_drt_fit.py, run against the drop-in class."""
    from hybdrt_b200.models import DRT
    c1 = load_golden('c1_golden.npz')
    drt = DRT(fit_inductance=True, fit_ohmic=True)
    hypers = dict(rp_scale=14, derivative_weights=np.array([1.5, 1.0, 0.5]), sigma_ds=np.array([1, 1000, 1000]),
                  l1_lambda_0=0, l2_lambda_0=142, s_alpha=np.array([5, 10, 25]), rho_alpha=np.array([0.15, 0.2, 0.25]),
                  iw_alpha=None, iw_beta=None, s_0=np.ones(3), rho_0=np.ones(3), outlier_p=None)
    drt.fit_eis(c1['freq'], c1['z'], **hypers)
    fp = drt.fit_parameters
    expected = dict(x=c1['expected_x'], R_inf=c1['expected_R_inf'], inductance=c1['expected_inductance'],
                    z_sigma_tot=c1['expected_z_sigma_tot'], q_vector=c1['expected_q_vector'])
    for key, val in expected.items():
        assert np.allclose(val, fp[key]), key
    assert fp['C_inv'] == 0 and fp['v_sigma_tot'] is None and fp['v_sigma_res'] is None and fp['vz_offset_eps'] == 1
    # tighter, against the reference run on the same machine as the fixtures
    assert rel_err(drt.cvx_result['x'], c1['cvx_x']) < FIT_TOL
    assert rel_err(drt.predict_z(c1['freq']), c1['z_pred']) < FIT_TOL
    assert rel_err(drt.qphb_params['rm'], c1['rm']) < MAT_TOL
    assert rel_err(drt.qphb_params['vmm'], c1['vmm']) < MAT_TOL
    assert rel_err(drt.basis_tau, c1['basis_tau']) < 1e-14
    # qphb_history (qphb.py:950-966): one entry per outer iteration, the last one populated
    hist = drt.qphb_history
    assert len(hist) == int(c1['n_outer']) and rel_err(hist[-1]['x'], c1['hist_x'][-1]) < FIT_TOL
    assert rel_err(hist[-1]['rho_vector'], c1['hist_rho'][-1]) < FIT_TOL and rel_err(np.array(hist[-1]['s_vectors']), c1['hist_s'][-1]) < 1e-5
    assert abs(hist[-1]['fun'] - float(c1['hist_fun'][-1])) < 1e-6 * abs(float(c1['hist_fun'][-1]))
    with pytest.raises(ValueError):
        drt.fit_eis(c1['freq'], c1['z'], not_a_hyper=1)


def test_model_eis_batch_fixtures():
    from hybdrt_b200.models import DRT
    c2 = load_golden('c2_eis.npz')
    drt = DRT()
    res = drt.fit_eis_batch(c2['freq'], c2['z'])
    h = res.host()
    fp = res.fit_parameters()
    assert np.array_equal(h['n_outer'], c2['n_outer'])
    assert np.array_equal(h['n_ipm'], c2['qp_log_total'])
    for b in range(len(c2['z'])):
        assert rel_err(h['x'][b], c2['cvx_x'][b]) < FIT_TOL
        assert rel_err(fp['x'][b], c2['x'][b]) < FIT_TOL
        assert abs(fp['R_inf'][b] - c2['R_inf'][b]) < FIT_TOL * abs(c2['R_inf'][b])
        assert rel_err(fp['z_sigma_tot'][b], c2['z_sigma_tot'][b]) < FIT_TOL
        assert rel_err(res.predict_z()[b], c2['z_pred'][b]) < FIT_TOL
    free = load_golden('c2_free.npz')
    drt.fit_eis(free['freq'], free['z'], nonneg=False)
    assert rel_err(drt.cvx_result['x'], free['cvx_x']) < FIT_TOL
    assert rel_err(drt.predict_z(free['freq']), free['z_pred']) < FIT_TOL


def test_model_trapz_mode():
    from hybdrt_b200.models import DRT
    g = load_golden('c2_trapz_fit.npz')
    drt = DRT(interpolate_integrals=False)
    drt.fit_eis(g['freq'], g['z'])
    assert drt.qphb_params['n_outer'] == int(g['n_outer'])
    assert rel_err(drt.cvx_result['x'], g['cvx_x']) < FIT_TOL
    assert rel_err(drt.predict_z(g['freq']), g['z_pred']) < FIT_TOL


def test_model_dop():
    from hybdrt_b200.models import DRT
    d = load_golden('dop.npz')
    drt = DRT(fit_dop=True)
    res = drt.fit_eis_batch(d['freq'], d['z'])
    h, fp = res.host(), res.fit_parameters()
    assert rel_err(_np(res.plan['rm']), d['rm']) < MAT_TOL
    assert rel_err(drt.dop_scale_vector, d['dop_scale_vector']) < 1e-13
    assert np.array_equal(h['n_outer'], d['n_outer'])
    for b in range(len(d['z'])):
        assert rel_err(h['x'][b], d['cvx_x'][b]) < FIT_TOL
        assert rel_err(fp['x_dop'][b], d['x_dop'][b]) < FIT_TOL
        assert rel_err(h['dop_rho'][b], d['dop_rho_vector'][b]) < FIT_TOL
        assert rel_err(res.predict_z()[b], d['z_pred'][b]) < FIT_TOL


def test_model_hybrid_and_chrono():
    from hybdrt_b200.models import DRT
    from hybdrt_b200 import synth
    hs = load_golden('hybrid_small.npz')
    drt = DRT()
    drt.fit_hybrid(hs['times'], hs['i_signal'], hs['v_signal'], hs['freq'], hs['z'])
    qp, fp = drt.qphb_params, drt.fit_parameters
    assert rel_err(qp['rv'], hs['rv']) < 1e-12
    assert rel_err(qp['vz_strength_vec'], hs['vz_strength_vec']) < 1e-12
    assert qp['n_outer'] == int(hs['n_outer']) and qp['n_ipm'] == int(hs['qp_log'].sum())
    assert rel_err(qp['rm'], hs['rm']) < 1e-9          # includes the final vz_offset column
    assert rel_err(drt.cvx_result['x'], hs['cvx_x']) < FIT_TOL
    assert rel_err(fp['x'], hs['x']) < FIT_TOL
    assert abs(fp['vz_offset'] - hs['vz_offset']) < FIT_TOL * max(1.0, abs(hs['vz_offset']))
    # v_baseline is a ~1e-6-sized parameter: its own relative error is the absolute 1e-6-class error of the fit over a tiny norm
    assert rel_err(fp['v_baseline'], hs['v_baseline']) < 1e-5
    assert rel_err(fp['z_sigma_tot'], hs['z_sigma_tot']) < FIT_TOL
    assert rel_err(fp['v_sigma_tot'], hs['v_sigma_tot']) < FIT_TOL
    assert rel_err(drt.predict_z(hs['freq']), hs['z_pred']) < FIT_TOL
    assert rel_err(drt.predict_response(), hs['v_pred']) < FIT_TOL

    cs = load_golden('chrono_small.npz')
    drt.fit_chrono(cs['times'], cs['i_signal'], cs['v_signal'])
    assert drt.qphb_params['n_outer'] == int(cs['n_outer'])
    assert rel_err(drt.qphb_params['rm'], cs['rm']) < MAT_TOL
    assert rel_err(drt.cvx_result['x'], cs['cvx_x']) < FIT_TOL
    assert rel_err(drt.predict_response(), cs['v_pred']) < FIT_TOL

    # full-size C3 spectra (2000 samples + 30 frequencies), outputs of the reference only
    hf = load_golden('hybrid_full.npz')
    t, i_sig, v, f, z = synth.make_hybrid_batch(2, seed=1)
    res = drt.fit_hybrid_batch(t, i_sig, v, f, z)
    h = res.host()
    assert np.array_equal(h['n_outer'], hf['n_outer'])
    for b in range(2):
        assert rel_err(h['x'][b], hf['cvx_x'][b]) < FIT_TOL
        assert rel_err(res.predict_z()[b], hf['z_pred'][b]) < FIT_TOL


def test_outlier_error_structure_and_removal(eng, orc, lookup_golden):
    """outlier_p (qphb.py:1497-1538, two-pass initialisation :1629-1655) and remove_outliers (drt1d.py:217-303)
    against the unmodified reference; the kernel alone against the oracle on the same problems."""
    from hybdrt_b200.models import DRT
    from hybdrt_b200 import engine as E
    g = load_golden('outlier_eis.npz')
    prep = orc.EisPrep(g['freq'], tables=lookup_golden)
    hyp = E.default_hypers()
    hyp.has_outlier_p, hyp.outlier_p = 1, float(g['outlier_p'])
    rvs = np.array([prep.problem(zz)[0]['rv'] for zz in g['z']])
    out = _eis_launch(eng, prep, rvs, hypers=hyp)
    for b in range(len(g['z'])):
        ref = prep.fit(g['z'][b], hypers=dict(outlier_p=float(g['outlier_p'])))
        assert int(out['n_outer'][b]) == ref['n_outer'] == int(g['n_outer'][b])
        assert int(out['n_ipm'][b]) == int(g['qp_log_total'][b])
        assert rel_err(out['x'][b], ref['x']) < FIT_TOL
        assert rel_err(out['outlier_t'][b], ref['outlier_t']) < FIT_TOL
        assert rel_err(out['est_weights'][b], g['est_weights'][b]) < FIT_TOL
        assert rel_err(out['x_overfit'][b], g['x_overfit_eis'][b]) < FIT_TOL
    drt = DRT()
    res = drt.fit_eis_batch(g['freq'], g['z'], outlier_p=float(g['outlier_p']))
    h = res.host()
    for b in range(len(g['z'])):
        assert rel_err(h['x'][b], g['cvx_x'][b]) < FIT_TOL
        assert rel_err(h['outlier_t'][b], g['outlier_t'][b]) < FIT_TOL
        assert rel_err(res.predict_z()[b], g['z_pred'][b]) < FIT_TOL
    drt.warn = False
    for b in range(len(g['z'])):
        drt.fit_eis(g['freq'], g['z'][b], outlier_p=float(g['outlier_p']), remove_outliers=True)
        assert np.array_equal(np.asarray(drt.eis_outlier_index), g['ro_index'][b])
        assert drt.qphb_params['n_outer'] == int(g['ro_n_outer'][b])
        assert rel_err(drt.fit_parameters['x'], g['ro_x'][b]) < FIT_TOL
        assert rel_err(drt.predict_z(g['freq']), g['ro_z_pred'][b]) < FIT_TOL
    drt.fit_eis(g['freq'], g['ext_z'], remove_extremes=True)          # drt1d.py:188-215
    assert np.array_equal(np.asarray(drt.f_fit), g['ext_f_fit']) and drt.qphb_params['n_outer'] == int(g['ext_n_outer'])
    assert rel_err(drt.fit_parameters['x'], g['ext_x']) < FIT_TOL
    assert rel_err(drt.predict_z(g['freq']), g['ext_z_pred']) < FIT_TOL
    with pytest.raises(ValueError):
        drt.fit_eis(g['freq'], g['z'][0], remove_outliers=True)
    with pytest.raises(E.EngineError):
        hyp.outlier_p = 1.5
        _eis_launch(eng, prep, rvs, hypers=hyp)


def test_flexible_chrono_error_structure(eng, orc):
    """chrono_error_structure=None (mat1d.py:455-490): the matrix builder, then fit_chrono / fit_hybrid with the
    dense chrono block, then the tutorial's flags (flexible chrono errors + outlier_p) against the reference."""
    from hybdrt_b200.models import DRT
    g = load_golden('chrono_flex.npz')
    v2 = _np(eng.build_chrono_vmm(g['t2'][None], g['st2'][None], 4.0)[0])
    assert rel_err(v2[::8], g['vmm_two_step_rows']) < MAT_TOL
    assert rel_err(v2, orc.chrono_vmm(g['t2'], g['st2'], 4.0)) < MAT_TOL
    assert np.allclose(v2.sum(axis=1), 1.0, atol=1e-13)
    vu = _np(eng.build_chrono_vmm(g['t2'][None], g['st2'][None], 4.0, uniform=True)[0])
    assert np.array_equal(vu, np.full_like(vu, 1.0 / len(g['t2'])))
    v1 = _np(eng.build_chrono_vmm(g['times'][None], g['step_times'][None], 4.0)[0])
    assert rel_err(v1[::8], g['vmm_chrono_rows']) < MAT_TOL
    drt = DRT()
    drt.fit_chrono(g['times'], g['i_signal'], g['v_signal'][1], error_structure=None)
    assert drt.qphb_params['n_outer'] == int(g['chrono_n_outer']) and drt.qphb_params['n_ipm'] == int(g['chrono_ipm'])
    assert rel_err(drt.cvx_result['x'], g['chrono_cvx_x']) < FIT_TOL
    assert rel_err(drt.qphb_params['est_weights'], g['chrono_est_weights']) < FIT_TOL
    assert rel_err(drt.predict_response(), g['chrono_v_pred']) < FIT_TOL
    drt.fit_hybrid(g['times'], g['i_signal'], g['v_signal'][0], g['freq'], g['z'][0], chrono_error_structure=None)
    assert drt.qphb_params['n_outer'] == int(g['hybrid_n_outer']) and drt.qphb_params['n_ipm'] == int(g['hybrid_ipm'])
    assert rel_err(drt.cvx_result['x'], g['hybrid_cvx_x']) < FIT_TOL
    assert rel_err(drt.predict_z(g['freq']), g['hybrid_z_pred']) < FIT_TOL
    drt.fit_hybrid(g['times'], g['i_signal'], g['tut_v_signal'], g['freq'], g['z'][0], chrono_error_structure=None,
                   outlier_p=0.01)
    assert drt.qphb_params['n_outer'] == int(g['tut_n_outer']) and drt.qphb_params['n_ipm'] == int(g['tut_ipm'])
    assert rel_err(drt.cvx_result['x'], g['tut_cvx_x']) < FIT_TOL
    assert rel_err(drt.qphb_params['outlier_t'], g['tut_outlier_t']) < FIT_TOL
    assert rel_err(drt.predict_z(g['freq']), g['tut_z_pred']) < FIT_TOL
    # the hybrid tutorial's exact call: discard_first_n too (drt1d.py:170-181)
    drt.fit_hybrid(g['times'], g['i_signal'], g['tut_v_signal'], g['freq'], g['z'][0], discard_first_n=1,
                   chrono_error_structure=None, chrono_vmm_epsilon=4, outlier_p=0.01)
    assert np.array_equal(drt.t_fit, g['tut2_t_fit']) and rel_err(drt.step_times, g['tut2_step_times']) < 1e-12
    assert drt.qphb_params['n_outer'] == int(g['tut2_n_outer']) and drt.qphb_params['n_ipm'] == int(g['tut2_ipm'])
    assert rel_err(drt.cvx_result['x'], g['tut2_cvx_x']) < FIT_TOL
    assert rel_err(drt.predict_z(g['freq']), g['tut2_z_pred']) < FIT_TOL
    assert rel_err(drt.predict_response(), g['tut2_v_pred']) < FIT_TOL
    with pytest.raises(ValueError):
        drt.fit_chrono(g['times'], g['i_signal'], g['v_signal'][1], error_structure='nonsense')


def test_pfrt_continuation(eng, orc, lookup_golden):
    """PFRT (drt1d.py:2558-2716): the initial fit and every warm-started continuation step inside one launch,
    against the unmodified reference (EIS and hybrid) and, kernel only, against the oracle."""
    from hybdrt_b200.models import DRT
    g = load_golden('pfrt.npz')
    prep = orc.EisPrep(g['freq'], tables=lookup_golden)
    rvs = np.array([prep.problem(zz)[0]['rv'] for zz in g['z']])
    hyp = E_default_hypers(max_iter=20)
    out = _eis_launch(eng, prep, rvs, hypers=hyp, pfrt=dict(factors=g['factors'], want_p=True))
    for b in range(2):
        assert int(out['pfrt_iters'][b].sum()) == int(g['n_hist'][b])
        assert int(out['n_outer'][b]) == int(g['init_n_outer'][b]) and int(out['n_ipm'][b]) == int(g['ipm'][b])
        assert rel_err(out['pfrt_x'][b], g['step_x'][b]) < FIT_TOL
        assert rel_err(np.diagonal(out['pfrt_p'][b], axis1=1, axis2=2), g['step_p_diag'][b]) < FIT_TOL
        assert rel_err(out['pfrt_p'][b, -1], g['step_p_last'][b]) < FIT_TOL
    drt = DRT()
    res = drt.pfrt_fit_eis_batch(g['freq'], g['z'])
    pr = res.pfrt_result()
    assert rel_err(pr['step_x'], g['step_x']) < FIT_TOL
    assert np.max(np.abs(pr['step_llh'] - g['step_llh']) / np.abs(g['step_llh'])) < FIT_TOL
    drt.pfrt_fit_eis(g['freq'], g['z'][1])
    assert rel_err(np.array(drt.pfrt_result['step_x']), g['step_x'][1]) < FIT_TOL
    assert rel_err(drt.pfrt_result['step_p_mat'][-1], g['step_p_last'][1]) < FIT_TOL
    assert drt.qphb_params['n_outer'] == int(g['init_n_outer'][1])               # attributes = the initial fit
    assert rel_err(drt.cvx_result['x'] , g['step_x'][1][0]) < FIT_TOL
    drt.pfrt_fit_hybrid(g['h_times'], g['h_i'], g['h_v'], g['h_freq'], g['h_z'], factors=g['h_factors'])
    assert int(np.sum(drt.pfrt_result['step_iters'])) == int(g['h_n_hist'])
    assert rel_err(np.array(drt.pfrt_result['step_x']), g['h_step_x']) < FIT_TOL
    assert np.max(np.abs(np.array(drt.pfrt_result['step_llh']) - g['h_step_llh']) / np.abs(g['h_step_llh'])) < FIT_TOL
    assert rel_err(np.array([np.diag(m) for m in drt.pfrt_result['step_p_mat']]), g['h_step_p_diag']) < FIT_TOL


def E_default_hypers(**kw):
    from hybdrt_b200 import engine as E
    hyp = E.default_hypers()
    for k, v in kw.items():
        setattr(hyp, k, v)
    return hyp


def test_solve_rp_and_update_scale():
    """solve_rp (drt1d.py:573-607, qphb.py:1684-1717) and update_scale (drt1d.py:914-936) against the unmodified
    reference: EIS, DRT + DOP (column rescale of the DOP block) and hybrid (chrono scale attributes)."""
    from hybdrt_b200.models import DRT
    g = load_golden('rescale.npz')
    drt = DRT()
    for tag, kw in (('us', dict(update_scale=True)), ('rp', dict(solve_rp=True)),
                    ('both', dict(solve_rp=True, update_scale=True))):
        res = drt.fit_eis_batch(g['freq'], g['z'], **kw)
        h, fp = res.host(), res.fit_parameters()
        assert np.array_equal(h['n_outer'], g[f'{tag}_n_outer']) and np.array_equal(h['n_ipm'], g[f'{tag}_ipm'])
        for b in range(2):
            assert rel_err(h['x'][b], g[f'{tag}_cvx_x'][b]) < FIT_TOL
            assert rel_err(fp['x'][b], g[f'{tag}_x'][b]) < FIT_TOL
            assert abs(res.scales['coefficient_scale'][b] / g[f'{tag}_coefficient_scale'][b] - 1) < FIT_TOL
            assert rel_err(res.predict_z()[b], g[f'{tag}_z_pred'][b]) < FIT_TOL
        drt.fit_eis(g['freq'], g['z'][1], **kw)
        qp = drt.qphb_params
        assert rel_err(qp['rv'], g[f'{tag}_rv'][1]) < FIT_TOL and rel_err(qp['q_vector'], g[f'{tag}_q_vector'][1]) < FIT_TOL
        assert rel_err(qp['est_weights'], g[f'{tag}_est_weights'][1]) < FIT_TOL
        assert rel_err(qp['init_weights'], g[f'{tag}_init_weights'][1]) < FIT_TOL
        assert rel_err(qp['x_overfit_eis'], g[f'{tag}_x_overfit_eis'][1]) < FIT_TOL
        assert rel_err(qp['xmx_norms'], g[f'{tag}_xmx_norms'][1]) < FIT_TOL
    dd = DRT(fit_dop=True)
    res = dd.fit_eis_batch(g['dop_freq'], g['dop_z'], solve_rp=True)
    h, fp = res.host(), res.fit_parameters()
    assert np.array_equal(h['n_outer'], g['dop_n_outer']) and np.array_equal(h['n_ipm'], g['dop_ipm'])
    for b in range(2):
        assert rel_err(h['x'][b], g['dop_cvx_x'][b]) < FIT_TOL
        assert rel_err(fp['x_dop'][b], g['dop_x_dop'][b]) < FIT_TOL
        assert rel_err(h['dop_rho'][b], g['dop_dop_rho_vector'][b]) < FIT_TOL
        assert rel_err(res.predict_z()[b], g['dop_z_pred'][b]) < FIT_TOL
    dd.fit_eis(g['dop_freq'], g['dop_z'][0], solve_rp=True)
    assert rel_err(dd.dop_scale_vector, g['dop_dop_scale_vector'][0]) < FIT_TOL
    assert rel_err(dd.qphb_params['rm'], g['dop_rm0']) < FIT_TOL
    hy = load_golden('chrono_flex.npz')           # the small hybrid trace
    drt.fit_hybrid(hy['times'], hy['i_signal'], hy['v_signal'][0], hy['freq'], hy['z'][0], solve_rp=True,
                   update_scale=True)
    assert drt.qphb_params['n_outer'] == int(g['hyb_n_outer']) and drt.qphb_params['n_ipm'] == int(g['hyb_ipm'])
    assert rel_err(drt.cvx_result['x'], g['hyb_cvx_x']) < FIT_TOL
    assert abs(drt.response_signal_scale / float(g['hyb_response_signal_scale']) - 1) < FIT_TOL
    assert abs(drt.scaled_response_offset / float(g['hyb_scaled_response_offset']) - 1) < FIT_TOL
    assert rel_err(drt.fit_parameters['v_baseline'], g['hyb_v_baseline']) < 1e-5
    assert rel_err(drt.predict_z(hy['freq']), g['hyb_z_pred']) < FIT_TOL
    assert rel_err(drt.predict_response(), g['hyb_v_pred']) < FIT_TOL


def test_dop_with_time_domain_data(eng, orc):
    """phasance.construct_phasor_v_matrix (phasance.py:121-144) and DRT + DOP fits of chrono / hybrid data."""
    from hybdrt_b200.models import DRT
    d = load_golden('dop_chrono.npz')
    nu_eps = float(d['nu_epsilon'])
    rm = _np(eng.build_dop_v(d['times'][None], d['basis_nu'], d['step_times'][None], d['step_sizes'][None], nu_eps)[0])
    assert rel_err(rm, d['rm_dop']) < MAT_TOL
    t2 = np.linspace(-0.01, 0.5, 37)
    st2, sa2 = np.array([0.0, 0.2]), np.array([0.01, -0.02])
    rm2 = _np(eng.build_dop_v(t2[None], d['basis_nu'], st2[None], sa2[None], nu_eps)[0])
    assert rel_err(rm2, orc.dop_v_matrix(t2, d['basis_nu'], nu_eps, st2, sa2)) < MAT_TOL
    drt = DRT(fit_dop=True)
    drt.fit_hybrid(d['times'], d['i_signal'], d['v_signal'][0], d['freq'], d['z'][0])
    assert rel_err(drt.qphb_params['rm'], d['hyb_rm']) < 1e-9      # includes the fitted vz_offset column (a fit output, not a 1e-10 matrix entry)
    assert drt.qphb_params['n_outer'] == int(d['hyb_n_outer']) and drt.qphb_params['n_ipm'] == int(d['hyb_ipm'])
    assert rel_err(drt.cvx_result['x'], d['hyb_cvx_x']) < FIT_TOL
    assert rel_err(drt.fit_parameters['x_dop'], d['hyb_x_dop']) < FIT_TOL
    assert rel_err(drt.predict_z(d['freq']), d['hyb_z_pred']) < FIT_TOL
    assert rel_err(drt.predict_response(), d['hyb_v_pred']) < FIT_TOL
    drt.fit_chrono(d['times'], d['i_signal'], d['v_signal'][1])
    assert drt.qphb_params['n_outer'] == int(d['chr_n_outer']) and drt.qphb_params['n_ipm'] == int(d['chr_ipm'])
    assert rel_err(drt.cvx_result['x'], d['chr_cvx_x']) < FIT_TOL
    assert rel_err(drt.predict_response(), d['chr_v_pred']) < FIT_TOL


def test_chrono_downsampling(eng):
    """preprocessing.downsample_data (preprocessing.py:335-468) for a batch of raw traces: decimation index on the
    host, antialiasing filter on the GPU, against the unmodified reference; then a fit with downsample=True."""
    from hybdrt_b200 import synth, preprocessing as pp
    from hybdrt_b200.models import DRT
    from oracle import chrono_oracle as co
    g = load_golden('downsample.npz')
    times, i_sig, v = synth.make_raw_chrono_batch(3, seed=5)
    cases = dict(tut=dict(decimation_interval=8, decimation_factor=2, prestep_samples=25, method='decimate'),
                 size=dict(target_size=300, decimation_factor=2, prestep_samples=10, decimation_max_period=0.05,
                           method='decimate'),
                 raw=dict(decimation_interval=20, decimation_factor=3, prestep_samples=5, antialiased=False,
                          method='decimate'))
    for tag, kw in cases.items():
        st, si, sv, idx = pp.downsample_data(times, i_sig, v, engine=eng, **kw)
        assert np.array_equal(idx, g[f'{tag}_index']) and np.array_equal(st, g[f'{tag}_times'])
        assert rel_err(si, g[f'{tag}_i']) < 1e-12 and rel_err(sv, g[f'{tag}_v']) < 1e-12
    st, si, sv1, idx = pp.downsample_data(times, i_sig, v[1], engine=eng, **cases['tut'])      # one trace
    assert sv1.shape == (len(idx),) and rel_err(sv1, g['tut_v'][1]) < 1e-12
    # size-independent properties on a larger batch: constants and (inside a segment) straight lines pass through
    plan = pp.filter_plan(times, pp.identify_steps(i_sig, allow_consecutive=False), g['tut_index'])
    big = np.concatenate([np.full((1, len(times)), 3.25), np.repeat(times[None] * 2.0 - 1.0, 2, axis=0),
                          np.random.default_rng(3).normal(size=(509, len(times)))])
    out = _np(eng.filter_gather(big, plan))
    assert np.allclose(out[0], 3.25, rtol=0, atol=1e-13)
    inner = (plan['idx'] - plan['seg_lo'] >= plan['lw']) & (plan['idx'] - plan['seg_lo'] + plan['lw'] < plan['seg_len'])
    assert np.allclose(out[1][inner], (times * 2.0 - 1.0)[plan['idx']][inner], rtol=0, atol=1e-12)
    for b in (3, 400, 511):
        ref = co.filter_chrono_signal(times, big[b], co.identify_steps(i_sig, False), decimate_index=g['tut_index'])
        assert rel_err(out[b], ref[g['tut_index']]) < 1e-12
    drt = DRT()
    drt.warn = False
    drt.fit_chrono(times, i_sig, v[0], downsample=True, downsample_kw=dict(cases['tut'], step_model='ideal'))
    assert np.array_equal(drt.t_fit, g['fit_t']) and np.array_equal(drt.sample_index, g['tut_index'])
    assert drt.qphb_params['n_outer'] == int(g['fit_n_outer']) and drt.qphb_params['n_ipm'] == int(g['fit_ipm'])
    assert rel_err(drt.cvx_result['x'], g['fit_cvx_x']) < FIT_TOL
    assert rel_err(drt.predict_response(), g['fit_v_pred']) < FIT_TOL


def test_predictions_away_from_the_fit_grids():
    """predict_drt (drt1d.py:3040-3061), predict_z at new frequencies, predict_response at new times and with
    explicit steps (drt1d.py:3363-3461), single spectrum and batched, against the unmodified reference."""
    from hybdrt_b200.models import DRT
    from hybdrt_b200 import synth
    g = load_golden('predict.npz')
    hy = load_golden('chrono_flex.npz')           # the small hybrid trace
    drt = DRT()
    drt.fit_hybrid(hy['times'], hy['i_signal'], hy['v_signal'][0], hy['freq'], hy['z'][0])
    assert rel_err(drt.predict_z(g['f_new']), g['z_new']) < FIT_TOL
    assert rel_err(drt.predict_z(g['f_new'], x=drt.cvx_result['x']), g['z_raw_x']) < FIT_TOL
    assert rel_err(drt.predict_response(g['t_new']), g['v_new']) < FIT_TOL
    assert rel_err(drt.predict_response(g['t_new'], step_times=np.array([0.0, 0.3]), step_sizes=np.array([0.01, -0.01])),
                   g['v_steps']) < FIT_TOL
    for order, key in ((0, 'drt0'), (1, 'drt1'), (2, 'drt2')):
        assert rel_err(drt.predict_drt(g['tau'], order=order), g[key]) < FIT_TOL
    assert rel_err(drt.get_tau_eval(20), g['tau_default']) < 1e-13
    assert rel_err(drt.predict_drt(), g['drt_default']) < FIT_TOL
    assert rel_err(drt.predict_drt(g['tau'], normalize=True), g['drt_norm']) < FIT_TOL
    for name in ('r_inf', 'r_p', 'r_tot'):                  # drt1d.py:3552-3584
        assert abs(getattr(drt, 'predict_' + name)() - float(g[name])) < FIT_TOL * abs(float(g[name]))
    freq, z = synth.make_eis_batch(3, seed=0)
    res = DRT().fit_eis_batch(freq, z)
    assert rel_err(res.predict_z(g['f_new']), g['b_z_new']) < FIT_TOL
    assert rel_err(res.predict_drt(g['tau']), g['b_drt0']) < FIT_TOL


def test_kramers_kronig_test(eng, orc, lookup_golden):
    """DRT.kk_test (drt1d.py:1370-1491, models/kk.py): per-row weight factors in the fit kernel, residual
    statistics and window selection on the host, against the unmodified reference."""
    from hybdrt_b200.models import DRT
    g = load_golden('kk.npz')
    drt = DRT()
    for b in range(2):
        out_idx, (f_min, f_max), (fc, zc) = drt.kk_test(g['freq'], g['z'][b])
        assert np.array_equal(np.asarray(out_idx), g[f'outlier_index_{b}'])
        assert f_min == g[f'limits_{b}'][0] and f_max == g[f'limits_{b}'][1] and len(fc) == int(g[f'n_clean_{b}'])
        assert rel_err(drt.basis_tau, g[f'basis_tau_{b}']) < 1e-13
        # the fit is deliberately under-regularised (l2_lambda_0 = 1e-2, free sign): its coefficients are huge
        # cancelling numbers that no two implementations share, its prediction is what the test uses
        # (nor, exactly, its prediction: 50 non-converged passes amplify rounding differences to ~1e-4, and the
        # numpy oracle is no closer to the reference than the kernel is); the decisions above are exact
        z_ref = g['z'][b] - g[f'resid_{b}'] * np.abs(g['z'][b]) / 100
        # (any change of summation order in the kernel moves this number by a factor of a few)
        assert rel_err(drt.predict_z(g['freq']), z_ref) < 1e-2
    assert drt.extend_basis_decades == 1
    with pytest.raises(ValueError):
        drt.fit_eis(g['freq'], g['z'][0], weight_factor=np.ones(5))
    # the per-row weight factor on a well-posed fit: kernel against the oracle, at the usual tolerance
    prep = orc.EisPrep(g['freq'], tables=lookup_golden)
    wf = np.random.default_rng(2).uniform(0.5, 1.5, 2 * len(g['freq']))
    wf[[30, 100]] = 1e-10
    prob, _ = prep.problem(g['z'][1])
    out = _eis_launch(eng, prep, prob['rv'][None], weight_factor_vec=eng.dev(wf))
    ref = prep.fit(g['z'][1], weight_factor=wf)
    assert int(out['n_outer'][0]) == ref['n_outer'] and int(out['n_ipm'][0]) == int(ref['ipm_iters'].sum())
    assert rel_err(out['x'][0], ref['x']) < FIT_TOL and rel_err(out['weights'][0], ref['weights'] / wf) < FIT_TOL


def test_per_spectrum_frequency_grids(orc, lookup_golden):
    """fit_eis_batch with one frequency grid per spectrum (ragged sweeps): all matrices of all spectra are built in
    one launch per kind and read per spectrum by the fit kernel; each member must equal its own single-grid fit."""
    from hybdrt_b200.models import DRT
    from hybdrt_b200 import synth
    nb_ = 48
    base, z0 = synth.make_eis_batch(nb_, seed=31)
    freqs = base[None, :] * (1 + 2e-3 * np.arange(nb_)[:, None] / nb_) * (1 + 1e-4 * np.sin(np.arange(70)))[None, :]
    p, _ = synth.zarc_params(nb_, seed=31)
    z = synth.zarc_impedance(freqs[0], p) * 0
    for b in range(nb_):                                        # spectra evaluated on their own grids
        pb = {k: v[b:b + 1] for k, v in p.items()}
        z[b] = synth.zarc_impedance(freqs[b], pb)[0]
    z = z + (z0 - synth.zarc_impedance(base, p))                # the same noise realisation as the shared-grid batch
    drt = DRT()
    res = drt.fit_eis_batch(freqs, z)
    h, fp = res.host(), res.fit_parameters()
    assert res.plan['rm'].shape[:2] == (nb_, 140) and np.all((h['status'] & 3) != 0)
    single = DRT()
    for b in (0, 17, 47):
        single.fit_eis(freqs[b], z[b])
        assert rel_err(single.basis_tau, drt.basis_tau[b]) < 1e-13
        assert single.qphb_params['n_outer'] == int(h['n_outer'][b])
        assert rel_err(h['x'][b], single.cvx_result['x']) < 1e-10
        assert rel_err(res.predict_z()[b], single.predict_z(freqs[b])) < 1e-10
        assert rel_err(fp['x'][b], single.fit_parameters['x']) < 1e-10
    ref = orc.EisPrep(freqs[5], tables=lookup_golden).fit(z[5])
    assert int(h['n_outer'][5]) == ref['n_outer'] and rel_err(h['x'][5], ref['x']) < FIT_TOL
    with pytest.raises(ValueError):                             # grids that need basis grids of different sizes
        bad = freqs.copy()
        bad[3] = np.logspace(6, -3, 70)
        drt.fit_eis_batch(bad, z)
    with pytest.raises(ValueError):
        drt.fit_eis_batch(freqs[:5], z)


def test_hybrid_fit_options():
    """hybrid_weight_factor_method='weight' / 'rp' (drt1d.py:745-800) and init_weights_separately (:647-669) against
    the unmodified reference."""
    from hybdrt_b200.models import DRT
    g = load_golden('hybrid_options.npz')
    hy = load_golden('chrono_flex.npz')           # the small hybrid trace
    args = (hy['times'], hy['i_signal'], hy['v_signal'][0], hy['freq'], hy['z'][0])
    drt = DRT()
    for tag, kw in (('weight', dict(hybrid_weight_factor_method='weight')), ('rp', dict(hybrid_weight_factor_method='rp')),
                    ('sep', dict(init_weights_separately=True))):
        drt.fit_hybrid(*args, **kw)
        qp = drt.qphb_params
        assert qp['n_outer'] == int(g[f'{tag}_n_outer']) and qp['n_ipm'] == int(g[f'{tag}_ipm']), tag
        assert rel_err(np.array([qp['chrono_weight_factor'], qp['eis_weight_factor']]), g[f'{tag}_factors']) < FIT_TOL, tag
        assert rel_err(drt.cvx_result['x'], g[f'{tag}_cvx_x']) < FIT_TOL, tag
        assert rel_err(qp['weights'], g[f'{tag}_weights']) < FIT_TOL, tag
        assert rel_err(qp['est_weights'], g[f'{tag}_est_weights']) < FIT_TOL, tag
        assert rel_err(drt.predict_z(hy['freq']), g[f'{tag}_z_pred']) < FIT_TOL, tag
    assert rel_err(drt.qphb_params['x_overfit_eis'], g['sep_x_overfit_eis']) < FIT_TOL
    assert rel_err(drt.qphb_params['x_overfit_chrono'], g['sep_x_overfit_chrono']) < FIT_TOL
    with pytest.raises(ValueError):
        drt.fit_hybrid(*args, hybrid_weight_factor_method='nonsense')


def test_builder_and_filter_edges(eng, orc, lookup_golden):
    """Empty and degenerate inputs of the L0 / L1 entry points: zero grids, a single frequency or basis function,
    an odd number of entries per grid (the 16-byte store path), filter windows several times longer than the step
    segment (repeated mirroring), an empty batch of traces."""
    from hybdrt_b200 import engine as E
    from oracle import chrono_oracle as co
    eps = float(lookup_golden['eps'])
    a_re, a_im = eng.build_impedance(np.zeros((0, 7)), np.zeros((0, 5)), eps, E.MODE_INTERP)
    assert a_re.shape == (0, 7, 5)
    for nf, nb in ((1, 1), (1, 9), (7, 1), (3, 5), (64, 33)):
        freq = np.logspace(5, -1, nf)[None] if nf > 1 else np.array([[37.0]])
        tau = np.logspace(-6, 1, nb)[None] if nb > 1 else np.array([[2e-3]])
        a_re, a_im = eng.build_impedance(freq, tau, eps, E.MODE_INTERP)
        assert rel_err(_np(a_re[0]), orc.impedance_matrix(freq[0], tau[0], eps, 'real', 'interp', lookup_golden)) < MAT_TOL
        assert rel_err(_np(a_im[0]), orc.impedance_matrix(freq[0], tau[0], eps, 'imag', 'interp', lookup_golden)) < MAT_TOL
    g3 = np.repeat(np.logspace(5, -1, 7)[None], 3, 0) * np.array([[1.0], [1.1], [1.2]])     # odd grid offsets: 7 x 5
    t3 = np.repeat(np.logspace(-6, 1, 5)[None], 3, 0)
    a_re, _ = eng.build_impedance(g3, t3, eps, E.MODE_INTERP)
    for i in range(3):
        assert rel_err(_np(a_re[i]), orc.impedance_matrix(g3[i], t3[i], eps, 'real', 'interp', lookup_golden)) < MAT_TOL
    # filter: one 5-sample and one 40-sample segment, windows of 24 taps each side
    y = np.random.default_rng(9).normal(size=(3, 45))
    lw, w = 24, np.exp(-0.5 / 36.0 * np.arange(-24, 25) ** 2)      # sigma 6, truncate 4
    w /= w.sum()
    idx = np.array([0, 2, 4, 5, 20, 44])
    plan = dict(idx=idx.astype(np.int32), seg_lo=np.where(idx < 5, 0, 5).astype(np.int32),
                seg_len=np.where(idx < 5, 5, 40).astype(np.int32), lw=np.full(6, lw, np.int32),
                woff=(np.arange(6) * 49).astype(np.int64), taps=np.tile(w, 6))
    out = _np(eng.filter_gather(y, plan))
    for b in range(3):
        ref = np.concatenate([co.gaussian_filter1d_reflect(y[b, :5], 6.0), co.gaussian_filter1d_reflect(y[b, 5:], 6.0)])
        assert rel_err(out[b], ref[idx]) < 1e-12
    assert eng.filter_gather(np.zeros((0, 45)), plan).shape == (0, 6)
    assert _np(eng.build_chrono_vmm(np.array([[0.0, 1.0]]), np.array([[0.5]]), 4.0)).shape == (1, 2, 2)


def test_unsupported_options_raise():
    from hybdrt_b200.models import DRT
    c2 = load_golden('c2_eis.npz')
    with pytest.raises(NotImplementedError):
        DRT(tau_basis_type='Cole-Cole')
    drt = DRT()
    for kw in (dict(series_neg=True), dict(penalty_type='discrete')):
        with pytest.raises(NotImplementedError):
            drt.fit_eis(c2['freq'], c2['z'][0], **kw)


def test_mapping_drtmd_against_the_reference():
    """hybdrt.mapping.DRTMD.fit_all (reference, one fit per observation) vs the batched dispatch."""
    from hybdrt_b200.mapping import DRTMD
    g = load_golden('drtmd_small.npz')
    md = DRTMD(tau_supergrid=g['tau_supergrid'], psi_dim_names=['row', 'col'], print_progress=False)
    assert abs(md.tau_epsilon - float(g['tau_epsilon'])) < 1e-12
    for b in range(len(g['z'])):
        md.add_observation(g['psi'][b], None, (g['freq'], g['z'][b]))
    assert md.num_obs == 6 and not md.obs_fit_status.any()
    md.fit_all()
    assert md.obs_fit_status.all() and list(md.fitted_obs_index) == list(range(6))
    assert np.array_equal(np.array(md.obs_tau_indices), g['obs_tau_indices'])
    for b in range(6):
        assert rel_err(md.obs_x[b], g['obs_x'][b]) < FIT_TOL
    assert rel_err(md.obs_special['R_inf'], g['special_R_inf']) < FIT_TOL
    assert rel_err(md.obs_special['inductance'], g['special_inductance']) < FIT_TOL
    # post-fit diagnostics (drtmd.py:256-279): distribution variance on the supergrid, llh, rss
    for b in range(6):
        # the distribution variance goes through P^-1 (drt1d.py:3063-3151): conditioning amplifies the 1e-6-class fit error
        assert rel_err(md.obs_drt_var[b], g['obs_drt_var'][b]) < 1e-5
    assert rel_err(md.obs_rss, g['obs_rss']) < FIT_TOL and rel_err(md.obs_llh, g['obs_llh']) < FIT_TOL
    # batched == one at a time, and refit / incremental adds keep the containers consistent
    md2 = DRTMD(tau_supergrid=g['tau_supergrid'], print_progress=False)
    for b in range(3):
        md2.add_observation(g['psi'][b], None, (g['freq'], g['z'][b]), fit=True)
    md2.add_observations(g['psi'][3:], g['freq'], g['z'][3:])
    assert md2.obs_fit_status.tolist() == [True] * 3 + [False] * 3
    md2.fit_all()
    assert np.allclose(md2.obs_x, md.obs_x, rtol=1e-12, atol=0) and md2.obs_special['R_inf'].shape == (6,)
    # a second measurement grid goes into its own batch
    md2.add_observation([9, 9], None, (g['freq'][::2], g['z'][0][::2]))
    md2.fit_all()
    assert md2.obs_fit_status.all() and md2.obs_tau_indices[6] is not None
    with pytest.raises(ValueError):
        md2.add_observation([0, 0], None, (g['freq'],))


def test_mapping_drtmd_pfrt_against_the_reference():
    """DRTMD(fit_type='pfrt') (drtmd.py:1140-1160, 1304-1342): one solution per factor and observation; the
    diagnostics are those of the initial fit."""
    from hybdrt_b200.mapping import DRTMD
    g = load_golden('drtmd_pfrt.npz')
    md = DRTMD(tau_supergrid=g['tau_supergrid'], psi_dim_names=['row', 'col'], print_progress=False, fit_type='pfrt')
    assert np.allclose(md.pfrt_factors, g['pfrt_factors'])
    md.add_observations(g['psi'], g['freq'], g['z'])
    md.fit_all()
    assert md.obs_fit_status.all() and md.obs_x.shape == g['obs_x'].shape
    assert np.array_equal(np.array(md.obs_tau_indices), g['obs_tau_indices'])
    assert rel_err(md.obs_x, g['obs_x']) < FIT_TOL
    assert rel_err(md.obs_special['R_inf'], g['special_R_inf']) < FIT_TOL
    assert rel_err(md.obs_special['inductance'], g['special_inductance']) < FIT_TOL
    assert rel_err(md.obs_drt_var, g['obs_drt_var']) < 1e-5
    assert rel_err(md.obs_rss, g['obs_rss']) < FIT_TOL and rel_err(md.obs_llh, g['obs_llh']) < FIT_TOL
    with pytest.raises(ValueError):
        DRTMD(tau_supergrid=g['tau_supergrid'], fit_type='nonsense')


def test_full_size_c2_batch_properties():
    """BASELINE config C2 at its full size (10,000 spectra) through the public API: size-independent properties,
    plus the oracle on a few members of the big batch."""
    from hybdrt_b200 import synth
    from hybdrt_b200.models import DRT
    from oracle import drt_oracle as orc
    freq, z = synth.make_eis_batch(10000, seed=0)
    drt = DRT()
    res = drt.fit_eis_batch(freq, z)
    h = res.host(['x', 'status', 'n_outer', 'n_ipm', 'weights'])
    st = h['status']
    assert np.all((st & 3) != 0) and np.all((st & 3) != 3) and not np.any(st & (8 | 16))
    # interior-point iterates: the bound holds to the solver's feasibility tolerance (cvxopt feastol 1e-7, relative)
    assert np.all(np.isfinite(h['x'])) and h['x'].min() > -1e-5 and np.all(h['weights'] > 0)
    assert 2 <= h['n_outer'].min() and h['n_outer'].max() <= 50
    idx = np.random.default_rng(1).permutation(10000)[:24]
    sub = drt.fit_eis_batch(freq, z[idx]).host(['x', 'n_ipm'])
    assert np.array_equal(sub['x'], h['x'][idx]) and np.array_equal(sub['n_ipm'], h['n_ipm'][idx])   # batch independent
    fp = res.fit_parameters()
    zp = res.predict_z()
    rel = np.linalg.norm(zp - z, axis=1) / np.linalg.norm(z, axis=1)
    assert np.median(rel) < 3e-2 and rel.max() < 0.2        # every spectrum is fitted to about its noise level (0.5 % of Rp)
    prep = orc.EisPrep(freq)
    for b in idx[:3]:
        ref = prep.fit(z[b])
        assert int(h['n_outer'][b]) == ref['n_outer']
        assert rel_err(h['x'][b], ref['x']) < FIT_TOL
        assert abs(fp['R_inf'][b] - ref['params']['R_inf']) < FIT_TOL * abs(ref['params']['R_inf'])


def test_full_size_c3_c4_c5_properties():
    """BASELINE configs C3 (4,096 hybrid fits, N = 2060), C4 (10,000 DRT + DOP fits, n = 153) and C5 (256 x 256 map
    through DRTMD) at their full sizes: size-independent properties, and members of the big batches against the
    same spectra fitted in small batches (bitwise: one code path, independent spectra)."""
    from hybdrt_b200 import synth
    from hybdrt_b200.models import DRT
    from hybdrt_b200.mapping import DRTMD
    rng = np.random.default_rng(4)
    # ---- C3
    t, i_sig, v, f, z = synth.make_hybrid_batch(4096, seed=1)
    drt = DRT()
    res = drt.fit_hybrid_batch(t, i_sig, v, f, z)
    h = res.host(['x', 'status', 'n_outer', 'weights'])
    st = h['status']
    assert res.plan['n_rows'] == 2060 and np.all((st & 3) != 0) and np.all((st & 3) != 3) and not np.any(st & (8 | 16))
    assert np.all(np.isfinite(h['x'])) and np.all(h['weights'] > 0)
    zp = res.predict_z()
    rel = np.linalg.norm(zp - z, axis=1) / np.linalg.norm(z, axis=1)
    assert np.median(rel) < 0.15 and rel.max() < 0.3         # (the reference's own fits of these spectra sit at ~8 %)
    idx = rng.permutation(4096)[:6]
    sub = drt.fit_hybrid_batch(t, i_sig, v[idx], f, z[idx]).host(['x', 'n_ipm'])
    assert np.array_equal(sub['x'], h['x'][idx])
    # ---- C4
    f4, z4 = synth.make_dop_batch(10000, seed=2)
    dd = DRT(fit_dop=True)
    res = dd.fit_eis_batch(f4, z4)
    h = res.host(['x', 'status', 'n_outer'])
    st = h['status']
    assert res.plan['n'] == 153 and np.all((st & 3) != 0) and not np.any(st & (8 | 16)) and np.all(np.isfinite(h['x']))
    assert h['x'].min() > -1e-5                               # DRT and DOP coefficients are non-negative
    rel = np.linalg.norm(res.predict_z() - z4, axis=1) / np.linalg.norm(z4, axis=1)
    assert np.median(rel) < 3e-2
    idx = rng.permutation(10000)[:6]
    assert np.array_equal(dd.fit_eis_batch(f4, z4[idx]).host(['x'])['x'], h['x'][idx])
    # ---- C5
    rows = cols = 256
    f5, z5 = synth.make_map_batch(rows, cols, seed=3)
    psi = np.stack(np.meshgrid(np.arange(rows), np.arange(cols), indexing='ij'), axis=-1).reshape(-1, 2).astype(float)
    md = DRTMD(tau_supergrid=np.logspace(-8, 3, 111), psi_dim_names=['row', 'col'], print_progress=False)
    md.add_observations(psi, f5, z5)
    md.fit_all()
    assert md.obs_fit_status.all() and np.all(np.isfinite(md.obs_x)) and np.all(np.isfinite(md.obs_drt_var))
    assert md.obs_x.min() > -1e-5 and np.all(md.obs_drt_var >= 0) and np.all(np.isfinite(md.obs_llh))
    # the map varies smoothly: R_inf is constant (1.0) over the map up to noise, and neighbours have close Rp
    assert abs(np.median(md.obs_special['R_inf']) - 1.0) < 0.02
    rp = (md.obs_x * md.tau_basis_area).sum(axis=1).reshape(rows, cols)
    assert np.median(np.abs(np.diff(rp, axis=0))) < 0.05 * np.median(rp)
    md2 = DRTMD(tau_supergrid=np.logspace(-8, 3, 111), print_progress=False)
    pick = rng.permutation(rows * cols)[:5]
    md2.add_observations(psi[pick], f5, z5[pick])
    md2.fit_all()
    assert np.array_equal(md2.obs_x, md.obs_x[pick])



# ---------------------------------------------------------------------------------------------------
# Parity at scale (hard part 1): the kernel against the CPU oracle on >= 1,024 C2 spectra and
# on >= 64 spectra of the hybrid (C3) and DRT + DOP (C4) configurations -- fraction of spectra within 1e-6, outer and
# interior-point iteration counts per spectrum, offenders listed by index.  The oracle runs in a process pool on
# the matrices of the kernel's own plan (the matrices themselves are pinned at 1e-10 above).
# ---------------------------------------------------------------------------------------------------
_POOL = {}


def _pool_init(prob):
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    _POOL['prob'] = prob


def _pool_fit(rv):
    from oracle import drt_oracle
    prob = dict(_POOL['prob'])
    prob['rv'] = rv
    res = drt_oracle.qphb_fit(prob, prob.pop('hypers', None))
    return res['n_outer'], int(np.sum(res['ipm_iters'])), np.asarray(res['x'], dtype=float)


def _oracle_pool_fits(prob, rvs):
    import multiprocessing as mp
    import os
    cores = max(1, min(32, len(os.sched_getaffinity(0))))
    with mp.get_context('spawn').Pool(cores, initializer=_pool_init, initargs=(prob,)) as pool:
        return pool.map(_pool_fit, list(rvs), chunksize=max(1, len(rvs) // (cores * 4)))


def _parity_report(ref, h, tag, min_frac=1.0):
    rel = np.array([np.max(np.abs(h['x'][i] - r[2])) / np.max(np.abs(r[2])) for i, r in enumerate(ref)])
    outer_bad = [i for i, r in enumerate(ref) if int(r[0]) != int(h['n_outer'][i])]
    ipm_bad = [i for i, r in enumerate(ref) if int(r[1]) != int(h['n_ipm'][i])]
    frac = float(np.mean(rel <= FIT_TOL))
    print(f'{tag}: n = {len(ref)}, fraction within 1e-6 = {frac:.4f}, worst rel = {rel.max():.3e} (spectrum {int(np.argmax(rel))}), '
          f'outer-iteration mismatches {outer_bad[:10]}, interior-point mismatches {ipm_bad[:10]}')
    assert not outer_bad, (tag, 'outer iteration counts differ', outer_bad[:10])
    assert not ipm_bad, (tag, 'interior-point iteration counts differ', ipm_bad[:10])
    assert frac >= min_frac, (tag, frac, [(int(i), float(rel[i])) for i in np.argsort(-rel)[:5]])


def test_parity_at_scale_c2():
    from hybdrt_b200 import synth
    from hybdrt_b200.models import DRT
    n = 1024
    freq, z = synth.make_eis_batch(10000, seed=0)
    drt = DRT()
    res = drt.fit_eis_batch(freq, z[:n])
    h = res.host(['x', 'n_outer', 'n_ipm'])
    plan = res.plan
    prob = dict(rm=_np(plan['rm']), vmm=_np(plan['vmm_eis']), pen=list(_np(plan['pen'])), h=_np(plan['h']), l1=_np(plan['l1']),
                n_special=plan['n_special'])
    zs = z[:n] / res.scales['coefficient_scale'][:, None]
    ref = _oracle_pool_fits(prob, np.concatenate([zs.real, zs.imag], axis=1))
    _parity_report(ref, h, 'C2 (70 f x 101 basis)')


def test_parity_at_scale_c3_c4():
    from hybdrt_b200 import synth
    from hybdrt_b200.models import DRT
    n = 64
    # ---- C3: hybrid fits, 2,000 chrono samples + 30 frequencies, vz_offset column rewritten every iteration
    t, i_sig, v, f, z = synth.make_hybrid_batch(n, seed=1)
    drt = DRT()
    res = drt.fit_hybrid_batch(t, i_sig, v, f, z)
    h = res.host(['x', 'n_outer', 'n_ipm'])
    plan = res.plan
    nc = plan['n_chrono']
    sp = plan['special_qp_params']
    hyp = drt._c_hypers(plan['opts'])
    prob = dict(rm=_np(plan['rm']), vmm=dict(n_chrono=nc, chrono=None, eis=_np(plan['vmm_eis'])), pen=list(_np(plan['pen'])),
                h=_np(plan['h']), l1=_np(plan['l1']), n_special=plan['n_special'], n_chrono=nc,
                vz_index=sp['vz_offset']['index'], vb_range=drt.get_special_indices('v_baseline'),
                vz_strength=plan['vz_strength_host'], chrono_weight_factor=float(hyp.chrono_weight_factor),
                eis_weight_factor=float(hyp.eis_weight_factor))
    prob['rm'][:, prob['vz_index']] = 0.0
    ref = _oracle_pool_fits(prob, _np(res.extra['rv_dev']))
    _parity_report(ref, h, 'C3 (hybrid, N = 2060)')
    # ---- C4: DRT + DOP, n = 153
    f4, z4 = synth.make_dop_batch(n, seed=2)
    dd = DRT(fit_dop=True)
    res = dd.fit_eis_batch(f4, z4)
    h = res.host(['x', 'n_outer', 'n_ipm'])
    plan = res.plan
    prob = dict(rm=_np(plan['rm']), vmm=_np(plan['vmm_eis']), pen=list(_np(plan['pen'])), h=_np(plan['h']), l1=_np(plan['l1']),
                n_special=plan['n_special'], dop_range=dd.get_special_indices('x_dop'))
    zs = z4 / res.scales['coefficient_scale'][:, None]
    ref = _oracle_pool_fits(prob, np.concatenate([zs.real, zs.imag], axis=1))
    _parity_report(ref, h, 'C4 (DRT + DOP, n = 153)')


# ---------------------------------------------------------------------------------------------------
# Multi-GPU: a map fitted through DRTMD.fit_all(shard=True) on two ranks (NCCL gather of the results) equals the
# single-rank result.  Needs two GPUs; the host logic alone is covered with gloo in tests/test_sharding.py.
# ---------------------------------------------------------------------------------------------------
def _map_worker(rank, ws, port, q):
    import os
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=ws, device_id=torch.device('cuda', rank))
    try:
        from hybdrt_b200 import synth
        from hybdrt_b200.mapping import DRTMD
        rows, cols = 12, 11
        f, z = synth.make_map_batch(rows, cols, seed=3)
        psi = np.array([(r, c) for r in range(rows) for c in range(cols)], dtype=float)
        md = DRTMD(tau_supergrid=np.logspace(-8, 3, 111), psi_dim_names=['row', 'col'], print_progress=False, device=rank)
        md.add_observations(psi, f, z)
        md.fit_all(shard=True)
        q.put((rank, md.obs_x.copy(), md.obs_drt_var.copy(), md.obs_llh.copy(), md.obs_outer_iterations.copy()))
    finally:
        dist.destroy_process_group()


def test_sharded_map_on_two_gpus_equals_single_rank():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (run with gpurun --gpus 2)')
    import socket
    import torch.multiprocessing as mp
    from hybdrt_b200 import synth
    from hybdrt_b200.mapping import DRTMD
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_map_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(60)
    rows, cols = 12, 11
    f, z = synth.make_map_batch(rows, cols, seed=3)
    psi = np.array([(r, c) for r in range(rows) for c in range(cols)], dtype=float)
    md = DRTMD(tau_supergrid=np.logspace(-8, 3, 111), psi_dim_names=['row', 'col'], print_progress=False)
    md.add_observations(psi, f, z)
    md.fit_all()
    for rank, x, var, llh, outer in got:
        assert np.array_equal(x, md.obs_x) and np.array_equal(var, md.obs_drt_var), rank
        assert np.array_equal(llh, md.obs_llh) and np.array_equal(outer, md.obs_outer_iterations), rank


def test_consecutive_steps_and_hybrid_map_against_the_reference():
    """(1) Finite-rise current steps with chrono_error_structure=None: the decorrelation blocks of the variance
    matrix follow the non-consecutive step times (drt1d.py:616).  (2) A hybrid map: obs_llh / obs_rss are evaluated
    with the vz_offset column rewritten from the final coefficients (drt1d.py:972-979, 4433-4496)."""
    from hybdrt_b200.models import DRT
    from hybdrt_b200.mapping import DRTMD
    g = load_golden('consec_steps.npz')
    drt = DRT()
    drt.fit_chrono(g['times'], g['i_signal'], g['v_signal'], error_structure=None)
    assert rel_err(drt.step_times, g['step_times']) < 1e-12 and rel_err(drt.nonconsec_step_times, g['nonconsec_step_times']) < 1e-12
    assert len(drt.step_times) > len(drt.nonconsec_step_times)
    assert drt.qphb_params['n_outer'] == int(g['n_outer']) and drt.qphb_params['n_ipm'] == int(g['ipm'])
    assert rel_err(drt.cvx_result['x'], g['cvx_x']) < FIT_TOL
    assert rel_err(drt.qphb_params['est_weights'], g['est_weights']) < FIT_TOL
    assert rel_err(drt.predict_response(), g['v_pred']) < FIT_TOL
    h = load_golden('drtmd_hybrid.npz')
    md = DRTMD(tau_supergrid=h['tau_supergrid'], psi_dim_names=['k'], print_progress=False)
    for b in range(3):
        md.add_observation([float(b)], (h['times'], h['i_signal'], h['v'][b]), (h['freq'], h['z'][b]))
    md.fit_all()
    assert md.obs_fit_status.all() and np.array_equal(np.array(md.obs_tau_indices), h['obs_tau_indices'])
    for b in range(3):
        assert rel_err(md.obs_x[b], h['obs_x'][b]) < FIT_TOL
        assert rel_err(md.obs_drt_var[b], h['obs_drt_var'][b]) < 1e-5
    for key in ('R_inf', 'inductance', 'v_baseline', 'vz_offset'):
        assert rel_err(md.obs_special[key], h['special_' + key]) < FIT_TOL, key
    assert rel_err(md.obs_rss, h['obs_rss']) < FIT_TOL and rel_err(md.obs_llh, h['obs_llh']) < FIT_TOL


def test_cross_observation_resolve_against_the_reference():
    """mapping/resolve.py:176-341 through DRTMD.resolve_observations / resolve_group (drtmd.py:432-560): one window of
    seven hybrid observations, then a group of nine in windows of seven with two shared neighbours (margin-weighted
    averaging).  Interior-point iteration counts identical to the reference's, resolved parameters within 1e-6."""
    from hybdrt_b200.mapping import DRTMD
    g = load_golden('drtmd_resolve.npz')
    md = DRTMD(tau_supergrid=g['tau_supergrid'], psi_dim_names=['k'], print_progress=False, keep_pq=True)
    for b in range(9):
        md.add_observation([float(b)], (g['times'], g['i_signal'], g['v'][b]), (g['freq'], g['z'][b]), group_id='g')
    md.fit_all()
    assert md.obs_fit_status.all()
    for b in range(9):
        assert rel_err(md.obs_x[b], g['obs_x'][b]) < FIT_TOL
    md.resolve_observations(np.arange(7), psi_sort_dims=['k'], sigma=1, lambda_psi=1)
    assert md.last_resolve['iters'].tolist() == g['win_ipm'].tolist() and not md.last_resolve['status'].any()
    assert md.obs_resolve_status[:7].all() and not md.obs_resolve_status[7:].any()
    for b in range(7):
        assert rel_err(md.obs_x_resolved[b], g['win_x_resolved'][b]) < FIT_TOL, b
    for key in ('R_inf', 'inductance'):
        assert rel_err(md.obs_special_resolved[key][:7], g['win_special_' + key]) < FIT_TOL, key
    md.resolve_group('g', batch_size=7, overlap=2, psi_sort_dims=['k'], sigma=1, lambda_psi=1)
    assert md.last_resolve['iters'].tolist() == g['grp_ipm'].tolist()
    assert md.obs_resolve_status.all()
    for b in range(9):
        assert rel_err(md.obs_x_resolved[b], g['grp_x_resolved'][b]) < FIT_TOL, b
    for key in ('R_inf', 'inductance'):
        assert rel_err(md.obs_special_resolved[key], g['grp_special_' + key]) < FIT_TOL, key
    assert np.all(np.isfinite(md.obs_x_resolved)) and md.obs_x_resolved.min() > -1e-6
    with pytest.raises(ValueError):
        DRTMD(tau_supergrid=g['tau_supergrid'], print_progress=False)._resolve_windows([np.arange(2)], False, 1, 1)
