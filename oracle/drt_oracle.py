"""CPU ORACLE (test infrastructure, not product code).

numpy restatement of the reference's hot path: response-matrix construction (hybdrt/matrices) and
the self-tuning hierarchical-Bayes QPHB loop (hybdrt/models/qphb.py + the loop in
hybdrt/models/drt1d.py:873-988).  Every function cites the reference lines it follows.  The QP is
oracle/coneqp.py (restatement of cvxopt's coneqp -- see that file's header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (hybrid-drt_b200/) never does; it fails loudly without its CUDA library.

Parity status: pinned.  tests/test_oracle_golden.py checks this module against fixtures produced by
the UNMODIFIED reference (oracle/make_golden.py), which include the reference's own golden vector
(tests/test_drt_fit.py:55-141).  fit_chrono / fit_hybrid / DOP / trapz have no golden vector in the
reference's tests; for those the fixtures are outputs of the reference itself run in the authoring
container on top of the coneqp restatement.
"""
import math

import numpy as np
from scipy.special import erf

from .coneqp import coneqp_orthant

QUAD_Y = np.linspace(-20, 20, 1000)      # basis.py:660, mat1d.py:350
RE_LIM = 2.7                             # basis.py:654
IM_LIM = 5.4                             # basis.py:655
TD_LIM = (-6.0, 2.0)                     # basis.py:679


# ------------------------------------------------------------------------------------------------
# L1: matrices
# ------------------------------------------------------------------------------------------------
def rbf(y, eps):
    """Gaussian radial basis function, basis.py:93-95."""
    return np.exp(-(eps * y) ** 2)


def z_integrand(y, ln_wt, eps, part):
    """Impedance integrands in y = ln(tau/tau_m), basis.py:562-570 (ln_wt = ln(omega_n tau_m))."""
    den = 1 + np.exp(2 * (y + ln_wt))
    if part == 'real':
        return rbf(y, eps) / den
    return -rbf(y, eps) * np.exp(y) * np.exp(ln_wt) / den


def v_integrand(y, td, eps):
    """Ideal galvanostatic step-response integrand, basis.py:616-618 (td = (t - t_step)/tau_m)."""
    return rbf(y, eps) * (1 - np.exp(-td / np.exp(y)))


def lookup_tables(eps, grid_points=2000):
    """basis.generate_impedance_lookup / generate_response_lookup, basis.py:648-689."""
    wt_re = np.logspace(-RE_LIM, RE_LIM, grid_points)
    wt_im = np.logspace(-IM_LIM, IM_LIM, grid_points)
    td = np.logspace(TD_LIM[0], TD_LIM[1], grid_points)
    re_v = np.array([np.trapezoid(z_integrand(QUAD_Y, np.log(w), eps, 'real'), x=QUAD_Y) for w in wt_re])
    im_v = np.array([np.trapezoid(z_integrand(QUAD_Y, np.log(w), eps, 'imag'), x=QUAD_Y) for w in wt_im])
    rs_v = np.array([np.trapezoid(v_integrand(QUAD_Y, t, eps), x=QUAD_Y) for t in td])
    return dict(re_x=np.log(wt_re), re_v=re_v, im_x=np.log(wt_im), im_v=im_v,
                resp_x=np.log(td), resp_v=rs_v)


def impedance_matrix(freq, tau, eps, part, mode='interp', tables=None):
    """mat1d.construct_impedance_matrix, mat1d.py:212-374, evaluated entry by entry.

    The reference's Toeplitz shortcut (mat1d.py:341-360) copies the first row/column along the
    diagonals; it agrees with entry-wise evaluation to 4.4e-16.
    """
    omega = 2 * np.pi * np.asarray(freq, dtype=float)
    tau = np.asarray(tau, dtype=float)
    out = np.empty((omega.size, tau.size))
    if mode == 'interp':
        gx, gv = (tables['re_x'], tables['re_v']) if part == 'real' else (tables['im_x'], tables['im_v'])
        for n, w in enumerate(omega):
            out[n] = np.interp(np.log(w * tau), gx, gv)
    elif mode == 'trapz':
        for n, w in enumerate(omega):
            for m, t in enumerate(tau):
                out[n, m] = np.trapezoid(z_integrand(QUAD_Y, np.log(w * t), eps, part), x=QUAD_Y)
    else:
        raise ValueError(mode)
    return out


def response_matrix(tau, times, step_times, step_sizes, eps, mode='interp', tables=None):
    """mat1d.construct_response_matrix (galvanostatic, ideal steps), mat1d.py:16-122."""
    tau = np.asarray(tau, dtype=float)
    times = np.asarray(times, dtype=float)
    out = np.zeros((times.size, tau.size))
    for st, sa in zip(step_times, step_sizes):
        post = times > st
        if not post.any():
            continue
        dt = times[post] - st
        if mode == 'interp':
            blk = np.array([np.interp(np.log(d / tau), tables['resp_x'], tables['resp_v']) for d in dt])
        else:
            blk = np.array([[np.trapezoid(v_integrand(QUAD_Y, d / t, eps), x=QUAD_Y) for t in tau] for d in dt])
        out[post] += sa * blk
    return out


def _pen_func(a, eps, order):
    """Closed-form integrated derivative products, basis.py:382-395."""
    g = np.exp(-(a ** 2 / 2))
    c = (np.pi / 2) ** 0.5
    if order == 0:
        return c * eps ** (-1) * g
    if order == 1:
        return -c * eps * (-1 + a ** 2) * g
    if order == 2:
        return c * eps ** 3 * (3 - 6 * a ** 2 + a ** 4) * g
    raise ValueError(order)


def is_uniform(x):
    """utils/array.py:142-151."""
    d = np.diff(x)
    return np.std(d) / np.mean(d) <= 0.01


def penalty_matrix(grid, order, eps):
    """mat1d.construct_integrated_derivative_matrix, mat1d.py:125-209 (gaussian, no limits).

    Uniform grids use the reference's symmetric-Toeplitz shortcut: entry (i, j) is the function of
    grid[|i-j|] - grid[0] (mat1d.py:158-168), not of grid[i] - grid[j].
    """
    grid = np.asarray(grid, dtype=float)
    n = grid.size
    if is_uniform(grid):
        col = _pen_func(eps * (grid[0] - grid), eps, order)
        idx = np.abs(np.arange(n)[:, None] - np.arange(n)[None, :])
        return col[idx]
    return _pen_func(eps * (grid[None, :] - grid[:, None]), eps, order)


def eis_vmm(freq, vmm_eps=0.25, reim_cor=0.25, structure=None):
    """mat1d.construct_eis_var_matrix, mat1d.py:493-515."""
    lf = np.log(np.asarray(freq, dtype=float))
    n = lf.size
    main = np.ones((n, n)) if structure == 'uniform' else rbf(lf[:, None] - lf[None, :], vmm_eps)
    vmm = np.block([[main, main * reim_cor], [main * reim_cor, main]])
    return vmm / vmm.sum(axis=1)[:, None]


def time_transform(times, step_times):
    """utils.chrono.get_time_transforms forward transform, chrono.py:5-39: linear before the first step, then
    log(time since the step) per segment, segments laid end to end."""
    times = np.asarray(times, dtype=float)
    start = np.asarray(step_times, dtype=float)
    t_sample = np.min(np.diff(times))
    base = np.log(t_sample / 4)
    offsets = np.concatenate([[0], np.cumsum(np.log(start[1:] - start[:-1]) - base)])
    tt = np.zeros_like(times)
    seg = np.zeros(times.size, dtype=int)
    pre = times < start[0]
    tt[pre] = times[pre] - start[0]
    for i, s0 in enumerate(start):
        s1 = np.inf if i == len(start) - 1 else start[i + 1]
        idx = (times >= s0) & (times < s1)
        tt[idx] = offsets[i] + np.log(np.maximum(times[idx] - s0, t_sample / 2)) - base
        seg[idx] = i + 1
    return tt, seg


def chrono_vmm(times, step_times, vmm_eps=4.0, structure=None):
    """mat1d.construct_chrono_var_matrix, mat1d.py:455-490."""
    nt = len(times)
    if structure == 'uniform':
        return np.ones((nt, nt)) / nt
    tt, seg = time_transform(times, step_times)
    vmm = rbf(tt[:, None] - tt[None, :], vmm_eps) * (seg[:, None] == seg[None, :])
    return vmm / vmm.sum(axis=1)[:, None]


def dop_z_matrix(freq, nu, nu_eps):
    """phasance.construct_phasor_z_matrix (gaussian, normalize=False), phasance.py:19-37,61-80,108-118."""
    omega = 2 * np.pi * np.asarray(freq, dtype=float)
    nn, ww = np.meshgrid(np.asarray(nu, dtype=float), omega)
    a = np.minimum(0, np.sign(nn))
    b = np.maximum(0, np.sign(nn))

    def prim(lim):
        out = 0.5 * np.sqrt(np.pi) * (1j * ww) ** nn / nu_eps
        out = out * (1j * ww) ** (np.log(1j * ww) / (4 * nu_eps ** 2))
        return out * erf(nu_eps * (lim - nn) - np.log(1j * ww) / (2 * nu_eps))

    return prim(b) - prim(a)


def dop_v_matrix(times, nu, nu_eps, step_times, step_sizes):
    """phasance.construct_phasor_v_matrix, gaussian basis, ideal galvanostatic steps, normalize=False
    (phasance.py:8-9,40-57,83-99,121-144)."""
    from scipy.special import erf, gamma
    times, nu = np.asarray(times, dtype=float), np.asarray(nu, dtype=float)
    a, b = np.minimum(0, np.sign(nu)), np.maximum(0, np.sign(nu))
    out = np.zeros((times.size, nu.size))
    for st, sa in zip(step_times, step_sizes):
        sel = times > st
        t = (times[sel] - st)[:, None]
        with np.errstate(divide='ignore'):
            pre = 0.5 * np.sqrt(np.pi) * (t ** -nu[None, :] / gamma(-nu[None, :] + 1)) / nu_eps
        pre = pre * t ** (np.log(t) / (4 * nu_eps ** 2))

        def prim(lim):
            return pre * erf(nu_eps * (lim[None, :] - nu[None, :]) + np.log(t) / (2 * nu_eps))
        out[sel] += sa * (prim(b) - prim(a))
    return out


def dop_scale_vector(nu, tau, quantiles=(0.25, 0.75)):
    """phasance.phasor_scale_vector, phasance.py:165-184."""
    lt = np.log(tau)
    lo, hi = lt.min(), lt.max()
    q1 = np.exp(lo + quantiles[0] * (hi - lo))
    q3 = np.exp(lo + quantiles[1] * (hi - lo))
    nu = np.asarray(nu, dtype=float)
    return np.where(nu <= 0, q3 ** nu, q1 ** nu)


def basis_tau_for(freq=None, times=None, step_times=None, ppd=10, extend=1):
    """preprocessing.get_basis_tau (no supergrid), preprocessing.py:948-1013."""
    lo, hi = np.inf, -np.inf
    if freq is not None:
        lo = min(lo, 1 / (2 * np.pi * np.max(freq)))
        hi = max(hi, 1 / (2 * np.pi * np.min(freq)))
    if times is not None:
        td = time_since_step(times, step_times)
        lo, hi = min(lo, td.min()), max(hi, td.max())
    lmin, lmax = np.log10(lo) - extend, np.log10(hi) + extend
    exact = (lmax - lmin) * ppd + 1
    num = int(np.ceil(exact))
    add = 0.5 * (num - exact) / ppd
    return np.logspace(lmin - add, lmax + add, num)


def time_since_step(times, step_times, prestep_value=None):
    """preprocessing.get_time_since_step, preprocessing.py:918-945."""
    times = np.asarray(times, dtype=float)
    t_sample = np.min(np.diff(times)) if times.size > 1 else times[0]
    parts = []
    if prestep_value is not None:
        parts.append(np.full(int(np.sum(times < step_times[0])), float(prestep_value)))
    for i, st in enumerate(step_times):
        en = step_times[i + 1] if i + 1 < len(step_times) else np.inf
        sel = (times >= st) & (times < en)
        if sel.any():
            parts.append(np.maximum(times[sel] - st, t_sample))
    return np.concatenate(parts)


# ------------------------------------------------------------------------------------------------
# L2: QPHB fit loop
# ------------------------------------------------------------------------------------------------
DEFAULT_HYPERS = dict(                      # qphb.get_default_hypers(eff_hp=True), qphb.py:208-255
    derivative_weights=np.array([1.5, 1.0, 0.5]), sigma_ds=np.array([1.0, 1000.0, 1000.0]),
    l1_lambda_0=0.0, l2_lambda_0=142.0, iw_alpha=None, iw_beta=None,
    s_alpha=np.array([5.0, 10.0, 25.0]), s_0=np.ones(3),
    rho_alpha=np.array([0.15, 0.2, 0.25]), rho_0=np.ones(3),
    dop_l2_lambda_0=10.0, dop_l1_lambda_0=0.0, dop_derivative_weights=np.array([0.5, 1.0, 0.5]),
    dop_s_alpha=np.array([5.0, 10.0, 25.0]), dop_rho_alpha=np.array([0.15, 0.2, 0.25]),
    dop_s_0=np.ones(3), dop_rho_0=np.ones(3), dop_sigma_ds=np.array([1.0, 1000.0, 1000.0]),
)


def l2_matrix(pen, s_vectors, rho, dop_rho, hyp, n_special, dop_range, l2_lambda_0=None):
    """qphb.calculate_qp_l2_matrix ('integral'), qphb.py:53-120."""
    lam0 = hyp['l2_lambda_0'] if l2_lambda_0 is None else l2_lambda_0
    dop_lam0 = hyp['dop_l2_lambda_0'] * (lam0 / hyp['l2_lambda_0'])   # drt1d.py:643-645
    out = np.zeros_like(pen[0])
    for k, dw in enumerate(hyp['derivative_weights']):
        if dw <= 0:
            continue
        mk = pen[k].copy()
        mk[n_special:, n_special:] *= lam0 * dw * rho[k]
        if dop_range is not None:
            a, b = dop_range
            mk[a:b, a:b] *= dop_lam0 * hyp['dop_derivative_weights'][k] * dop_rho[k]
        r = np.sqrt(s_vectors[k])
        out += r[:, None] * mk * r[None, :]
    return out


def update_s(m, x, s_in, rho_eff, alpha, beta, g_mat, sigma):
    """qphb.solve_s ('integral' branch), qphb.py:320-338,354 and the floor at :780."""
    gamma = rho_eff * (x[:, None] * m * x[None, :]) + g_mat / (2 * sigma ** 2) + beta * np.eye(x.size)
    u = np.sqrt(s_in)
    gu = gamma * u[None, :]
    np.fill_diagonal(gu, 0)
    gd = np.diag(gamma)
    if np.max(np.abs(gu)) > 1e-10:
        b = gu.sum(axis=1)
        with np.errstate(invalid='ignore', divide='ignore'):
            u_hat = (-b + np.sign(b) * np.sqrt(b ** 2 + 4 * gd * (alpha - 1))) / (2 * gd)
        s_hat = u_hat ** 2
    else:
        s_hat = (alpha - 1) / gd
    s_hat[np.isnan(s_hat)] = 1
    s_hat[s_hat <= 0] = 1e-15
    return s_hat


def update_rho(m, x, s, alpha, beta, xmx_norm):
    """qphb.solve_rho, qphb.py:385-401."""
    r = np.sqrt(s) * x
    return alpha / ((r @ m @ r) / xmx_norm + beta)


def apply_vmm(vmm, r2):
    """vmm @ r^2 where vmm is dense or the structured (n_chrono uniform block + dense EIS block)."""
    if isinstance(vmm, dict):
        nc = vmm['n_chrono']
        out = np.empty_like(r2)
        if nc:
            out[:nc] = vmm['chrono'] @ r2[:nc] if vmm.get('chrono') is not None else np.mean(r2[:nc])
        if vmm.get('eis') is not None:
            out[nc:] = vmm['eis'] @ r2[nc:]
        return out
    return vmm @ r2


def vmm_diag(vmm, n):
    """diag(vmm) for the dense or structured form."""
    if isinstance(vmm, dict):
        nc = vmm['n_chrono']
        d = np.empty(n)
        if nc:
            d[:nc] = np.diag(vmm['chrono']) if vmm.get('chrono') is not None else 1.0 / nc
        if vmm.get('eis') is not None:
            d[nc:] = np.diag(vmm['eis'])
        return d
    return np.diag(vmm).copy()


def apply_vmm_base(vmm, u):
    """(vmm - diag(vmm)) / (1 - diag(vmm))[:, None] @ u: the self-excluding averaging matrix of the first
    initialisation pass with outliers, qphb.py:1644-1649."""
    d = vmm_diag(vmm, u.size)
    return (apply_vmm(vmm, u) - d * u) / (1 - d)


def outlier_t_vector(s_bar, resid, outlier_p):
    """qphb.solve_outlier_t, qphb.py:1497-1519 (pdf_normal: utils/stats.py:11-12)."""
    with np.errstate(invalid='ignore', divide='ignore'):
        sd = np.sqrt(s_bar)
        pdf_in = 1 / (sd * np.sqrt(2 * np.pi)) * np.exp(-0.5 * resid ** 2 / sd ** 2)
        ar = np.abs(resid)
        pdf_out = 1 / (ar * np.sqrt(2 * np.pi)) * np.exp(-0.5 * resid ** 2 / ar ** 2)
        t = 1 - outlier_p * pdf_out / ((1 - outlier_p) * pdf_in + outlier_p * pdf_out)
    t[sd > ar] = 1
    return t


def estimate_weights(x, y, vmm, rm, est_weights=None, outlier_p=None, base=False, return_t=False):
    """qphb.estimate_weights, qphb.py:1545-1594.  With outlier_p the averaging matrix becomes
    T^1/2 vmm T^1/2 + (I - T) (qphb.outlier_tvt :1522-1538), applied here without forming it;
    base=True uses the self-excluding matrix of the first initialisation pass (:1644-1649)."""
    resid = rm @ x - y
    r2 = resid ** 2
    mv = apply_vmm_base if base else apply_vmm
    if outlier_p is not None:
        t = outlier_t_vector(mv(vmm, r2), resid, outlier_p)
        sq = t ** 0.5
        s_hat = sq * mv(vmm, sq * r2) + (1 - t) * r2
    else:
        t = np.ones(len(y))
        s_hat = mv(vmm, r2)
    floor = np.var(y) * 1e-7
    s_hat = np.where(s_hat < floor, floor, s_hat)
    w = s_hat ** -0.5
    if est_weights is not None:
        frac = w / (w + est_weights)
        w = frac * w + (1 - frac) * est_weights
    w = np.maximum(w, 1e-10)
    return (w, t) if return_t else w


def converged(x_in, x_out, atol, rtol):
    """qphb.is_converged, qphb.py:597-603."""
    d = x_out - x_in
    with np.errstate(invalid='ignore', divide='ignore'):
        return bool(np.max(np.abs(d / (x_in + 1e-15))) <= rtol or np.max(np.abs(d)) <= atol)


def update_hypers(x, s_vec, rho, dop_rho, pen, hyp, ns, dop_range, xmx, dop_xmx):
    """The s / rho updates of one qphb.iterate_qphb call (qphb.py:722-933) for the DRT block and the DOP block."""
    kk = len(hyp['derivative_weights'])
    s_vec = [s.copy() for s in s_vec]
    rho = rho.copy()
    xd = x[ns:]
    for k, dw in enumerate(hyp['derivative_weights']):                       # qphb.py:722-803
        if dw <= 0:
            continue
        m = pen[k][ns:, ns:]
        alpha = hyp['s_alpha'][k]
        beta = (alpha - 1) / hyp['s_0'][k]
        if k == 0:
            xh = np.sign(xd) * np.abs(xd) ** 0.5
            g = xh[:, None] * pen[1][ns:, ns:] * xh[None, :]
        else:
            g = 0
        s_new = update_s(m, xd, s_vec[k][ns:], 1.0, alpha, beta, g, hyp['sigma_ds'][k])
        s_vec[k][ns:] = s_new
        ra = hyp['rho_alpha'][k]
        rho[k] = update_rho(m, xd, s_new, ra, ra / hyp['rho_0'][k], xmx[k])
    if dop_range is not None:                                                # qphb.py:822-933
        a, b = dop_range
        dop_rho = dop_rho.copy()
        xp = x[a:b]
        for k, dw in enumerate(hyp['dop_derivative_weights']):
            if dw <= 0:
                continue
            m = pen[k][a:b, a:b]
            alpha = hyp['dop_s_alpha'][k]
            beta = (alpha - 1) / hyp['dop_s_0'][k]
            s_new = update_s(m, xp, s_vec[k][a:b], 1.0, alpha, beta, 0, hyp['dop_sigma_ds'][k])
            s_vec[k][a:b] = s_new
            ra = hyp['dop_rho_alpha'][k]
            dop_rho[k] = update_rho(m, xp, s_new, ra, ra / hyp['dop_rho_0'][k], dop_xmx[k])
    return s_vec, rho, dop_rho


def qphb_fit(prob, hypers=None, record_history=False):
    """The loop of DRT._qphb_fit_core, drt1d.py:556-1008, on prepared matrices.

    ``prob`` keys: rm (N,n), rv (N,), vmm (dense or structured dict), pen (3,n,n), h (n,), l1 (n,),
    n_special; optional: dop_range, vz_index, vb_range, vz_strength (N,), n_chrono,
    chrono_weight_factor, eis_weight_factor, weight_factor, xtol, max_iter, iw_l1, iw_l2.
    """
    hyp = dict(DEFAULT_HYPERS)
    if hypers:
        hyp.update(hypers)
    rm = np.array(prob['rm'], dtype=float)
    rv = np.asarray(prob['rv'], dtype=float)
    vmm = prob['vmm']
    pen = [np.asarray(m, dtype=float) for m in prob['pen']]
    h = np.asarray(prob['h'], dtype=float)
    l1 = np.asarray(prob['l1'], dtype=float)
    ns = int(prob['n_special'])
    dop_range = prob.get('dop_range')
    n = rm.shape[1]
    xtol = prob.get('xtol', 1e-2)
    max_iter = prob.get('max_iter', 50)
    wf = prob.get('weight_factor', 1.0)
    nc = int(prob.get('n_chrono', 0))
    cwf = prob.get('chrono_weight_factor', 1.0)
    ewf = prob.get('eis_weight_factor', 1.0)
    vz_index = prob.get('vz_index')
    hybrid = vz_index is not None or (nc > 0 and nc < rm.shape[0])
    kk = len(hyp['derivative_weights'])

    rho = np.array(hyp['rho_0'], dtype=float)
    dop_rho = np.array(hyp['dop_rho_0'], dtype=float) if dop_range is not None else None
    s_vec = [np.ones(n) * hyp['s_0'][k] for k in range(kk)]
    x = np.zeros(n) + 1e-6                                                       # drt1d.py:612

    if vz_index is not None:                                                     # drt1d.py:503-519
        rm_vz = rm.copy()
        a, b = prob['vb_range']
        rm_vz[:, a:b] = 0
        vz_strength = np.asarray(prob['vz_strength'], dtype=float)

    ipm_log = []

    def qp(wrm, wrv, l2, l1v):
        res = coneqp_orthant(wrm.T @ wrm + l2, -wrm.T @ wrv + l1v, h)           # qphb.py:465-519
        ipm_log.append(res['iterations'])
        return res

    # solve_rp (drt1d.py:573-607, qphb.estimate_x_rp :1684-1717, DRT._solve_data_scale drt1d.py:5421-5437)
    rp_factor, us_factor, dop_cs = 1.0, 1.0, 1.0
    rp_scale, area = hyp.get('rp_scale', 14.0), prob.get('basis_area')
    if prob.get('solve_rp'):
        l2_rp = l2_matrix(pen, s_vec, rho, dop_rho, hyp, ns, dop_range, l2_lambda_0=1e-4)
        x_rp = qp(rm, rv, l2_rp, 1e-3)['x']
        rp_factor = rp_scale / (np.sum(np.abs(x_rp[ns:])) * area)
        rv = rv * rp_factor
        if dop_range is not None and prob.get('normalize_dop', True):
            a, b = dop_range
            dop_cs = 1.0 / (np.max(np.abs(x_rp[ns:])) / np.max(np.abs(x_rp[a:b])))
            rm[:, a:b] *= dop_cs
            if vz_index is not None:
                rm_vz[:, a:b] *= dop_cs

    # initialize_weights, qphb.py:1609-1681 with iw hypers of drt1d.py:640-645
    l2_iw = l2_matrix(pen, s_vec, rho, dop_rho, hyp, ns, dop_range, l2_lambda_0=prob.get('iw_l2', 1e-4))
    outlier_p = hyp.get('outlier_p')
    if outlier_p is not None:                                                    # qphb.py:1629-1655
        est_w = np.ones(rm.shape[0])
        for _ in range(2):
            res = qp(est_w[:, None] * rm, est_w * rv, l2_iw, prob.get('iw_l1', 1e-4))
            x_overfit = res['x']
            est_w, outlier_t = estimate_weights(x_overfit, rv, vmm, rm, outlier_p=outlier_p, base=True,
                                                return_t=True)
    elif prob.get('init_weights_separately') and 0 < nc < rm.shape[0]:         # drt1d.py:647-669
        def vmm_part(a, b):
            if isinstance(vmm, dict):
                return dict(n_chrono=(nc if a == 0 else 0), chrono=vmm.get('chrono') if a == 0 else None,
                            eis=None if a == 0 else vmm.get('eis'))
            return vmm[a:b, a:b]
        parts, x_parts = [], []
        for a, b in ((0, nc), (nc, rm.shape[0])):
            res = qp(rm[a:b], rv[a:b], l2_iw, prob.get('iw_l1', 1e-4))
            x_parts.append(res['x'])
            parts.append(estimate_weights(res['x'], rv[a:b], vmm_part(a, b), rm[a:b]))
        est_w = np.concatenate(parts)
        x_overfit, x_overfit_eis = x_parts
        outlier_t = np.ones(rm.shape[0])
    else:
        res = qp(rm, rv, l2_iw, prob.get('iw_l1', 1e-4))
        x_overfit = res['x']
        est_w, outlier_t = estimate_weights(x_overfit, rv, vmm, rm, return_t=True)
    init_outlier_t = outlier_t
    if hyp['iw_alpha'] is not None:                                              # qphb.py:1471-1479
        bq = 0.5 - hyp['iw_alpha'] + 1
        s_hat = (-bq + np.sqrt(bq ** 2 + 2 * hyp['iw_beta'] * est_w ** -2.0)) / (2 * hyp['iw_beta'])
        init_w = s_hat ** -0.5
    else:
        init_w = est_w.copy()
    w = init_w.copy()
    if hybrid and prob.get('hybrid_weight_factor_method') == 'weight':            # drt1d.py:749-759
        ratio = (np.mean(est_w[nc:] ** -2.0) ** -0.5 / np.mean(est_w[:nc] ** -2.0) ** -0.5) ** 0.25
        ewf, cwf = 1 / ratio, ratio

    xmx = np.ones(kk)
    dop_xmx = np.ones(kk)
    history = []
    it = 0
    conv = False
    while it < max_iter:
        x_in = x.copy()
        if hybrid:                                                               # drt1d.py:882-884
            w[:nc] *= cwf
            w[nc:] *= ewf
        if it > 0:
            w = w * wf
        if it > 1 and prob.get('update_scale'):                                  # drt1d.py:914-936
            sf = (rp_scale / (np.sum(np.abs(x[ns:])) * area)) ** 0.5
            x_in = x_in * sf
            x_overfit = x_overfit * sf
            rv = rv * sf
            xmx = xmx * sf ** 0.5
            dop_xmx = dop_xmx * sf ** 0.5
            est_w = est_w / sf
            init_w = init_w / sf
            w = w / sf
            us_factor *= sf
        wrm = w[:, None] * rm
        wrv = w * rv
        l2 = l2_matrix(pen, s_vec, rho, dop_rho, hyp, ns, dop_range)
        res = qp(wrm, wrv, l2, l1)
        x = res['x']

        s_vec, rho, dop_rho = update_hypers(x, s_vec, rho, dop_rho, pen, hyp, ns, dop_range, xmx, dop_xmx)
        w, outlier_t = estimate_weights(x, rv, vmm, rm, est_w, outlier_p=outlier_p, return_t=True)   # qphb.py:938
        conv = converged(x_in, x, np.mean(x_in) * 1e-3, xtol)                    # qphb.py:969-970
        if record_history:
            history.append(dict(x=x.copy(), s=np.array(s_vec), rho=rho.copy(), w=w.copy(),
                                fun=res['primal objective'], ipm=res['iterations']))
        if it == 0:                                                              # drt1d.py:946-962
            xd = x[ns:]
            xmx = np.array([xd @ pen[k][ns:, ns:] @ xd for k in range(kk)])
            if dop_range is not None:
                a, b = dop_range
                dop_xmx = np.array([x[a:b] @ pen[k][a:b, a:b] @ x[a:b] for k in range(kk)])
        if vz_index is not None:                                                 # drt1d.py:972-979
            sep = rm_vz @ x
            sep[nc:] *= -1
            rm[:, vz_index] = sep * vz_strength
        it += 1
        if conv:
            break

    w_true = w * wf                                                              # drt1d.py:991
    w_scaled = w_true.copy()
    if hybrid:
        w_scaled[:nc] *= cwf
        w_scaled[nc:] *= ewf
    l2 = l2_matrix(pen, s_vec, rho, dop_rho, hyp, ns, dop_range)                 # qphb.calculate_pq :1154
    wrm = w_scaled[:, None] * rm
    out = dict(
        x=x, fun=res['primal objective'], weights=w_scaled, true_weights=w_true,
        est_weights=est_w, init_weights=init_w, x_overfit=x_overfit,
        s_vectors=np.array(s_vec), rho=rho, dop_rho=dop_rho, xmx_norms=xmx, dop_xmx_norms=dop_xmx,
        n_outer=it, converged=conv, ipm_iters=np.array(ipm_log), outlier_t=outlier_t,
        chrono_weight_factor=cwf, eis_weight_factor=ewf, init_outlier_t=init_outlier_t, scale_factors=np.array([rp_factor, us_factor, dop_cs]),
        p_matrix=l2 + wrm.T @ wrm, q_vector=-wrm.T @ (w_scaled * rv) + l1, rm_final=rm,
    )
    if record_history:
        out['history'] = history
    return out


def qphb_continue(prob, state, hypers, max_iter=10, min_iter=2):
    """DRT._continue_from_init, drt1d.py:1270-1365: iterate_qphb warm-started from ``state`` (x, s_vectors, rho,
    dop_rho, weights, est_weights, xmx_norms, dop_xmx_norms, rm with the current vz_offset column) under modified
    hyper-parameters.  The weights are multiplied by the weight factors on every pass, xmx norms stay fixed, and
    convergence only ends the loop from pass ``min_iter`` on."""
    hyp = dict(DEFAULT_HYPERS)
    hyp.update(hypers)
    rm = np.array(state['rm'], dtype=float)
    rv = np.asarray(prob['rv'], dtype=float)
    vmm, h, l1 = prob['vmm'], np.asarray(prob['h'], dtype=float), np.asarray(prob['l1'], dtype=float)
    pen = [np.asarray(m, dtype=float) for m in prob['pen']]
    ns, dop_range = int(prob['n_special']), prob.get('dop_range')
    nc = int(prob.get('n_chrono', 0))
    vz_index = prob.get('vz_index')
    hybrid = vz_index is not None or (nc > 0 and nc < rm.shape[0])
    wf, cwf, ewf = prob.get('weight_factor', 1.0), prob.get('chrono_weight_factor', 1.0), prob.get('eis_weight_factor', 1.0)
    xtol = prob.get('xtol', 1e-2)
    x, w = state['x'].copy(), state['weights'].copy()
    s_vec, rho, dop_rho = [v.copy() for v in state['s_vectors']], state['rho'].copy(), state.get('dop_rho')
    if vz_index is not None:
        # the prediction matrix is copied once, *with* the vz_offset column of the incoming state
        # (drt1d.py:1296-1302; the plain fit copies it while that column is still zero, :503-511)
        a, b = prob['vb_range']
        rm_vz = rm.copy()
        rm_vz[:, a:b] = 0
        vz_strength = np.asarray(prob['vz_strength'], dtype=float)
    history = []
    it = 0
    while it < max_iter:
        x_in = x.copy()
        if hybrid:
            w[:nc] *= cwf
            w[nc:] *= ewf
        w = w * wf
        l2 = l2_matrix(pen, s_vec, rho, dop_rho, hyp, ns, dop_range)
        wrm = w[:, None] * rm
        res = coneqp_orthant(wrm.T @ wrm + l2, -wrm.T @ (w * rv) + l1, h)
        x = res['x']
        s_vec, rho, dop_rho = update_hypers(x, s_vec, rho, dop_rho, pen, hyp, ns, dop_range, state['xmx_norms'],
                                            state.get('dop_xmx_norms'))
        w = estimate_weights(x, rv, vmm, rm, state['est_weights'], outlier_p=hyp.get('outlier_p'))
        conv = converged(x_in, x, np.mean(x_in) * 1e-3, xtol)
        history.append(dict(x=x.copy(), s_vectors=np.array(s_vec), rho=rho.copy(), dop_rho=dop_rho, weights=w.copy(),
                            ipm=res['iterations']))
        if vz_index is not None:
            sep = rm_vz @ x
            sep[nc:] *= -1
            rm[:, vz_index] = sep * vz_strength
        if conv and it >= min_iter - 1:
            break
        it += 1
    return history, rm


def llh_terms(x, rm, rv, weights):
    """(weighted rss, sum log w) -- the two data terms of qphb.evaluate_llh, qphb.py:1347-1377."""
    r = weights * (rm @ x - rv)
    return float(r @ r), float(np.sum(np.log(weights)))


def marginal_llh(wrss, sum_log_w, n_data, alpha_0=2.0, beta_0=1.0):
    """qphb.evaluate_llh with marginalize_weights=True, qphb.py:1359-1370."""
    from scipy.special import loggamma
    alpha_n = alpha_0 - 1 + n_data / 2
    beta_n = beta_0 + 0.5 * wrss
    return alpha_0 * np.log(beta_0) - alpha_n * np.log(beta_n) + loggamma(alpha_n) - loggamma(alpha_0) + sum_log_w


def pfrt_fit(prob, factors=None, max_iter_per_step=10, max_init_iter=20, hypers=None):
    """DRT._pfrt_fit_core, drt1d.py:2558-2698: a fit at s_0 * f, l2_lambda_0 / f for the first factor, then one
    warm-started continuation per remaining factor; per step the final x, the marginal log-likelihood under
    weights re-estimated from x alone, and calculate_pq's P under those weights and the *initial* hypers."""
    base = dict(DEFAULT_HYPERS)
    if hypers:
        base.update(hypers)
    factors = np.logspace(-1, 1, 11) if factors is None else np.asarray(factors, dtype=float)

    def step_hypers(f):
        return dict(s_0=np.asarray(base['s_0'], dtype=float) * f, l2_lambda_0=base['l2_lambda_0'] / f)

    init_h = dict(base, **step_hypers(factors[0]))
    p0 = dict(prob, max_iter=max_init_iter)
    fit = qphb_fit(p0, init_h, record_history=True)
    rv = np.asarray(prob['rv'], dtype=float)
    ns, dop_range = int(prob['n_special']), prob.get('dop_range')
    pen = [np.asarray(m, dtype=float) for m in prob['pen']]
    steps = []

    def record(x, s_vec, rho, dop_rho, rm, n_iter):
        w = estimate_weights(x, rv, prob['vmm'], rm)
        wrss, slw = llh_terms(x, rm, rv, w)
        l2 = l2_matrix(pen, list(s_vec), rho, dop_rho, init_h, ns, dop_range)
        wrm = w[:, None] * rm
        steps.append(dict(x=x.copy(), llh=marginal_llh(wrss, slw, rv.size), p_matrix=l2 + wrm.T @ wrm, n_iter=n_iter,
                          wrss=wrss, sum_log_w=slw))

    last = fit['history'][-1]
    rm = fit['rm_final']
    record(last['x'], last['s'], last['rho'], fit['dop_rho'], rm, fit['n_outer'])
    state = dict(x=last['x'], s_vectors=list(last['s']), rho=last['rho'], dop_rho=fit['dop_rho'], weights=last['w'],
                 est_weights=fit['est_weights'], xmx_norms=fit['xmx_norms'], dop_xmx_norms=fit['dop_xmx_norms'], rm=rm)
    for f in factors[1:]:
        hist, rm = qphb_continue(prob, state, dict(init_h, **step_hypers(f)), max_iter=max_iter_per_step)
        last = hist[-1]
        record(last['x'], last['s_vectors'], last['rho'], last['dop_rho'], rm, len(hist))
        state.update(x=last['x'], s_vectors=list(last['s_vectors']), rho=last['rho'], dop_rho=last['dop_rho'],
                     weights=last['weights'], rm=rm)
    return dict(init=fit, steps=steps, factors=factors)


# ------------------------------------------------------------------------------------------------
# Preparation of an EIS problem (DRT defaults) -- enough for the benchmark's CPU baseline
# ------------------------------------------------------------------------------------------------
class EisPrep:
    """Spectrum-independent part of DRT.fit_eis on a shared frequency grid.

    Restates DRTBase.__init__ (drtbase.py:127-156), _prep_for_fit (drt1d.py:5439-5555),
    _prep_impedance_fit_matrix (:5625-5671), _prep_penalty_matrices (:5673-5734) and
    _format_qp_matrices (:5736-5963) for fit_ohmic = fit_inductance = True, gaussian basis.
    The reference caches A across fits on one DRT instance the same way (drt1d.py:5627,5652-5657).
    """

    def __init__(self, freq, mode='interp', tables=None, nonneg=True, ppd=10,
                 inductance_scale=1e-5, special_penalty=1e-6, vmm_eps=0.25, reim_cor=0.25,
                 error_structure=None, extend_basis_decades=1):
        self.freq = np.asarray(freq, dtype=float)
        self.eps = 1 / np.log(10 ** (1 / ppd))                                   # preprocessing.py:1016
        self.tables = tables if (tables is not None or mode != 'interp') else lookup_tables(self.eps)
        self.tau = basis_tau_for(self.freq, ppd=ppd, extend=extend_basis_decades)
        nb, nf = self.tau.size, self.freq.size
        a_re = impedance_matrix(self.freq, self.tau, self.eps, 'real', mode, self.tables)
        a_im = impedance_matrix(self.freq, self.tau, self.eps, 'imag', mode, self.tables)
        self.zm = a_re + 1j * a_im
        ns = 2
        rm = np.zeros((2 * nf, ns + nb))
        rm[:nf, 0] = 1.0                                                         # R_inf, drt1d.py:5837
        rm[nf:, 1] = 2 * np.pi * self.freq * inductance_scale                    # L, :5834 + mat1d.py:446
        rm[:nf, ns:] = a_re
        rm[nf:, ns:] = a_im
        pen = []
        for k in range(3):
            m = np.zeros((ns + nb, ns + nb))
            m[0, 0] = m[1, 1] = special_penalty                                  # drt1d.py:5886-5889
            m[ns:, ns:] = penalty_matrix(np.log(self.tau), k, self.eps)
            pen.append(m)
        self.inductance_scale = inductance_scale
        self.n_special = ns
        self.rm = rm
        self.pen = pen
        self.vmm = eis_vmm(self.freq, vmm_eps, reim_cor, error_structure)
        self.h = np.zeros(ns + nb) if nonneg else np.concatenate([np.zeros(ns), 1e5 * np.ones(nb)])
        self.l1 = np.zeros(ns + nb)

    def problem(self, z, rp_scale=14.0):
        z = np.asarray(z)
        scale = (np.max(z.real) - np.min(z.real)) / rp_scale                     # preprocessing.py:828-841
        zs = z / scale
        prob = dict(rm=self.rm, rv=np.concatenate([zs.real, zs.imag]), vmm=self.vmm, pen=self.pen,
                    h=self.h, l1=self.l1, n_special=self.n_special)
        return prob, scale

    def fit(self, z, hypers=None, **kw):
        prob, scale = self.problem(z)
        prob.update(kw)
        res = qphb_fit(prob, hypers)
        x = res['x']
        res['coefficient_scale'] = scale
        res['params'] = dict(                                                    # drt1d.py:6228-6289
            x=x[2:] * scale, R_inf=x[0] * scale, inductance=x[1] * scale * self.inductance_scale)
        sig = (1 / res['true_weights']) * scale                                  # drt1d.py:1083-1088
        nf = self.freq.size
        res['params']['z_sigma_tot'] = sig[:nf] + 1j * sig[nf:]
        return res

    def predict_z(self, params):
        """DRT.predict_z at the fit frequencies, drt1d.py:3500-3542."""
        return self.zm @ params['x'] + params['R_inf'] + params['inductance'] * 2j * np.pi * self.freq
