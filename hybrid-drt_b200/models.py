"""Drop-in mirror of hybrid-drt's ``DRT`` model for the QPHB fit path, dispatching to the B200 engine.

Mirrors ``hybdrt.models.DRT`` (reference hybdrt/models/drt1d.py:38, drtbase.py:20): same constructor
keywords, same ``fit_eis / fit_chrono / fit_hybrid`` signatures (drt1d.py:1197-1268), same result
attributes (``fit_parameters``, ``qphb_params``, ``cvx_result['x']``, ``special_qp_params``,
``basis_tau``, ``coefficient_scale`` ...).  New: ``fit_eis_batch / fit_hybrid_batch / fit_chrono_batch``
fit whole batches of spectra that share a measurement grid in one kernel launch.

Host code here only lays out parameters, scales data (O(N) numpy per spectrum) and unpacks results; the
response matrices and the whole fit loop run in the CUDA library (engine.py).  Options of the reference
that lie outside the hot-path scope raise NotImplementedError -- nothing silently falls back to CPU.
"""
import warnings

import numpy as np
import torch

from . import engine as _engine
from . import kk as _kk


def _not_supported(what):
    raise NotImplementedError(f'{what} is outside the B200 hot-path scope of hybdrt_b200 (see DESIGN.md)')


# ---- small host helpers (restate reference utilities; O(N) scalar work) ---------------------------
def is_uniform(x):
    """utils/array.py:142-151"""
    d = np.diff(x)
    return bool(np.std(d) / np.mean(d) <= 0.01)


def unit_step(t, ts=0.0):
    """utils/array.py:172-181"""
    return (np.asarray(t) >= ts).astype(float)


def nearest_index(arr, val, constraint=None):
    """utils/array.py:207-242"""
    arr = np.asarray(arr, dtype=float)
    if constraint is None:
        return int(np.argmin(np.abs(arr - val)))
    obj = np.full(arr.shape, np.inf)
    ok = constraint * arr >= constraint * val
    obj[ok] = constraint * (arr - val)[ok]
    idx = int(np.argmin(obj))
    if obj[idx] == np.inf:
        raise ValueError(f'No index satisfying {constraint} constraint for value {val}')
    return idx


def identify_steps(y, allow_consecutive=True, rthresh=50, athresh=1e-10):
    """preprocessing.identify_steps, preprocessing.py:17-38"""
    dy = np.diff(y)
    idx = np.where((np.abs(dy) >= np.median(np.abs(dy)) * rthresh) & (np.abs(dy) >= athresh))[0] + 1
    if not allow_consecutive:
        gap = np.concatenate(([2], np.diff(idx)))
        idx = idx[gap > 1]
    return idx


def step_indices_from_times(times, step_times):
    """preprocessing.get_step_indices_from_step_times, preprocessing.py:161-178"""
    out = []
    for st in step_times:
        d = np.where(times >= st, times - st, np.inf)
        out.append(int(np.argmin(d)))
    return np.array(out, dtype=int)


def step_sizes_from_signal(times, y, step_times, step_index=None):
    """preprocessing.get_step_sizes, preprocessing.py:105-129"""
    if step_index is None:
        step_index = step_indices_from_times(times, step_times)
    n = len(step_times)
    out = np.zeros(n)
    for k in range(n):
        end = len(y) if k == n - 1 else step_index[k + 1]
        prev = 0 if k == 0 else step_index[k - 1]
        out[k] = np.mean(y[step_index[k]:end]) - np.mean(y[prev:step_index[k]])
    return out


def step_info(times, y, offset_step_times=True, offset_size=None, rthresh=50):
    """preprocessing.get_step_info (ideal step model), preprocessing.py:57-102"""
    idx = identify_steps(y, True, rthresh)
    st = times[idx].copy()
    if offset_step_times:
        if offset_size is None:
            offset_size = -np.min(np.diff(times)) * (1 - 1e-8)
        st = st + offset_size
    return st, step_sizes_from_signal(times, y, st, step_index=idx)


def time_since_step(times, step_times, prestep_value=None):
    """preprocessing.get_time_since_step, preprocessing.py:918-945"""
    t_sample = np.min(np.diff(times)) if len(times) > 1 else times[0]
    parts = []
    if prestep_value is not None:
        parts.append(np.full(int(np.sum(times < step_times[0])), float(prestep_value)))
    for i, st in enumerate(step_times):
        en = step_times[i + 1] if i + 1 < len(step_times) else np.inf
        sel = (times >= st) & (times < en)
        if sel.any():
            parts.append(np.maximum(times[sel] - st, t_sample))
    return np.concatenate(parts)


def get_basis_tau(frequencies, times, step_times, ppd=10, extend_decades=1, tau_grid=None):
    """preprocessing.get_tau_lim + get_basis_tau, preprocessing.py:948-1013"""
    lo, hi = np.inf, -np.inf
    if frequencies is not None:
        lo, hi = 1 / (2 * np.pi * np.max(frequencies)), 1 / (2 * np.pi * np.min(frequencies))
    if times is not None:
        td = time_since_step(times, step_times)
        lo, hi = min(lo, np.min(td)), max(hi, np.max(td))
    lmin, lmax = np.log10(lo) - extend_decades, np.log10(hi) + extend_decades
    if tau_grid is not None:
        tau_grid = np.asarray(tau_grid)
        left = 0 if 10 ** lmin < np.min(tau_grid) else nearest_index(tau_grid, 10 ** lmin, -1)
        right = len(tau_grid) if 10 ** lmax > np.max(tau_grid) else nearest_index(tau_grid, 10 ** lmax, 1) + 1
        return tau_grid[left:right]
    exact = (lmax - lmin) * ppd + 1
    num = int(np.ceil(exact))
    add = 0.5 * (num - exact) / ppd
    return np.logspace(lmin - add, lmax + add, num)


def estimate_rp_batch(times, step_times, step_sizes, response, z):
    """preprocessing.estimate_rp (ideal steps), preprocessing.py:764-841, vectorised over the batch.

    response [B, Nt] or None; z [B, Nf] complex or None.  Returns rp_est [B].
    """
    r_min = np.full(1, np.inf)
    r_max = np.zeros(1)
    if times is not None:
        step_times = np.asarray(step_times, dtype=float)
        step_sizes = np.asarray(step_sizes, dtype=float)
        new_idx = np.concatenate(([0], np.where(np.diff(step_times) > 2e-5)[0] + 1))
        if len(new_idx) < len(step_times):
            bounds = list(new_idx) + [len(step_sizes)]
            step_sizes = np.array([np.sum(step_sizes[a:b]) for a, b in zip(bounds[:-1], bounds[1:])])
            step_times = step_times[new_idx]
        sidx = step_indices_from_times(times, step_times)
        mins, maxs = [], []
        for i, start in enumerate(sidx):
            end = len(times) if i == len(sidx) - 1 else sidx[i + 1]
            if start == end:
                mins.append(np.full(response.shape[0], np.nan))
                maxs.append(np.full(response.shape[0], np.nan))
                continue
            prev = response[:, start - 1:start] if start > 0 else 0.0        # a step at the first sample: no pre-step value
            r = (response[:, start:end] - prev) / step_sizes[i]
            mins.append(np.min(r, axis=1))
            maxs.append(np.max(r, axis=1))
        r_min = np.nanmean(np.array(mins), axis=0)
        r_max = np.nanpercentile(np.array(maxs), 99, axis=0)
    if z is not None:
        r_min = np.minimum(r_min, np.min(z.real, axis=1))
        r_max = np.maximum(r_max, np.max(z.real, axis=1))
    return r_max - r_min


_HYPER_KEYS = ('rp_scale', 'derivative_weights', 'sigma_ds', 'l1_lambda_0', 'l2_lambda_0', 'iw_alpha', 'iw_beta',
               's_alpha', 's_0', 'rho_alpha', 'rho_0', 'outlier_p')
_DOP_HYPER_KEYS = ('dop_l2_lambda_0', 'dop_l1_lambda_0', 'dop_derivative_weights', 'dop_s_alpha', 'dop_rho_alpha',
                   'dop_s_0', 'dop_rho_0', 'dop_sigma_ds')


def get_default_hypers(eff_hp=True, fit_dop=False):
    """qphb.get_default_hypers, qphb.py:208-255"""
    if not eff_hp:
        s_alpha, rho_alpha = np.array([1.05, 1.15, 2.5]), np.array([0.05, 0.1, 0.05])
    else:
        s_alpha, rho_alpha = np.array([5.0, 10.0, 25.0]), np.array([0.15, 0.2, 0.25])
    hyp = dict(rp_scale=14, derivative_weights=np.array([1.5, 1.0, 0.5]), sigma_ds=np.array([1.0, 1000.0, 1000.0]),
               l1_lambda_0=0, l2_lambda_0=142, iw_alpha=None, iw_beta=None, s_alpha=s_alpha, s_0=np.ones(3),
               rho_alpha=rho_alpha, rho_0=np.ones(3), outlier_p=None)
    if fit_dop:
        hyp.update(dop_l2_lambda_0=10, dop_l1_lambda_0=0, dop_derivative_weights=np.array([0.5, 1.0, 0.5]),
                   dop_s_alpha=np.array([5.0, 10.0, 25.0]), dop_rho_alpha=np.array([0.15, 0.2, 0.25]),
                   dop_s_0=np.ones(3), dop_rho_0=np.ones(3), dop_sigma_ds=np.array([1.0, 1000.0, 1000.0]))
    return hyp


class BatchFit:
    """Results of one batched fit: device tensors + the plan that produced them + host-side unpacking."""

    def __init__(self, plan, raw, scales, extra):
        self.plan = plan
        self.raw = raw              # dict of device tensors from Engine.qphb_fit_batch
        self.scales = scales        # dict of per-spectrum host arrays (coefficient_scale, ...)
        self.extra = extra
        self._host = None

    def host(self, keys=None):
        """Device -> host copy of the named outputs (all of them by default); cached per key."""
        if self._host is None:
            self._host = {}
        want = list(self.raw.keys()) if keys is None else keys
        missing = [k for k in want if k not in self._host and k in self.raw and isinstance(self.raw[k], torch.Tensor)]
        if missing:
            # through page-locked staging tensors: a pageable .cpu() of the big per-spectrum arrays runs at ~2 GB/s
            # (the staging buffers are the engine's cached ones -- page-locking memory costs about as much as the copy --
            # so the arrays handed out are copies of them, not views)
            eng = self.plan['model'].engine
            stream = torch.cuda.current_stream(self.raw[missing[0]].device)
            for k in missing:
                t = self.raw[k]
                if t.numel() * t.element_size() >= (1 << 20):
                    buf = eng.pinned(*t.shape, dtype=t.dtype)
                    buf.copy_(t, non_blocking=True)
                    stream.synchronize()
                    self._host[k] = buf.numpy().copy()
                else:
                    stream.synchronize()
                    self._host[k] = t.cpu().numpy()
        return self._host

    @property
    def x_raw(self):
        return self.host()['x']

    def pfrt_result(self):
        """The reference's DRT.pfrt_result (drt1d.py:2687-2694), batched: factors [F], step_x [B,F,n] (raw, scaled
        space, special parameters included), step_llh [B,F] (qphb.evaluate_llh with marginalised weights,
        qphb.py:1359-1373, alpha_0=2, beta_0=1), step_p_mat [B,F,n,n] if it was requested, step_iters [B,F]."""
        from math import lgamma
        h = self.host(['pfrt_x', 'pfrt_llh', 'pfrt_iters', 'pfrt_factors'] + (['pfrt_p'] if 'pfrt_p' in self.raw else []))
        n_data = self.plan['n_rows']
        alpha_0, beta_0 = 2.0, 1.0
        alpha_n = alpha_0 - 1 + n_data / 2
        beta_n = beta_0 + 0.5 * h['pfrt_llh'][..., 0]
        llh = alpha_0 * np.log(beta_0) - alpha_n * np.log(beta_n) + lgamma(alpha_n) - lgamma(alpha_0) \
            + h['pfrt_llh'][..., 1]
        return dict(factors=h['pfrt_factors'], step_x=h['pfrt_x'], step_llh=llh, step_p_mat=h.get('pfrt_p'),
                    step_iters=h['pfrt_iters'])

    def extract_parameters(self, x):
        """DRT.extract_qphb_parameters (drt1d.py:6228-6289) for raw coefficient arrays x [B, n] or [B, F, n]
        (scaled space, special parameters first): dict of unscaled arrays with the same leading shape."""
        pl, sc = self.plan, self.scales
        x = np.asarray(x)
        lead = x.shape[:-1]

        def per_b(a):                    # [B] -> broadcastable against x[..., i]
            return np.asarray(a).reshape((len(a),) + (1,) * (len(lead) - 1))
        cs = per_b(sc['coefficient_scale'])
        ns = pl['n_special']
        out = {'x': x[..., ns:] * cs[..., None]}
        sp = pl['special_qp_params']
        out['R_inf'] = x[..., sp['R_inf']['index']] * cs if 'R_inf' in sp else np.zeros(lead)
        out['inductance'] = (x[..., sp['inductance']['index']] * cs * pl['inductance_scale']
                             if 'inductance' in sp else np.zeros(lead))
        out['C_inv'] = (x[..., sp['C_inv']['index']] * cs * pl['capacitance_scale']
                        if 'C_inv' in sp else np.zeros(lead))
        if 'v_baseline' in sp:
            a = sp['v_baseline']['index']
            b = a + sp['v_baseline']['size']
            vb = x[..., a:b] * (1.0 / pl['v_baseline_scale'])
            vb[..., 0] -= per_b(sc['scaled_response_offset'])
            out['v_baseline'] = vb * per_b(sc['response_signal_scale'])[..., None]
        if 'vz_offset' in sp:
            out['vz_offset'] = x[..., sp['vz_offset']['index']]
        if 'x_dop' in sp:
            a = sp['x_dop']['index']
            b = a + sp['x_dop']['size']
            dsv = pl['dop_scale_vector'] * np.ones(lead + (1,))
            if 'dop_column_scale' in sc:
                dsv = dsv * per_b(sc['dop_column_scale'])[..., None]
            out['x_dop'] = x[..., a:b] * (dsv * cs[..., None])
        return out

    def fit_parameters(self):
        """DRT.extract_qphb_parameters (drt1d.py:6228-6289) for the whole batch: dict of [B, ...] arrays."""
        pl, h, sc = self.plan, self.host(['x', 'weights']), self.scales
        out = self.extract_parameters(h['x'])
        cs = sc['coefficient_scale'][:, None]
        # sigma from the unscaled weights (drt1d.py:1082-1092)
        w_true = h['weights'] * pl['weight_factor']
        sig = 1.0 / w_true
        nc, nf = pl['n_chrono'], pl['n_freq']
        out['v_sigma_tot'] = sig[:, :nc] * sc['response_signal_scale'][:, None] if nc else None
        out['z_sigma_tot'] = (sig[:, nc:nc + nf] + 1j * sig[:, nc + nf:]) * cs if nf else None
        return out

    def predict_drt(self, tau=None, ppd=20, order=0):
        """DRT.predict_drt (drt1d.py:3040-3061) for the whole batch: gamma(tau) [B, len(tau)]."""
        drt = self.plan['model']
        tau = drt.get_tau_eval(ppd) if tau is None else np.asarray(tau, dtype=float)
        return self.fit_parameters()['x'] @ drt._basis_eval_matrix(tau, order, self.plan['basis_tau']).T

    # ---- post-fit diagnostics of the mapping path (drtmd.py:256-279); need diag_tau at fit time
    def _uniform_weights(self):
        """weights='uniform': the mean estimated weight of each domain (drt1d.py:4433-4447)."""
        if getattr(self, '_uw', None) is None:       # evaluate_rss and evaluate_llh both ask for it
            ew = self.host(['est_weights'])['est_weights']
            nc = self.plan['n_chrono']
            wc = ew[:, :nc].mean(axis=1) if nc else np.zeros(len(ew))
            we = ew[:, nc:].mean(axis=1) if ew.shape[1] > nc else np.zeros(len(ew))
            self._uw = (wc, we)
        return self._uw

    def evaluate_rss(self, normalize=True):
        """DRT.evaluate_rss(weights='uniform') (drt1d.py:4433-4459, qphb.evaluate_rss :1347-1352)."""
        ss = self.host(['resid_ss'])['resid_ss']
        wc, we = self._uniform_weights()
        rss = wc ** 2 * ss[:, 0] + we ** 2 * ss[:, 1]
        return rss / self.plan['n_rows'] if normalize else rss

    def evaluate_llh(self, normalize=True, alpha_0=2, beta_0=1):
        """DRT.evaluate_llh(weights='uniform', marginalize_weights=True) (drt1d.py:4461-4496, qphb.py:1355-1377)."""
        from scipy.special import loggamma
        n_rows, nc = self.plan['n_rows'], self.plan['n_chrono']
        wc, we = self._uniform_weights()
        rss = self.evaluate_rss(normalize=False)
        alpha_n = alpha_0 - 1 + n_rows / 2
        beta_n = beta_0 + 0.5 * rss
        llh = alpha_0 * np.log(beta_0) - alpha_n * np.log(beta_n) + loggamma(alpha_n) - loggamma(alpha_0)
        with np.errstate(divide='ignore', invalid='ignore'):
            llh = llh + (nc * np.log(wc) if nc else 0.0) + ((n_rows - nc) * np.log(we) if n_rows > nc else 0.0)
        return llh / n_rows if normalize else llh

    def distribution_var(self, extend_var=True):
        """diag of DRT.estimate_distribution_cov(tau=diag_tau) (drt1d.py:3063-3151) for the whole batch; NaN rows
        where the final P was not positive definite (reference: None + 'Singular P matrix' warning)."""
        pl = self.plan
        h = self.host(['dist_var', 'status'])
        var = h['dist_var'] * self.scales['coefficient_scale'][:, None] ** 2
        var[(h['status'] & _engine.ST_COV_FAIL) != 0] = np.nan
        if extend_var:
            tau = pl['diag_tau']
            lo, hi = np.inf, -np.inf
            if pl['frequencies'] is not None:
                lo, hi = 1 / (2 * np.pi * np.max(pl['frequencies'])), 1 / (2 * np.pi * np.min(pl['frequencies']))
            if pl['times'] is not None:
                td = time_since_step(pl['times'], pl['step_times'])
                lo, hi = min(lo, np.min(td)), max(hi, np.max(td))
            li, ri = nearest_index(tau, lo) + 1, nearest_index(tau, hi)
            var[:, :li] = np.maximum(var[:, :li], var[:, li:li + 1])
            var[:, ri:] = np.maximum(var[:, ri:], var[:, ri:ri + 1])
        return var

    def predict_z(self, frequencies=None):
        """DRT.predict_z at the fit frequencies (drt1d.py:3500-3542) for the whole batch."""
        pl = self.plan
        fp = self.fit_parameters()
        if pl.get('multi'):
            if frequencies is not None:
                _not_supported('predict_z at new frequencies for a fit with per-spectrum frequency grids')
            dev = pl['model'].engine.dev
            xd = dev(fp['x'])
            z = torch.complex(torch.einsum('bfm,bm->bf', pl['a_re'], xd), torch.einsum('bfm,bm->bf', pl['a_im'], xd))
            f = pl['frequencies']
            return (z.cpu().numpy() + fp['R_inf'][:, None] + fp['inductance'][:, None] * 2j * np.pi * f
                    + fp['C_inv'][:, None] * (2j * np.pi * f) ** -1)
        same = frequencies is None or (pl['frequencies'] is not None and np.array_equal(frequencies, pl['frequencies']))
        if same:
            f, zm, zd = pl['frequencies'], pl['zm_drt_host'], pl.get('zm_dop_host')
            es = pl.get('eis_vz_strength')
        else:                           # matrices of the requested grid are built on the GPU
            drt = pl['model']
            f = np.asarray(frequencies, dtype=float)
            a_re, a_im = drt.engine.build_impedance(f[None], pl['basis_tau'][None], drt.tau_epsilon, drt._mode(),
                                                    drt.interpolate_lookups)
            zm = (a_re[0] + 1j * a_im[0]).cpu().numpy()
            zd = (drt.engine.build_dop_z(f[None], drt.basis_nu, drt.nu_epsilon)[0].cpu().numpy()
                  if 'x_dop' in fp else None)
            es = drt._vz_strength(None, f)[1] if 'vz_offset' in fp else None
        z = fp['x'] @ zm.T + fp['R_inf'][:, None] + fp['inductance'][:, None] * 2j * np.pi * f[None, :]
        z = z + fp['C_inv'][:, None] * (2j * np.pi * f[None, :]) ** -1
        if 'x_dop' in fp:
            z = z + fp['x_dop'] @ zd.T
        if 'vz_offset' in fp:
            z = z * (1 - fp['vz_offset'][:, None] * es[None, :])
        return z


class DRT:
    """Mirror of hybdrt.models.DRT for the QPHB fit path (see module docstring)."""

    def __init__(self, fixed_basis_tau=None, tau_supergrid=None, tau_basis_type='gaussian', tau_epsilon=None,
                 basis_tau_ppd=10, extend_basis_decades=1,
                 step_model='ideal', chrono_mode='galv', interpolate_integrals=True, chrono_tau_rise=None,
                 fixed_basis_nu=None, nu_basis_type='gaussian', nu_epsilon=None, fit_dop=False, normalize_dop=True,
                 fit_inductance=True, fit_ohmic=True, fit_capacitance=False,
                 time_precision=10, input_signal_precision=10, frequency_precision=10,
                 print_diagnostics=False, warn=True, device=0):
        if tau_basis_type != 'gaussian':
            _not_supported(f"tau_basis_type '{tau_basis_type}'")
        if nu_basis_type != 'gaussian':
            _not_supported(f"nu_basis_type '{nu_basis_type}'")
        if step_model != 'ideal':
            _not_supported(f"step_model '{step_model}'")
        if chrono_mode != 'galv':
            _not_supported(f"chrono_mode '{chrono_mode}'")
        if fixed_basis_tau is not None and tau_supergrid is not None:
            warnings.warn('If fixed_basis_tau is provided, tau_supergrid will be ignored')
        self.engine = _engine.get_engine(device)
        self.fixed_basis_tau = None if fixed_basis_tau is None else np.asarray(fixed_basis_tau, dtype=float)
        self.tau_supergrid = None if tau_supergrid is None else np.asarray(tau_supergrid, dtype=float)
        self.tau_basis_type, self.nu_basis_type = tau_basis_type, nu_basis_type
        self.tau_epsilon = tau_epsilon
        self.extend_basis_decades = extend_basis_decades
        self.basis_tau_ppd = basis_tau_ppd
        self.step_model, self.chrono_mode = step_model, chrono_mode
        self.fixed_basis_nu = fixed_basis_nu
        self.nu_epsilon = nu_epsilon
        self.fit_dop, self.normalize_dop = fit_dop, normalize_dop
        self.fit_inductance, self.fit_ohmic, self.fit_capacitance = fit_inductance, fit_ohmic, fit_capacitance
        self.time_precision, self.input_signal_precision = time_precision, input_signal_precision
        self.frequency_precision = frequency_precision
        self.print_diagnostics, self.warn = print_diagnostics, warn
        self.basis_tau = self.basis_nu = None
        self.dop_scale_vector = None
        self.special_qp_params = {}
        self.fit_parameters = self.qphb_params = self.qphb_history = self.cvx_result = None
        self.fit_type = self.fit_kwargs = None
        self.coefficient_scale = self.impedance_scale = 1.0
        self.input_signal_scale = self.response_signal_scale = 1.0
        self.inductance_scale = self.capacitance_scale = None
        self.step_times = self.step_sizes = self.nonconsec_step_times = None
        self.t_fit, self.f_fit = [], []
        self.last_batch = None
        # drtbase.py:127-135
        if self.tau_epsilon is None:
            if self.fixed_basis_tau is not None:
                self.tau_epsilon = 1 / np.mean(np.diff(np.log(self.fixed_basis_tau)))
            elif self.tau_supergrid is not None:
                self.tau_epsilon = 1 / np.mean(np.diff(np.log(self.tau_supergrid)))
            elif basis_tau_ppd is not None:
                self.tau_epsilon = 1 / np.log(10 ** (1 / basis_tau_ppd))
        # drtbase.py:137-159: lookup tables (device resident) or direct quadrature
        if interpolate_integrals:
            self.integrate_method = 'interp'
            self.interpolate_lookups = self.engine.build_lookup(self.tau_epsilon)
        else:
            self.integrate_method = 'trapz'
            self.interpolate_lookups = None

    # ------------------------------------------------------------------------------------------------
    # reference-compatible bookkeeping
    # ------------------------------------------------------------------------------------------------
    def _add_special_qp_param(self, sp, name, nonneg, size=1):
        sp[name] = {'index': int(sum(v.get('size', 1) for v in sp.values())), 'nonneg': nonneg, 'size': size}

    def get_qp_mat_offset(self):
        return int(sum(v.get('size', 1) for v in self.special_qp_params.values()))

    def get_special_indices(self, key):
        a = self.special_qp_params[key]['index']
        return a, a + self.special_qp_params[key].get('size', 1)

    @property
    def dop_indices(self):
        if 'x_dop' in self.special_qp_params:
            return self.get_special_indices('x_dop')
        return None, None

    @property
    def nu_basis_area(self):
        return np.sqrt(np.pi) / self.nu_epsilon

    @property
    def tau_basis_area(self):
        return np.sqrt(np.pi) / self.tau_epsilon

    def get_fit_times(self):
        return self.t_fit

    def get_fit_frequencies(self):
        return self.f_fit

    # ------------------------------------------------------------------------------------------------
    # plan: everything that is shared by the spectra of one batch (one measurement grid)
    # ------------------------------------------------------------------------------------------------
    def _mode(self):
        return _engine.MODE_INTERP if self.integrate_method == 'interp' else _engine.MODE_TRAPZ

    def _build_plan(self, times, i_signal, frequencies, kw):
        eng = self.engine
        dev = eng.dev
        hyp = kw['hypers']
        data_type = 'eis' if times is None else ('chrono' if frequencies is None else 'hybrid')

        # special parameter layout, drt1d.py:374-410
        sp = {}
        if times is not None:
            self._add_special_qp_param(sp, 'v_baseline', False, kw['v_baseline_deg'] + 1 + int(kw['v_baseline_sqrt']))
        if kw['vz_offset'] and data_type == 'hybrid':
            self._add_special_qp_param(sp, 'vz_offset', False)
        if self.fit_ohmic:
            self._add_special_qp_param(sp, 'R_inf', True)
        if self.fit_inductance:
            self._add_special_qp_param(sp, 'inductance', True)
        if self.fit_capacitance:
            self._add_special_qp_param(sp, 'C_inv', True)
        if self.fit_dop:
            if self.fixed_basis_nu is None:
                self.basis_nu = np.concatenate([np.linspace(-1, -0.4, 25), np.linspace(0.4, 1, 25)])
            else:
                self.basis_nu = np.asarray(self.fixed_basis_nu, dtype=float)
            if self.nu_epsilon is None:
                self.nu_epsilon = 1 / np.median(np.diff(np.sort(self.basis_nu)))
            self._add_special_qp_param(sp, 'x_dop', True, size=len(self.basis_nu))
        else:
            self.basis_nu = None
        self.special_qp_params = sp
        ns = self.get_qp_mat_offset()

        # chrono signal processing, drtbase.py:285-373 (no downsampling)
        if times is not None:
            times = np.asarray(times, dtype=float)
            i_signal = np.asarray(i_signal, dtype=float)
            step_times, step_sizes = kw['step_times'], kw['step_sizes']
            if step_times is None:
                step_times, step_sizes = step_info(times, i_signal, kw['offset_steps'], kw['step_offset_size'])
            else:
                step_times = np.asarray(step_times, dtype=float)
                if step_sizes is None:
                    step_sizes = step_sizes_from_signal(times, i_signal, step_times)
            if len(step_times) == 0:
                raise ValueError('no steps found in the input signal')
            if len(step_times) > 1:
                t_sample = np.min(np.diff(times))
                nonconsec = step_times[1:][np.diff(step_times) > 1.1 * t_sample]
                self.nonconsec_step_times = np.insert(nonconsec, 0, step_times[0])
            else:
                self.nonconsec_step_times = step_times
            self.step_times, self.step_sizes = step_times.copy(), np.asarray(step_sizes, dtype=float).copy()
            self.t_fit = times
            self.raw_input_signal = i_signal.copy()
        else:
            step_times = step_sizes = None
            self.t_fit = []
        if frequencies is not None:
            frequencies = np.asarray(frequencies, dtype=float)
            self.f_fit = frequencies
        else:
            self.f_fit = []

        # basis grid, drt1d.py:5469-5485
        if self.fixed_basis_tau is not None:
            self.basis_tau = self.fixed_basis_tau
        else:
            self.basis_tau = get_basis_tau(frequencies, times, step_times, tau_grid=self.tau_supergrid,
                                           extend_decades=self.extend_basis_decades)
        if self.tau_epsilon is None:
            self.tau_epsilon = 1 / np.mean(np.diff(np.log(self.basis_tau)))
        tau, eps = self.basis_tau, self.tau_epsilon
        nb = len(tau)
        n = ns + nb
        nc = 0 if times is None else len(times)
        nf = 0 if frequencies is None else len(frequencies)
        n_rows = nc + 2 * nf
        mode = self._mode()

        rm = torch.zeros(n_rows, n, dtype=torch.float64, device=eng.device)
        plan = dict(data_type=data_type, special_qp_params=sp, n_special=ns, n=n, n_rows=n_rows, n_chrono=nc,
                    n_freq=nf, frequencies=frequencies, times=times, basis_tau=tau,
                    inductance_scale=kw['inductance_scale'], capacitance_scale=kw['capacitance_scale'],
                    weight_factor=kw['weight_factor'], hypers=hyp, step_times=step_times)
        self.inductance_scale, self.capacitance_scale = kw['inductance_scale'], kw['capacitance_scale']

        # DOP scale vector, drt1d.py:5767-5788
        if self.fit_dop:
            if self.normalize_dop:
                ev = self.tau_supergrid if self.tau_supergrid is not None else tau
                lt = np.log(ev)
                q1 = np.exp(lt.min() + 0.25 * (lt.max() - lt.min()))
                q3 = np.exp(lt.min() + 0.75 * (lt.max() - lt.min()))
                nu = self.basis_nu
                self.dop_scale_vector = np.where(nu <= 0, q3 ** nu, q1 ** nu) / self.nu_basis_area
            else:
                self.dop_scale_vector = np.ones(len(self.basis_nu))
            plan['dop_scale_vector'] = self.dop_scale_vector
            dop_a, dop_b = self.dop_indices
        else:
            self.dop_scale_vector = None
            dop_a = dop_b = None

        # ---- chrono block, drt1d.py:5557-5623 + 5792-5819
        if times is not None:
            in_scale = float(np.max(np.abs(step_sizes)))        # drtbase.py:463, applied at drt1d.py:5543-5549
            plan['input_signal_scale'] = in_scale if kw['scale_data'] else 1.0
            iscale = plan['input_signal_scale']
            rm_drt = eng.build_response(times[None], tau[None], np.asarray(step_times)[None],
                                        np.asarray(step_sizes)[None], eps, mode, self.interpolate_lookups)[0]
            rm[:nc, ns:] = rm_drt / iscale
            if kw['smooth_inf_response']:                       # mat1d.py:399-421
                inf_rv = np.zeros(nc)
                for st, sa in zip(step_times, step_sizes):
                    inf_rv += sa * unit_step(times, st)
            else:
                inf_rv = i_signal - np.mean(i_signal[times < step_times[0]])
            cap_rv = np.zeros(nc)                               # mat1d.py:424-443
            for st, sa in zip(step_times, step_sizes):
                cap_rv[times >= st] += sa * (times[times >= st] - st)
            a, b = sp['v_baseline']['index'], sp['v_baseline']['index'] + sp['v_baseline']['size']
            vb = np.zeros((nc, b - a))                          # background.get_baseline_matrix :23-37
            for d in range(kw['v_baseline_deg'] + 1):
                vb[:, d] = (times - times[0]) ** d
            if kw['v_baseline_sqrt']:
                vb[:, -1] = (times - times[0]) ** 0.5
            vb_scale = np.max(vb, axis=0)
            plan['v_baseline_scale'] = vb_scale
            rm[:nc, a:b] = dev(vb / vb_scale[None, :])
            if 'R_inf' in sp:
                rm[:nc, sp['R_inf']['index']] = dev(inf_rv / iscale)
            if 'C_inv' in sp:
                rm[:nc, sp['C_inv']['index']] = dev(cap_rv / iscale * kw['capacitance_scale'])
            if self.fit_dop:                                    # drt1d.py:5612-5619, :5548-5549, :5816
                rm_dop = eng.build_dop_v(times[None], self.basis_nu, np.asarray(step_times)[None],
                                         np.asarray(step_sizes)[None], self.nu_epsilon)[0]
                rm[:nc, dop_a:dop_b] = rm_dop / iscale * dev(self.dop_scale_vector)[None, :]
                plan['rm_dop_chrono'] = rm_dop
            # inductance response is identically zero for ideal steps (mat1d.py:378-396)
            plan['rm_drt_chrono'] = rm_drt
            plan['inf_rv'], plan['cap_rv'] = inf_rv, cap_rv

        # ---- EIS block, drt1d.py:5625-5671 + 5824-5857
        if frequencies is not None:
            a_re, a_im = eng.build_impedance(frequencies[None], tau[None], eps, mode, self.interpolate_lookups)
            rm[nc:nc + nf, ns:] = a_re[0]
            rm[nc + nf:, ns:] = a_im[0]
            omega = 2 * np.pi * frequencies
            if 'R_inf' in sp:
                rm[nc:nc + nf, sp['R_inf']['index']] = 1.0
            if 'inductance' in sp:
                rm[nc + nf:, sp['inductance']['index']] = dev(omega * kw['inductance_scale'])
            if 'C_inv' in sp:
                rm[nc + nf:, sp['C_inv']['index']] = dev(-1.0 / omega * kw['capacitance_scale'])
            if self.fit_dop:
                zd = eng.build_dop_z(frequencies[None], self.basis_nu, self.nu_epsilon)[0]
                zd = zd * dev(self.dop_scale_vector)[None, :]
                rm[nc:nc + nf, dop_a:dop_b] = zd.real
                rm[nc + nf:, dop_a:dop_b] = zd.imag
                plan['zm_dop_dev'] = zd / dev(self.dop_scale_vector)[None, :]
            plan['a_re'], plan['a_im'] = a_re[0], a_im[0]

        # ---- penalty matrices, drt1d.py:5673-5734 + 5863-5910
        pen = torch.zeros(3, n, n, dtype=torch.float64, device=eng.device)
        m_drt = eng.build_penalty(np.log(tau)[None], eps, is_uniform(np.log(tau)))[0]
        pen[:, ns:, ns:] = m_drt
        diag = {'R_inf': kw['ohmic_penalty'], 'inductance': kw['inductance_penalty'],
                'C_inv': kw['capacitance_penalty']}
        for name, val in diag.items():
            if name in sp:
                i = sp[name]['index']
                pen[:, i, i] = val
        if 'v_baseline' in sp:
            a, b = self.get_special_indices('v_baseline')
            vbp = kw['v_baseline_penalty']
            vals = np.full(b - a, vbp) if np.isscalar(vbp) else np.asarray(vbp, dtype=float)
            if len(vals) != b - a:
                raise ValueError('If v_baseline_penalty is iterable, it must match the number of v_baseline '
                                 f'parameters. Number of v_baseline parameters is {b - a}')
            for j, i in enumerate(range(a, b)):
                pen[:, i, i] = float(vals[j])
        if 'vz_offset' in sp:
            i = sp['vz_offset']['index']
            pen[:, i, i] = 1 / kw['vz_offset_scale']
        if self.fit_dop:
            m_dop = eng.build_penalty(self.basis_nu[None], self.nu_epsilon, is_uniform(self.basis_nu))[0]
            pen[:, dop_a:dop_b, dop_a:dop_b] = m_dop
        plan['pen'] = pen
        plan['pen_hint'] = eng.penalty_hint(pen, ns) if not self.fit_dop else None
        plan['rm'] = rm

        # ---- variance-estimation matrices, drt1d.py:614-636
        plan['vmm_chrono'] = None             # 'uniform': rank one (ones / Nt), never materialised
        if times is not None and kw['chrono_error_structure'] != 'uniform':
            if kw['chrono_error_structure'] is not None:
                raise ValueError(f"Invalid error structure {kw['chrono_error_structure']}")
            # the decorrelation blocks follow the NON-consecutive steps (drt1d.py:616): a finite-rise step made of
            # consecutive samples is one block, not several
            plan['vmm_chrono'] = eng.build_chrono_vmm(times[None], np.asarray(self.nonconsec_step_times, dtype=float)[None],
                                                      kw['chrono_vmm_epsilon'])[0]
        if frequencies is not None:
            es = kw['eis_error_structure']
            if es not in (None, 'uniform'):
                raise ValueError(f'Invalid error structure {es}')
            plan['vmm_eis'] = eng.build_eis_vmm(frequencies[None], kw['eis_vmm_epsilon'], kw['eis_reim_cor'],
                                                es == 'uniform')[0]
        else:
            plan['vmm_eis'] = None

        # ---- constraints and l1 vector, qphb.py:521-557, drt1d.py:556-561
        h = np.zeros(n)
        if kw['nonneg']:
            for v in sp.values():
                if not v['nonneg']:
                    h[v['index']:v['index'] + v.get('size', 1)] = 1000
        else:
            rng = kw['neg_allowed_tau_range']
            if rng is not None:
                for v in sp.values():
                    if not v['nonneg']:
                        h[v['index']:v['index'] + v.get('size', 1)] = 1000
                lo, hi = rng                                    # drt1d._get_neg_allowed_indices
                idx = np.where((tau >= lo) & (tau <= hi))[0] + ns
                h[idx] = 1e5
            else:
                h[:] = 1e5
            for v in sp.values():
                if v['nonneg']:
                    h[v['index']:v['index'] + v.get('size', 1)] = 0
        l1 = np.zeros(n)
        l1[ns:] = hyp['l1_lambda_0']
        if self.fit_dop:
            l1[dop_a:dop_b] = hyp['dop_l1_lambda_0']
        plan['h_host'], plan['l1_host'] = h, l1
        plan['h'], plan['l1'] = dev(h), dev(l1)

        # ---- hybrid vz_offset strength, drt1d.py:503-522 + 6173-6226
        if 'vz_offset' in sp:
            fit_td = time_since_step(times, self.nonconsec_step_times, prestep_value=-1)
            chrono_tau_min = np.min(fit_td[fit_td > 0])
            eis_tau_max = np.max(1 / (2 * np.pi * frequencies))
            veps = kw['vz_offset_eps']
            cs = np.ones(nc)
            if veps is not None:
                sel = fit_td >= eis_tau_max
                cs[sel] = np.exp(-(veps * np.log(fit_td[sel] / eis_tau_max)) ** 2)
                cs[fit_td == -1] = 0
                f_inv = 1 / (2 * np.pi * frequencies)
                es_ = np.ones(nf)
                sel = f_inv <= chrono_tau_min
                es_[sel] = np.exp(-(veps * np.log(f_inv[sel] / chrono_tau_min)) ** 2)
            else:
                es_ = np.ones(nf)
            plan['chrono_vz_strength'], plan['eis_vz_strength'] = cs, es_
            plan['vz_strength_host'] = np.concatenate([cs, es_, es_])
            plan['vz_strength'] = dev(plan['vz_strength_host'])
        return plan

    def _build_plan_multi_eis(self, freqs, kw):
        """fit_eis_batch with one frequency grid PER SPECTRUM (freqs [B, Nf], e.g. instruments that log slightly
        different frequencies per sweep): every matrix of every spectrum is built on the GPU in one launch per
        kind (hdrt_build_* with n_grids = B) and the fit kernel reads them per spectrum (non-zero strides).  All
        spectra must lead to the same basis size; DOP and tau_supergrid are not available in this mode."""
        if self.fit_dop or self.tau_supergrid is not None or self.fixed_basis_tau is not None:
            _not_supported('per-spectrum frequency grids together with fit_dop / tau_supergrid / fixed_basis_tau')
        eng, dev, hyp = self.engine, self.engine.dev, kw['hypers']
        freqs = np.asarray(freqs, dtype=float)
        nbat, nf = freqs.shape
        sp = {}
        if self.fit_ohmic:
            self._add_special_qp_param(sp, 'R_inf', True)
        if self.fit_inductance:
            self._add_special_qp_param(sp, 'inductance', True)
        if self.fit_capacitance:
            self._add_special_qp_param(sp, 'C_inv', True)
        self.special_qp_params, self.basis_nu = sp, None
        ns = self.get_qp_mat_offset()
        # preprocessing.get_tau_lim + get_basis_tau (preprocessing.py:948-1013), vectorised over the spectra
        ext, ppd = self.extend_basis_decades, self.basis_tau_ppd
        lmin = np.log10(1 / (2 * np.pi * freqs.max(axis=1))) - ext
        lmax = np.log10(1 / (2 * np.pi * freqs.min(axis=1))) + ext
        exact = (lmax - lmin) * ppd + 1
        num = np.ceil(exact).astype(int)
        if np.any(num != num[0]):
            raise ValueError('per-spectrum frequency grids must lead to basis grids of one size; '
                             f'got sizes {sorted(set(num.tolist()))}: fit the groups separately')
        nb = int(num[0])
        add = 0.5 * (num - exact) / ppd
        taus = 10 ** ((lmin - add)[:, None] + ((lmax + add) - (lmin - add))[:, None] * np.linspace(0, 1, nb)[None, :])
        if self.tau_epsilon is None:
            self.tau_epsilon = 1 / np.log(10 ** (1 / ppd))
        eps, n, n_rows = self.tau_epsilon, ns + nb, 2 * nf
        self.basis_tau, self.f_fit, self.t_fit = taus, freqs, []
        a_re, a_im = eng.build_impedance(freqs, taus, eps, self._mode(), self.interpolate_lookups)
        rm = torch.zeros(nbat, n_rows, n, dtype=torch.float64, device=eng.device)
        rm[:, :nf, ns:], rm[:, nf:, ns:] = a_re, a_im
        omega = dev(2 * np.pi * freqs)
        if 'R_inf' in sp:
            rm[:, :nf, sp['R_inf']['index']] = 1.0
        if 'inductance' in sp:
            rm[:, nf:, sp['inductance']['index']] = omega * kw['inductance_scale']
        if 'C_inv' in sp:
            rm[:, nf:, sp['C_inv']['index']] = -1.0 / omega * kw['capacitance_scale']
        pen = torch.zeros(nbat, 3, n, n, dtype=torch.float64, device=eng.device)
        pen[:, :, ns:, ns:] = eng.build_penalty(np.log(taus), eps, True)      # log-uniform by construction
        for name, val in (('R_inf', kw['ohmic_penalty']), ('inductance', kw['inductance_penalty']),
                          ('C_inv', kw['capacitance_penalty'])):
            if name in sp:
                pen[:, :, sp[name]['index'], sp[name]['index']] = val
        es = kw['eis_error_structure']
        if es not in (None, 'uniform'):
            raise ValueError(f'Invalid error structure {es}')
        h = np.zeros(n)
        if not kw['nonneg']:
            if kw['neg_allowed_tau_range'] is not None:
                _not_supported('neg_allowed_tau_range with per-spectrum frequency grids')
            h[:] = 1e5
            for v in sp.values():
                if v['nonneg']:
                    h[v['index']:v['index'] + v.get('size', 1)] = 0
        l1 = np.zeros(n)
        l1[ns:] = hyp['l1_lambda_0']
        self.inductance_scale, self.capacitance_scale = kw['inductance_scale'], kw['capacitance_scale']
        return dict(data_type='eis', special_qp_params=sp, n_special=ns, n=n, n_rows=n_rows, n_chrono=0, n_freq=nf,
                    frequencies=freqs, times=None, basis_tau=taus, inductance_scale=kw['inductance_scale'],
                    capacitance_scale=kw['capacitance_scale'], weight_factor=kw['weight_factor'], hypers=hyp,
                    step_times=None, rm=rm, pen=pen, a_re=a_re, a_im=a_im, multi=True, vmm_chrono=None,
                    vmm_eis=eng.build_eis_vmm(freqs, kw['eis_vmm_epsilon'], kw['eis_reim_cor'], es == 'uniform'),
                    h_host=h, l1_host=l1, h=dev(h), l1=dev(l1))

    def _c_hypers(self, kw):
        hyp = kw['hypers']
        ch = _engine.default_hypers()
        for name in ('derivative_weights', 'sigma_ds', 's_alpha', 's_0', 'rho_alpha', 'rho_0'):
            v = np.broadcast_to(np.asarray(hyp[name], dtype=float), (3,))
            setattr(ch, name, (_engine._D3)(*v))
        ch.l2_lambda_0 = float(hyp['l2_lambda_0'])
        if self.fit_dop:
            for name in ('dop_derivative_weights', 'dop_sigma_ds', 'dop_s_alpha', 'dop_s_0', 'dop_rho_alpha',
                         'dop_rho_0'):
                v = np.broadcast_to(np.asarray(hyp[name], dtype=float), (3,))
                setattr(ch, name, (_engine._D3)(*v))
            ch.dop_l2_lambda_0 = float(hyp['dop_l2_lambda_0'])
        ch.iw_l1_lambda_0, ch.iw_l2_lambda_0 = float(kw['iw_l1_lambda_0']), float(kw['iw_l2_lambda_0'])
        if hyp['iw_alpha'] is not None:
            ch.has_iw_prior, ch.iw_alpha, ch.iw_beta = 1, float(hyp['iw_alpha']), float(hyp['iw_beta'])
        ch.xtol, ch.max_iter = float(kw['xtol']), int(kw['max_iter'])
        ch.weight_factor = float(kw['weight_factor']) if np.isscalar(kw['weight_factor']) else 1.0
        ch.chrono_weight_factor = float(kw['chrono_weight_factor'])
        ch.eis_weight_factor = float(kw['eis_weight_factor'])
        if kw.get('solve_rp') or kw.get('update_scale'):          # drt1d.py:573-607, :914-936
            ch.solve_rp, ch.update_scale = int(bool(kw.get('solve_rp'))), int(bool(kw.get('update_scale')))
            ch.normalize_dop = int(bool(self.fit_dop and self.normalize_dop))
            ch.rp_scale = float(hyp['rp_scale'])
            ch.basis_area = float(np.sqrt(np.pi) / self.tau_epsilon)
        if kw.get('init_weights_separately'):
            ch.init_weights_separately = 1
        if kw.get('wf_method') == 'weight':
            ch.hybrid_wf_method = 1
        if hyp.get('outlier_p') is not None:                     # qphb.py:232; error structure qphb.py:1497-1538
            ch.has_outlier_p, ch.outlier_p = 1, float(hyp['outlier_p'])
        return ch

    def _eval_matrix(self, plan, tau_eval):
        """basis.construct_func_eval_matrix (basis.py:488-514, order 0, gaussian) on ``tau_eval``, embedded in the
        full coefficient vector (zero columns for the special parameters).  None when no diagnostics are wanted."""
        if tau_eval is None:
            return None
        d = np.log(np.asarray(tau_eval, dtype=float))[:, None] - np.log(plan['basis_tau'])[None, :]
        em = np.zeros((d.shape[0], plan['n']))
        em[:, plan['n_special']:] = np.exp(-(self.tau_epsilon * d) ** 2)
        return self.engine.dev(em)

    # ------------------------------------------------------------------------------------------------
    # the batched fit core
    # ------------------------------------------------------------------------------------------------
    def _fit_core_batch(self, times, i_signal, v_batch, frequencies, z_batch, want_pq=False, diag_tau=None,
                        pfrt=None,
                        step_times=None,
                        step_sizes=None, nonneg=True, neg_allowed_tau_range=None, series_neg=False,
                        scale_data=True, update_scale=False, solve_rp=False,
                        offset_steps=True, step_offset_size=None, discard_first_n=None,
                        offset_baseline=True, v_baseline_deg=0, v_baseline_sqrt=False,
                        downsample=False, downsample_kw=None, subtract_background=False, background_type='static',
                        background_corr_power=None, estimate_background_kw=None, smooth_inf_response=True,
                        v_baseline_penalty=1e-6, ohmic_penalty=1e-6, inductance_penalty=1e-6,
                        capacitance_penalty=1e-6, inductance_scale=1e-5, capacitance_scale=1e-3,
                        background_penalty=1, penalty_type='integral', remove_extremes=False, extreme_kw=None,
                        init_weights_separately=False, chrono_error_structure='uniform', eis_error_structure=None,
                        remove_outliers=False, return_outlier_index=False, outlier_thresh=0.75,
                        chrono_vmm_epsilon=4, eis_vmm_epsilon=0.25, eis_reim_cor=0.25,
                        iw_l1_lambda_0=1e-4, iw_l2_lambda_0=1e-4,
                        vz_offset=True, vz_offset_scale=1, vz_offset_eps=1,
                        eis_weight_factor=None, chrono_weight_factor=None, hybrid_weight_factor_method=None,
                        eff_hp=True, weight_factor=1, xtol=1e-2, max_iter=50, peak_locations=None, **kw):
        # options outside the hot path (same keyword names as drt1d.py:102-137)
        for flag, name in ((series_neg, 'series_neg'),
                           (subtract_background, 'subtract_background'),
                           (remove_extremes, 'remove_extremes in a batched call (the flagged points differ per '
                                             'spectrum; use the single-spectrum fit_* methods)'),
                           (peak_locations is not None, 'peak_locations'),
                           ):
            if flag:
                _not_supported(f'{name}')
        if penalty_type != 'integral':
            _not_supported(f"penalty_type '{penalty_type}'")
        if not eff_hp:
            _not_supported('eff_hp=False')
        if nonneg and neg_allowed_tau_range is not None:        # drt1d.py:83-84
            raise ValueError('If nonneg==True, neg_allowed_tau_range cannot be specified')
        hypers = get_default_hypers(eff_hp, self.fit_dop)
        for key in kw:                                          # drt1d.py:415-419
            if key not in hypers:
                raise ValueError(f'Invalid keyword argument {key}')
        hypers.update(kw)
        if remove_outliers:
            _not_supported('remove_outliers in a batched call (the flagged points differ per spectrum; '
                           'use the single-spectrum fit_* methods or group the spectra yourself)')
        if return_outlier_index:             # drt1d.py:817-835: stop after initialize_weights
            max_iter = 0
        if (eis_weight_factor is None) != (chrono_weight_factor is None):
            warnings.warn('Both eis_weight_factor and chrono_weight_factor must be provided. '
                          'If only one is provided, it will be ignored.')
        if hybrid_weight_factor_method not in (None, 'weight', 'rp'):
            raise ValueError(f"Invalid hybrid_weight_factor_method argument {hybrid_weight_factor_method}. "
                             "Options: 'weight', 'rp', None")
        wf_method = None
        if eis_weight_factor is None or chrono_weight_factor is None:
            eis_weight_factor = chrono_weight_factor = 1
            wf_method = hybrid_weight_factor_method            # drt1d.py:745-800: only when no factors are given
        opts = dict(hypers=hypers, step_times=step_times, step_sizes=step_sizes, nonneg=nonneg,
                    neg_allowed_tau_range=neg_allowed_tau_range, scale_data=scale_data, offset_steps=offset_steps,
                    step_offset_size=step_offset_size, offset_baseline=offset_baseline,
                    v_baseline_deg=v_baseline_deg, v_baseline_sqrt=v_baseline_sqrt,
                    smooth_inf_response=smooth_inf_response, v_baseline_penalty=v_baseline_penalty,
                    ohmic_penalty=ohmic_penalty, inductance_penalty=inductance_penalty,
                    capacitance_penalty=capacitance_penalty, inductance_scale=inductance_scale,
                    capacitance_scale=capacitance_scale, chrono_error_structure=chrono_error_structure,
                    eis_error_structure=eis_error_structure, eis_vmm_epsilon=eis_vmm_epsilon,
                    chrono_vmm_epsilon=chrono_vmm_epsilon,
                    eis_reim_cor=eis_reim_cor, iw_l1_lambda_0=iw_l1_lambda_0, iw_l2_lambda_0=iw_l2_lambda_0,
                    vz_offset=vz_offset, vz_offset_scale=vz_offset_scale, vz_offset_eps=vz_offset_eps,
                    eis_weight_factor=eis_weight_factor, chrono_weight_factor=chrono_weight_factor,
                    weight_factor=weight_factor, xtol=xtol, max_iter=max_iter,
                    solve_rp=bool(solve_rp and scale_data), update_scale=bool(update_scale and scale_data),
                    init_weights_separately=bool(init_weights_separately), wf_method=wf_method)
        if solve_rp and not scale_data and self.warn:
            warnings.warn('solve_rp is ignored if scale_data=False')
        self.v_baseline_deg, self.v_baseline_sqrt = v_baseline_deg, v_baseline_sqrt
        multi = frequencies is not None and np.ndim(frequencies) == 2
        if z_batch is not None:
            z_batch = np.asarray(z_batch)
            if z_batch.ndim != 2 or z_batch.shape[1] != np.shape(frequencies)[-1]:
                raise ValueError('z must have shape [batch, len(frequencies)]')
            if multi and (times is not None or np.shape(frequencies)[0] != len(z_batch)):
                raise ValueError('per-spectrum frequencies must have shape [batch, Nf] (EIS fits only)')
        if v_batch is not None:
            v_batch = np.asarray(v_batch, dtype=float)
            if v_batch.ndim != 2 or v_batch.shape[1] != len(times):
                raise ValueError('v_signal must have shape [batch, len(times)]')
        batch = len(z_batch) if z_batch is not None else len(v_batch)

        if discard_first_n is not None and times is not None:
            # drt1d.py:170-181 + pp.discard_first_n_chrono (preprocessing.py:473-504): drop the first n samples of
            # every segment (the pre-step one included) and move the assumed step time back accordingly
            times = np.asarray(times, dtype=float)
            i_signal = np.asarray(i_signal, dtype=float)
            dt_short = np.min(np.diff(times))
            starts = np.insert(identify_steps(i_signal, False), 0, 0)
            ends = np.append(starts[1:], len(times))
            keep = np.concatenate([np.arange(a + discard_first_n, b) for a, b in zip(starts, ends)])
            times, i_signal, v_batch = times[keep], i_signal[keep], v_batch[:, keep]
            if step_offset_size is None:
                step_offset_size = -(dt_short + np.min(np.diff(times)) * (discard_first_n - 1e-8))
                opts['step_offset_size'] = step_offset_size
        self.sample_index = None if times is None else np.arange(len(times))
        if downsample and times is not None:
            # DRTBase.process_chrono_signals (drtbase.py:296-339): step data come from the raw signal, then the
            # traces are filtered and decimated (on the GPU, all of them at once) before anything else happens
            from . import preprocessing as _pp
            times = np.asarray(times, dtype=float)
            i_signal = np.asarray(i_signal, dtype=float)
            if step_times is None:
                step_times, step_sizes = step_info(times, i_signal, offset_steps, step_offset_size)
            elif step_sizes is None:
                step_sizes = step_sizes_from_signal(times, i_signal, np.asarray(step_times, dtype=float))
            step_times = np.asarray(step_times, dtype=float)
            nonconsec = step_times
            if len(step_times) > 1:
                keep = np.diff(step_times) > 1.1 * np.min(np.diff(times))
                nonconsec = np.insert(step_times[1:][keep], 0, step_times[0])
            dkw = dict(downsample_kw) if downsample_kw is not None else {'prestep_samples': 10, 'target_times': None}
            times, i_signal, v_batch, self.sample_index = _pp.downsample_data(
                times, i_signal, v_batch, stepwise_sample_times=True, step_times=nonconsec, op_mode=self.chrono_mode,
                engine=self.engine, **dkw)
            opts['step_times'], opts['step_sizes'] = step_times, np.asarray(step_sizes, dtype=float)

        plan = self._build_plan_multi_eis(frequencies, opts) if multi else self._build_plan(times, i_signal, frequencies, opts)
        plan['opts'] = opts
        plan['model'] = self
        sp, nc, nf = plan['special_qp_params'], plan['n_chrono'], plan['n_freq']
        self.fit_kwargs = dict(smooth_inf_response=smooth_inf_response, offset_steps=offset_steps,
                               step_offset_size=step_offset_size, nonneg=nonneg, eff_hp=eff_hp,
                               penalty_type=penalty_type, subtract_background=False, background_type=background_type,
                               background_corr_power=background_corr_power,
                               neg_allowed_tau_range=neg_allowed_tau_range, **hypers)

        # ---- scale_data (drtbase.py:439-514), vectorised over the batch
        if scale_data:
            rp_est = estimate_rp_batch(plan['times'], self.step_times, self.step_sizes, v_batch, z_batch)
            cscale = rp_est / hypers['rp_scale']
        else:
            rp_est = np.ones(batch)
            cscale = np.ones(batch)
        scales = dict(coefficient_scale=cscale)
        rv_pin = self.engine.pinned(batch, plan['n_rows'])     # pinned host staging buffer for the H2D copy
        rv = rv_pin.numpy()
        if nc:
            rscale = plan['input_signal_scale'] * rp_est / hypers['rp_scale'] if scale_data else np.ones(batch)
            v_scaled = v_batch / rscale[:, None]
            pre = plan['times'] < self.step_times[0]
            baseline = np.median(v_scaled[:, pre], axis=1)                       # drt1d.py:5534-5538
            offset = -baseline if offset_baseline else np.zeros(batch)          # drt1d.py:525-531
            rv[:, :nc] = v_scaled + offset[:, None]
            scales.update(response_signal_scale=rscale, scaled_response_offset=offset)
        if nf:
            zs = z_batch / cscale[:, None]
            rv[:, nc:nc + nf] = zs.real
            rv[:, nc + nf:] = zs.imag
        eng = self.engine
        rv_dev = rv_pin.to(eng.device, non_blocking=True)
        copied = torch.cuda.Event()
        copied.record()
        dop_range = self.dop_indices if self.fit_dop else None
        wf_vec = None
        if not np.isscalar(weight_factor):          # per-row factors (kk_fit, drt1d.py:1394-1405)
            wfa = np.asarray(weight_factor, dtype=float)
            if wfa.shape != (plan['n_rows'],):
                raise ValueError(f"weight_factor must be a scalar or have one entry per data row ({plan['n_rows']})")
            wf_vec = eng.dev(wfa)
        hybrid_wf = None
        if opts['wf_method'] == 'rp' and plan['data_type'] == 'hybrid':
            # hybrid_weight_factor_method='rp' (drt1d.py:761-791): balance by the Rp each domain sees on its own
            rp_eis = estimate_rp_batch(None, None, None, None, z_batch)
            rp_chrono = estimate_rp_batch(plan['times'], self.step_times, self.step_sizes, v_batch, None)
            rp_tot = cscale * hypers['rp_scale']
            hybrid_wf = eng.dev(np.stack([rp_chrono ** 0.75 / (rp_eis ** 0.25 * rp_tot ** 0.5),
                                          rp_eis ** 0.75 / (rp_chrono ** 0.25 * rp_tot ** 0.5)], axis=1))
        vz_index = sp['vz_offset']['index'] if 'vz_offset' in sp else -1
        vb_range = self.get_special_indices('v_baseline') if 'vz_offset' in sp else (-1, -1)
        # host copies of the (small, shared) impedance matrices for predict_z: fetched before the fit is launched, so
        # that nothing between the launch and the caller's first read of a result waits for the fit kernel
        if nf and not plan.get('multi'):
            plan['zm_drt_host'] = (plan['a_re'] + 1j * plan['a_im']).cpu().numpy()
            if self.fit_dop:
                plan['zm_dop_host'] = plan['zm_dop_dev'].cpu().numpy()
        c_hyp, eval_mat = self._c_hypers(opts), self._eval_matrix(plan, diag_tau)

        def launch(out=None):
            return eng.qphb_fit_batch(plan['rm'], rv_dev, plan['pen'], plan['h'], plan['l1'], plan['n_special'],
                                      vmm_eis=plan['vmm_eis'], vmm_chrono=plan['vmm_chrono'], n_chrono=nc,
                                      dop_range=dop_range, vz_index=vz_index, vb_range=vb_range,
                                      vz_strength=plan.get('vz_strength'), hybrid=(plan['data_type'] == 'hybrid'),
                                      hypers=c_hyp, want_pq=want_pq, eval_mat=eval_mat, want_resid=diag_tau is not None,
                                      pfrt=pfrt, weight_factor_vec=wf_vec, hybrid_wf=hybrid_wf, out=out,
                                      pen_hint=plan.get('pen_hint'))
        raw = launch()
        plan['diag_tau'] = None if diag_tau is None else np.asarray(diag_tau, dtype=float)
        copied.synchronize()        # the pinned staging buffer may be refilled by the next call from here on
        rv_host = rv.copy() if want_pq else None
        if opts['solve_rp'] or opts['update_scale']:
            # DRTBase.update_data_scale (drtbase.py:516-536, galvanostatic) with the factors the kernel applied
            torch.cuda.current_stream(eng.device).synchronize()
            sf = raw['scale_factors'].cpu().numpy()
            tot = sf[:, 0] * sf[:, 1]
            scales['coefficient_scale'] = scales['coefficient_scale'] / tot
            if nc:
                scales['response_signal_scale'] = scales['response_signal_scale'] / tot
                scales['scaled_response_offset'] = scales['scaled_response_offset'] * tot
            scales['update_scale_factor'] = sf[:, 1]
            scales['dop_column_scale'] = sf[:, 2]          # dop_scale_vector /= dop_rescale_factor (drt1d.py:589-592)
            if rv_host is not None:
                rv_host *= tot[:, None]
        # 'relaunch' repeats the fit kernel on the inputs already resident in HBM (bench.py times it)
        res = BatchFit(plan, raw, scales, dict(rv=rv_host, rv_dev=rv_dev, h2d_bytes=rv.nbytes, relaunch=launch))
        self.last_batch = res
        self.fit_type = f"qphb_{plan['data_type']}"
        return res

    # ------------------------------------------------------------------------------------------------
    # public batched API (new; the reference fits one spectrum per call, drtmd.py:311-312)
    # ------------------------------------------------------------------------------------------------
    def fit_eis_batch(self, frequencies, z, nonneg=True, neg_allowed_tau_range=None, scale_data=True,
                      update_scale=False, error_structure=None, vmm_epsilon=0.25, vmm_reim_cor=0.25, **kwargs):
        """fit_eis for z [batch, len(frequencies)] on a shared frequency grid.  Returns a BatchFit."""
        return self._fit_core_batch(None, None, None, frequencies, z, nonneg=nonneg,
                                    neg_allowed_tau_range=neg_allowed_tau_range, scale_data=scale_data,
                                    update_scale=update_scale, eis_error_structure=error_structure,
                                    eis_vmm_epsilon=vmm_epsilon, eis_reim_cor=vmm_reim_cor, **kwargs)
    # every *_batch entry accepts diag_tau=<tau grid>: the kernel then also returns what DRTMD.fit_observation
    # computes after each fit (distribution variance on that grid, residual sums for llh / rss)

    def fit_chrono_batch(self, times, i_signal, v_signal, step_times=None, step_sizes=None, nonneg=True,
                         error_structure='uniform', vmm_epsilon=4, **kwargs):
        """fit_chrono for v_signal [batch, len(times)] with a shared time grid and input signal."""
        return self._fit_core_batch(times, i_signal, v_signal, None, None, step_times=step_times,
                                    step_sizes=step_sizes, nonneg=nonneg, chrono_error_structure=error_structure,
                                    chrono_vmm_epsilon=vmm_epsilon, **kwargs)

    def fit_hybrid_batch(self, times, i_signal, v_signal, frequencies, z, step_times=None, step_sizes=None,
                         nonneg=True, **kwargs):
        """fit_hybrid for v_signal [batch, Nt], z [batch, Nf] with shared grids and input signal."""
        return self._fit_core_batch(times, i_signal, v_signal, frequencies, z, step_times=step_times,
                                    step_sizes=step_sizes, nonneg=nonneg, **kwargs)

    # ------------------------------------------------------------------------------------------------
    # reference single-spectrum API (drt1d.py:1197-1268)
    # ------------------------------------------------------------------------------------------------
    def _qphb_fit_core(self, times, i_signal, v_signal, frequencies, z, remove_outliers=False, outlier_thresh=0.75,
                       **kw):
        """Single-spectrum entry (drt1d.py:102-1104) = a batch of one, plus the remove_outliers two-pass flow
        (drt1d.py:217-303) on the host."""
        if times is not None:
            times, i_signal, v_signal = np.array(times), np.array(i_signal), np.array(v_signal)
        if frequencies is not None:
            frequencies, z = np.array(frequencies), np.array(z)
        if kw.pop('remove_extremes', False):                # drt1d.py:188-215, preprocessing.py:844-857
            ekw = kw.pop('extreme_kw', None) or {'qr_size': 0.8, 'qr_thresh': 1.5}

            def extreme(y):
                q_lo = np.percentile(y, 50 - 100 * ekw['qr_size'] / 2)
                q_hi = np.percentile(y, 50 + 100 * ekw['qr_size'] / 2)
                qr = q_hi - q_lo
                return (y < q_lo - qr * ekw['qr_thresh']) | (y > q_hi + qr * ekw['qr_thresh'])
            if times is not None:
                flag = extreme(i_signal) | extreme(v_signal)
                if flag.any():
                    if self.warn:
                        warnings.warn('Identified extreme values in chrono data at the following '
                                      f'indices: {np.where(flag)[0].tolist()}. '
                                      'These data points will be removed before fitting')
                    times, i_signal, v_signal = times[~flag], i_signal[~flag], v_signal[~flag]
            if frequencies is not None:
                flag = extreme(z.real) | extreme(z.imag)
                if flag.any():
                    if self.warn:
                        warnings.warn('Identified extreme values in EIS data at the following '
                                      f'indices: {np.where(flag)[0].tolist()}. '
                                      'These data points will be removed before fitting')
                    frequencies, z = frequencies[~flag], z[~flag]
        else:
            kw.pop('extreme_kw', None)
        self.eis_outlier_index = self.eis_outliers = None
        self.chrono_outlier_index = self.chrono_outliers = None
        if remove_outliers:
            if 'outlier_p' not in kw:
                raise ValueError('If remove_outliers is True, the prior probability of outlier presence, outlier_p, '
                                 'must be specified. A good starting value might be 0.01-0.05')
            chrono_idx, eis_idx = self._outlier_index(times, i_signal, v_signal, frequencies, z, outlier_thresh, kw)
            self.eis_outlier_index, self.chrono_outlier_index = eis_idx, chrono_idx
            kw = dict(kw)
            if kw.get('step_times') is None and times is not None:
                kw['step_times'] = self.step_times          # step times determined before the outlier removal
            if times is not None and np.sum(chrono_idx) > 0:
                if self.warn:
                    warnings.warn('Found outliers in chrono data at the following '
                                  f'indices: {np.where(chrono_idx)[0].tolist()}. '
                                  'These data points will be removed before fitting')
                self.chrono_outliers = (times[chrono_idx], i_signal[chrono_idx], v_signal[chrono_idx])
                times, i_signal, v_signal = times[~chrono_idx], i_signal[~chrono_idx], v_signal[~chrono_idx]
            if frequencies is not None and np.sum(eis_idx) > 0:
                if self.warn:
                    warnings.warn('Found outliers in EIS data at the following '
                                  f'indices: {np.where(eis_idx)[0].tolist()}. '
                                  'These data points will be removed before fitting')
                self.eis_outliers = (frequencies[eis_idx], z[eis_idx])
                frequencies, z = frequencies[~eis_idx], z[~eis_idx]
            kw['outlier_p'] = None                           # drt1d.py:297-298
        v_b = None if v_signal is None else np.asarray(v_signal, dtype=float)[None, :]
        z_b = None if z is None else np.asarray(z)[None, :]
        res = self._fit_core_batch(times, i_signal, v_b, frequencies, z_b, want_pq=True, **kw)
        self.z_fit = None if z is None else np.asarray(z).copy()
        self._store_single(res)

    def _outlier_index(self, times, i_signal, v_signal, frequencies, z, outlier_thresh, kw):
        """First pass of remove_outliers: (chrono_outlier_index, eis_outlier_index) from the outlier_t of the
        weight initialisation (drt1d.py:817-835)."""
        v_b = None if v_signal is None else np.asarray(v_signal, dtype=float)[None, :]
        z_b = None if z is None else np.asarray(z)[None, :]
        res = self._fit_core_batch(times, i_signal, v_b, frequencies, z_b, return_outlier_index=True, **kw)
        t = res.host(['outlier_t'])['outlier_t'][0]
        idx = (1 - t) > outlier_thresh
        nc = res.plan['n_chrono']
        chrono_idx = idx[:nc] if nc else None
        eis_idx = None
        if res.plan['n_freq']:
            nf = res.plan['n_freq']
            eis_idx = idx[nc:nc + nf] | idx[nc + nf:]        # bad real OR imag value
        return chrono_idx, eis_idx

    def fit_eis(self, frequencies, z, nonneg=True, neg_allowed_tau_range=None, scale_data=True, update_scale=False,
                error_structure=None, vmm_epsilon=0.25, vmm_reim_cor=0.25, **kwargs):
        self._qphb_fit_core(None, None, None, frequencies, z, nonneg=nonneg,
                            neg_allowed_tau_range=neg_allowed_tau_range, scale_data=scale_data,
                            update_scale=update_scale, eis_error_structure=error_structure,
                            eis_vmm_epsilon=vmm_epsilon, eis_reim_cor=vmm_reim_cor, **kwargs)

    def fit_chrono(self, times, i_signal, v_signal, step_times=None, step_sizes=None, nonneg=True,
                   neg_allowed_tau_range=None, scale_data=True, update_scale=False, offset_baseline=True,
                   offset_steps=True, step_offset_size=None, discard_first_n=None, subtract_background=False,
                   estimate_background_kw=None, downsample=False, downsample_kw=None, smooth_inf_response=True,
                   error_structure='uniform', vmm_epsilon=4, **kwargs):
        self._qphb_fit_core(times, i_signal, v_signal, None, None, step_times=step_times, step_sizes=step_sizes,
                            nonneg=nonneg, neg_allowed_tau_range=neg_allowed_tau_range, scale_data=scale_data,
                            update_scale=update_scale, offset_steps=offset_steps, step_offset_size=step_offset_size,
                            discard_first_n=discard_first_n, offset_baseline=offset_baseline, downsample=downsample,
                            downsample_kw=downsample_kw, subtract_background=subtract_background,
                            estimate_background_kw=estimate_background_kw, smooth_inf_response=smooth_inf_response,
                            chrono_error_structure=error_structure, chrono_vmm_epsilon=vmm_epsilon, **kwargs)

    def fit_hybrid(self, times, i_signal, v_signal, frequencies, z, step_times=None, step_sizes=None, nonneg=True,
                   neg_allowed_tau_range=None, scale_data=True, update_scale=False, offset_steps=True,
                   step_offset_size=None, discard_first_n=None, offset_baseline=True, subtract_background=False,
                   estimate_background_kw=None, downsample=False, downsample_kw=None, smooth_inf_response=True,
                   vz_offset=True, vz_offset_scale=1, vz_offset_eps=1, chrono_error_structure='uniform',
                   eis_error_structure=None, chrono_vmm_epsilon=4, eis_vmm_epsilon=0.25, eis_reim_cor=0.25,
                   eis_weight_factor=None, chrono_weight_factor=None, **kwargs):
        self._qphb_fit_core(times, i_signal, v_signal, frequencies, z, step_times=step_times, step_sizes=step_sizes,
                            nonneg=nonneg, neg_allowed_tau_range=neg_allowed_tau_range, scale_data=scale_data,
                            update_scale=update_scale, offset_steps=offset_steps, step_offset_size=step_offset_size,
                            discard_first_n=discard_first_n, offset_baseline=offset_baseline, downsample=downsample,
                            downsample_kw=downsample_kw, subtract_background=subtract_background,
                            estimate_background_kw=estimate_background_kw, smooth_inf_response=smooth_inf_response,
                            chrono_error_structure=chrono_error_structure, eis_error_structure=eis_error_structure,
                            chrono_vmm_epsilon=chrono_vmm_epsilon, eis_vmm_epsilon=eis_vmm_epsilon,
                            eis_reim_cor=eis_reim_cor, vz_offset=vz_offset, vz_offset_scale=vz_offset_scale,
                            vz_offset_eps=vz_offset_eps, eis_weight_factor=eis_weight_factor,
                            chrono_weight_factor=chrono_weight_factor, **kwargs)

    # ------------------------------------------------------------------------------------------------
    # PFRT (drt1d.py:2558-2716): one launch runs the initial fit and every continuation step
    # ------------------------------------------------------------------------------------------------
    def _pfrt_setup(self, factors, max_iter_per_step, max_init_iter, xtol, kw):
        hyp0 = get_default_hypers(True, self.fit_dop)
        factors = np.logspace(-1, 1, 11) if factors is None else np.asarray(factors, dtype=float)
        # the step hypers always start from the *default* s_0 / l2_lambda_0 (drt1d.py:2575-2580)
        kw = dict(kw, s_0=np.asarray(hyp0['s_0'], dtype=float), l2_lambda_0=hyp0['l2_lambda_0'])
        pfrt = dict(factors=factors, max_iter_per_step=max_iter_per_step, min_iter=2)
        return factors, pfrt, dict(kw, max_iter=max_init_iter, xtol=xtol)

    def _pfrt_fit_core_batch(self, times, i_signal, v_batch, frequencies, z_batch, factors=None, max_iter_per_step=10,
                             max_init_iter=20, xtol=1e-2, nonneg=True, want_p=False, diag_tau=None, **kw):
        factors, pfrt, kw = self._pfrt_setup(factors, max_iter_per_step, max_init_iter, xtol, kw)
        pfrt['want_p'] = want_p
        return self._fit_core_batch(times, i_signal, v_batch, frequencies, z_batch, nonneg=nonneg, pfrt=pfrt,
                                    diag_tau=diag_tau, **kw)

    def pfrt_fit_eis_batch(self, frequencies, z, **kw):
        """pfrt_fit_eis for z [batch, Nf]; BatchFit.pfrt_result() holds the per-factor outputs."""
        return self._pfrt_fit_core_batch(None, None, None, frequencies, z, **kw)

    def pfrt_fit_chrono_batch(self, times, i_signal, v_signal, **kw):
        return self._pfrt_fit_core_batch(times, i_signal, v_signal, None, None, **kw)

    def pfrt_fit_hybrid_batch(self, times, i_signal, v_signal, frequencies, z, **kw):
        return self._pfrt_fit_core_batch(times, i_signal, v_signal, frequencies, z, **kw)

    def _pfrt_fit_core(self, times, i_signal, v_signal, frequencies, z, factors=None, max_iter_per_step=10,
                       max_init_iter=20, xtol=1e-2, nonneg=True, **kw):
        """Single-spectrum PFRT.  As in the reference the object's fit attributes afterwards are those of the
        initial fit (first factor); pfrt_result holds the per-factor outputs."""
        factors_a, _, kw0 = self._pfrt_setup(factors, max_iter_per_step, max_init_iter, xtol, kw)
        f0 = factors_a[0]
        init_kw = dict(kw0, s_0=kw0['s_0'] * f0, l2_lambda_0=kw0['l2_lambda_0'] / f0)
        v_b = None if v_signal is None else np.asarray(v_signal, dtype=float)[None, :]
        z_b = None if z is None else np.asarray(z)[None, :]
        res = self._pfrt_fit_core_batch(times, i_signal, v_b, frequencies, z_b, factors=factors,
                                        max_iter_per_step=max_iter_per_step, max_init_iter=max_init_iter, xtol=xtol,
                                        nonneg=nonneg, want_p=True, **kw)
        pr = res.pfrt_result()
        self._qphb_fit_core(times, i_signal, v_signal, frequencies, z, nonneg=nonneg, **init_kw)
        s0 = np.asarray(get_default_hypers(True, self.fit_dop)['s_0'], dtype=float)
        l20 = get_default_hypers(True, self.fit_dop)['l2_lambda_0']
        self.pfrt_history = None        # per-iteration history is not exported by the batched kernel
        self.pfrt_result = {
            'factors': factors_a, 'step_x': list(pr['step_x'][0]), 'step_llh': list(pr['step_llh'][0]),
            'step_p_mat': list(pr['step_p_mat'][0]),
            'step_hypers': [{'s_0': s0 * f, 'l2_lambda_0': l20 / f} for f in factors_a],
            'step_backgrounds': [None] * len(factors_a), 'step_iters': pr['step_iters'][0],
        }

    def pfrt_fit_eis(self, frequencies, z, factors=None, max_iter_per_step=10, max_init_iter=20, xtol=1e-2,
                     nonneg=True, **kw):
        self._pfrt_fit_core(None, None, None, frequencies, z, factors=factors, max_iter_per_step=max_iter_per_step,
                            max_init_iter=max_init_iter, xtol=xtol, nonneg=nonneg, **kw)

    def pfrt_fit_chrono(self, times, i_signal, v_signal, factors=None, max_iter_per_step=10, max_init_iter=20,
                        xtol=1e-2, nonneg=True, **kw):
        self._pfrt_fit_core(times, i_signal, v_signal, None, None, factors=factors,
                            max_iter_per_step=max_iter_per_step, max_init_iter=max_init_iter, xtol=xtol,
                            nonneg=nonneg, **kw)

    def pfrt_fit_hybrid(self, times, i_signal, v_signal, frequencies, z, factors=None, max_iter_per_step=10,
                        max_init_iter=20, xtol=1e-2, nonneg=True, **kw):
        self._pfrt_fit_core(times, i_signal, v_signal, frequencies, z, factors=factors,
                            max_iter_per_step=max_iter_per_step, max_init_iter=max_init_iter, xtol=xtol,
                            nonneg=nonneg, **kw)

    # ------------------------------------------------------------------------------------------------
    # Kramers-Kronig test (drt1d.py:1370-1491, models/kk.py): a lightly regularised free-sign DRT fit on an
    # extended basis; points the fit cannot follow are flagged and the clean frequency window is returned
    # ------------------------------------------------------------------------------------------------
    def kk_fit(self, frequencies, z, nonneg=False, l2_lambda_0=1e-2, extend_basis_decades=2, outlier_index=None):
        keep = self.extend_basis_decades
        self.extend_basis_decades = extend_basis_decades
        try:
            weight_factor = 1
            if outlier_index is not None:           # zero weight, but the points stay in the residuals
                nf = len(frequencies)
                weight_factor = np.ones(2 * nf)
                weight_factor[np.asarray(outlier_index, dtype=int)] = 1e-10
                weight_factor[np.asarray(outlier_index, dtype=int) + nf] = 1e-10
            self.fit_eis(frequencies, z, nonneg=nonneg, l2_lambda_0=l2_lambda_0, weight_factor=weight_factor)
        finally:
            self.extend_basis_decades = keep

    def eval_kk_residuals(self, norm='modulus'):
        """drt1d.py:1472-1481: residuals of the last fit, in % of |Z| by default."""
        return _kk.normalize_residuals(self.z_fit, self.predict_z(np.asarray(self.f_fit)), norm=norm)

    def get_kk_outliers(self, norm='modulus', n_iter=2, p_thresh=1e-4, n_sigma=None, std_sample_fraction=0.6):
        """drt1d.py:1483-1486"""
        return _kk.get_outliers(self.eval_kk_residuals(norm=norm), n_iter, p_thresh, n_sigma=n_sigma,
                                std_sample_fraction=std_sample_fraction)

    def get_kk_limits(self, outlier_index, max_num_outliers=2):
        """drt1d.py:1488-1491"""
        return _kk.get_limits(np.asarray(self.f_fit), outlier_index, max_num_outliers=max_num_outliers)

    def kk_test(self, frequencies, z, nonneg=False, l2_lambda_0=1e-2, extend_basis_decades=2, norm='modulus',
                max_num_outliers=2, p_thresh=1e-4, n_sigma=None, std_sample_fraction=0.6, n_iter=2,
                n_outlier_iter=2, show_plot=False):
        """drt1d.py:1370-1390.  Returns (outlier_index, (f_min, f_max), (frequencies, z) inside the limits).
        Plotting is not part of this package: show_plot is accepted for signature compatibility only."""
        frequencies, z = np.asarray(frequencies, dtype=float), np.asarray(z)
        outlier_index = None
        for _ in range(n_iter):
            self.kk_fit(frequencies, z, nonneg=nonneg, l2_lambda_0=l2_lambda_0,
                        extend_basis_decades=extend_basis_decades, outlier_index=outlier_index)
            outlier_index = self.get_kk_outliers(norm=norm, p_thresh=p_thresh, n_iter=n_outlier_iter, n_sigma=n_sigma,
                                                 std_sample_fraction=std_sample_fraction)
            f_min, f_max = self.get_kk_limits(outlier_index, max_num_outliers=max_num_outliers)
            fz_clean = _kk.trim_data(frequencies, z, f_min, f_max)
        if show_plot and self.warn:
            warnings.warn('kk_test: plotting is outside hybdrt_b200; show_plot is ignored')
        return outlier_index, (f_min, f_max), fz_clean

    def _store_single(self, res):
        """Populate the attributes DRT._qphb_fit_core leaves behind (drt1d.py:1040-1104)."""
        pl, h, sc, opts = res.plan, res.host(), res.scales, res.plan['opts']
        st = int(h['status'][0])
        if st & (_engine.ST_NAN | _engine.ST_KKT_FAIL) and not np.all(np.isfinite(h['x'][0])):
            raise ValueError('Rank(A) < p or Rank([P; A; G]) < n')          # what cvxopt raises
        if (st & _engine.ST_MAXITER) and self.warn:
            warnings.warn(f"Solution did not converge within {opts['max_iter']} iterations. "
                          'This is usually not an issue.')
        nc = pl['n_chrono']
        self.coefficient_scale = float(sc['coefficient_scale'][0])
        self.impedance_scale = self.coefficient_scale if pl['n_freq'] else 1.0
        if nc:
            self.input_signal_scale = pl['input_signal_scale']
            self.response_signal_scale = float(sc['response_signal_scale'][0])
            self.scaled_response_offset = float(sc['scaled_response_offset'][0])
            self.v_baseline_scale = pl['v_baseline_scale']
        fp_b = res.fit_parameters()
        fp = {}
        for k, v in fp_b.items():
            fp[k] = None if v is None else (v[0] if np.ndim(v) >= 1 else v)
        fp['v_sigma_res'] = None
        fp['vz_offset_eps'] = opts['vz_offset_eps']
        fp['p_matrix'] = h['p_matrix'][0]
        fp['q_vector'] = h['q_vector'][0]
        self.fit_parameters = fp
        self.cvx_result = {'x': h['x'][0], 'primal objective': float(h['fun'][0]),
                           'status': 'unknown' if st & (_engine.ST_QP_MAXITERS | _engine.ST_KKT_FAIL) else 'optimal'}
        w_true = h['weights'][0] * opts['weight_factor']
        w_scaled = w_true.copy()
        cwf, ewf = opts['chrono_weight_factor'], opts['eis_weight_factor']
        if 'hybrid_wf' in h:                                 # factors chosen per spectrum (drt1d.py:745-800)
            cwf, ewf = float(h['hybrid_wf'][0, 0]), float(h['hybrid_wf'][0, 1])
        if pl['data_type'] == 'hybrid':
            w_scaled[:nc] *= cwf
            w_scaled[nc:] *= ewf
        rm = pl['rm'].cpu().numpy()
        if 'vz_offset' in pl['special_qp_params']:
            rm[:, pl['special_qp_params']['vz_offset']['index']] = h['vz_col'][0]
        pen = pl['pen'].cpu().numpy()
        x_of = h['x_overfit'][0]
        init_w = h['init_weights'][0]
        if 'update_scale_factor' in sc:                     # drt1d.py:924-933, :589-596
            x_of = x_of * sc['update_scale_factor'][0]
            init_w = init_w / sc['update_scale_factor'][0]
            if self.fit_dop:
                a, b = self.dop_indices
                rm[:, a:b] *= sc['dop_column_scale'][0]
                self.dop_scale_vector = self.dop_scale_vector * sc['dop_column_scale'][0]
        self.qphb_params = {
            'est_weights': h['est_weights'][0], 'init_weights': init_w, 'weights': w_scaled,
            'true_weights': w_true, 'xmx_norms': h['xmx_norms'][0],
            'dop_xmx_norms': h['dop_xmx_norms'][0] if 'dop_xmx_norms' in h else np.ones(3),
            'x_overfit_chrono': x_of if (pl['data_type'] == 'chrono' or 'x_overfit_eis' in h) else (x_of[:nc] if nc else None),
            'x_overfit_eis': (h['x_overfit_eis'][0] if 'x_overfit_eis' in h else
                              (x_of if pl['data_type'] == 'eis' else (x_of[nc:] if pl['n_freq'] else None))),
            'p_matrix': fp['p_matrix'], 'q_vector': fp['q_vector'], 'rho_vector': h['rho'][0],
            'dop_rho_vector': h['dop_rho'][0] if 'dop_rho' in h else None,
            's_vectors': [h['s_vectors'][0, k] for k in range(3)],
            'vmm': None if pl['vmm_eis'] is None else pl['vmm_eis'].cpu().numpy(),
            'l1_lambda_vector': pl['l1_host'], 'rm': rm, 'rv': res.extra['rv'][0],
            'penalty_matrices': {f'm{k}': pen[k] for k in range(3)}, 'hypers': pl['hypers'],
            'num_eis': pl['n_freq'], 'num_chrono': nc,
            'chrono_weight_factor': cwf, 'eis_weight_factor': ewf,
            'vz_strength_vec': pl.get('vz_strength_host', 1),
            'n_outer': int(h['n_outer'][0]), 'n_ipm': int(h['n_ipm'][0]), 'status': st,
            'outlier_t': h['outlier_t'][0] if 'outlier_t' in h else np.ones(pl['n_rows']),
        }
        # qphb.iterate_qphb appends one entry per outer iteration (qphb.py:950-966).  The kernel keeps the iterates on
        # chip, so only the LAST entry is populated (what the reference's own callers read: qphb_history[-1]['x'],
        # drt1d.py:1517, 4161, 4447; len(qphb_history) = outer iterations); the earlier ones are None.
        n_out = int(h['n_outer'][0])
        last = {'x': h['x'][0].copy(), 's_vectors': [h['s_vectors'][0, k].copy() for k in range(3)],
                'rho_vector': h['rho'][0].copy(), 'dop_rho_vector': h['dop_rho'][0].copy() if 'dop_rho' in h else None,
                'weights': None,            # the reference stores the weights that ENTERED the iteration; not exported
                'outlier_t': self.qphb_params['outlier_t'], 'fun': float(h['fun'][0]), 'cvx_result': self.cvx_result}
        self.qphb_history = [None] * max(n_out - 1, 0) + ([last] if n_out > 0 else [])

    # ------------------------------------------------------------------------------------------------
    # prediction (drt1d.py:3363-3542); matrices are rebuilt on the GPU for the requested grid
    # ------------------------------------------------------------------------------------------------
    def _vz_strength(self, times=None, frequencies=None):
        """drt1d._get_vz_strength_vec, drt1d.py:6173-6226"""
        veps = self.fit_parameters.get('vz_offset_eps', None) if self.fit_parameters else None
        have = len(self.t_fit) > 0 and len(self.f_fit) > 0 and veps is not None
        cs = es = None
        if have:
            fit_td = time_since_step(self.t_fit, self.nonconsec_step_times, prestep_value=-1)
            chrono_tau_min = np.min(fit_td[fit_td > 0])
            eis_tau_max = np.max(1 / (2 * np.pi * np.asarray(self.f_fit)))
        if times is not None:
            cs = np.ones(len(times))
            if have:
                td = time_since_step(times, self.nonconsec_step_times, prestep_value=-1)
                sel = td >= eis_tau_max
                cs[sel] = np.exp(-(veps * np.log(td[sel] / eis_tau_max)) ** 2)
                cs[td == -1] = 0
        if frequencies is not None:
            es = np.ones(len(frequencies))
            if have:
                f_inv = 1 / (2 * np.pi * frequencies)
                sel = f_inv <= chrono_tau_min
                es[sel] = np.exp(-(veps * np.log(f_inv[sel] / chrono_tau_min)) ** 2)
        return cs, es

    def extract_qphb_parameters(self, x):
        """drt1d.py:6228-6289 for one raw coefficient vector (scaled space, special parameters first)."""
        fp = self.last_batch.extract_parameters(np.asarray(x, dtype=float)[None])
        return {k: (v[0] if np.ndim(v) >= 1 else v) for k, v in fp.items()}

    def get_tau_eval(self, ppd):
        """drtbase.py:263-283: one decade beyond the basis grid on either side."""
        bt = self.fixed_basis_tau if self.fixed_basis_tau is not None else self.basis_tau
        if bt is None:
            raise ValueError('basis_tau must be set, either by specifying fixed_basis_tau or by fitting the '
                             'DRT instanceto data, before using get_tau_eval')
        lo, hi = np.min(np.log10(bt)) - 1, np.max(np.log10(bt)) + 1
        return np.logspace(lo, hi, int((hi - lo) * ppd) + 1)

    def _basis_eval_matrix(self, tau, order=0, basis_tau=None):
        """basis.construct_func_eval_matrix, gaussian basis (basis.py:488-514, derivatives :218-233)."""
        bt = self.basis_tau if basis_tau is None else basis_tau
        y = np.log(np.asarray(tau, dtype=float))[:, None] - np.log(bt)[None, :]
        e = self.tau_epsilon
        phi = np.exp(-(e * y) ** 2)
        if order == 0:
            return phi
        if order == 1:
            return -2 * e ** 2 * y * phi
        if order == 2:
            return (-2 * e ** 2 + 4 * e ** 4 * y ** 2) * phi
        if order == 3:
            return (12 * e ** 4 * y - 8 * e ** 6 * y ** 3) * phi
        _not_supported(f'distribution derivative order {order}')

    def predict_drt(self, tau=None, ppd=20, x=None, order=0, sign=1, normalize=False, normalize_by=None,
                    abs_norm=False):
        """drt1d.py:3040-3061: the distribution (or a derivative of it) on ``tau``."""
        tau = self.get_tau_eval(ppd) if tau is None else np.asarray(tau, dtype=float)
        if x is None:
            xd = self.fit_parameters['x']
        elif isinstance(x, dict):
            xd = x['x']
        else:
            x = np.asarray(x, dtype=float)
            xd = x if len(x) <= len(self.basis_tau) else self.extract_qphb_parameters(x)['x']
        if normalize_by is not None:
            normalize = True
        if normalize and normalize_by is None:
            normalize_by = float(np.sum(np.abs(xd) if abs_norm else xd) * self.tau_basis_area)
        return self._basis_eval_matrix(tau, order) @ xd / (normalize_by if normalize else 1)

    def predict_distribution(self, *args, **kw):
        warnings.warn('predict_distribution is deprecated and will be removed in the future. '
                      'Please use predict_drt instead', DeprecationWarning)
        return self.predict_drt(*args, **kw)

    def predict_z(self, frequencies, include_vz_offset=True, x=None, include_dop=True, include_drt=True,
                  include_inductance=True, include_ohmic=True, include_cap=True):
        frequencies = np.asarray(frequencies, dtype=float)
        fp = self.fit_parameters if x is None else x
        if not isinstance(fp, dict):
            fp = self.extract_qphb_parameters(fp)
        a_re, a_im = self.engine.build_impedance(frequencies[None], self.basis_tau[None], self.tau_epsilon,
                                                 self._mode(), self.interpolate_lookups)
        zm = (a_re[0] + 1j * a_im[0]).cpu().numpy()
        z = np.zeros(len(frequencies), dtype=complex)
        if include_drt:
            z += zm @ fp['x']
        if include_ohmic:
            z += fp.get('R_inf', 0)
        if include_inductance:
            z += fp.get('inductance', 0) * 2j * np.pi * frequencies
        if include_cap:
            z += fp.get('C_inv', 0) * (2j * np.pi * frequencies) ** -1
        if fp.get('x_dop') is not None and include_dop:
            zd = self.engine.build_dop_z(frequencies[None], self.basis_nu, self.nu_epsilon)[0].cpu().numpy()
            z += zd @ fp['x_dop']
        if include_vz_offset:
            _, es = self._vz_strength(None, frequencies)
            z *= (1 - fp.get('vz_offset', 0) * es)
        return z

    def predict_response(self, times=None, input_signal=None, step_times=None, step_sizes=None, x=None,
                         include_vz_offset=True, include_dop=True, include_drt=True, include_ohmic=True,
                         include_cap=True, v_baseline=None):
        """Voltage response (drt1d.py:3363-3461): at the fit times from the matrices of the fit, at any other
        times (with the fitted steps, or explicit step_times / step_sizes) from matrices rebuilt on the GPU."""
        pl = self.last_batch.plan
        if x is None:
            fp = self.fit_parameters
        else:
            fp = x if isinstance(x, dict) else self.extract_qphb_parameters(x)
        if input_signal is not None:
            _not_supported('predict_response from a new input signal (pass step_times and step_sizes)')
        if step_times is not None and step_sizes is None:
            raise ValueError('If input signal steps are provided, both step_times and step_sizes must be provided; '
                             'received step_times only')
        if times is None and step_times is None:
            times = pl['times']
            rm_drt = pl['rm_drt_chrono'].cpu().numpy()
            inf_rv, cap_rv = pl['inf_rv'], pl['cap_rv']
            rm_dop = pl['rm_dop_chrono'].cpu().numpy() if 'rm_dop_chrono' in pl else None
        else:
            times = np.asarray(pl['times'] if times is None else times, dtype=float)
            st = np.asarray(self.step_times if step_times is None else step_times, dtype=float)
            sa = np.asarray(self.step_sizes if step_times is None else step_sizes, dtype=float)
            eng = self.engine
            rm_drt = eng.build_response(times[None], self.basis_tau[None], st[None], sa[None], self.tau_epsilon,
                                        self._mode(), self.interpolate_lookups)[0].cpu().numpy()
            inf_rv, cap_rv = np.zeros(len(times)), np.zeros(len(times))
            for t0, a0 in zip(st, sa):                          # mat1d.py:399-443 (smooth_inf_response)
                inf_rv += a0 * unit_step(times, t0)
                cap_rv[times >= t0] += a0 * (times[times >= t0] - t0)
            rm_dop = (eng.build_dop_v(times[None], self.basis_nu, st[None], sa[None], self.nu_epsilon)[0].cpu().numpy()
                      if self.fit_dop else None)
        resp = np.zeros(len(times))
        if include_drt:
            resp += rm_drt @ fp['x']
        if include_ohmic:
            resp += inf_rv * fp.get('R_inf', 0)
        if include_cap:
            resp += fp.get('C_inv', 0) * cap_rv
        if fp.get('x_dop') is not None and include_dop and rm_dop is not None:     # drt1d.py:3431-3432
            resp += rm_dop @ fp['x_dop']
        if include_vz_offset:
            cs, _ = self._vz_strength(times, None)
            resp = resp * (1 + fp.get('vz_offset', 0) * cs)
        if v_baseline is None:                                  # predict_v_baseline, drt1d.py:3466-3473
            vb = np.zeros((len(times), len(fp['v_baseline'])))
            for d in range(self.v_baseline_deg + 1):
                vb[:, d] = (times - times[0]) ** d
            if self.v_baseline_sqrt:
                vb[:, -1] = (times - times[0]) ** 0.5
            v_baseline = vb @ fp['v_baseline']
        return resp + v_baseline

    def predict_r_inf(self):
        """drt1d.py:3573-3581 (gaussian DOP basis: the ohmic resistance is R_inf itself)."""
        return self.fit_parameters.get('R_inf', 0)

    def predict_r_tot(self):
        """drt1d.py:3583-3584"""
        return self.predict_r_inf() + self.predict_r_p()

    def predict_r_p(self, absolute=False):
        """Polarisation resistance: sum of DRT coefficients times basis area (drt1d.py:3552-3590)."""
        x = self.fit_parameters['x']
        return float(np.sum(np.abs(x) if absolute else x) * self.tau_basis_area)

    def predict_sigma(self, measurement):
        key = 'v_sigma_tot' if measurement == 'chrono' else 'z_sigma_tot'
        return self.fit_parameters.get(key, None)
