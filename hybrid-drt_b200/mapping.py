"""Mirror of ``hybdrt.mapping.DRTMD`` (hybdrt/mapping/drtmd.py:22-430) for the multi-observation fit path.

The reference fits one observation per ``fit_observation`` call in a Python loop (drtmd.py:303-319); every fit is
independent of the others.  Here ``fit_observations`` groups the observations by measurement grid (data type,
frequencies / times / input signal) and sends each group to the GPU as ONE batch through
``DRT.fit_eis_batch / fit_chrono_batch / fit_hybrid_batch``; the per-observation bookkeeping the reference does
after each fit (``obs_x`` on the tau supergrid, ``obs_special``, ``obs_tau_indices``, ``obs_fit_status``, error
flags) is then filled for the whole group with array operations.

With ``shard=True`` under ``torchrun`` every rank fits an interleaved slice of each group and the results are
gathered on all ranks (sharding.gather_results); no collective runs during the fits.

The post-fit diagnostics the reference computes after every fit (drtmd.py:256-279: ``obs_drt_var`` =
diag of ``estimate_distribution_cov`` on the tau supergrid with ``extend_var``, ``obs_llh``, ``obs_rss`` with uniform
weights, normalised) come out of the same kernel launch (``diag_tau``).

``fit_type='pfrt'`` (drtmd.py:1140-1160, 1304-1342): every observation gets one solution per factor
(``obs_x`` [n, F, n_tau], ``obs_special[key]`` [n, F]); the whole group still goes to the GPU as one launch
(initial fit + all continuation steps inside the kernel).

Not mirrored (outside SURVEY.md section 8): file readers, resolve / filter / badness scoring, prediction
helpers.
"""
import time

import numpy as np

from . import engine as _engine
from . import sharding as _sharding
from .models import DRT, nearest_index


class DRTMD:
    def __init__(self, tau_supergrid, psi_dim_names=None, store_attr_categories=None, extend_basis_decades=1,
                 tau_basis_type='gaussian', tau_epsilon=None, step_model='ideal', chrono_mode='galv',
                 fit_inductance=True, fit_ohmic=True, fit_capacitance=False, fixed_basis_nu=None, fit_dop=False,
                 normalize_dop=True, nu_basis_type='gaussian', nu_epsilon=None, time_precision=10,
                 input_signal_precision=10, frequency_precision=10, fit_kw=None, fit_type='drt',
                 pfrt_factors=None, print_diagnostics=False, print_progress=True, warn=False, llh_kw=None,
                 rss_kw=None, device=0):
        for kw_dict in (llh_kw, rss_kw):
            if kw_dict and (kw_dict.get('normalize', True) is not True or kw_dict.get('weights', 'uniform') != 'uniform'):
                raise NotImplementedError('hybdrt_b200: llh_kw / rss_kw other than the DRTMD defaults')
        if fit_type not in ('drt', 'pfrt'):
            raise ValueError(f"Invalid fit_type {fit_type}. Options: ['drt', 'pfrt']")
        self.pfrt_factors = np.logspace(-0.7, 0.7, 11) if pfrt_factors is None else np.asarray(pfrt_factors, dtype=float)
        self.tau_supergrid = np.asarray(tau_supergrid, dtype=float)
        self.drt1d = DRT(interpolate_integrals=True, tau_supergrid=self.tau_supergrid, tau_epsilon=tau_epsilon,
                         tau_basis_type=tau_basis_type, fixed_basis_nu=fixed_basis_nu, nu_epsilon=nu_epsilon,
                         nu_basis_type=nu_basis_type, extend_basis_decades=extend_basis_decades,
                         step_model=step_model, chrono_mode=chrono_mode, fit_dop=fit_dop, normalize_dop=normalize_dop,
                         fit_inductance=fit_inductance, fit_ohmic=fit_ohmic, fit_capacitance=fit_capacitance,
                         warn=warn, device=device)
        self.psi_dim_names = psi_dim_names
        self.store_attr_categories = store_attr_categories or ['config', 'fit_core']
        self.tau_basis_type, self.tau_epsilon = tau_basis_type, self.drt1d.tau_epsilon
        self.fit_inductance, self.fit_ohmic, self.fit_capacitance = fit_inductance, fit_ohmic, fit_capacitance
        self.fit_dop, self.normalize_dop = fit_dop, normalize_dop
        self.fit_type = fit_type
        self.fit_kw = dict({'nonneg': True}, **(fit_kw or {}))
        self.print_progress, self.warn, self.print_diagnostics = print_progress, warn, print_diagnostics
        self.clear_obs()

    # ---- containers (drtmd.py:101-142, 379-430) ------------------------------------------------------
    def clear_obs(self):
        self.obs_psi = np.zeros((0, len(self.psi_dim_names))) if self.psi_dim_names is not None else None
        self.obs_data, self.obs_group_id = [], []
        self.obs_data_badness = np.zeros(0)
        self.obs_ignore_flag = np.zeros(0, dtype=bool)
        self._n = 0
        self.clear_fits()

    def clear_fits(self):
        n = self._n
        self.obs_fit_attr = [None] * n
        self.obs_fit_status = np.zeros(n, dtype=bool)
        self.obs_fit_errors = [None] * n
        self.obs_fit_badness = np.zeros(n)
        self.obs_tau_indices = [None] * n
        self.obs_x = np.zeros((n, *self.drt_param_shape()))
        self.obs_drt_var = np.zeros((n, *self.drt_param_shape()))
        self.obs_llh, self.obs_rss = np.zeros(n), np.zeros(n)
        self.obs_special = None
        self.obs_outer_iterations = np.zeros(n, dtype=int)   # not in the reference: iteration count per fit
        self.obs_status = np.zeros(n, dtype=int)             # HDRT_ST_* bits per fit

    @property
    def num_obs(self):
        return self._n

    @property
    def fitted_obs_index(self):
        return np.where(self.obs_fit_status)[0]

    @property
    def tau_basis_area(self):
        return self.drt1d.tau_basis_area

    def drt_param_shape(self, factor_index=None):
        """drtmd.py:1304-1316"""
        if self.fit_type == 'pfrt':
            if factor_index is None:
                return [len(self.pfrt_factors), len(self.tau_supergrid)]
            nf = len(np.atleast_1d(factor_index))
            return [nf, len(self.tau_supergrid)] if nf > 1 else [len(self.tau_supergrid)]
        return [len(self.tau_supergrid)]

    def special_param_shape(self, key):
        """drtmd.py:1318-1337"""
        size = self.drt1d.special_qp_params[key].get('size', 1)
        lead = [len(self.pfrt_factors)] if self.fit_type == 'pfrt' else []
        return lead if size == 1 else lead + [size]

    # ---- observations (drtmd.py:186-243) -------------------------------------------------------------
    def add_observation(self, psi, chrono_data, eis_data, group_id=None, fit=False):
        psi = np.atleast_1d(psi).flatten()
        if self.obs_psi is None:
            self.obs_psi = np.zeros((0, len(psi)))
        if len(psi) != self.obs_psi.shape[1]:
            raise ValueError(f'psi must have length {self.obs_psi.shape[1]}')
        for name, data, k in (('chrono', chrono_data, 3), ('eis', eis_data, 2)):
            if data is not None and (not isinstance(data, tuple) or len(data) != k):
                raise ValueError(f'Expected {name} data tuple to contain {k} arrays')
        self.obs_psi = np.vstack([self.obs_psi, psi[None, :]])
        self.obs_data.append((chrono_data, eis_data))
        self.obs_group_id.append(group_id)
        self.obs_data_badness = np.append(self.obs_data_badness, 0)
        self.obs_ignore_flag = np.append(self.obs_ignore_flag, False)
        self._n += 1
        nt = len(self.tau_supergrid)
        self.obs_fit_attr.append(None)
        self.obs_fit_errors.append(None)
        self.obs_tau_indices.append(None)
        self.obs_fit_status = np.append(self.obs_fit_status, False)
        self.obs_fit_badness = np.append(self.obs_fit_badness, 0)
        self.obs_x = np.concatenate([self.obs_x, np.zeros((1, *self.drt_param_shape()))], axis=0)
        self.obs_drt_var = np.concatenate([self.obs_drt_var, np.zeros((1, *self.drt_param_shape()))], axis=0)
        self.obs_llh, self.obs_rss = np.append(self.obs_llh, 0), np.append(self.obs_rss, 0)
        self.obs_outer_iterations = np.append(self.obs_outer_iterations, 0)
        self.obs_status = np.append(self.obs_status, 0)
        if self.obs_special is not None:
            for key in list(self.obs_special):
                pad = np.zeros((1,) + self.obs_special[key].shape[1:])
                self.obs_special[key] = np.concatenate([self.obs_special[key], pad], axis=0)
        if fit:
            self.fit_observation(self._n - 1)

    def add_observations(self, psi, eis_frequencies, z):
        """Bulk form of add_observation for EIS maps on a shared frequency grid: psi [B, d], z [B, Nf].  The
        containers grow once (add_observation re-allocates them per call, as the reference does)."""
        z = np.asarray(z)
        nb = len(z)
        psi = np.asarray(psi, dtype=float).reshape(nb, -1)
        if self.obs_psi is None:
            self.obs_psi = np.zeros((0, psi.shape[1]))
        if psi.shape[1] != self.obs_psi.shape[1]:
            raise ValueError(f'psi must have length {self.obs_psi.shape[1]}')
        f = np.asarray(eis_frequencies, dtype=float)
        nt = len(self.tau_supergrid)
        self.obs_psi = np.vstack([self.obs_psi, psi])
        self.obs_data.extend((None, (f, z[b])) for b in range(nb))
        self.obs_group_id.extend([None] * nb)
        self.obs_data_badness = np.concatenate([self.obs_data_badness, np.zeros(nb)])
        self.obs_ignore_flag = np.concatenate([self.obs_ignore_flag, np.zeros(nb, dtype=bool)])
        self._n += nb
        self.obs_fit_attr.extend([None] * nb)
        self.obs_fit_errors.extend([None] * nb)
        self.obs_tau_indices.extend([None] * nb)
        self.obs_fit_status = np.concatenate([self.obs_fit_status, np.zeros(nb, dtype=bool)])
        self.obs_fit_badness = np.concatenate([self.obs_fit_badness, np.zeros(nb)])
        self.obs_x = np.concatenate([self.obs_x, np.zeros((nb, *self.drt_param_shape()))], axis=0)
        self.obs_drt_var = np.concatenate([self.obs_drt_var, np.zeros((nb, *self.drt_param_shape()))], axis=0)
        self.obs_llh = np.concatenate([self.obs_llh, np.zeros(nb)])
        self.obs_rss = np.concatenate([self.obs_rss, np.zeros(nb)])
        self.obs_outer_iterations = np.concatenate([self.obs_outer_iterations, np.zeros(nb, dtype=int)])
        self.obs_status = np.concatenate([self.obs_status, np.zeros(nb, dtype=int)])
        if self.obs_special is not None:
            for key in list(self.obs_special):
                pad = np.zeros((nb,) + self.obs_special[key].shape[1:])
                self.obs_special[key] = np.concatenate([self.obs_special[key], pad], axis=0)

    def get_obs_data(self, obs_index):
        chrono_data, eis_data = self.obs_data[obs_index]
        return (chrono_data if chrono_data is not None else (None, None, None),
                eis_data if eis_data is not None else (None, None))

    # ---- fitting (drtmd.py:245-329) ------------------------------------------------------------------
    @staticmethod
    def _grid_key(chrono, eis):
        key = []
        if chrono[0] is not None:
            t, i_sig = np.asarray(chrono[0], dtype=float), np.asarray(chrono[1], dtype=float)
            key += [t.shape, t.tobytes(), i_sig.tobytes()]
        if eis[0] is not None:
            f = np.asarray(eis[0], dtype=float)
            key += [f.shape, f.tobytes()]
        return tuple(key)

    def _group(self, obs_index):
        """Observations by measurement grid.  Arrays shared by identity (bulk adds) are hashed once."""
        groups, by_id = {}, {}
        for idx in obs_index:
            chrono, eis = self.get_obs_data(idx)
            ident = (id(chrono[0]), id(chrono[1]), id(eis[0]))
            if ident not in by_id:
                by_id[ident] = self._grid_key(chrono, eis)
            groups.setdefault(by_id[ident], []).append(idx)
        return groups

    def fit_observation(self, obs_index, ignore_errors=False):
        self.fit_observations([obs_index], ignore_errors=ignore_errors, _quiet=True)

    def fit_observations(self, obs_index, print_interval=None, ignore_errors=False, shard=False, _quiet=False):
        obs_index = [int(i) for i in obs_index]
        verbose = self.print_progress and not _quiet
        if verbose:
            print(f'Found {len(obs_index)} observations to fit')
        start = time.time()
        for members in self._group(obs_index).values():
            self._fit_group(np.asarray(members), ignore_errors, shard)
        if verbose and obs_index:
            el = time.time() - start
            print('Fitted {} observations in {:.1f} minutes'.format(len(obs_index), el / 60))
            print('{:.4f} seconds per observation'.format(el / len(obs_index)))

    def fit_all(self, refit=False, print_interval=None, ignore_errors=False, shard=False):
        if refit:
            idx = np.arange(self.num_obs)
        else:
            idx = np.where(~self.obs_fit_status & ~self.obs_ignore_flag)[0]
        self.fit_observations(idx, print_interval, ignore_errors, shard=shard)

    def _fit_group(self, members, ignore_errors, shard):
        """One measurement grid, one batch (or one shard of it per rank)."""
        chrono0, eis0 = self.get_obs_data(members[0])
        rank, ws = _sharding.world() if shard else (0, 1)
        mine = _sharding.shard_indices(len(members), ws, rank, interleave=True)
        local = members[mine]
        z = v = None
        if eis0[0] is not None:
            z = np.stack([np.asarray(self.obs_data[i][1][1]) for i in local]) if len(local) else \
                np.zeros((0, len(eis0[0])), dtype=complex)
        if chrono0[0] is not None:
            v = np.stack([np.asarray(self.obs_data[i][0][2], dtype=float) for i in local]) if len(local) else \
                np.zeros((0, len(chrono0[0])))
        drt = self.drt1d
        if len(local) or ws > 1:
            # every rank builds the plan (cheap) so that special_qp_params / basis_tau agree everywhere; a rank whose
            # shard of the group is empty fits the group's first spectrum (a benign stand-in whose result is dropped)
            zz = z if (z is None or len(z)) else np.asarray(eis0[1])[None]
            vv = v if (v is None or len(v)) else np.asarray(chrono0[2], dtype=float)[None]
            if self.fit_type == 'pfrt':
                # drtmd.py:1339-1342 calls DRT._pfrt_fit_core(*chrono, *eis, **fit_kw): the factors are those of
                # fit_kw (default logspace(-1, 1, 11)); DRTMD.pfrt_factors only sizes the containers
                res = drt._pfrt_fit_core_batch(chrono0[0], chrono0[1], vv if chrono0[0] is not None else None,
                                               eis0[0], zz if eis0[0] is not None else None,
                                               diag_tau=self.tau_supergrid, **self.fit_kw)
            elif chrono0[0] is None:
                res = drt.fit_eis_batch(eis0[0], zz, diag_tau=self.tau_supergrid, **self.fit_kw)
            elif eis0[0] is None:
                res = drt.fit_chrono_batch(chrono0[0], chrono0[1], vv, diag_tau=self.tau_supergrid, **self.fit_kw)
            else:
                res = drt.fit_hybrid_batch(chrono0[0], chrono0[1], vv, eis0[0], zz, diag_tau=self.tau_supergrid,
                                           **self.fit_kw)
        if self.fit_type == 'pfrt':
            step_x = res.pfrt_result()['step_x']
            if step_x.shape[1] != len(self.pfrt_factors):
                raise ValueError(f'{step_x.shape[1]} PFRT factors were fitted but pfrt_factors has '
                                 f'{len(self.pfrt_factors)} entries')
            fp = res.extract_parameters(step_x)                 # format_1d_params, drtmd.py:1145-1158
        else:
            fp = res.fit_parameters()
        host = res.host(['status', 'n_outer'])
        left = nearest_index(self.tau_supergrid, drt.basis_tau[0])
        right = nearest_index(self.tau_supergrid, drt.basis_tau[-1]) + 1
        nloc = len(local)
        out = {'x': fp['x'][:nloc], 'status': host['status'][:nloc], 'n_outer': host['n_outer'][:nloc],
               'drt_var': res.distribution_var(extend_var=True)[:nloc], 'llh': res.evaluate_llh()[:nloc],
               'rss': res.evaluate_rss()[:nloc]}
        sp_keys = list(drt.special_qp_params.keys())
        for key in sp_keys:
            out['sp_' + key] = np.asarray(fp[key])[:nloc]
        if ws > 1:
            out = _sharding.gather_results(out, len(members), interleave=True)
        # ---- scatter into the observation arrays (drtmd.py:256-287)
        if self.obs_special is None:
            self.obs_special = {}
        for key in sp_keys:
            if key not in self.obs_special:
                self.obs_special[key] = np.zeros([self.num_obs, *self.special_param_shape(key)])
        # a singular P (ST_COV_FAIL) makes the reference's estimate_distribution_cov fail, which fit_observation reports
        # as a fit error of the observation (drtmd.py:256-287)
        bad = (out['status'] & (_engine.ST_NAN | _engine.ST_KKT_FAIL | _engine.ST_COV_FAIL)) != 0
        bad |= ~np.all(np.isfinite(out['x']).reshape(len(out['x']), -1), axis=1)
        if bad.any() and not ignore_errors:
            raise ValueError(f'Error encountered at obs_index {int(members[np.argmax(bad)])}: '
                             'Rank(A) < p or Rank([P; A; G]) < n')
        good = members[~bad]
        self.obs_x[good] = 0.0
        self.obs_x[good, ..., left:right] = out['x'][~bad]
        for key in sp_keys:       # a vector parameter of size one (v_baseline with a single step) is stored as a scalar per observation
            val = out['sp_' + key][~bad]
            self.obs_special[key][good] = val.reshape((len(val), *self.special_param_shape(key)))
        dv = out['drt_var'][~bad]                               # of the initial fit; one row per factor (drtmd.py:270)
        self.obs_drt_var[good] = dv[:, None, :] if self.fit_type == 'pfrt' else dv
        self.obs_llh[good] = out['llh'][~bad]
        self.obs_rss[good] = out['rss'][~bad]
        self.obs_fit_status[good] = True
        self.obs_outer_iterations[members] = out['n_outer']
        self.obs_status[members] = out['status']
        for i in members:
            self.obs_tau_indices[i] = (left, right)
        for i in members[bad]:
            self.obs_fit_status[i] = False
            self.obs_ignore_flag[i] = True
            self.obs_fit_errors[i] = ValueError('Rank(A) < p or Rank([P; A; G]) < n')
