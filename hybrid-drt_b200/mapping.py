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
import warnings

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
                 rss_kw=None, device=0, keep_pq=False):
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
        # keep_pq: every fit also returns its P matrix / q vector (qphb.calculate_pq), kept on the device per fit group;
        # resolve_observations / resolve_group need them (the reference restores them from obs_fit_attr, drtmd.py:366-377)
        self.keep_pq = bool(keep_pq)
        self.clear_obs()

    # ---- containers (drtmd.py:101-142, 379-430) ------------------------------------------------------
    def clear_obs(self):
        self.obs_psi = np.zeros((0, len(self.psi_dim_names))) if self.psi_dim_names is not None else None
        self.obs_data, self.obs_group_id = [], []
        self._blocks = []
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
        self._pq_store, self._pq_ref = [], {}                # fit groups' (p, q, plan scalars) on the device; obs -> (store, row)
        self.obs_scales = {}                                 # coefficient_scale, response_signal_scale, ... per observation
        self.obs_resolve_status = np.zeros(n, dtype=bool)
        self.obs_x_resolved = np.zeros((n, *self.drt_param_shape()))
        self.obs_special_resolved = None
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
        self._blocks.append((self._n, nb, f, z))        # bulk block: grouping and input stacking work on the arrays directly
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
        obs_index = np.asarray(obs_index, dtype=int)
        if self._blocks and len(obs_index) > 64:
            # members of bulk blocks (add_observations): one key per block, no per-observation Python work
            starts = np.array([b[0] for b in self._blocks])
            ends = starts + np.array([b[1] for b in self._blocks])
            blk = np.searchsorted(starts, obs_index, side='right') - 1
            inside = (blk >= 0) & (obs_index < ends[np.maximum(blk, 0)])
            for bi in np.unique(blk[inside]):
                f = self._blocks[bi][2]
                key = self._grid_key((None, None, None), (f, None))
                groups.setdefault(key, []).extend(obs_index[inside & (blk == bi)].tolist())
            obs_index = obs_index[~inside]
        for idx in obs_index:
            idx = int(idx)
            chrono, eis = self.get_obs_data(idx)
            ident = (id(chrono[0]), id(chrono[1]), id(eis[0]))
            if ident not in by_id:
                by_id[ident] = self._grid_key(chrono, eis)
            groups.setdefault(by_id[ident], []).append(idx)
        return groups

    def _block_z(self, local):
        """z [len(local), Nf] straight from the bulk block all of `local` belongs to (None if they do not share one)."""
        if not self._blocks or len(local) == 0:
            return None
        for start, nb, _, z in self._blocks:
            if local[0] >= start and local[0] < start + nb:
                if local.min() >= start and local.max() < start + nb:
                    return np.asarray(z)[local - start]
                return None
        return None

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
            z = self._block_z(local)
            if z is None:
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
                res = drt.fit_eis_batch(eis0[0], zz, diag_tau=self.tau_supergrid, want_pq=self.keep_pq, **self.fit_kw)
            elif eis0[0] is None:
                res = drt.fit_chrono_batch(chrono0[0], chrono0[1], vv, diag_tau=self.tau_supergrid, want_pq=self.keep_pq, **self.fit_kw)
            else:
                res = drt.fit_hybrid_batch(chrono0[0], chrono0[1], vv, eis0[0], zz, diag_tau=self.tau_supergrid,
                                           want_pq=self.keep_pq, **self.fit_kw)
        if self.fit_type == 'pfrt':
            step_x = res.pfrt_result()['step_x']
            if step_x.shape[1] != len(self.pfrt_factors):
                raise ValueError(f'{step_x.shape[1]} PFRT factors were fitted but pfrt_factors has '
                                 f'{len(self.pfrt_factors)} entries')
            fp = res.extract_parameters(step_x)                 # format_1d_params, drtmd.py:1145-1158
        else:
            fp = res.extract_parameters(res.host(['x'])['x'])      # the map keeps coefficients, not per-point sigmas: no weights copy
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
        if self.keep_pq and self.fit_type != 'pfrt':
            if ws > 1:
                raise NotImplementedError('hybdrt_b200: keep_pq with shard=True (the P matrices stay on the rank that fitted them)')
            self._ensure_len()
            pl = res.plan
            self._pq_store.append(dict(p=res.raw['p_matrix'], q=res.raw['q_vector'], special=dict(drt.special_qp_params),
                                       v_baseline_scale=pl.get('v_baseline_scale'), inductance_scale=pl['inductance_scale'],
                                       capacitance_scale=pl['capacitance_scale']))
            for row, i in enumerate(members):
                self._pq_ref[int(i)] = (len(self._pq_store) - 1, row)
            for key in ('coefficient_scale', 'response_signal_scale', 'scaled_response_offset'):
                if key in res.scales:
                    self.obs_scales.setdefault(key, np.zeros(self._n))[members] = np.asarray(res.scales[key])[:nloc]
        if ws > 1:
            out = _sharding.gather_results(out, len(members), interleave=True, copy=False)   # scattered into obs_* right below
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
        # the usual case -- every fit of the group fine, the group a run of consecutive observations (a whole map is one
        # group) -- scatters with slices: block copies instead of index arrays over the full map on every rank
        m0, nm = int(members[0]), len(members)
        run = int(members[-1]) - m0 + 1 == nm and bool(np.all(np.diff(members) == 1))
        clean = not bad.any()
        sel = slice(None) if clean else ~bad
        good = (slice(m0, m0 + nm) if run else members) if clean else members[~bad]
        allm = slice(m0, m0 + nm) if run else members
        self.obs_x[good] = 0.0
        self.obs_x[good, ..., left:right] = out['x'][sel]
        for key in sp_keys:       # a vector parameter of size one (v_baseline with a single step) is stored as a scalar per observation
            val = out['sp_' + key][sel]
            self.obs_special[key][good] = val.reshape((len(val), *self.special_param_shape(key)))
        dv = out['drt_var'][sel]                                # of the initial fit; one row per factor (drtmd.py:270)
        self.obs_drt_var[good] = dv[:, None, :] if self.fit_type == 'pfrt' else dv
        self.obs_llh[good] = out['llh'][sel]
        self.obs_rss[good] = out['rss'][sel]
        self.obs_fit_status[good] = True
        self.obs_outer_iterations[allm] = out['n_outer']
        self.obs_status[allm] = out['status']
        if run:
            self.obs_tau_indices[m0:m0 + nm] = [(left, right)] * nm
        else:
            for i in members:
                self.obs_tau_indices[i] = (left, right)
        for i in members[bad]:
            self.obs_fit_status[i] = False
            self.obs_ignore_flag[i] = True
            self.obs_fit_errors[i] = ValueError('Rank(A) < p or Rank([P; A; G]) < n')

    # ---- cross-observation resolve (drtmd.py:432-560, mapping/resolve.py) ----------------------------------------
    def _ensure_len(self):
        """The resolve containers and per-observation scales follow num_obs (they are allocated lazily)."""
        n = self._n
        if len(self.obs_x_resolved) != n:
            keep = min(len(self.obs_x_resolved), n)
            new = np.zeros((n, *self.drt_param_shape()))
            new[:keep] = self.obs_x_resolved[:keep]
            self.obs_x_resolved = new
            st = np.zeros(n, dtype=bool)
            st[:keep] = self.obs_resolve_status[:keep]
            self.obs_resolve_status = st
        for key, arr in list(self.obs_scales.items()):
            if len(arr) != n:
                new = np.zeros(n)
                new[:min(len(arr), n)] = arr[:n]
                self.obs_scales[key] = new
        if self.obs_special is not None:
            if self.obs_special_resolved is None:
                self.obs_special_resolved = {}
            for key, arr in self.obs_special.items():                  # drtmd.py:1160-1166
                cur = self.obs_special_resolved.get(key)
                if cur is None or len(cur) != n:
                    new = np.zeros_like(arr)
                    if cur is not None:
                        new[:min(len(cur), n)] = cur[:n]
                    self.obs_special_resolved[key] = new

    def get_group_index(self, group_id):
        """drtmd.py:1194-1210 (unsorted)"""
        gids = np.array(self.obs_group_id, dtype=object)
        if isinstance(group_id, str) or group_id is None:
            return np.where(np.array([g == group_id for g in gids]))[0]
        return np.where(np.array([g in group_id for g in gids]))[0]

    def _sorted(self, obs_index, psi_sort_dims, psi_distance_dims=None):
        dims = psi_sort_dims if psi_sort_dims is not None else psi_distance_dims
        if dims is None:
            return obs_index
        vals = [self.obs_psi[obs_index, self.psi_dim_names.index(d)] for d in dims][::-1]
        return obs_index[np.lexsort(vals)]

    def _resolve_windows(self, windows, truncate, sigma, lambda_psi):
        """The QPs of `windows` (lists of nr observation indices each) in one launch.  Returns, per window,
        (x_drt [nr, n_tau], x_special dict, tau_indices), unpacked as resolve.unpack_resolved_x does."""
        from scipy.ndimage import gaussian_filter1d, median_filter
        if truncate:
            raise NotImplementedError('hybdrt_b200: resolve with truncate=True')
        if self.fit_dop:
            raise NotImplementedError('hybdrt_b200: resolve of DRT + DOP fits')
        eng = self.drt1d.engine
        nr = len(windows[0])
        flat = np.concatenate(windows)
        missing = [int(i) for i in flat if int(i) not in self._pq_ref]
        if missing:
            raise ValueError(f'observation {missing[0]} has no stored P matrix: construct DRTMD(keep_pq=True) before fitting')
        st0 = self._pq_store[self._pq_ref[int(flat[0])][0]]
        sp = st0['special']
        removed = [k for k in ('v_baseline', 'vz_offset') if k in sp]
        if removed != ['v_baseline', 'vz_offset'] or sp['v_baseline']['index'] != 0:
            # resolve.get_offset_pq indexes both keys (resolve.py:24-25): the reference resolves hybrid fits only
            raise KeyError('vz_offset' if 'v_baseline' in sp else 'v_baseline')
        k_rm = sum(sp[k].get('size', 1) for k in removed)
        special_dict = {k: dict(v, index=v['index'] - k_rm) for k, v in sp.items() if k not in removed}     # resolve.py:138-164
        so = sum(v.get('size', 1) for v in special_dict.values())
        # ---- per (window, observation): trimmed P / q, resized to the window's tau range (resolve.py:11-135)
        import torch
        p_rows, q_rows, xrem = [], [], []
        tau_win = []
        for w in windows:
            lo = min(self.obs_tau_indices[i][0] for i in w)
            hi = max(self.obs_tau_indices[i][1] for i in w)
            tau_win.append((lo, hi))
        nc = so + max(hi - lo for lo, hi in tau_win)
        if any(so + hi - lo != nc for lo, hi in tau_win):
            raise NotImplementedError('hybdrt_b200: resolve windows with different tau ranges in one call')
        cs = self.obs_scales['coefficient_scale']
        rs_, ro_ = self.obs_scales['response_signal_scale'], self.obs_scales['scaled_response_offset']
        p_res = torch.zeros(len(flat), nc, nc, dtype=torch.float64, device=eng.device)
        q_res = torch.zeros(len(flat), nc, dtype=torch.float64, device=eng.device)
        for pos, (w, (lo, hi)) in enumerate(zip(windows, tau_win)):
            for r, i in enumerate(w):
                sid, row = self._pq_ref[int(i)]
                st = self._pq_store[sid]
                vb = np.atleast_1d(np.asarray(self.obs_special['v_baseline'][i], dtype=float)) / rs_[i]
                vb[0] += ro_[i]
                vb = vb * np.atleast_1d(st['v_baseline_scale'])
                x_rm = eng.dev(np.concatenate([vb, [float(self.obs_special['vz_offset'][i])]]))
                p_full, q_full = st['p'][row], st['q'][row]
                p_t, q_t = p_full[k_rm:, k_rm:], q_full[k_rm:] + x_rm @ p_full[:k_rm, k_rm:]
                tl, tr_ = self.obs_tau_indices[i]
                a, b = so + (tl - lo), nc + (tr_ - hi)                      # expand (resolve.py:84-100)
                dst = pos * nr + r
                p_res[dst, :so, :so], q_res[dst, :so] = p_t[:so, :so], q_t[:so]
                p_res[dst, a:b, a:b], q_res[dst, a:b] = p_t[so:, so:], q_t[so:]
                p_res[dst, a:b, :so], p_res[dst, :so, a:b] = p_t[so:, :so], p_t[:so, so:]
        # ---- coupling and parameter scales per window (resolve.py:223-273)
        my = np.zeros((len(windows), nr, nr))
        pscale = np.ones((len(windows), nc))
        ly = gaussian_filter1d(np.eye(nr), sigma=sigma, mode='reflect', order=2)
        for pos, w in enumerate(windows):
            scale_vec = cs[w]
            smooth = gaussian_filter1d(median_filter(scale_vec, 3), 2)
            lys = ly @ np.diag(scale_vec / smooth)
            my[pos] = (lys.T @ lys) * lambda_psi
            if 'R_inf' in special_dict:
                x_inf = self.obs_special['R_inf'][w] / scale_vec
                pscale[pos, special_dict['R_inf']['index']] = (5 * np.std(x_inf)) ** -2
        h = np.zeros(nc) if self.fit_kw['nonneg'] else 10.0 * np.ones(nc)
        for v in special_dict.values():
            if v['nonneg']:
                h[v['index']:v['index'] + v.get('size', 1)] = 0.0
        out = eng.resolve_qp_batch(p_res, q_res, np.arange(len(windows)) * nr, my, pscale, h, nr)
        x_all = out['x'].cpu().numpy()
        self.last_resolve = dict(iters=out['iters'].cpu().numpy(), status=out['status'].cpu().numpy())
        results = []
        for pos, (w, tw) in enumerate(zip(windows, tau_win)):
            x = x_all[pos]
            scale_vec = cs[w]
            x_special = {}
            for key, info in special_dict.items():                       # resolve.unpack_resolved_x :344-375
                xk = x[:, info['index']:info['index'] + info.get('size', 1)] * scale_vec[:, None]
                if key == 'inductance':
                    xk = xk * self._pq_store[self._pq_ref[int(w[0])][0]]['inductance_scale']
                elif key == 'C_inv':
                    xk = xk * self._pq_store[self._pq_ref[int(w[0])][0]]['capacitance_scale']
                x_special[key] = xk.flatten() if info.get('size', 1) == 1 else xk
            results.append((x[:, so:] * scale_vec[:, None], x_special, tw))
        return results

    def _insert_resolved(self, obs_index, x_drt, x_special, tau_indices):
        self.obs_x_resolved[obs_index, tau_indices[0]:tau_indices[1]] = x_drt
        for key, val in x_special.items():
            self.obs_special_resolved[key][obs_index] = val
        self.obs_resolve_status[obs_index] = True

    def resolve_observations(self, obs_index, psi_sort_dims=None, psi_distance_dims=None, truncate=False, sigma=1,
                             lambda_psi=1, tau_filter_sigma=0, special_filter_sigma=0):
        """drtmd.py:432-484"""
        if tau_filter_sigma or special_filter_sigma:
            raise NotImplementedError('hybdrt_b200: resolve with tau_filter_sigma / special_filter_sigma')
        self._ensure_len()
        obs_index = np.asarray(obs_index)
        obs_index = obs_index[self.obs_fit_status[obs_index] & ~self.obs_ignore_flag[obs_index]]
        if psi_sort_dims is not None:
            obs_index = self._sorted(obs_index, psi_sort_dims)
        if len(obs_index) == 1:
            warnings.warn('Only one observation included in resolution group; raw parameters will be copied')
            ti = self.obs_tau_indices[obs_index[0]]
            self._insert_resolved(obs_index, self.obs_x[obs_index, ti[0]:ti[1]],
                                  {k: v[obs_index] for k, v in self.obs_special.items()}, ti)
        elif len(obs_index) > 1:
            (x_drt, x_special, ti), = self._resolve_windows([obs_index], truncate, sigma, lambda_psi)
            self._insert_resolved(obs_index, x_drt, x_special, ti)
        else:
            warnings.warn('No valid observations included in resolution group')

    def resolve_group(self, group_id, batch_size=7, overlap=2, psi_sort_dims=None, psi_distance_dims=None,
                      truncate=False, sigma=1, lambda_psi=1, tau_filter_sigma=0, special_filter_sigma=0):
        """drtmd.py:486-560: windows of batch_size observations with `overlap` shared neighbours, every window one QP
        (all of them in ONE launch here), the overlapping solutions averaged with margin weights."""
        if tau_filter_sigma or special_filter_sigma:
            raise NotImplementedError('hybdrt_b200: resolve with tau_filter_sigma / special_filter_sigma')
        self._ensure_len()
        obs_index = self.get_group_index(group_id)
        obs_index = obs_index[self.obs_fit_status[obs_index] & ~self.obs_ignore_flag[obs_index]]
        obs_index = self._sorted(obs_index, psi_sort_dims, psi_distance_dims)
        self.obs_x_resolved[obs_index] = 0
        num_obs = len(obs_index)
        if num_obs < 2:
            return self.resolve_observations(obs_index, psi_sort_dims, psi_distance_dims, truncate, sigma, lambda_psi)
        batch_size = min(batch_size, num_obs)
        stride = max(batch_size - overlap, 1)
        spans = []
        for start in range(0, num_obs, stride):
            if num_obs - start < batch_size:
                start = max(0, num_obs - batch_size)                       # a full batch for the last one
            spans.append((start, start + batch_size))
            if start + batch_size >= num_obs:
                break
        results = self._resolve_windows([obs_index[a:b] for a, b in spans], truncate, sigma, lambda_psi)
        nb = len(spans)
        x_batch = np.zeros((nb, *self.obs_x_resolved[obs_index].shape))
        x_special = {k: np.zeros((nb, *v[obs_index].shape)) for k, v in self.obs_special_resolved.items()}
        margins = np.full((nb, num_obs), -1.0)
        for i, ((a, b), (x_drt, xs, ti)) in enumerate(zip(spans, results)):
            self._insert_resolved(obs_index[a:b], x_drt, xs, ti)
            x_batch[i, a:b] = self.obs_x_resolved[obs_index[a:b]]
            for key in self.obs_special_resolved:
                x_special[key][i, a:b] = self.obs_special_resolved[key][obs_index[a:b]]
            margins[i, a:b] = np.minimum(np.arange(batch_size), np.arange(batch_size)[::-1])
        if overlap > 0 and num_obs > 1:
            wts = margins + 0.1
            wts[wts < 0] = 0
            xw = np.moveaxis(np.tile(wts, (x_batch.shape[-1], 1, 1)), 0, -1)
            self.obs_x_resolved[obs_index] = np.average(x_batch, axis=0, weights=xw)
            for key, val in x_special.items():
                kw_ = np.moveaxis(np.tile(wts, (val.shape[-1], 1, 1)), 0, -1) if np.ndim(val) > 2 else wts
                self.obs_special_resolved[key][obs_index] = np.average(val, axis=0, weights=kw_)
