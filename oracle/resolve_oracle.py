"""CPU ORACLE (test infrastructure, not product code): numpy restatement of the cross-observation resolve QP of
hybdrt/mapping/resolve.py (get_offset_pq :11-62, resize_pq :65-135, resolve_observations :176-341, unpack_resolved_x
:344-375) for the default options (tau_filter_sigma = special_filter_sigma = 0, no DOP block).

Pinned by tests/golden/drtmd_resolve.npz, which the unmodified reference produced (oracle/make_golden_map.py); only
tests/ may import this module."""
import numpy as np
from scipy.ndimage import gaussian_filter1d, median_filter

from .coneqp import coneqp_orthant


def offset_pq(p, q, x_remove):
    """resolve.get_offset_pq: drop the data-dependent leading parameters (v_baseline, vz_offset), folding their fitted
    values into q.  x_remove: their values in the scaled space (resolve.py:44-52)."""
    k = len(x_remove)
    return p[k:, k:], q[k:] + x_remove @ p[:k, k:]


def scaled_v_baseline(v_baseline, response_signal_scale, scaled_response_offset, v_baseline_scale):
    """resolve.py:44-49: raw -> scaled space (the offset applies to the zero-degree coefficient only)."""
    s = np.array(v_baseline, dtype=float) / response_signal_scale
    s[0] += scaled_response_offset
    return s * v_baseline_scale


def resize_pq(p, q, special_offset, tau_indices, match):
    """resolve.resize_pq, the expand case (resolve.py:84-100): embed the DRT block into the common tau window."""
    num_drt, match_num = tau_indices[1] - tau_indices[0], match[1] - match[0]
    new = p.shape[0] + (match_num - num_drt)
    lo, ro = tau_indices[0] - match[0], tau_indices[1] - match[1]
    if lo < 0 or ro > 0:
        raise NotImplementedError('truncate branches of resize_pq')
    po, qo = np.zeros((new, new)), np.zeros(new)
    so = special_offset
    po[:so, :so], qo[:so] = p[:so, :so], q[:so]
    left, right = so + lo, new + ro
    po[left:right, left:right], qo[left:right] = p[so:, so:], q[so:]
    po[left:right, :so], po[:so, left:right] = p[so:, :so], p[:so, so:]
    return po, qo


def coupling(scale_vec, sigma=1.0):
    """My = (Ly S)^T (Ly S), Ly = second derivative of a Gaussian along the observation axis with reflected edges,
    S = coefficient scales over their smoothed trend (resolve.py:223-273)."""
    nr = len(scale_vec)
    ly = gaussian_filter1d(np.eye(nr), sigma=sigma, mode='reflect', order=2)
    smooth = gaussian_filter1d(median_filter(np.asarray(scale_vec, dtype=float), 3), 2)
    lys = ly @ np.diag(scale_vec / smooth)
    return lys.T @ lys


def resolve_window(p_list, q_list, scale_vec, x_inf_scaled, r_inf_index, nonneg, special_nonneg_index, sigma=1.0,
                   lambda_psi=1.0):
    """The QP of one window (resolve.py:242-334): blockdiag(P_i) + lambda My (x) diag(param_scale), G = -I."""
    nr, nc = len(p_list), len(q_list[0])
    my = coupling(scale_vec, sigma)
    param_scale = np.ones(nc)
    if r_inf_index is not None:
        param_scale[r_inf_index] = (5 * np.std(x_inf_scaled)) ** -2                       # resolve.py:239-242
    big = np.zeros((nr * nc, nr * nc))
    for i in range(nr):
        big[i * nc:(i + 1) * nc, i * nc:(i + 1) * nc] = p_list[i]
    for i in range(nr):
        for j in range(nr):
            big[i * nc:(i + 1) * nc, j * nc:(j + 1) * nc] += np.diag(param_scale * my[i, j]) * lambda_psi
    h = np.zeros(nr * nc) if nonneg else 10.0 * np.ones(nr * nc)
    for idx in special_nonneg_index:
        h[idx::nc] = 0.0
    res = coneqp_orthant(big.T, np.concatenate(q_list), h)
    return np.asarray(res['x']).reshape(nr, nc), res['iterations'], dict(my=my, param_scale=param_scale)
