"""Host-side statistics of the Kramers-Kronig test: mirror of hybdrt/models/kk.py (same function names and
arguments).  The fit behind the test runs in the CUDA engine (DRT.kk_fit); these functions only look at its
residuals."""
import numpy as np


def normalize_residuals(z_meas, z_pred, norm='modulus'):
    """models/kk.py:9-19: residuals in % of |Z| ('modulus') or divided by ``norm``."""
    z_err = z_meas - z_pred
    return 100 * z_err / np.abs(z_meas) if isinstance(norm, str) and norm == 'modulus' else z_err / norm


def std_normal_quantile(quantile):
    """utils/stats.py:108-116 (interpolated inverse of the standard normal cdf)."""
    from scipy.stats import norm
    s_interp = np.linspace(0, 14, 2000)
    q = np.asarray(quantile, dtype=float)
    return np.interp(np.abs(q - 0.5) + 0.5, norm.cdf(s_interp), s_interp) * np.sign(q - 0.5)


def robust_std(x, sample_fraction=0.5):
    """utils/stats.py:124-134: standard deviation from an inter-quantile range."""
    if sample_fraction > 1:
        raise ValueError('sample_fraction must be no greater than 1')
    q_lo = np.percentile(x, 50 - 100 * sample_fraction / 2)
    q_hi = np.percentile(x, 50 + 100 * sample_fraction / 2)
    return (q_hi - q_lo) / (2 * std_normal_quantile(0.5 + sample_fraction / 2))


def get_outliers(z_err_norm, n_iter=2, p_thresh=1e-4, n_sigma=None, std_sample_fraction=0.6):
    """models/kk.py:21-53: the squared error modulus against a chi-squared law (2 degrees of freedom) whose scale is
    a robust standard deviation, re-estimated without the points flagged so far."""
    from scipy.stats import chi2
    z_err_norm = np.asarray(z_err_norm)
    mask = np.zeros(len(z_err_norm), dtype=bool)
    for _ in range(n_iter):
        kept = z_err_norm[~mask]
        std = robust_std(np.concatenate([kept.real, kept.imag]), sample_fraction=std_sample_fraction)
        if n_sigma is None:
            mask = (1 - chi2.cdf(np.abs(z_err_norm) ** 2, 2, loc=0, scale=std ** 2)) < p_thresh
        else:
            mask = np.abs(z_err_norm) > std * n_sigma
    return np.where(mask)[0]


def get_limits(f_fit, outlier_index, max_num_outliers=2, return_index=False):
    """models/kk.py:56-123: the widest frequency window whose ends are clean points with a clean neighbour and that
    holds at most ``max_num_outliers`` flagged points."""
    f_fit = np.asarray(f_fit, dtype=float)
    order = np.argsort(f_fit)[::-1]
    f_sorted = f_fit[order]
    pos = {int(i): k for k, i in enumerate(order)}
    is_out = np.zeros(len(f_fit))
    is_out[[pos[int(i)] for i in outlier_index]] = 1
    padded = np.concatenate(([is_out[0]], is_out, [is_out[-1]]))       # ndimage.uniform_filter1d, size 3, 'reflect'
    badness = (padded[:-2] + padded[1:-1] + padded[2:]) / 3
    clean = np.where(badness == 0)[0]
    i_left, i_right = clean[0], clean[-1]
    n_bad = np.sum(is_out[i_left:i_right])
    if n_bad > max_num_outliers:
        need = n_bad - max_num_outliers
        from_left = np.cumsum(is_out[i_left:i_right + 1])
        from_right = np.cumsum(is_out[i_left:i_right + 1][::-1])
        ll, rr = np.meshgrid(from_left, from_right)
        index = np.argwhere(ll + rr >= need)
        r, l = index[np.argmin(np.sum(index, axis=1))]
        i_left, i_right = i_left + l, i_right - r
    if is_out[i_left] == 1:
        i_left = np.min(clean[clean >= i_left])
    if is_out[i_right] == 1:
        i_right = np.max(clean[clean <= i_right])
    f_max, f_min = f_sorted[i_left], f_sorted[i_right]
    return ((f_min, f_max), (i_left, i_right)) if return_index else (f_min, f_max)


def trim_data(frequencies, z, f_min, f_max):
    """models/kk.py:125-128"""
    mask = (frequencies <= f_max) & (frequencies >= f_min)
    return frequencies[mask], z[mask]
