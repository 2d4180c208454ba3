"""CPU ORACLE (test infrastructure, not product code): chrono conditioning ahead of fit_chrono / fit_hybrid.

numpy restatement of hybdrt/preprocessing.py downsample_data (:335-468) with its helpers get_decimation_index
(:620-689), select_decimation_interval (:603-617), sigma_from_decimate_index (:577-591), filter_chrono_signal
(:507-574) and of filters/_filters.py nonuniform_gaussian_filter1d (:261-341), whose inner
scipy.ndimage.gaussian_filter1d (mode='reflect', order 0) is written out explicitly here.
Pinned by tests/golden/downsample.npz (outputs of the unmodified reference).
"""
import numpy as np


def identify_steps(y, allow_consecutive=True, rthresh=50, athresh=1e-10):
    """preprocessing.py:17-38"""
    dy = np.diff(y)
    idx = np.where((np.abs(dy) >= np.median(np.abs(dy)) * rthresh) & (np.abs(dy) >= athresh))[0] + 1
    if not allow_consecutive:
        idx = idx[np.concatenate(([2], np.diff(idx))) > 1]
    return idx


def step_indices_from_times(times, step_times):
    """preprocessing.py:161-178: first sample at or after each step time."""
    return np.array([int(np.argmin(np.where(times >= st, times - st, np.inf))) for st in step_times])


def split_steps(x, step_index):
    """preprocessing.py:41-54"""
    si = np.array(step_index)
    if si[0] > 0:
        si = np.insert(si, 0, 0)
    if si[-1] < len(x):
        si = np.append(si, len(x))
    return [x[a:b] for a, b in zip(si[:-1], si[1:])]


def decimation_index(times, step_times, t_sample, prestep_points, interval, factor, max_t_sample):
    """preprocessing.get_decimation_index, preprocessing.py:620-689."""
    n_pre = int(np.sum(times < np.min(step_times)))
    keep = [np.linspace(0, n_pre - 1, prestep_points).round(0).astype(int)]
    step_index = step_indices_from_times(times, step_times)
    max_interval = np.inf if max_t_sample is None else int(max_t_sample / t_sample)
    for i, start in enumerate(step_index):
        nxt = len(times) if start == step_index[-1] else step_index[i + 1]
        undec = np.arange(start, min(start + interval + 1, nxt), dtype=int)
        keep.append(undec)
        last, j = undec[-1], 1
        while last < nxt - 1:
            si = min(int(factor ** j), max_interval)
            end = nxt if si == max_interval else min(last + interval * si + 1, nxt)
            k = np.arange(last + si, end, si, dtype=int)
            if len(k) == 0:
                k = np.array([end - 1])
            if end == nxt and k[-1] < nxt - 1:
                k = np.append(k, nxt - 1)
            keep.append(k)
            last = k[-1]
            j += 1
    return np.unique(np.concatenate(keep))


def select_interval(times, step_times, t_sample, prestep_points, factor, max_t_sample, target_size):
    """preprocessing.select_decimation_interval, preprocessing.py:603-617."""
    intervals = np.logspace(np.log10(2), np.log10(1000), 12).astype(int)
    sizes = [len(decimation_index(times, step_times, t_sample, prestep_points, iv, factor, max_t_sample))
             for iv in intervals]
    return int(np.interp(target_size, sizes, intervals))


def sigma_from_decimate_index(n, dec_index, truncate=4.0):
    """preprocessing.py:577-591: the filter reaches halfway to the nearest kept sample."""
    sig = np.zeros(n)
    d = np.diff(dec_index)
    md = np.minimum(np.insert(d, 0, d[0]), np.append(d, d[-1]))
    sd = md / (2 * truncate)
    sd[md < 2] = 0
    sig[dec_index] = sd
    return sig


def gaussian_filter1d_reflect(a, sigma, truncate=4.0):
    """scipy.ndimage.gaussian_filter1d(a, sigma, mode='reflect', order=0): weights exp(-k^2 / 2 sigma^2) on
    k = -lw..lw with lw = int(truncate sigma + 0.5), normalised; the signal is mirrored about its edges
    (d c b a | a b c d | d c b a), repeatedly if the window is longer than the signal."""
    n = len(a)
    lw = int(truncate * float(sigma) + 0.5)
    k = np.arange(-lw, lw + 1)
    w = np.exp(-0.5 / (sigma * sigma) * k ** 2)
    w = w / w.sum()
    pos = (np.arange(n)[:, None] + k[None, :]) % (2 * n)
    pos = np.where(pos >= n, 2 * n - 1 - pos, pos)
    return (a[pos] * w[None, :]).sum(axis=1)


def sigma_nodes_for(sigma, node_factor=1.5, min_sigma=0.25):
    """The log-spaced filter widths of nonuniform_gaussian_filter1d (filters/_filters.py:264-295).
    Returns (clipped sigma, nodes, node_delta)."""
    sigma = np.maximum(sigma, 1e-8)
    lo = max(np.min(np.log10(sigma)), np.log10(min_sigma))
    hi = max(np.max(np.log10(sigma)), np.log10(min_sigma))
    num = int(np.ceil((hi - lo) / np.log10(node_factor))) + 1
    nodes = np.logspace(lo, hi, num)
    if np.min(sigma) < min_sigma:
        factor = nodes[-1] / nodes[-2] if len(nodes) > 1 else node_factor
        sigma = sigma.copy()
        sigma[sigma < min_sigma / factor ** 2] = min_sigma / factor ** 2
        while nodes[0] > np.min(sigma) * 1.001:
            nodes = np.insert(nodes, 0, nodes[0] / factor)
    delta = np.log(nodes[-1] / nodes[-2]) if len(nodes) > 1 else 1
    return sigma, nodes, delta


def nonuniform_gaussian_filter1d(a, sigma, truncate=4, node_factor=1.5, min_sigma=0.25):
    """filters/_filters.py:261-341 (empty=False, mode='reflect', order=0): Gaussian filters at log-spaced widths,
    blended per sample with hat weights in ln(sigma); widths below min_sigma pass the signal through."""
    sigma = np.array(sigma, dtype=float)
    if not np.max(sigma) > 0:
        return a
    sigma, nodes, delta = sigma_nodes_for(sigma, node_factor, min_sigma)
    outs = np.empty((len(nodes), len(a)))
    for i, sn in enumerate(nodes):
        outs[i] = a if sn < min_sigma else gaussian_filter1d_reflect(a, sn, truncate)
    nw = np.abs(np.log(sigma[None, :] / nodes[:, None])) / delta
    nw = 1 - np.minimum(nw, 1)
    return np.sum(outs * nw, axis=0)


def filter_chrono_signal(times, y, step_index, decimate_index=None, sigma_factor=0.01, max_sigma=None):
    """preprocessing.filter_chrono_signal (no outlier removal, no median prefilter), preprocessing.py:507-574."""
    t_sample = np.median(np.diff(times))
    if max_sigma is None:
        max_sigma = sigma_factor / t_sample
    dec_sig = None
    if decimate_index is not None:
        dec_sig = split_steps(sigma_from_decimate_index(len(y), decimate_index), step_index)
    out = []
    for i, (ts, ys) in enumerate(zip(split_steps(times, step_index), split_steps(y, step_index))):
        sig = sigma_factor * (np.exp(1) * (ts - (ts[0] - t_sample)) / 2 / t_sample)
        sig[sig > max_sigma] = max_sigma
        if dec_sig is not None:
            sig = np.minimum(dec_sig[i], sig)
        out.append(nonuniform_gaussian_filter1d(ys, sig))
    return np.concatenate(out)


def downsample_data(times, i_signal, v_signal, target_size=None, step_times=None, method='decimate',
                    decimation_interval=10, decimation_factor=2, decimation_max_period=None, antialiased=True,
                    prestep_samples=20, filter_kw=None):
    """preprocessing.downsample_data, galvanostatic, ideal steps, stepwise sample times, preprocessing.py:335-468
    ('decimate', and 'match' without target times = keep everything)."""
    if step_times is None:
        step_times = times[identify_steps(i_signal, True)]
    if method == 'decimate':
        t_sample = np.min(np.diff(times))
        if target_size is not None:
            decimation_interval = select_interval(times, step_times, t_sample, prestep_samples, decimation_factor,
                                                  decimation_max_period, target_size)
        idx = decimation_index(times, step_times, t_sample, prestep_samples, decimation_interval,
                               decimation_factor, decimation_max_period)
    else:
        idx = np.arange(len(times))
    if antialiased:
        si = identify_steps(i_signal, allow_consecutive=False)
        i_signal = filter_chrono_signal(times, i_signal, si, decimate_index=idx, **(filter_kw or {}))
        v_signal = filter_chrono_signal(times, v_signal, si, decimate_index=idx, **(filter_kw or {}))
    return times[idx], i_signal[idx], v_signal[idx], idx
