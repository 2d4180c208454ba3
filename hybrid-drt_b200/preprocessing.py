"""Chrono conditioning ahead of fit_chrono / fit_hybrid: mirror of hybdrt.preprocessing.downsample_data
(reference preprocessing.py:335-468) for whole batches of traces recorded on one time grid.

Split of work: everything that depends on the time grid alone -- the decimation index (preprocessing.py:603-689)
and the filter taps of every kept sample (filter_chrono_signal :507-574, sigma_from_decimate_index :577-591,
filters.nonuniform_gaussian_filter1d filters/_filters.py:261-341) -- is laid out once on the host; the
filtering itself runs on the GPU for all traces at once (csrc/chrono_kernels.cu, hdrt_filter_gather).
There is no CPU filtering path.
"""
import warnings

import numpy as np

from . import engine as _engine


def identify_steps(y, allow_consecutive=True, rthresh=50, athresh=1e-10):
    """preprocessing.identify_steps, preprocessing.py:17-38"""
    dy = np.abs(np.diff(y))
    idx = np.flatnonzero((dy >= np.median(dy) * rthresh) & (dy >= athresh)) + 1
    if not allow_consecutive and len(idx) > 1:
        idx = idx[np.concatenate(([True], np.diff(idx) > 1))]
    return idx


def step_indices_from_times(times, step_times):
    """preprocessing.get_step_indices_from_step_times, preprocessing.py:161-178 (times ascending)."""
    return np.minimum(np.searchsorted(times, np.asarray(step_times, dtype=float), side='left'), len(times) - 1)


def get_decimation_index(times, step_times, t_sample, prestep_points, decimation_interval, decimation_factor,
                         max_t_sample):
    """preprocessing.get_decimation_index, preprocessing.py:620-689: after each step keep ``decimation_interval``
    consecutive samples, then the same number at every stride decimation_factor ** j (capped by max_t_sample)."""
    n_pre = int(np.searchsorted(times, np.min(step_times), side='left'))
    pieces = [np.linspace(0, n_pre - 1, prestep_points).round(0).astype(int)]
    starts = step_indices_from_times(times, step_times)
    cap = np.inf if max_t_sample is None else int(max_t_sample / t_sample)
    for s, start in enumerate(starts):
        stop = len(times) if start == starts[-1] else starts[s + 1]
        head = np.arange(start, min(start + decimation_interval + 1, stop))
        pieces.append(head)
        last, j = head[-1], 1
        while last < stop - 1:
            stride = min(int(decimation_factor ** j), cap)
            end = stop if stride == cap else min(last + decimation_interval * stride + 1, stop)
            run = np.arange(last + stride, end, stride)
            if run.size == 0:
                run = np.array([end - 1])
            if end == stop and run[-1] < stop - 1:
                run = np.append(run, stop - 1)
            pieces.append(run)
            last = run[-1]
            j += 1
    return np.unique(np.concatenate(pieces)).astype(int)


def select_decimation_interval(times, step_times, t_sample, prestep_points, decimation_factor, max_t_sample,
                               target_size):
    """preprocessing.select_decimation_interval, preprocessing.py:603-617"""
    intervals = np.logspace(np.log10(2), np.log10(1000), 12).astype(int)
    sizes = [len(get_decimation_index(times, step_times, t_sample, prestep_points, iv, decimation_factor,
                                      max_t_sample)) for iv in intervals]
    if target_size > sizes[-1]:
        warnings.warn(f'Cannot achieve target size of {target_size} with selected decimation factor of '
                      f'{decimation_factor}. Decrease the decimation factor and/or decrease the maximum period')
    if target_size < sizes[0]:
        warnings.warn(f'Cannot achieve target size of {target_size} with selected decimation factor of '
                      f'{decimation_factor}. Increase the decimation factor and/or increase the maximum period')
    return int(np.interp(target_size, sizes, intervals))


def _gauss_taps(sigma, truncate):
    """scipy.ndimage._filters._gaussian_kernel1d, order 0."""
    lw = int(truncate * float(sigma) + 0.5)
    k = np.arange(-lw, lw + 1)
    w = np.exp(-0.5 / (sigma * sigma) * k ** 2)
    return lw, w / w.sum()


def filter_plan(times, step_index, decimate_index, sigma_factor=0.01, max_sigma=None, truncate=4,
                sigma_node_factor=1.5, min_sigma=0.25):
    """Filter taps of every kept sample: what filter_chrono_signal + nonuniform_gaussian_filter1d would apply at
    ``decimate_index`` (per step segment: sigma = sigma_factor e dt / 2 capped by max_sigma and by half the
    distance to the neighbouring kept samples; Gaussians at log-spaced node widths blended with hat weights in
    ln sigma; nodes below min_sigma pass the sample through).  Returns a dict of host arrays for
    Engine.filter_gather."""
    times = np.asarray(times, dtype=float)
    n = len(times)
    dec = np.asarray(decimate_index, dtype=int)
    t_sample = np.median(np.diff(times))
    if max_sigma is None:
        max_sigma = sigma_factor / t_sample
    # sigma_from_decimate_index, preprocessing.py:577-591
    gap = np.diff(dec)
    near = np.minimum(np.insert(gap, 0, gap[0]), np.append(gap, gap[-1]))
    dec_sigma = np.where(near < 2, 0.0, near / (2.0 * 4.0))
    bounds = np.unique(np.concatenate(([0], np.asarray(step_index, dtype=int), [n])))
    m = len(dec)
    seg_lo, seg_len, lw_out = np.zeros(m, np.int32), np.zeros(m, np.int32), np.zeros(m, np.int32)
    woff = np.zeros(m, np.int64)
    taps = []
    cursor = 0
    for a, b in zip(bounds[:-1], bounds[1:]):
        sel = np.flatnonzero((dec >= a) & (dec < b))
        if sel.size == 0:
            continue
        ts = times[a:b]
        # the reference derives the node widths from the sigma of *every* sample of the segment
        sig_all = np.minimum(sigma_factor * (np.e * (ts - (ts[0] - t_sample)) / 2 / t_sample), max_sigma)
        dsig = np.zeros(b - a)
        dsig[dec[sel] - a] = dec_sigma[sel]
        sig_all = np.minimum(dsig, sig_all)
        seg_lo[sel], seg_len[sel] = a, b - a
        if not np.max(sig_all) > 0:                      # nothing to filter in this segment
            for j in sel:
                woff[j], lw_out[j] = cursor, 0
                taps.append(np.ones(1))
                cursor += 1
            continue
        sig_all = np.maximum(sig_all, 1e-8)
        lo = max(np.min(np.log10(sig_all)), np.log10(min_sigma))
        hi = max(np.max(np.log10(sig_all)), np.log10(min_sigma))
        nodes = np.logspace(lo, hi, int(np.ceil((hi - lo) / np.log10(sigma_node_factor))) + 1)
        if np.min(sig_all) < min_sigma:
            factor = nodes[-1] / nodes[-2] if len(nodes) > 1 else sigma_node_factor
            sig_all = np.where(sig_all < min_sigma / factor ** 2, min_sigma / factor ** 2, sig_all)
            while nodes[0] > np.min(sig_all) * 1.001:
                nodes = np.insert(nodes, 0, nodes[0] / factor)
        delta = np.log(nodes[-1] / nodes[-2]) if len(nodes) > 1 else 1
        kernels = {}
        for j in sel:
            s = sig_all[dec[j] - a]
            nw = 1 - np.minimum(np.abs(np.log(s / nodes)) / delta, 1)
            parts = []
            for q in np.flatnonzero(nw > 0):
                if nodes[q] < min_sigma:
                    parts.append((0, nw[q] * np.ones(1)))
                else:
                    if q not in kernels:
                        kernels[q] = _gauss_taps(nodes[q], truncate)
                    parts.append((kernels[q][0], nw[q] * kernels[q][1]))
            lw = max(p[0] for p in parts) if parts else 0
            tap = np.zeros(2 * lw + 1)
            for r, w in parts:
                tap[lw - r:lw + r + 1] += w
            woff[j], lw_out[j] = cursor, lw
            taps.append(tap)
            cursor += tap.size
    return dict(idx=dec.astype(np.int32), seg_lo=seg_lo, seg_len=seg_len, lw=lw_out, woff=woff,
                taps=np.concatenate(taps), n_raw=n)


def downsample_data(times, i_signal, v_signal, target_times=None, target_size=None, stepwise_sample_times=True,
                    step_times=None, step_model='ideal', method='match', decimation_interval=10,
                    decimation_factor=2, decimation_max_period=None, antialiased=True, filter_kw=None,
                    discard_first_n_points=None, discard_only=False, op_mode='galv', prestep_samples=20,
                    engine=None, return_device=False):
    """hybdrt.preprocessing.downsample_data (preprocessing.py:335-468), same arguments; ``v_signal`` may be one
    trace [Nt] or a batch [B, Nt] on the shared time grid (the input signal is shared).  Galvanostatic, ideal
    steps, stepwise sample times.  Returns (sample_times, sample_i, sample_v, sample_index)."""
    if op_mode != 'galv' or step_model != 'ideal' or not stepwise_sample_times:
        raise NotImplementedError('downsample_data: galvanostatic ideal-step traces with stepwise sample times only')
    if discard_only or discard_first_n_points is not None:
        raise NotImplementedError('downsample_data: discard_first_n_points / discard_only')
    times = np.asarray(times, dtype=float)
    i_signal = np.asarray(i_signal, dtype=float)
    v_arr = np.asarray(v_signal, dtype=float)
    single = v_arr.ndim == 1
    v_b = v_arr[None] if single else v_arr
    if step_times is None:
        step_times = times[identify_steps(i_signal, True)]
    step_index = step_indices_from_times(times, step_times)
    if method == 'match':
        if target_times is not None:
            tt = np.unique(np.concatenate([np.asarray(target_times, dtype=float) + ts for ts in step_times]))
            sample_index = np.unique([int(np.argmin(np.abs(times - t))) for t in tt])
        else:
            sample_index = np.arange(step_index[0], len(times), dtype=int)
        if step_index[0] > 0 and prestep_samples > 0:
            sample_index = np.unique(np.concatenate((np.arange(0, step_index[0], dtype=int), sample_index)))
    elif method == 'decimate':
        t_sample = np.min(np.diff(times))
        if target_size is not None:
            decimation_interval = select_decimation_interval(times, step_times, t_sample, prestep_samples,
                                                             decimation_factor, decimation_max_period, target_size)
        sample_index = get_decimation_index(times, step_times, t_sample, prestep_samples, decimation_interval,
                                            decimation_factor, decimation_max_period)
    else:
        raise ValueError(f"Invalid downsample method {method}. Options: 'match', 'decimate'")
    if antialiased:
        eng = engine or _engine.get_engine(0)
        plan = filter_plan(times, identify_steps(i_signal, allow_consecutive=False), sample_index, **(filter_kw or {}))
        sample_i = eng.filter_gather(i_signal, plan)[0].cpu().numpy()
        sv = eng.filter_gather(v_b, plan)
        sample_v = sv if return_device else sv.cpu().numpy()
    else:
        sample_i = i_signal[sample_index]
        sample_v = v_b[:, sample_index]
    if single and not return_device:
        sample_v = sample_v[0]
    return times[sample_index], sample_i, sample_v, sample_index
