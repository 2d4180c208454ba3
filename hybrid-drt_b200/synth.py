"""Synthetic spectra for the benchmark configurations.

Pure numpy; used by bench.py, the tests and oracle/make_golden.py so that every arm sees the same
seeded inputs.  Element formulas restate the reference's equivalent-circuit impedances
(hybdrt/models/elements.py:2054-2058 for the ZARC/RQ element, :2124-2137 for the RC step response)
and the tutorial's "mixed" noise structure (tutorials/Probabilistic_DRT_fitting.ipynb cell 3).
"""
import numpy as np

C2_FREQ = np.logspace(6, -2, 70)
C3_FREQ = np.logspace(6, 2, 30)
C3_TIMES = np.arange(-10, 1990) * 1e-3


def _mixed_noise(rng, z_exact, rp, sigma_rel=5e-3, uniform_frac=0.1):
    z_mod = np.abs(z_exact)
    sigma = uniform_frac * sigma_rel * rp[:, None] \
        + (1 - uniform_frac) * sigma_rel * rp[:, None] * z_mod / np.mean(z_mod, axis=1, keepdims=True)
    return rng.normal(0, sigma) + 1j * rng.normal(0, sigma)


def zarc_params(batch, seed=0):
    """Random 2-ZARC parameter draws of config C2 (one row per spectrum)."""
    rng = np.random.default_rng(seed)
    p = dict(
        r_inf=rng.uniform(0.5, 2.0, batch),
        r=rng.uniform(0.2, 2.0, (batch, 2)),
        log_tau=np.stack([rng.uniform(-5, -2, batch), rng.uniform(-2, 0.5, batch)], axis=1),
        beta=rng.uniform(0.6, 1.0, (batch, 2)),
        induc=10 ** rng.uniform(-8, -6.5, batch),
    )
    return p, rng


def zarc_impedance(freq, p):
    jw = 2j * np.pi * np.asarray(freq)[None, :]
    z = p['r_inf'][:, None] + jw * p['induc'][:, None]
    for i in range(p['r'].shape[1]):
        tau = 10 ** p['log_tau'][:, i, None]
        z = z + p['r'][:, i, None] / (1 + (jw * tau) ** p['beta'][:, i, None])
    return z


def make_eis_batch(batch, freq=None, seed=0):
    """Config C2: ``batch`` noisy 2-ZARC spectra on a shared frequency grid. Returns (freq, z[B,Nf])."""
    freq = C2_FREQ if freq is None else np.asarray(freq, dtype=float)
    p, rng = zarc_params(batch, seed)
    z = zarc_impedance(freq, p)
    z = z + _mixed_noise(rng, z, p['r'].sum(axis=1))
    return freq, z


def make_map_batch(rows=256, cols=256, freq=None, seed=3):
    """Config C5: rows x cols map; R_1 varies along rows, log tau_1 along columns, rest fixed."""
    freq = C2_FREQ if freq is None else np.asarray(freq, dtype=float)
    batch = rows * cols
    rng = np.random.default_rng(seed)
    rr, cc = np.meshgrid(np.linspace(0, 1, rows), np.linspace(0, 1, cols), indexing='ij')
    p = dict(
        r_inf=np.full(batch, 1.0),
        r=np.stack([0.3 + 1.2 * rr.ravel(), np.full(batch, 0.8)], axis=1),
        log_tau=np.stack([-4.5 + 2.0 * cc.ravel(), np.full(batch, -0.5)], axis=1),
        beta=np.stack([np.full(batch, 0.85), np.full(batch, 0.75)], axis=1),
        induc=np.full(batch, 1e-7),
    )
    z = zarc_impedance(freq, p)
    z = z + _mixed_noise(rng, z, p['r'].sum(axis=1))
    return freq, z


def make_dop_batch(batch, freq=None, seed=2):
    """Config C4: one ZARC + pseudo-capacitive 0.05 (jw)^-0.6 + pseudo-inductive 1e-5 (jw)^0.7."""
    freq = C2_FREQ if freq is None else np.asarray(freq, dtype=float)
    rng = np.random.default_rng(seed)
    r_inf = rng.uniform(0.5, 2.0, batch)
    r = rng.uniform(0.2, 2.0, batch)
    log_tau = rng.uniform(-4, -1, batch)
    beta = rng.uniform(0.6, 1.0, batch)
    jw = 2j * np.pi * freq[None, :]
    z = r_inf[:, None] + r[:, None] / (1 + (jw * 10 ** log_tau[:, None]) ** beta[:, None])
    z = z + 0.05 * jw ** -0.6 + 1e-5 * jw ** 0.7
    z = z + _mixed_noise(rng, z, r)
    return freq, z


def make_hybrid_batch(batch, freq=None, times=None, seed=1, i0=1e-2, n_rc=4):
    """Config C3: RC-ladder step response + high-frequency EIS.

    Returns (times, i_signal[Nt], v[B,Nt], freq, z[B,Nf]).  ZARC step responses need the absent
    ``mitlef`` package, hence RC elements.
    """
    freq = C3_FREQ if freq is None else np.asarray(freq, dtype=float)
    times = C3_TIMES if times is None else np.asarray(times, dtype=float)
    rng = np.random.default_rng(seed)
    r_inf = rng.uniform(0.5, 2.0, batch)
    r = rng.uniform(0.2, 1.0, (batch, n_rc))
    edges = np.linspace(np.log10(1e-5), np.log10(0.5), n_rc + 1)
    log_tau = rng.uniform(edges[:-1], edges[1:], (batch, n_rc))
    tau = 10 ** log_tau
    i_signal = i0 * (times >= 0)
    tpos = np.maximum(times, 0.0)[None, None, :]
    v = r_inf[:, None] * i_signal[None, :] \
        + i0 * np.sum(r[:, :, None] * (1 - np.exp(-tpos / tau[:, :, None])), axis=1) * (times >= 0)
    v = v + rng.normal(0, 2e-6, v.shape)
    jw = 2j * np.pi * freq[None, None, :]
    z = r_inf[:, None] + np.sum(r[:, :, None] / (1 + jw * tau[:, :, None]), axis=1)
    z = z + rng.normal(0, 1e-3, z.shape) + 1j * rng.normal(0, 1e-3, z.shape)
    return times, i_signal, v, freq, z


def make_raw_chrono_batch(batch=3, seed=5, n_pre=300, n_post=24000, dt=1e-4):
    """Densely sampled two-step galvanostatic traces (RC ladder responses + noise), as an instrument records them
    before downsampling.  Returns (times[Nt], i_signal[Nt], v[B,Nt])."""
    rng = np.random.default_rng(seed)
    times = (np.arange(-n_pre, n_post) + 0.0) * dt
    i_signal = 1e-2 * (times >= 0) - 0.6e-2 * (times >= 1.2)
    r_inf = rng.uniform(0.5, 2.0, batch)
    r = rng.uniform(0.2, 1.0, (batch, 3))
    tau = 10 ** rng.uniform([-4, -2.5, -1], [-3, -1.5, 0], (batch, 3))
    v = np.zeros((batch, times.size))
    for st, sa in ((0.0, 1e-2), (1.2, -0.6e-2)):
        tp = np.maximum(times - st, 0.0)[None, None, :]
        v += sa * (r_inf[:, None] + np.sum(r[:, :, None] * (1 - np.exp(-tp / tau[:, :, None])), axis=1)) * (times >= st)
    v += rng.normal(0, 5e-6, v.shape)
    i_noisy = i_signal + rng.normal(0, 2e-7, times.shape)
    return times, i_noisy, v
