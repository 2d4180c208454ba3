"""CPU ORACLE support (test infrastructure, not product code).

Golden fixture for the mapping path: runs the UNMODIFIED reference ``hybdrt.mapping.DRTMD`` (on top of
oracle/refshim.py) on a small synthetic map and stores what ``fit_all`` leaves behind.  Run in the authoring
container only (the reference cannot travel):

    python -m oracle.make_golden_map
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

refshim.install()
warnings.filterwarnings('ignore')
from hybdrt.mapping.drtmd import DRTMD  # noqa: E402  (the reference)
from hybdrt_b200 import synth  # noqa: E402


def main():
    rows, cols = 2, 3
    freq, z = synth.make_map_batch(rows, cols, seed=3)
    supergrid = np.logspace(-8, 3, 111)
    md = DRTMD(tau_supergrid=supergrid, psi_dim_names=['row', 'col'], print_progress=False)
    psi = np.array([(r, c) for r in range(rows) for c in range(cols)], dtype=float)
    for b in range(len(z)):
        md.add_observation(psi[b], None, (freq, z[b]))
    md.fit_all()
    assert md.obs_fit_status.all()
    out = dict(freq=freq, z=z, psi=psi, tau_supergrid=supergrid, obs_x=md.obs_x,
               obs_tau_indices=np.array(md.obs_tau_indices), obs_drt_var=md.obs_drt_var, obs_llh=md.obs_llh,
               obs_rss=md.obs_rss, tau_epsilon=md.tau_epsilon)
    for key, val in md.obs_special.items():
        out['special_' + key] = np.asarray(val)
    path = os.path.join(ROOT, 'tests', 'golden', 'drtmd_small.npz')
    if '--hybrid-only' in sys.argv:
        return hybrid_map(supergrid)
    if '--resolve-only' in sys.argv:
        return resolve_map(supergrid)
    if '--pfrt-only' not in sys.argv:
        np.savez_compressed(path, **out)
        print('wrote', path, {k: np.shape(v) for k, v in out.items()})
    # fit_type='pfrt': one solution per factor and observation (drtmd.py:1140-1160)
    mp = DRTMD(tau_supergrid=supergrid, psi_dim_names=['row', 'col'], print_progress=False, fit_type='pfrt')
    for b in range(3):
        mp.add_observation(psi[b], None, (freq, z[b]))
    mp.fit_all()
    assert mp.obs_fit_status.all()
    outp = dict(freq=freq, z=z[:3], psi=psi[:3], tau_supergrid=supergrid, obs_x=mp.obs_x,
                obs_tau_indices=np.array(mp.obs_tau_indices), obs_drt_var=mp.obs_drt_var, obs_llh=mp.obs_llh,
                obs_rss=mp.obs_rss, pfrt_factors=mp.pfrt_factors)
    for key, val in mp.obs_special.items():
        outp['special_' + key] = np.asarray(val)
    path = os.path.join(ROOT, 'tests', 'golden', 'drtmd_pfrt.npz')
    np.savez_compressed(path, **outp)
    print('wrote', path, {k: np.shape(v) for k, v in outp.items()})
    hybrid_map(supergrid)
    resolve_map(supergrid)


def hybrid_map(supergrid):
    """Hybrid observations (chrono + EIS): obs_llh / obs_rss are evaluated with the design matrix whose vz_offset column
    has been rewritten from the final coefficients (drt1d.py:972-979, 4433-4496)."""
    times = np.concatenate([np.linspace(-0.01, -1e-4, 25), np.logspace(-4, 0, 220)])
    t, i_sig, v, freq, z = synth.make_hybrid_batch(3, times=times, seed=11)
    mh = DRTMD(tau_supergrid=supergrid, psi_dim_names=['k'], print_progress=False)
    for b in range(3):
        mh.add_observation([float(b)], (t, i_sig, v[b]), (freq, z[b]))
    mh.fit_all()
    assert mh.obs_fit_status.all()
    outh = dict(times=t, i_signal=i_sig, v=v, freq=freq, z=z, tau_supergrid=supergrid, obs_x=mh.obs_x,
                obs_tau_indices=np.array(mh.obs_tau_indices), obs_drt_var=mh.obs_drt_var, obs_llh=mh.obs_llh, obs_rss=mh.obs_rss)
    for key, val in mh.obs_special.items():
        outh['special_' + key] = np.asarray(val)
    path = os.path.join(ROOT, 'tests', 'golden', 'drtmd_hybrid.npz')
    np.savez_compressed(path, **outh)
    print('wrote', path, {k: np.shape(v) for k, v in outh.items()})



def resolve_map(supergrid=np.logspace(-8, 3, 111)):
    """Cross-observation resolve (mapping/resolve.py:176-341 through DRTMD.resolve_observations / resolve_group,
    drtmd.py:432-560) on nine hybrid observations whose parameters drift along psi."""
    times = np.concatenate([np.linspace(-0.01, -1e-4, 25), np.logspace(-4, 0, 220)])
    t, i_sig, v, freq, z = synth.make_hybrid_batch(9, times=times, seed=17)
    # a smooth drift along psi, so that the coupling between neighbours has something to do
    order = np.argsort(z.real.max(axis=1))
    v, z = v[order], z[order]
    mr = DRTMD(tau_supergrid=supergrid, psi_dim_names=['k'], print_progress=False)
    for b in range(9):
        mr.add_observation([float(b)], (t, i_sig, v[b]), (freq, z[b]), group_id='g')
    mr.fit_all()
    assert mr.obs_fit_status.all()
    from oracle import refshim as _rs
    fits = [mr.get_fit(i) for i in range(9)]
    sp = fits[0].special_qp_params
    per_obs = dict(
        fit_p=np.array([f.fit_parameters['p_matrix'] for f in fits]), fit_q=np.array([f.fit_parameters['q_vector'] for f in fits]),
        fit_coefficient_scale=np.array([f.coefficient_scale for f in fits]),
        fit_response_signal_scale=np.array([f.response_signal_scale for f in fits]),
        fit_scaled_response_offset=np.array([f.scaled_response_offset for f in fits]),
        fit_v_baseline_scale=np.array([np.ravel(f.v_baseline_scale) for f in fits]),
        fit_v_baseline=np.array([np.ravel(f.fit_parameters['v_baseline']) for f in fits]),
        fit_vz_offset=np.array([f.fit_parameters['vz_offset'] for f in fits]),
        fit_R_inf=np.array([f.fit_parameters['R_inf'] for f in fits]),
        fit_inductance_scale=np.array([f.inductance_scale for f in fits]),
        special_names=np.array(list(sp.keys())), special_index=np.array([v['index'] for v in sp.values()]),
        special_size=np.array([v.get('size', 1) for v in sp.values()]), special_nonneg=np.array([v['nonneg'] for v in sp.values()]))
    _rs.QP_LOG.clear()
    mr.resolve_observations(np.arange(7), psi_sort_dims=['k'], sigma=1, lambda_psi=1)
    out = dict(times=t, i_signal=i_sig, v=v, freq=freq, z=z, tau_supergrid=supergrid, obs_x=mr.obs_x,
               obs_tau_indices=np.array(mr.obs_tau_indices), win_ipm=np.array(_rs.QP_LOG),
               win_x_resolved=mr.obs_x_resolved[:7].copy())
    for key, val in mr.obs_special_resolved.items():
        out['win_special_' + key] = np.asarray(val)[:7].copy()
    _rs.QP_LOG.clear()
    mr.resolve_group('g', batch_size=7, overlap=2, psi_sort_dims=['k'], sigma=1, lambda_psi=1)
    out.update(per_obs)
    out.update(grp_ipm=np.array(_rs.QP_LOG), grp_x_resolved=mr.obs_x_resolved.copy(), grp_status=mr.obs_resolve_status.copy())
    for key, val in mr.obs_special_resolved.items():
        out['grp_special_' + key] = np.asarray(val).copy()
    path = os.path.join(ROOT, 'tests', 'golden', 'drtmd_resolve.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: np.shape(v) for k, v in out.items()}, 'ipm', out['win_ipm'], out['grp_ipm'])

if __name__ == '__main__':
    main()
