"""CPU-side checks: the C-ABI library loads and exports what include/hybdrt_b200.h declares, struct
layouts agree, and the host-side logic of the model class matches reference-generated fixtures."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden, rel_err


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    g.build()
    from hybdrt_b200 import engine
    return engine.load_library()


def test_library_exports_every_declared_symbol(lib):
    from hybdrt_b200 import engine
    hdr = open(os.path.join(ROOT, 'include', 'hybdrt_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(hdrt_[a-z0-9_]+)\s*\(', hdr)))
    assert declared, 'no declarations parsed'
    for sym in declared:
        assert hasattr(lib, sym), f'{sym} missing from libhybdrt_b200.so'
    assert sorted(engine.EXPORTED_SYMBOLS) == declared
    assert lib.hdrt_version() >= 100


def test_struct_defaults_and_smem_budget(lib):
    from hybdrt_b200 import engine
    h = engine.default_hypers()
    assert list(h.derivative_weights) == [1.5, 1.0, 0.5]
    assert list(h.s_alpha) == [5.0, 10.0, 25.0]
    assert h.l2_lambda_0 == 142.0 and h.dop_l2_lambda_0 == 10.0
    assert h.max_iter == 50 and h.xtol == 1e-2 and h.has_iw_prior == 0
    # C2 (140 x 103) must allow three CTAs per SM (228 KB per SM, 1 KB reserved per CTA); C4 (140 x 153) and
    # C3 (2060 x 96) must fit at all
    assert 0 < lib.hdrt_qphb_smem_bytes(140, 103) <= (228 * 1024 - 3 * 1024) // 3
    assert 0 < lib.hdrt_qphb_smem_bytes(140, 153) <= 227 * 1024
    assert 0 < lib.hdrt_qphb_smem_bytes(2060, 96) <= (228 * 1024 - 2 * 1024) // 2     # hybrid: two CTAs per SM
    assert lib.hdrt_qphb_smem_bytes(140, 300) < 0


def test_argument_errors_do_not_need_a_gpu(lib):
    from hybdrt_b200 import engine
    assert lib.hdrt_build_lookup(4.0, 1, 1000, None, None, None, None, None, None, None) == -1
    assert b'invalid' in lib.hdrt_last_error()
    p = engine.Problem()
    assert lib.hdrt_qphb_fit_batch(None, ctypes.byref(p), None) == -1


def test_engine_refuses_to_run_without_gpu():
    import torch
    from hybdrt_b200 import engine
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(engine.EngineError):
        engine.Engine(0)


def test_host_helpers_match_reference_fixtures():
    from hybdrt_b200 import models, synth
    c2 = load_golden('c2_eis.npz')
    assert rel_err(models.get_basis_tau(c2['freq'], None, None), c2['basis_tau']) < 1e-14
    rp = models.estimate_rp_batch(None, None, None, None, c2['z'])
    assert rel_err(rp / 14, c2['coefficient_scale']) < 1e-14
    hs = load_golden('hybrid_small.npz')
    st, sa = models.step_info(hs['times'], hs['i_signal'])
    assert rel_err(st, hs['step_times']) < 1e-14 and rel_err(sa, hs['step_sizes']) < 1e-14
    assert rel_err(models.get_basis_tau(hs['freq'], hs['times'], st), hs['basis_tau']) < 1e-13
    rp = models.estimate_rp_batch(hs['times'], st, sa, hs['v_signal'][None], hs['z'][None])
    assert rel_err(rp / 14, hs['coefficient_scale']) < 1e-13
    # supergrid selection
    grid = np.logspace(-8, 3, 111)
    sub = models.get_basis_tau(synth.C2_FREQ, None, None, tau_grid=grid)
    assert sub[0] <= 1 / (2 * np.pi * 1e6) / 10 * 1.0001 and sub[-1] >= 1 / (2 * np.pi * 1e-2) * 10 * 0.9999


def test_synthetic_generators_are_seeded():
    from hybdrt_b200 import synth
    f, z = synth.make_eis_batch(12, seed=0)
    c2 = load_golden('c2_eis.npz')
    assert np.array_equal(z, c2['z']) and np.array_equal(f, c2['freq'])


def test_kk_statistics_match_reference_fixture():
    """models/kk.py: outlier flags and frequency window from the residuals of the reference's own KK fits."""
    from hybdrt_b200 import kk
    g = load_golden('kk.npz')
    for b in range(2):
        out = kk.get_outliers(g[f'resid_{b}'])
        assert np.array_equal(out, g[f'outlier_index_{b}'])
        f_min, f_max = kk.get_limits(g['freq'], out)
        assert f_min == g[f'limits_{b}'][0] and f_max == g[f'limits_{b}'][1]
        assert len(kk.trim_data(g['freq'], g['z'][b], f_min, f_max)[0]) == int(g[f'n_clean_{b}'])
    # hand-made cases for the window logic: a clean run is kept, a cluster at one end is cut off
    f = np.logspace(5, 0, 21)
    assert kk.get_limits(f, np.array([], dtype=int)) == (f[-1], f[0])
    f_min, f_max = kk.get_limits(f, np.array([0, 1, 2, 10]), max_num_outliers=1)
    assert f_max == f[4] and f_min == f[-1]        # the first clean point with clean neighbours
    (f_min, f_max), (il, ir) = kk.get_limits(f, np.array([18, 19, 20]), return_index=True)
    assert f_max == f[0] and f_min == f[16] and (il, ir) == (0, 16)
