"""ctypes binding of libhybdrt_b200.so (include/hybdrt_b200.h) with torch as the device allocator.

torch is plumbing here: it owns device memory and the CUDA stream.  Every number is produced by the
CUDA kernels behind the C ABI; if the library is missing or no GPU is present this module raises --
there is no CPU path.
"""
import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_lib', 'libhybdrt_b200.so')

MODE_INTERP, MODE_TRAPZ = 0, 1
ST_CONVERGED, ST_MAXITER, ST_QP_MAXITERS, ST_KKT_FAIL, ST_NAN, ST_COV_FAIL = 1, 2, 4, 8, 16, 32

_D3 = C.c_double * 3


class Hypers(C.Structure):
    """struct hdrt_hypers"""
    _fields_ = [
        ('derivative_weights', _D3), ('sigma_ds', _D3), ('s_alpha', _D3), ('s_0', _D3),
        ('rho_alpha', _D3), ('rho_0', _D3), ('l2_lambda_0', C.c_double),
        ('dop_derivative_weights', _D3), ('dop_sigma_ds', _D3), ('dop_s_alpha', _D3), ('dop_s_0', _D3),
        ('dop_rho_alpha', _D3), ('dop_rho_0', _D3), ('dop_l2_lambda_0', C.c_double),
        ('iw_l1_lambda_0', C.c_double), ('iw_l2_lambda_0', C.c_double),
        ('iw_alpha', C.c_double), ('iw_beta', C.c_double),
        ('xtol', C.c_double), ('weight_factor', C.c_double),
        ('chrono_weight_factor', C.c_double), ('eis_weight_factor', C.c_double),
        ('has_iw_prior', C.c_int), ('max_iter', C.c_int),
        ('outlier_p', C.c_double), ('has_outlier_p', C.c_int), ('solve_rp', C.c_int),
        ('update_scale', C.c_int), ('normalize_dop', C.c_int), ('init_weights_separately', C.c_int),
        ('hybrid_wf_method', C.c_int), ('rp_scale', C.c_double), ('basis_area', C.c_double),
    ]


_P = C.c_void_p


class Problem(C.Structure):
    """struct hdrt_qphb_problem"""
    _fields_ = [
        ('batch', C.c_int), ('n_rows', C.c_int), ('n_cols', C.c_int), ('n_special', C.c_int),
        ('n_chrono', C.c_int), ('dop_start', C.c_int), ('dop_end', C.c_int), ('vz_index', C.c_int),
        ('vb_start', C.c_int), ('vb_end', C.c_int), ('hybrid', C.c_int),
        ('rm', _P), ('rm_stride', C.c_longlong), ('rv', _P),
        ('vmm_eis', _P), ('vmm_eis_stride', C.c_longlong),
        ('vmm_chrono', _P), ('vmm_chrono_stride', C.c_longlong),
        ('pen', _P), ('pen_stride', C.c_longlong),
        ('h', _P), ('l1', _P), ('vz_strength', _P),
        ('hyp', Hypers),
        ('x', _P), ('weights', _P), ('est_weights', _P), ('init_weights', _P), ('x_overfit', _P),
        ('s_vectors', _P), ('rho', _P), ('dop_rho', _P), ('xmx_norms', _P), ('dop_xmx_norms', _P),
        ('fun', _P), ('vz_col', _P), ('p_matrix', _P), ('q_vector', _P),
        ('n_outer', _P), ('n_ipm', _P), ('status', _P),
        ('eval_mat', _P), ('n_eval', C.c_int), ('dist_var', _P), ('resid_ss', _P), ('outlier_t', _P), ('scale_factors', _P),
        ('n_pfrt', C.c_int), ('pfrt_max_iter', C.c_int), ('pfrt_min_iter', C.c_int), ('pfrt_factors', _P),
        ('pfrt_x', _P), ('pfrt_llh', _P), ('pfrt_p', _P), ('pfrt_iters', _P), ('hybrid_wf_in', _P), ('hybrid_wf_out', _P), ('x_overfit_eis', _P),
        ('weight_factor_vec', _P), ('vz_scratch', _P),
        ('pen_toeplitz', _P), ('pen_band', C.c_int),
    ]


class ResolveProblem(C.Structure):
    """struct hdrt_resolve_problem"""
    _fields_ = [('n_windows', C.c_int), ('nr', C.c_int), ('nc', C.c_int), ('p', _P), ('q', _P), ('first_obs', _P),
                ('my', _P), ('param_scale', _P), ('h', _P), ('x', _P), ('iters', _P), ('status', _P)]


class EngineError(RuntimeError):
    pass


_lib = None


def load_library():
    """dlopen the in-tree CUDA library; raise (never fall back) if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(f'{LIB_PATH} not found: run `python -c "import __graft_entry__ as g; g.build()"` '
                          f'(or python hybrid-drt_b200/build.py). hybdrt_b200 has no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    lib.hdrt_version.restype = C.c_int
    lib.hdrt_last_error.restype = C.c_char_p
    lib.hdrt_create.argtypes = [C.POINTER(_P), C.c_int]
    lib.hdrt_destroy.argtypes = [_P]
    lib.hdrt_sm_count.argtypes = [_P]
    lib.hdrt_build_lookup.argtypes = [C.c_double, C.c_int, C.c_int] + [_P] * 6 + [_P]
    lib.hdrt_build_impedance.argtypes = [C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, C.c_double, _P, _P, _P, _P,
                                         C.c_int, C.c_int, _P, _P, _P]
    lib.hdrt_build_response.argtypes = [C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                        _P, _P, C.c_int, C.c_int, _P, _P]
    lib.hdrt_build_penalty.argtypes = [_P, C.c_int, C.c_int, C.c_double, C.c_int, _P, _P]
    lib.hdrt_build_eis_vmm.argtypes = [_P, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, _P, _P]
    lib.hdrt_build_chrono_vmm.argtypes = [_P, _P, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, _P, _P]
    lib.hdrt_build_dop_z.argtypes = [_P, _P, C.c_int, C.c_int, C.c_int, C.c_double, _P, _P]
    lib.hdrt_build_dop_v.argtypes = [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _P, _P]
    lib.hdrt_filter_gather.argtypes = [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, C.c_int, _P, _P]
    lib.hdrt_default_hypers.argtypes = [C.POINTER(Hypers)]
    lib.hdrt_default_hypers.restype = None
    lib.hdrt_qphb_smem_bytes.argtypes = [C.c_int, C.c_int]
    lib.hdrt_qphb_smem_bytes.restype = C.c_longlong
    lib.hdrt_qphb_fit_batch.argtypes = [_P, C.POINTER(Problem), _P]
    lib.hdrt_probe_fp64.argtypes = [_P, C.POINTER(C.c_double), _P]
    lib.hdrt_resolve_qp_batch.argtypes = [_P, C.POINTER(ResolveProblem), _P, _P]
    lib.hdrt_resolve_work_bytes.argtypes = [_P, C.c_int, C.c_int, C.c_int]
    lib.hdrt_resolve_work_bytes.restype = C.c_longlong
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    'hdrt_version', 'hdrt_last_error', 'hdrt_create', 'hdrt_destroy', 'hdrt_sm_count', 'hdrt_build_lookup',
    'hdrt_build_impedance', 'hdrt_build_response', 'hdrt_build_penalty', 'hdrt_build_eis_vmm', 'hdrt_build_chrono_vmm', 'hdrt_build_dop_z', 'hdrt_build_dop_v', 'hdrt_filter_gather',
    'hdrt_default_hypers', 'hdrt_qphb_smem_bytes', 'hdrt_qphb_fit_batch', 'hdrt_probe_fp64', 'hdrt_resolve_qp_batch',
    'hdrt_resolve_work_bytes',
]


def default_hypers():
    h = Hypers()
    load_library().hdrt_default_hypers(C.byref(h))
    return h


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class Engine:
    """One per device: owns the C-side handle, allocates through torch, launches on torch's stream."""

    def __init__(self, device=0):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise EngineError('hybdrt_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        self.device = torch.device('cuda', device if isinstance(device, int) else torch.device(device).index or 0)
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)      # make sure the primary context exists
        h = _P()
        self._check(self.lib.hdrt_create(C.byref(h), self.device.index))
        self.handle = h
        self.sm_count = self.lib.hdrt_sm_count(h)
        self._lookups = {}
        self._pinned = {}
        self.launches = 0           # kernels launched through this engine (bench.py reports it)
        self.max_pinned_shapes = 16

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.hdrt_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------------------------------
    def _check(self, code):
        if code != 0:
            raise EngineError(f'hybdrt_b200 error {code}: {self.lib.hdrt_last_error().decode()}')

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def dev(self, a, dtype=torch.float64):
        """numpy / tensor -> contiguous device tensor of dtype."""
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.device)

    def empty(self, *shape, dtype=torch.float64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def pinned(self, *shape, dtype=torch.float64):
        """Page-locked host staging buffer, cached per shape (reused across calls)."""
        key = (tuple(shape), dtype)
        buf = self._pinned.pop(key, None)
        if buf is None:
            buf = torch.empty(*shape, dtype=dtype, pin_memory=True)
        self._pinned[key] = buf                       # most recently used last
        while len(self._pinned) > self.max_pinned_shapes:     # maps with many group sizes: do not hoard page-locked memory
            self._pinned.pop(next(iter(self._pinned)))
        return buf

    # -- L0 --------------------------------------------------------------------------------------
    def filter_gather(self, y, plan):
        """Antialiasing filter + decimation of a batch of traces y [S, nt] (device or host) with the tap layout
        ``plan`` of hybdrt_b200.preprocessing.filter_plan (cached on the device inside the plan).  -> [S, m]."""
        y = self.dev(y)
        if y.dim() == 1:
            y = y[None]
        if 'dev' not in plan or plan['dev'][0] != self.device:
            plan['dev'] = (self.device, {k: self.dev(plan[k], dtype=torch.int32) for k in ('idx', 'seg_lo', 'seg_len', 'lw')},
                           self.dev(plan['woff'], dtype=torch.int64), self.dev(plan['taps']))
        _, ints, woff, taps = plan['dev']
        m = len(plan['idx'])
        out = self.empty(y.shape[0], m)
        self._check(self.lib.hdrt_filter_gather(_ptr(y), y.shape[0], y.shape[1], _ptr(ints['idx']), _ptr(ints['seg_lo']),
                                                _ptr(ints['seg_len']), _ptr(woff), _ptr(ints['lw']), _ptr(taps), m,
                                                _ptr(out), self._stream()))
        self.launches += 1
        return out

    # -- L1 --------------------------------------------------------------------------------------
    def build_lookup(self, eps, grid_points=2000, quad_points=1000):
        """basis.generate_impedance_lookup + generate_response_lookup (basis.py:648-689); cached per eps."""
        key = (float(eps), grid_points, quad_points)
        if key not in self._lookups:
            t = self.empty(6, grid_points)
            self._check(self.lib.hdrt_build_lookup(float(eps), grid_points, quad_points,
                                                   *[_ptr(t[i]) for i in range(6)], self._stream()))
            self.launches += 1
            self._lookups[key] = dict(re_x=t[0], re_v=t[1], im_x=t[2], im_v=t[3], resp_x=t[4], resp_v=t[5],
                                      grid_points=grid_points)
        return self._lookups[key]

    def build_impedance(self, freq, tau, eps, mode=MODE_INTERP, tables=None, quad_points=1000):
        """mat1d.construct_impedance_matrix for both parts. freq [G,nf], tau [G,nb] -> a_re, a_im [G,nf,nb]."""
        freq = self.dev(freq).reshape(-1, np.shape(freq)[-1])
        tau = self.dev(tau).reshape(-1, np.shape(tau)[-1])
        g, nf = freq.shape
        nb = tau.shape[1]
        if tau.shape[0] != g:
            raise ValueError('freq and tau must have the same number of grids')
        a_re, a_im = self.empty(g, nf, nb), self.empty(g, nf, nb)
        if mode == MODE_INTERP:
            tables = tables or self.build_lookup(eps)
            targs = (_ptr(tables['re_x']), _ptr(tables['re_v']), _ptr(tables['im_x']), _ptr(tables['im_v']),
                     tables['grid_points'])
        else:
            targs = (None, None, None, None, 0)
        self._check(self.lib.hdrt_build_impedance(mode, _ptr(freq), _ptr(tau), g, nf, nb, float(eps), *targs,
                                                  quad_points, _ptr(a_re), _ptr(a_im), self._stream()))
        self.launches += 1
        return a_re, a_im

    def build_response(self, times, tau, step_times, step_sizes, eps, mode=MODE_INTERP, tables=None,
                       quad_points=1000):
        """mat1d.construct_response_matrix. times [G,nt], tau [G,nb], steps [G,ns] -> rm [G,nt,nb]."""
        times = self.dev(times).reshape(-1, np.shape(times)[-1])
        tau = self.dev(tau).reshape(-1, np.shape(tau)[-1])
        st = self.dev(step_times).reshape(-1, np.shape(step_times)[-1])
        sa = self.dev(step_sizes).reshape(-1, np.shape(step_sizes)[-1])
        g, nt = times.shape
        nb = tau.shape[1]
        rm = self.empty(g, nt, nb)
        if mode == MODE_INTERP:
            tables = tables or self.build_lookup(eps)
            targs = (_ptr(tables['resp_x']), _ptr(tables['resp_v']), tables['grid_points'])
        else:
            targs = (None, None, 0)
        self._check(self.lib.hdrt_build_response(mode, _ptr(times), _ptr(tau), _ptr(st), _ptr(sa), g, nt, nb,
                                                 st.shape[1], float(eps), *targs, quad_points, _ptr(rm),
                                                 self._stream()))
        self.launches += 1
        return rm

    def build_penalty(self, grid, eps, toeplitz):
        """mat1d.construct_integrated_derivative_matrix orders 0..2. grid [G,nb] -> [G,3,nb,nb]."""
        grid = self.dev(grid).reshape(-1, np.shape(grid)[-1])
        g, nb = grid.shape
        m = self.empty(g, 3, nb, nb)
        self._check(self.lib.hdrt_build_penalty(_ptr(grid), g, nb, float(eps), int(bool(toeplitz)), _ptr(m),
                                                self._stream()))
        self.launches += 1
        return m

    def build_eis_vmm(self, freq, vmm_eps=0.25, reim_cor=0.25, uniform=False):
        """mat1d.construct_eis_var_matrix. freq [G,nf] -> [G,2nf,2nf]."""
        freq = self.dev(freq).reshape(-1, np.shape(freq)[-1])
        g, nf = freq.shape
        vmm = self.empty(g, 2 * nf, 2 * nf)
        self._check(self.lib.hdrt_build_eis_vmm(_ptr(freq), g, nf, float(vmm_eps), float(reim_cor),
                                                int(bool(uniform)), _ptr(vmm), self._stream()))
        self.launches += 1
        return vmm

    def build_chrono_vmm(self, times, step_times, vmm_eps=4, uniform=False):
        """mat1d.construct_chrono_var_matrix. times [G,nt], step_times [G,ns] -> [G,nt,nt]."""
        times = self.dev(times).reshape(-1, np.shape(times)[-1])
        st = self.dev(step_times).reshape(-1, np.shape(step_times)[-1])
        g, nt = times.shape
        vmm = self.empty(g, nt, nt)
        self._check(self.lib.hdrt_build_chrono_vmm(_ptr(times), _ptr(st), g, nt, st.shape[1], float(vmm_eps),
                                                   int(bool(uniform)), _ptr(vmm), self._stream()))
        self.launches += 1
        return vmm

    def build_dop_z(self, freq, nu, nu_eps):
        """phasance.construct_phasor_z_matrix (gaussian). freq [G,nf], nu [n_nu] -> complex128 [G,nf,n_nu]."""
        freq = self.dev(freq).reshape(-1, np.shape(freq)[-1])
        nu = self.dev(nu).reshape(-1)
        g, nf = freq.shape
        zm = self.empty(g, nf, nu.numel(), 2)
        self._check(self.lib.hdrt_build_dop_z(_ptr(freq), _ptr(nu), g, nf, nu.numel(), float(nu_eps), _ptr(zm),
                                              self._stream()))
        self.launches += 1
        return torch.view_as_complex(zm)

    def build_dop_v(self, times, nu, step_times, step_sizes, nu_eps):
        """phasance.construct_phasor_v_matrix (gaussian). times [G,nt], steps [G,ns], nu [n_nu] -> [G,nt,n_nu]."""
        times = self.dev(times).reshape(-1, np.shape(times)[-1])
        st = self.dev(step_times).reshape(-1, np.shape(step_times)[-1])
        sa = self.dev(step_sizes).reshape(-1, np.shape(step_sizes)[-1])
        nu = self.dev(nu).reshape(-1)
        g, nt = times.shape
        rm = self.empty(g, nt, nu.numel())
        self._check(self.lib.hdrt_build_dop_v(_ptr(times), _ptr(nu), _ptr(st), _ptr(sa), g, nt, nu.numel(),
                                              st.shape[1], float(nu_eps), _ptr(rm), self._stream()))
        self.launches += 1
        return rm

    # -- L2 --------------------------------------------------------------------------------------
    def smem_bytes(self, n_rows, n_cols):
        return int(self.lib.hdrt_qphb_smem_bytes(n_rows, n_cols))

    def qphb_fit_batch(self, rm, rv, pen, h, l1, n_special, vmm_eis=None, vmm_chrono=None, n_chrono=0,
                       dop_range=None, vz_index=-1, vb_range=(-1, -1), vz_strength=None, hybrid=False,
                       hypers=None, want_pq=False, out=None, eval_mat=None, want_resid=False, pfrt=None,
                       weight_factor_vec=None, hybrid_wf=None, pen_hint=None):
        """Launch the batched QPHB solver.  All inputs are device float64 tensors.

        rm [N,n] (shared) or [B,N,n]; rv [B,N]; pen [3,n,n] or [B,3,n,n]; h, l1 [n].
        Returns a dict of device tensors (asynchronous; synchronise before reading on the host).
        """
        rv = rv.contiguous()
        b, n_rows = rv.shape
        n = rm.shape[-1]
        assert rm.shape[-2] == n_rows and pen.shape[-1] == n and pen.shape[-3] == 3
        for t in (rm, rv, pen, h, l1):
            assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        hyp = hypers if hypers is not None else default_hypers()
        o = out if out is not None else {}

        def buf(name, *shape, dtype=torch.float64):
            t = o.get(name)
            if t is None or tuple(t.shape) != tuple(shape):
                t = self.empty(*shape, dtype=dtype)
                o[name] = t
            return t

        p = Problem()
        p.batch, p.n_rows, p.n_cols, p.n_special, p.n_chrono = b, n_rows, n, int(n_special), int(n_chrono)
        p.dop_start, p.dop_end = (dop_range if dop_range is not None else (-1, -1))
        p.vz_index = int(vz_index)
        p.vb_start, p.vb_end = vb_range
        p.hybrid = int(bool(hybrid))
        p.rm, p.rm_stride = _ptr(rm), (n_rows * n if rm.dim() == 3 else 0)
        p.rv = _ptr(rv)
        if vmm_eis is not None:
            p.vmm_eis, p.vmm_eis_stride = _ptr(vmm_eis), (vmm_eis.shape[-1] ** 2 if vmm_eis.dim() == 3 else 0)
        if vmm_chrono is not None:
            p.vmm_chrono, p.vmm_chrono_stride = _ptr(vmm_chrono), (n_chrono ** 2 if vmm_chrono.dim() == 3 else 0)
        p.pen, p.pen_stride = _ptr(pen), (3 * n * n if pen.dim() == 4 else 0)
        if pen_hint is not None and pen.dim() == 3:     # (first rows [3, n - n_special], band) from penalty_hint
            tz, band = pen_hint
            assert tz.is_cuda and tz.dtype == torch.float64 and tz.is_contiguous() and tuple(tz.shape) == (3, n - int(n_special))
            p.pen_toeplitz, p.pen_band = _ptr(tz), int(band)
        p.h, p.l1 = _ptr(h), _ptr(l1)
        p.vz_strength = _ptr(vz_strength)
        p.hyp = hyp
        p.x = _ptr(buf('x', b, n))
        p.weights = _ptr(buf('weights', b, n_rows))
        p.est_weights = _ptr(buf('est_weights', b, n_rows))
        p.init_weights = _ptr(buf('init_weights', b, n_rows))
        p.x_overfit = _ptr(buf('x_overfit', b, n))
        p.s_vectors = _ptr(buf('s_vectors', b, 3, n))
        p.rho = _ptr(buf('rho', b, 3))
        p.xmx_norms = _ptr(buf('xmx_norms', b, 3))
        if dop_range is not None:
            p.dop_rho = _ptr(buf('dop_rho', b, 3))
            p.dop_xmx_norms = _ptr(buf('dop_xmx_norms', b, 3))
        p.fun = _ptr(buf('fun', b))
        if vz_index >= 0:
            p.vz_col = _ptr(buf('vz_col', b, n_rows))
        if want_pq:
            p.p_matrix = _ptr(buf('p_matrix', b, n, n))
            p.q_vector = _ptr(buf('q_vector', b, n))
        if eval_mat is not None:       # post-fit diagnostics of the mapping path
            assert eval_mat.is_cuda and eval_mat.dtype == torch.float64 and eval_mat.is_contiguous() and eval_mat.shape[1] == n
            p.eval_mat, p.n_eval = _ptr(eval_mat), int(eval_mat.shape[0])
            p.dist_var = _ptr(buf('dist_var', b, eval_mat.shape[0]))
        if want_resid:
            p.resid_ss = _ptr(buf('resid_ss', b, 2))
        if hyp.has_outlier_p:
            p.outlier_t = _ptr(buf('outlier_t', b, n_rows))
        if hybrid and (hybrid_wf is not None or hyp.hybrid_wf_method or hyp.init_weights_separately):
            p.hybrid_wf_out = _ptr(buf('hybrid_wf', b, 2))
            if hybrid_wf is not None:          # per-spectrum chrono / EIS factors [B, 2]
                assert hybrid_wf.is_cuda and hybrid_wf.dtype == torch.float64 and tuple(hybrid_wf.shape) == (b, 2)
                p.hybrid_wf_in = _ptr(hybrid_wf.contiguous())
            if hyp.init_weights_separately:
                p.x_overfit_eis = _ptr(buf('x_overfit_eis', b, n))
        if weight_factor_vec is not None:
            assert weight_factor_vec.is_cuda and weight_factor_vec.dtype == torch.float64 and weight_factor_vec.numel() == n_rows
            p.weight_factor_vec = _ptr(weight_factor_vec)
        if hyp.solve_rp or hyp.update_scale:
            p.scale_factors = _ptr(buf('scale_factors', b, 3))
        if pfrt is not None:           # dict(factors=..., max_iter_per_step=10, min_iter=2, want_p=False)
            fac = self.dev(np.asarray(pfrt['factors'], dtype=float))
            o['pfrt_factors'] = fac
            nfac = fac.numel()
            p.n_pfrt, p.pfrt_factors = nfac, _ptr(fac)
            p.pfrt_max_iter, p.pfrt_min_iter = int(pfrt.get('max_iter_per_step', 10)), int(pfrt.get('min_iter', 2))
            p.pfrt_x = _ptr(buf('pfrt_x', b, nfac, n))
            p.pfrt_llh = _ptr(buf('pfrt_llh', b, nfac, 2))
            p.pfrt_iters = _ptr(buf('pfrt_iters', b, nfac, dtype=torch.int32))
            if pfrt.get('want_p'):
                p.pfrt_p = _ptr(buf('pfrt_p', b, nfac, n, n))
            if vz_index >= 0:
                p.vz_scratch = _ptr(buf('vz_scratch', b, n_rows))
        p.n_outer = _ptr(buf('n_outer', b, dtype=torch.int32))
        p.n_ipm = _ptr(buf('n_ipm', b, dtype=torch.int32))
        p.status = _ptr(buf('status', b, dtype=torch.int32))
        self._check(self.lib.hdrt_qphb_fit_batch(self.handle, C.byref(p), self._stream()))
        self.launches += 1
        return o

    def penalty_hint(self, pen, n_special):
        """Structure hint for the fit kernel (pen_toeplitz / pen_band of the ABI): when the DRT block of every penalty
        matrix in `pen` [3, n, n] is an exactly symmetric Toeplitz matrix, returns (first rows [3, nb] on the device,
        band), band = the largest |i - j| whose entry reaches 1e-45 of the largest one; otherwise None."""
        if pen.dim() != 3:
            return None
        blk = pen[:, n_special:, n_special:]
        nb = blk.shape[-1]
        if nb < 2:
            return None
        toep = bool(torch.equal(blk[:, 1:, 1:], blk[:, :-1, :-1])) and bool(torch.equal(blk, blk.transpose(1, 2)))
        if not toep:
            return None
        tz = blk[:, 0, :].contiguous()
        a = tz.abs().cpu().numpy()
        big = np.nonzero(a >= 1e-45 * a.max(axis=1, keepdims=True))[1]
        return tz, int(big.max()) if len(big) else 0

    def resolve_qp_batch(self, p, q, first_obs, my, param_scale, h, nr):
        """Batched cross-observation resolve QP (hdrt_resolve_qp_batch).  p [n_obs, nc, nc], q [n_obs, nc] device;
        first_obs [W] int; my [W, nr, nr]; param_scale [W, nc]; h [nc].  Returns dict(x [W, nr, nc], iters, status)."""
        p, q = p.contiguous(), q.contiguous()
        nc = q.shape[1]
        fo = self.dev(np.asarray(first_obs, dtype=np.int32), dtype=torch.int32)
        my_d, ps_d, h_d = self.dev(np.asarray(my, dtype=float)), self.dev(np.asarray(param_scale, dtype=float)), self.dev(np.asarray(h, dtype=float))
        nw = int(fo.numel())
        assert p.is_cuda and p.dtype == torch.float64 and tuple(p.shape[1:]) == (nc, nc) and tuple(my_d.shape) == (nw, nr, nr)
        x = self.empty(nw, nr, nc)
        iters = self.empty(nw, dtype=torch.int32)
        status = self.empty(nw, dtype=torch.int32)
        wb = int(self.lib.hdrt_resolve_work_bytes(self.handle, nw, nr, nc))
        if wb < 0:
            raise EngineError(f'resolve window of {nr} x {nc} unknowns is not supported')
        work = self.empty(max(wb // 8, 1))
        pr = ResolveProblem()
        pr.n_windows, pr.nr, pr.nc = nw, int(nr), int(nc)
        pr.p, pr.q, pr.first_obs, pr.my, pr.param_scale, pr.h = _ptr(p), _ptr(q), _ptr(fo), _ptr(my_d), _ptr(ps_d), _ptr(h_d)
        pr.x, pr.iters, pr.status = _ptr(x), _ptr(iters), _ptr(status)
        self._check(self.lib.hdrt_resolve_qp_batch(self.handle, C.byref(pr), _ptr(work), self._stream()))
        self.launches += 1
        return dict(x=x, iters=iters, status=status, _keep=(fo, my_d, ps_d, h_d, work))

    def probe_fp64(self):
        """Achieved DFMA TFLOP/s of this GPU (register-resident FMA loop on every SM)."""
        v = C.c_double()
        self._check(self.lib.hdrt_probe_fp64(self.handle, C.byref(v), self._stream()))
        self.launches += 2
        return v.value


_engines = {}


def get_engine(device=0):
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]
