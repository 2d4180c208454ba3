"""Benchmark of the DRT hot path: hierarchical-Bayes DRT fits/s on synthetic EIS spectra.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on all host cores

Workload (BASELINE.json configs[1], SURVEY.md section 8d "C2"): 10,000 synthetic 2-ZARC spectra per GPU,
70 frequencies 1e6..1e-2 Hz x 101 Gaussian RBF basis functions, DRT() defaults (hierarchical-Bayes,
non-negative, interp-mode matrices), FP64.  One step = one pass of the hot path over that batch.
Scaling is weak (every rank fits its own 10,000 spectra; no data-path collective).

One JSON line on stdout (rank 0): value = fits/s with inputs resident in HBM (CUDA events on the launching
stream, max over ranks); e2e = the same through DRT.fit_eis_batch with HOST buffers (scaling on the host,
pinned H2D, fit, D2H of coefficients and weights inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH_PER_GPU = 10000
N_FREQ, N_BASIS, N_SPECIAL = 70, 101, 2
L2_FLUSH_BYTES = 256 << 20


def fit_flops(n_rows, n, nb, n_outer, n_ipm):
    """Algorithmic FP64 flops of a batch of fits from the per-spectrum iteration counts the kernel reports
    (SURVEY.md section 8d): per QP a Gram (2Nn^2 + 4Nn) and an interior-point start (n^3/3 + 2n^2); per
    interior-point iteration a Cholesky and four triangular solves (n^3/3 + 8n^2); per outer iteration the
    s/rho updates (24 Nb^2), the L2 assembly (6n^2) and the weights (2N^2 + 2Nn)."""
    n_qp = n_outer + 1.0                                   # + the initialize_weights QP
    per_qp = 2.0 * n_rows * n * n + 4.0 * n_rows * n + n ** 3 / 3.0 + 2.0 * n * n
    per_ipm = n ** 3 / 3.0 + 8.0 * n * n
    per_outer = 24.0 * nb * nb + 6.0 * n * n + 2.0 * n_rows * n_rows + 2.0 * n_rows * n
    return float(np.sum(n_qp * per_qp + n_ipm * per_ipm + n_outer * per_outer))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '200', '-i', str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: [self.lines.append(ln) for ln in self.proc.stdout], daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            p = [s.strip() for s in ln.split(',')]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), p[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (the reference itself is Python and cannot travel; see
# DESIGN.md).  Also used, on a bounded sample, for the cpu_baseline object of the GPU line.
# ---------------------------------------------------------------------------------------------------
_WORKER = {}


def _cpu_init(freq):
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle import drt_oracle as orc
    _WORKER['prep'] = orc.EisPrep(freq)


def _cpu_fit(z):
    res = _WORKER['prep'].fit(z)
    return res['n_outer'], int(res['ipm_iters'].sum()), np.asarray(res['x'], dtype=np.float64)


def parity_summary(cpu_out, x_gpu, n_outer_gpu, n_ipm_gpu):
    """The CPU oracle against the kernel on the same spectra (scaled space): fraction of spectra whose coefficient
    vector agrees within 1e-6 (norm-wise relative, SURVEY.md section 7 hard part 1), outer / interior-point iteration
    count mismatches, and the offenders."""
    n = len(cpu_out)
    rel = np.empty(n)
    out_mis, ipm_mis = [], []
    for i, (no, ni, x) in enumerate(cpu_out):
        rel[i] = np.max(np.abs(x_gpu[i] - x)) / np.max(np.abs(x))
        if int(no) != int(n_outer_gpu[i]):
            out_mis.append(i)
        if int(ni) != int(n_ipm_gpu[i]):
            ipm_mis.append(i)
    worst = np.argsort(-rel)[:5]
    return {'n': n, 'frac_x_within_1e-6': float(np.mean(rel <= 1e-6)), 'n_outer_mismatch': len(out_mis),
            'n_ipm_mismatch': len(ipm_mis), 'worst_rel': float(rel.max()), 'median_rel': float(np.median(rel)),
            'worst_spectra': [{'index': int(i), 'rel': float(rel[i])} for i in worst if rel[i] > 1e-6],
            'outer_mismatch_spectra': out_mis[:10], 'ipm_mismatch_spectra': ipm_mis[:10],
            'against': 'oracle/drt_oracle.py (numpy restatement + coneqp restatement) on the first n spectra of the batch'}


REF_ROOT = os.path.join(ROOT, 'baseline', '_ref')


def _ref_init(freq):
    """Worker of the reference arm: the UNMODIFIED reference package installed under baseline/_ref (pip --no-deps),
    imported on top of oracle/refshim.py (stubs for its absent plotting / file-format dependencies; cvxopt.solvers.qp
    -> the coneqp restatement).  One DRT() per worker, as BASELINE.md section 3 prescribes."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    import warnings
    warnings.filterwarnings('ignore')
    from oracle import refshim
    refshim.install(REF_ROOT)
    t0 = time.perf_counter()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):          # the reference prints while it builds its lookup tables
        from hybdrt.models import DRT
        drt = DRT()
    _WORKER['ref_drt'] = drt
    _WORKER['ref_freq'] = np.asarray(freq, dtype=float)
    _WORKER['ref_setup_s'] = time.perf_counter() - t0


def _ref_fit(z):
    drt = _WORKER['ref_drt']
    drt.fit_eis(_WORKER['ref_freq'], z)
    hist = drt.qphb_history
    return (len(hist), int(sum(e['cvx_result']['iterations'] for e in hist)),
            np.asarray(drt.cvx_result['x'], dtype=np.float64).ravel())


def _ref_setup(_):
    return _WORKER['ref_setup_s']


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, 'hybdrt'))


class CpuPool:
    """One worker process per core, each holding its own EisPrep (lookup tables built once per worker, as
    DRT() construction does in the reference; excluded from the timing)."""

    def __init__(self, freq, cores, reference=False):
        import multiprocessing as mp
        self.cores = cores
        self.fit = _ref_fit if reference else _cpu_fit
        self.pool = mp.get_context('spawn').Pool(cores, initializer=_ref_init if reference else _cpu_init, initargs=(freq,))
        self.setup_s = float(np.mean(self.pool.map(_ref_setup, range(cores)))) if reference else None

    def fits_per_second(self, z):
        self.pool.map(self.fit, list(z[:self.cores]))      # touch every worker (imports, tables)
        t0 = time.perf_counter()
        out = self.pool.map(self.fit, list(z), chunksize=max(1, len(z) // (self.cores * 8)))
        dt = time.perf_counter() - t0
        return len(z) / dt, dt, out

    def close(self):
        self.pool.close()
        self.pool.join()


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    """CPU arm.  The unmodified reference (baseline/_ref, under the import shim) on every host core when it is
    installed; the oracle port otherwise.  Every step fits the same fixed sample: the first n spectra of the seeded
    C2 batch the GPU arm fits."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from hybdrt_b200 import synth
    cores = host_cores()
    use_ref = reference_available() and not args.port
    n_sample = int(min(BATCH_PER_GPU, max(64, (16 if use_ref else 64) * cores)))
    freq, z = synth.make_eis_batch(BATCH_PER_GPU, seed=0)
    z = z[:n_sample]
    vals = []
    pool = CpuPool(freq, cores, reference=use_ref)
    out = []
    for i in range(args.warmup + args.steps):
        v, dt, out = pool.fits_per_second(z)
        if i >= args.warmup:
            vals.append((v, dt))
    setup_s = pool.setup_s
    pool.close()
    value = float(np.mean([v for v, _ in vals]))
    if use_ref:
        kind = 'reference'
        how = (f'the unmodified reference package (baseline/_ref, pip --no-deps) under oracle/refshim.py: DRT().fit_eis per '
               f'spectrum, one DRT() per worker process ({setup_s:.2f} s construction incl. lookup tables, not timed), BLAS '
               f'threads = 1; cvxopt is not installable here: cvxopt.solvers.qp = the coneqp restatement (oracle/coneqp.py)')
    else:
        kind = 'port'
        how = ('numpy oracle (oracle/drt_oracle.py), one process per core, BLAS threads = 1; QP = coneqp restatement '
               '(cvxopt not installable)')
    line = {
        'impl': 'reference', 'metric': 'DRT fits/sec (70f x 101 basis, hierarchical-Bayes, FP64)', 'value': value,
        'unit': 'fits/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': float(np.mean([dt for _, dt in vals]) * 1e3), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'C2: 10,000 synthetic 2-ZARC EIS spectra, 70 freqs x 101 RBF basis, DRT() defaults',
                   'sample': f'the first {n_sample} of the 10,000 spectra, every step',
                   'mean_outer_iters': float(np.mean([o[0] for o in out])), 'mean_ipm_iters': float(np.mean([o[1] for o in out]))},
        'cpu_baseline': {'value': value, 'unit': 'fits/s', 'cores': cores, 'kind': kind,
                         'sample': f'first {n_sample} spectra of the seeded C2 batch; {how}'},
        'e2e': {'value': value, 'unit': 'fits/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def builder_rooflines(eng, drt, flush, hbm_peak):
    """Achieved HBM write bandwidth of the other matrix builders (algorithmic bytes of SURVEY.md section 8d over the
    CUDA-event time of one launch on n_grids independent grids, L2 flushed): the chrono step-response matrix (C3 shape),
    the dense chrono variance matrix, the penalty matrices, the DOP columns, the trapz-mode impedance matrices (FP64 /
    transcendental bound: reported with its evaluation rate)."""
    import torch
    from hybdrt_b200 import engine as E, synth

    def timed(fn, nbytes, reps=3):
        fn()
        fn()
        ms = []
        for k in range(reps):
            flush.fill_(k)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        sec = float(np.mean(ms)) * 1e-3
        return {'achieved': nbytes / sec / 1e9, 'unit': 'GB/s', 'frac': nbytes / sec / 1e9 / hbm_peak, 'ms': sec * 1e3,
                'bytes_per_launch': nbytes}

    res = {}
    eps = drt.tau_epsilon
    tab = drt.interpolate_lookups
    t3, i3, _, _, _ = synth.make_hybrid_batch(1, seed=1)
    tau = np.logspace(-7, 2, 94)
    g = 512
    times = eng.dev(np.repeat(t3[None], g, 0) * (1 + 1e-4 * np.arange(g)[:, None] / g))
    taus = eng.dev(np.repeat(tau[None], g, 0))
    st, sa = eng.dev(np.zeros((g, 1))), eng.dev(np.full((g, 1), 1e-2))
    res['response_interp_kernel'] = dict(timed(lambda: eng.build_response(times, taus, st, sa, eps, E.MODE_INTERP, tab), 8.0 * g * len(t3) * len(tau)),
                                         workload=f'{g} grids, {len(t3)} samples x {len(tau)} basis (C3 shape)')
    g = 16
    res['chrono_vmm_kernel'] = dict(timed(lambda: eng.build_chrono_vmm(times[:g], st[:g], 4.0), 8.0 * g * len(t3) ** 2),
                                    workload=f'{g} grids, {len(t3)}^2 dense variance matrix')
    g = 8192
    grid = eng.dev(np.repeat(np.log(np.logspace(-7, 3, 101))[None], g, 0) * (1 + 1e-6 * np.arange(g)[:, None] / g))
    res['penalty_kernel'] = dict(timed(lambda: eng.build_penalty(grid, eps, False), 24.0 * g * 101 ** 2),
                                 workload=f'{g} grids, M0..M2 101 x 101 (general path, no Toeplitz fill)')
    g = 8192
    freq = eng.dev(np.repeat(synth.C2_FREQ[None], g, 0) * (1 + 1e-3 * np.arange(g)[:, None] / g))
    nu = np.linspace(-1, 1, 51)
    res['dop_z_kernel'] = dict(timed(lambda: eng.build_dop_z(freq, nu, 5.0), 16.0 * g * 70 * len(nu)),
                               workload=f'{g} grids, 70 freqs x {len(nu)} phasance basis (complex erf per entry)')
    g = 64
    tz = timed(lambda: eng.build_impedance(freq[:g], eng.dev(np.repeat(np.logspace(-7, 3, 101)[None], g, 0)), eps, E.MODE_TRAPZ), 16.0 * g * 70 * 101)
    tz['integrand_evaluations_per_s'] = 2.0 * g * 70 * 101 * 1000 / (tz['ms'] * 1e-3)
    res['impedance_trapz_kernel'] = dict(tz, bound='fp64 + transcendental (1000-point quadrature per entry), not HBM',
                                         workload=f'{g} grids, 70 x 101, A_re + A_im')
    return res


def other_configs(eng, fp64_peak, flush):
    """BASELINE configs C3 (hybrid) and C4 (DRT+DOP) at full size: the fit kernel re-launched on the inputs already
    resident in HBM (the launch the public API made), CUDA events, FP64 roofline fraction from the same flop model."""
    import torch
    from hybdrt_b200 import synth
    from hybdrt_b200.models import DRT
    res = {}

    def measure(name, fit, workload):
        r = fit()
        torch.cuda.synchronize()
        relaunch, raw = r.extra['relaunch'], {}
        relaunch(raw)
        ms = []
        for k in range(3):
            flush.fill_(k)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            relaunch(raw)
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        n_outer = raw['n_outer'].cpu().numpy().astype(np.float64)
        n_ipm = raw['n_ipm'].cpu().numpy().astype(np.float64)
        plan = r.plan
        flops = fit_flops(plan['n_rows'], plan['n'], plan['n'] - plan['n_special'], n_outer, n_ipm)
        sec = float(np.mean(ms)) * 1e-3
        nb = len(n_outer)
        res[name] = {'workload': workload, 'fits_per_s': nb / sec, 'ms_per_launch': sec * 1e3, 'n_rows': plan['n_rows'],
                     'n_cols': plan['n'], 'mean_outer_iters': float(n_outer.mean()), 'mean_ipm_iters': float(n_ipm.mean()),
                     'frac_converged': float(np.mean((raw['status'].cpu().numpy() & 1) > 0)),
                     'roofline': {'bound': 'fp64', 'achieved': flops / sec / 1e12, 'peak': fp64_peak, 'unit': 'TFLOP/s',
                                  'frac': flops / sec / 1e12 / fp64_peak, 'flops_per_launch': flops}}

    times, i_sig, v, freq3, z3 = synth.make_hybrid_batch(4096, seed=1)
    d3 = DRT()
    measure('C3', lambda: d3.fit_hybrid_batch(times, i_sig, v, freq3, z3),
            '4096 hybrid fits: 2000-sample chrono step response + 30 high-frequency EIS points each')
    del d3
    freq4, z4 = synth.make_dop_batch(10000, seed=2)
    d4 = DRT(fit_dop=True)
    measure('C4', lambda: d4.fit_eis_batch(freq4, z4), '10000 DRT+DOP fits, 70 frequencies')
    return res


# ---------------------------------------------------------------------------------------------------
# C5: the 256 x 256 map through DRTMD, sharded over the ranks (strong scaling), result gather inside the timed region
# ---------------------------------------------------------------------------------------------------
def run_c5(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as graft
    graft.build()
    from hybdrt_b200 import engine as E, synth
    from hybdrt_b200.mapping import DRTMD

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    eng = E.get_engine(local)
    rows = cols = args.map_size
    freq, z = synth.make_map_batch(rows, cols, seed=3)
    psi = np.array([(r, c) for r in range(rows) for c in range(cols)], dtype=float)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        md = DRTMD(tau_supergrid=np.logspace(-8, 3, 111), psi_dim_names=['row', 'col'], print_progress=False, device=local)
        md.add_observations(psi, freq, z)
        md.fit_all(ignore_errors=True, shard=world > 1)       # every rank ends with the full map (all-gather over NCCL)
        return md

    for _ in range(max(1, min(args.warmup, 2))):
        md = step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = eng.launches
    t0 = time.perf_counter()
    for _ in range(args.steps):
        md = step()
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=eng.device)
    launches = eng.launches - launches0
    clocks = sampler.stop()
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    total_s = float(dt.item())
    n_obs = rows * cols
    value = n_obs * args.steps / total_s
    if rank == 0:
        per_obs = (md.obs_x.shape[-1] + md.obs_drt_var.shape[-1] + 6) * 8
        line = {
            'metric': 'DRT fits/sec (70f x 101 basis, hierarchical-Bayes, FP64)', 'value': value, 'unit': 'fits/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_s / args.steps * 1e3,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': f'C5: {rows} x {cols} map of synthetic 2-ZARC spectra through DRTMD.fit_all (fit + '
                                   f'distribution variance + llh + rss per observation), interleaved shards, results '
                                   f'all-gathered over NCCL inside the timed region', 'observations': n_obs,
                       'fitted': int(md.obs_fit_status.sum()), 'mean_outer_iters': float(md.obs_outer_iterations.mean()),
                       'timing': 'host wall clock around fit_all + barrier, max over ranks (host buffers in, host arrays out)'},
            'e2e': {'value': value, 'unit': 'fits/s', 'h2d_bytes_per_step': int(z.shape[1] * 16 * n_obs // world),
                    'd2h_bytes_per_step': int(per_obs * n_obs // world)},
            'gpu_launches': launches, 'clocks': clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as graft
    graft.build()
    from hybdrt_b200 import engine as E, synth
    from hybdrt_b200.models import DRT

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    eng = E.get_engine(local)
    dev = eng.device

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.batch
    freq, z = synth.make_eis_batch(B, seed=rank)            # every rank fits its own spectra (weak scaling)
    drt = DRT(device=local)

    # ---- resident-input path: same plan + launch the public API uses, inputs already in HBM
    res0 = drt.fit_eis_batch(freq, z)
    plan = res0.plan
    scale = res0.scales['coefficient_scale']
    zs = z / scale[:, None]
    rv_dev = eng.dev(np.concatenate([zs.real, zs.imag], axis=1))
    hyp = drt._c_hypers(plan['opts'])
    out = {}
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def step_resident():
        eng.qphb_fit_batch(plan['rm'], rv_dev, plan['pen'], plan['h'], plan['l1'], plan['n_special'],
                           vmm_eis=plan['vmm_eis'], hypers=hyp, out=out, pen_hint=plan.get('pen_hint'))

    from hybdrt_b200 import sharding

    def step_e2e():
        r = drt.fit_eis_batch(freq, z)
        fp = r.fit_parameters()                             # D2H of x and weights + unscaling on the host
        if world > 1:                                       # the path's only collective: results to rank 0 (NCCL)
            sharding.gather_results({'x': r.raw['x'], 'weights': r.raw['weights']}, world * B, dst=0, copy=False)
        return r, fp

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.fill_(k)                                      # evict L2 between timed steps
        ev[k][0].record()
        step_resident()
        ev[k][1].record()
    barrier()
    launches = eng.launches - launches0
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) * 1e-3
    value = world * B * args.steps / total_s

    n_outer = out['n_outer'].cpu().numpy().astype(np.float64)
    n_ipm = out['n_ipm'].cpu().numpy().astype(np.float64)
    status = out['status'].cpu().numpy()
    flops = fit_flops(plan['n_rows'], plan['n'], plan['n'] - plan['n_special'], n_outer, n_ipm)
    kernel_s = float(np.mean(step_ms)) * 1e-3               # one qphb launch per step

    # ---- end-to-end path through the public API with host buffers
    for _ in range(min(2, args.warmup)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        r_e2e, fp = step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(e2e_s.item())
    h2d = int(r_e2e.extra['h2d_bytes'])
    d2h = int(sum(v.nbytes for k, v in r_e2e.host(['x', 'weights']).items() if k in ('x', 'weights')))

    # final gather of per-rank summaries (the only collective; off the timed path)
    summ = torch.tensor([float(np.mean(n_outer)), float(np.mean(n_ipm)), float(np.mean((status & 1) > 0))],
                        dtype=torch.float64, device=dev)
    if world > 1:
        gathered = [torch.zeros_like(summ) for _ in range(world)] if rank == 0 else None
        dist.gather(summ, gathered, dst=0)
        if rank == 0:
            summ = torch.stack(gathered).mean(dim=0)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        fp64_peak = eng.probe_fp64()
        # secondary roofline: matrix construction with per-spectrum grids (HBM-write bound)
        g = B
        f_dev = eng.dev(np.repeat(freq[None], g, 0) * (1 + 1e-3 * np.arange(g)[:, None] / g))
        t_dev = eng.dev(np.repeat(plan['basis_tau'][None], g, 0))
        for _ in range(3):
            eng.build_impedance(f_dev, t_dev, drt.tau_epsilon, E.MODE_INTERP, drt.interpolate_lookups)
        mb = []
        for k in range(5):
            flush.fill_(k)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.build_impedance(f_dev, t_dev, drt.tau_epsilon, E.MODE_INTERP, drt.interpolate_lookups)
            b.record()
            torch.cuda.synchronize()
            mb.append(a.elapsed_time(b))
        mat_bytes = 16.0 * g * N_FREQ * N_BASIS
        hbm_peak = peaks.get('hbm_gbs', 6650.0)
        mat_gbs = mat_bytes / (float(np.mean(mb)) * 1e-3) / 1e9

        # secondary roofline: chrono conditioning (antialiasing filter + decimation) of a batch of raw traces
        # (HBM-read bound: every raw sample is read once, ~1 % of them are written back)
        from hybdrt_b200 import preprocessing as pp
        rt, ri, rvv = synth.make_raw_chrono_batch(1, seed=5)
        n_tr = 4096
        raw_dev = eng.dev(np.repeat(rvv, n_tr, 0) + 1e-6 * np.arange(n_tr)[:, None])
        st_raw = rt[pp.identify_steps(ri, True)]
        dec = pp.get_decimation_index(rt, st_raw, np.min(np.diff(rt)), 25, 8, 2, None)
        fplan = pp.filter_plan(rt, pp.identify_steps(ri, allow_consecutive=False), dec)
        for _ in range(3):
            eng.filter_gather(raw_dev, fplan)
        fb = []
        for k in range(5):
            flush.fill_(k)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.filter_gather(raw_dev, fplan)
            b.record()
            torch.cuda.synchronize()
            fb.append(a.elapsed_time(b))
        filt_bytes = 8.0 * n_tr * (len(rt) + len(dec))
        filt_gbs = filt_bytes / (float(np.mean(fb)) * 1e-3) / 1e9
        del raw_dev

        mat_traffic = filt_traffic = None
        try:        # ncu captures of the two secondary kernels, scaled to these launches
            bt = json.load(open(os.path.join(ROOT, 'profiles', 'builder_traffic.json')))
            mi, fi = bt['impedance_interp_kernel'], bt['filter_gather_kernel']
            mat_traffic = (mi['dram_bytes_read'] + mi['dram_bytes_write']) / mi['grids_in_profiled_launch'] * g
            filt_traffic = (fi['dram_bytes_read'] + fi['dram_bytes_write']) / fi['traces_in_profiled_launch'] * n_tr
        except Exception:
            pass
        traffic = None
        prof = os.path.join(ROOT, 'profiles', 'qphb_traffic.json')
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get('dram_bytes_per_fit') * B   # ncu capture, scaled to this launch
            except Exception:
                traffic = None

        cpu = parity = None
        if world == 1 and not args.no_cpu_baseline:
            cores = host_cores()
            n_sample = int(min(B, max(1024, 96 * cores)))
            pool = CpuPool(freq, cores)
            v, dt, cpu_out = pool.fits_per_second(z[:n_sample])
            pool.close()
            cpu = {'value': v, 'unit': 'fits/s', 'cores': cores, 'kind': 'port',
                   'sample': f'first {n_sample} spectra of the same batch ({dt:.1f} s); numpy oracle, one process per core, '
                             f'BLAS threads = 1; QP = coneqp restatement (cvxopt not installable here)'}
            parity = parity_summary(cpu_out, out['x'][:n_sample].cpu().numpy(), n_outer[:n_sample], n_ipm[:n_sample])
            if reference_available():      # second figure: the unmodified reference under the import shim, small sample
                n_ref = int(min(B, 8 * cores))
                rpool = CpuPool(freq, cores, reference=True)
                rv_, rdt, rout = rpool.fits_per_second(z[:n_ref])
                rpool.close()
                rpar = parity_summary(rout, out['x'][:n_ref].cpu().numpy(), n_outer[:n_ref], n_ipm[:n_ref])
                cpu['reference_shim'] = {'value': rv_, 'unit': 'fits/s', 'cores': cores, 'kind': 'reference',
                                         'sample': f'first {n_ref} spectra ({rdt:.1f} s): unmodified reference package '
                                                   f'(baseline/_ref) under oracle/refshim.py, one DRT() per worker',
                                         'parity_vs_kernel': {k: rpar[k] for k in ('n', 'frac_x_within_1e-6', 'n_outer_mismatch',
                                                                                    'worst_rel')}}

        achieved = flops / kernel_s / 1e12
        line = {
            'metric': 'DRT fits/sec (70f x 101 basis, hierarchical-Bayes, FP64)', 'value': value, 'unit': 'fits/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': float(total_s / args.steps * 1e3),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': f'C2: {B} synthetic 2-ZARC EIS spectra per GPU, 70 freqs 1e6-1e-2 Hz x 101 RBF basis, '
                                   f'DRT() defaults (non-negative hierarchical-Bayes, interp matrices)',
                       'batch_per_gpu': B, 'n_rows': plan['n_rows'], 'n_cols': plan['n'],
                       'l2': f'flushed between timed steps ({L2_FLUSH_BYTES >> 20} MiB write)',
                       'mean_outer_iters': float(summ[0]), 'mean_ipm_iters': float(summ[1]),
                       'frac_converged': float(summ[2])},
            'e2e': {'value': e2e_value, 'unit': 'fits/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
            'gpu_launches': launches,
            'clocks': clocks,
            'roofline': {'bound': 'fp64', 'achieved': achieved, 'peak': fp64_peak, 'unit': 'TFLOP/s',
                         'frac': achieved / fp64_peak, 'traffic': traffic,
                         'traffic_source': 'static: ncu dram bytes per fit of profiles/qphb_traffic.json x this batch '
                                           '(not a same-run counter)',
                         'kernel': 'qphb_warp_kernel (DMMA.8x8x4 + DFMA, one warp per spectrum)',
                         'peak_source': 'measured in this run: register-resident DFMA loop on every SM '
                                        '(MEASURED_PEAKS.json has no FP64 entry)',
                         'flops_per_launch': flops},
            'roofline_matrix_build': {'bound': 'hbm', 'achieved': mat_gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                                      'frac': mat_gbs / hbm_peak, 'traffic': mat_traffic, 'kernel': 'impedance_interp_kernel',
                                      'bytes_per_launch': mat_bytes,
                                      'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if 'hbm_gbs' in peaks else 'fallback',
                                      'workload': f'{g} per-spectrum (freq, tau) grids, A_re + A_im 70 x 101 each'},
            'roofline_chrono_filter': {'bound': 'hbm', 'achieved': filt_gbs, 'peak': hbm_peak, 'unit': 'GB/s',
                                       'frac': filt_gbs / hbm_peak, 'traffic': filt_traffic, 'kernel': 'filter_gather_kernel',
                                       'bytes_per_launch': filt_bytes,
                                       'workload': f'{n_tr} raw traces x {len(rt)} samples -> {len(dec)} kept samples each '
                                                   f'(downsample_data, decimation_interval=8, factor 2)'},
            'roofline_builders': builder_rooflines(eng, drt, flush, hbm_peak) if (world == 1 and not args.no_configs) else None,
            'cpu_baseline': cpu,
            'parity': parity,
            'configs': other_configs(eng, fp64_peak, flush) if (world == 1 and not args.no_configs) else None,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=BATCH_PER_GPU, help='spectra per GPU (default: the C2 workload)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--port', action='store_true', help='--impl reference: time the oracle port even if baseline/_ref exists')
    ap.add_argument('--no-configs', action='store_true', help='skip the C3 / C4 measurements of the default line')
    ap.add_argument('--config', default='c2', choices=['c2', 'c5'],
                    help='c2: weak-scaled batch of 10,000 spectra per GPU (default); c5: 256 x 256 map, strong scaling')
    ap.add_argument('--map-size', type=int, default=256)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'b200':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    elif args.config == 'c5':
        run_c5(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
