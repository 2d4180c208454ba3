"""CPU ORACLE (test infrastructure, not product code).

Restatement of the interior-point method that the reference reaches through
``cvxopt.solvers.qp`` (reference call site: hybdrt/models/qphb.py:512-519, with
``G = -I`` built at qphb.py:472 and ``h`` from ``make_h_constraint`` qphb.py:521-557).

cvxopt itself is a third-party dependency that is NOT vendored under /root/reference and is
not installable here (requirements.txt:4 / setup.py:15 list a bare, unpinned ``cvxopt``).
This module restates the published algorithm of ``cvxopt.coneprog.coneqp`` (cvxopt 1.3.x) for
the only cone the hot path uses: the non-negative orthant with ``G = -I``, no equality
constraints, default options (``abstol 1e-7, reltol 1e-6, feastol 1e-7, maxiters 100``,
``refinement 0``, KKT solver ``chol2``).

The method is Mehrotra predictor-corrector in Nesterov-Todd scaled variables:

    W = diag(d), d = sqrt(s/z), lambda = sqrt(s*z)
    KKT:  (P + diag(1/d^2)) ux = bx - (bz/d)/d ;  W uz = -(ux)/d - bz/d
    step length 0.99 / t, centering exponent 3.

The reference answer is the early-stopped central-path iterate this method returns, NOT the
exact QP optimum, so the operation order below is kept close to cvxopt's.

Parity status: pinned by the reference's own golden vector (tests/test_drt_fit.py:55-141) when
the unmodified reference runs on top of this solver -- see oracle/make_golden.py and
tests/test_oracle_golden.py.
"""
import math

import numpy as np
from scipy.linalg import cholesky, solve_triangular

ABSTOL = 1e-7
RELTOL = 1e-6
FEASTOL = 1e-7
MAXITERS = 100
STEP = 0.99
EXPON = 3


class KKTError(ArithmeticError):
    pass


def _factor(P, di):
    """Cholesky of S = P + Gs'Gs with Gs = -diag(di) (cvxopt misc.kkt_chol2 'factor')."""
    S = P.copy()
    S[np.diag_indices_from(S)] += di * di
    try:
        return cholesky(S, lower=True, check_finite=False)
    except np.linalg.LinAlgError as err:  # pragma: no cover - depends on data
        raise KKTError(str(err))


def _kkt_solve(L, di, bx, bz):
    """cvxopt misc.kkt_chol2 'solve' for G = -I: returns (ux, W*uz)."""
    zs = di * bz                      # z := W^{-T} bz
    x = bx - di * zs                  # x := bx + Gs' z
    x = solve_triangular(L, x, lower=True, check_finite=False)
    x = solve_triangular(L, x, lower=True, trans='T', check_finite=False)
    z = -di * x - zs                  # z := Gs x - z
    return x, z


def coneqp_orthant(P, q, h, abstol=ABSTOL, reltol=RELTOL, feastol=FEASTOL, maxiters=MAXITERS,
                   trace=None):
    """Solve  min 1/2 x'Px + q'x  s.t.  -x <= h  the way cvxopt.coneqp does.

    Returns a dict with the cvxopt result keys the reference reads ('x', 'primal objective',
    'status', ...) plus 'iterations'.
    """
    P = np.asarray(P, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64).ravel()
    h = np.asarray(h, dtype=np.float64).ravel()
    n = q.size

    resx0 = max(1.0, math.sqrt(q @ q))
    resz0 = max(1.0, math.sqrt(h @ h))

    # Initial point: W = I
    ones = np.ones(n)
    try:
        L = _factor(P, ones)
    except KKTError:
        raise ValueError("Rank(A) < p or Rank([P; A; G]) < n")
    x, z = _kkt_solve(L, ones, -q, h)
    s = -z
    nrms = math.sqrt(s @ s)
    ts = -s.min()
    if ts >= -1e-8 * max(nrms, 1.0):
        s = s + (1.0 + ts)
    nrmz = math.sqrt(z @ z)
    tz = -z.min()
    if tz >= -1e-8 * max(nrmz, 1.0):
        z = z + (1.0 + tz)

    gap = float(s @ z)
    d = di = lmbda = None
    status = 'unknown'
    iters = 0
    for iters in range(maxiters + 1):
        # rx = P x + q + G'z ;  f0 = 1/2 x'Px + q'x
        rx = P @ x + q
        f0 = 0.5 * (x @ rx + x @ q)
        rx = rx - z
        resx = math.sqrt(rx @ rx)
        # rz = s + G x - h
        rz = s - h - x
        resz = math.sqrt(rz @ rz)

        pcost = f0
        dcost = f0 + z @ rz - gap
        if pcost < 0.0:
            relgap = gap / -pcost
        elif dcost > 0.0:
            relgap = gap / dcost
        else:
            relgap = None
        pres = resz / resz0
        dres = resx / resx0
        if trace is not None:
            trace.append(dict(it=iters, pcost=pcost, dcost=dcost, gap=gap, pres=pres, dres=dres,
                              x=x.copy()))

        converged = (pres <= feastol and dres <= feastol and
                     (gap <= abstol or (relgap is not None and relgap <= reltol)))
        if converged or iters == maxiters:
            status = 'optimal' if converged else 'unknown'
            break

        if iters == 0:
            d = np.sqrt(s / z)
            di = d ** -1
            lmbda = np.sqrt(s * z)
        lmbdasq = lmbda * lmbda

        try:
            L = _factor(P, di)
        except KKTError:
            status = 'unknown'
            break

        mu = gap / n
        sigma = 0.0
        step = 1.0
        ws3 = None
        for i in (0, 1):
            ds = np.zeros(n)
            if i == 1:
                ds = ds - ws3
            ds = ds - lmbdasq
            ds = ds + sigma * mu
            dx = -rx
            dz = -rz
            # f4_no_ir
            ds = ds / lmbda
            dz = dz - d * ds
            dx, dz = _kkt_solve(L, di, dx, dz)
            ds = ds - dz

            dsdz = float(ds @ dz)
            if i == 0:
                ws3 = ds * dz
            ds = ds / lmbda
            dz = dz / lmbda
            ts = -ds.min()
            tz = -dz.min()
            t = max(0.0, ts, tz)
            if t == 0:
                step = 1.0
            elif i == 0:
                step = min(1.0, 1.0 / t)
            else:
                step = min(1.0, STEP / t)
            if i == 0:
                sigma = min(1.0, max(0.0, 1.0 - step + dsdz / gap * step ** 2)) ** EXPON

        x = x + step * dx
        ds = step * ds + 1.0
        dz = step * dz + 1.0
        ds = ds * lmbda
        dz = dz * lmbda
        # update_scaling
        sq_s = np.sqrt(ds)
        sq_z = np.sqrt(dz)
        d = d * sq_s / sq_z
        di = d ** -1
        lmbda = sq_s * sq_z
        s = lmbda * d
        z = lmbda * di
        gap = float(lmbda @ lmbda)

    return {
        'x': x, 's': s, 'z': z, 'status': status, 'gap': gap, 'relative gap': relgap,
        'primal objective': pcost, 'dual objective': dcost,
        'primal infeasibility': pres, 'dual infeasibility': dres,
        'iterations': iters,
    }
