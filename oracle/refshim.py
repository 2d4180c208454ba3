"""CPU ORACLE support (test infrastructure, not product code).

Lets the UNMODIFIED reference package (``/root/reference/hybdrt``) import and run in the
authoring container, where four of its dependencies are absent:

* ``matplotlib``, ``skimage``, ``mitlef``, ``galvani`` -- import-time only on the gaussian-basis
  fit path; replaced by inert stub modules.
* ``cvxopt`` -- replaced by a shim whose ``solvers.qp`` is oracle/coneqp.py (the coneqp
  restatement), accepting exactly the call shape of hybdrt/models/qphb.py:512-519.

Only oracle/make_golden.py (run here, where /root/reference exists) uses this module. Nothing
that runs on the GPU box imports it.
"""
import importlib.abc
import importlib.machinery
import sys
import types

import numpy as np

from . import coneqp as _coneqp

_STUB_ROOTS = ('matplotlib', 'skimage', 'mitlef', 'galvani', 'mpl_toolkits')

# every solvers.qp call appends its iteration count here (read by make_golden.py)
QP_LOG = []


class _Inert:
    """Absorbs any attribute access / call made at import time."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return _Inert()

    def __call__(self, *a, **k):
        return _Inert()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return _Inert()


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split('.')[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = _StubModule(spec.name)
        mod.__path__ = []
        return mod

    def exec_module(self, module):
        pass


def _qp(P, q, G=None, h=None, A=None, b=None, solver=None, kktsolver=None, initvals=None, **kw):
    if A is not None or initvals is not None:
        raise NotImplementedError('shim covers the hot-path call shape only')
    P = np.asarray(P, dtype=float)
    q = np.asarray(q, dtype=float).ravel()
    G = np.asarray(G, dtype=float)
    h = np.asarray(h, dtype=float).ravel()
    n = q.size
    if G.shape != (n, n) or not np.array_equal(G, -np.eye(n)):
        raise NotImplementedError('shim covers G = -I only (qphb.py:472)')
    res = _coneqp.coneqp_orthant(P, q, h)
    QP_LOG.append(res['iterations'])
    return res


def install(reference_root='/root/reference'):
    """Register the stubs + cvxopt shim and put the reference on sys.path."""
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _StubFinder())
    if 'cvxopt' not in sys.modules:
        cvx = types.ModuleType('cvxopt')
        cvx.matrix = lambda a, *args, **kw: np.array(a, dtype=float)
        solvers = types.ModuleType('cvxopt.solvers')
        solvers.options = {}
        solvers.qp = _qp
        cvx.solvers = solvers
        sys.modules['cvxopt'] = cvx
        sys.modules['cvxopt.solvers'] = solvers
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
