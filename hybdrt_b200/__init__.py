"""Import alias for the ``hybrid-drt_b200/`` package directory.

The package lives in ``hybrid-drt_b200/`` (the name the layout contract fixes); a hyphen is not
importable, so this stub extends its own ``__path__`` onto that directory and executes the real
``__init__``.  ``import hybdrt_b200`` therefore gives the engine, ``hybdrt_b200.models.DRT`` the
drop-in model class, and so on.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'hybrid-drt_b200')
__path__.insert(0, _real)
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
del _f
