"""Importable alias of the ``csmri-refinement_b200/`` package directory.

The package lives in ``csmri-refinement_b200/`` (the name the layout contract
asks for); a hyphen cannot appear in a Python module name, so this shim points
``__path__`` at that directory and runs its ``__init__``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      'csmri-refinement_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
del _os, _f, _real
