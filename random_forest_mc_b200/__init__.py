"""Importable alias for the engine package.

The product lives in ``random-forest-mc_b200/`` (the directory name the project layout mandates;
a hyphen is not a valid Python identifier).  This shim makes ``import random_forest_mc_b200``
resolve every submodule from that directory and re-exports its public names.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "random-forest-mc_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"), globals())
del _f
