"""pyflation - import name of the reference package (/pyflation/__init__.py), resolved to the
B200-native implementation of its first-order path.

``from pyflation import cosmomodels`` (and ``rk4``, ``cmpotentials``, ``helpers``,
``configuration``, ``run_config``, ``sohelpers``, ``analysis`` with ``analysis.adiabatic`` /
``analysis.utilities``) gives the very module objects of :mod:`pyflation_b200`, so code written
against the reference runs on the GPU path without editing its imports.  Modules of the
reference outside the first- / second-order solver path (``sourceterm``, ``solutions``,
``cosmographs``, ``analysis.nonadiabatic``) are out of scope and raise ImportError.
"""
import importlib
import sys

import pyflation_b200 as _impl

__version__ = "0.2.3"                      # the reference version whose API is mirrored
__b200_version__ = _impl.__version__

for _name in ("configuration", "helpers", "cmpotentials", "rk4", "analysis", "cosmomodels", "run_config", "sohelpers"):
    _mod = importlib.import_module("pyflation_b200." + _name)
    sys.modules[__name__ + "." + _name] = _mod
    globals()[_name] = _mod
for _name in ("adiabatic", "utilities"):
    sys.modules[__name__ + ".analysis." + _name] = importlib.import_module("pyflation_b200.analysis." + _name)
del _name, _mod
