"""pyflation_b200 - B200-native first-order perturbation hot path of ihuston/pyflation.

Module names follow the reference package so that ``from pyflation_b200 import cosmomodels``
is a drop-in for ``from pyflation import cosmomodels`` on the first-order path.
"""
__version__ = "0.1.0"

from . import configuration, helpers, cmpotentials, analysis, rk4, cosmomodels  # noqa: F401
