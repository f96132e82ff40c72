"""Loader for the UNMODIFIED ihuston/pyflation reference (test infrastructure only).

The reference is Python-2 era source (``raise X, "msg"``, ``except X, e``, implicit
relative imports, ``StandardError``, removed NumPy aliases, a hard ``import tables``).
This module executes the reference's files *where they lie* under ``/root/reference``
through a handful of textual rewrites so they run on Python 3.12 / NumPy 2 without a
single byte of the reference being copied into this repository.

It exists for two jobs only:
  * ``tests/golden/make_golden.py`` -- generating the committed golden vectors;
  * ad-hoc validation of ``oracle/pyflation_oracle.py`` in the build container.

``/root/reference`` does not exist on the GPU box, so nothing that runs there (gpu tests,
``smoke()``, ``bench.py``) may import this file.  It is never imported by the product
package ``pyflation_b200``.
"""
import builtins
import os
import re
import sys
import types
import warnings

import numpy as np

REFERENCE_ROOT = os.environ.get("PYFLATION_REFERENCE", "/root/reference")

_LOAD_ORDER = [
    ("configuration", "configuration.py", False),
    ("cmpotentials", "cmpotentials.py", False),
    ("rk4", "rk4.py", False),
    ("helpers", "helpers.py", False),
    ("analysis", "analysis/__init__.py", True),
    ("cosmomodels", "cosmomodels.py", False),
]
_ANALYSIS_SUBMODULES = ["utilities", "nonadiabatic", "adiabatic"]

_REWRITES = [
    # raise ValueError, "text"   ->   raise ValueError("text")
    (re.compile(r'raise\s+(\w+)\s*,\s*("[^"]*")'), r"raise \1(\2)"),
    # except IOError, er:        ->   except IOError as er:
    (re.compile(r"except\s+([\w\.]+)\s*,\s*(\w+)\s*:"), r"except \1 as \2:"),
    # NumPy 2 rejects an index holding several Ellipsis objects
    (re.compile(r"\[Ellipsis\]\*len\(y\.shape\)"), "[slice(None)]*len(y.shape)"),
]


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyflation"))


def _patch_environment():
    for name, value in (("int", int), ("float", float), ("complex", complex),
                        ("NaN", np.nan)):
        if name not in np.__dict__:
            setattr(np, name, value)
    if "asscalar" not in np.__dict__:
        np.asscalar = lambda a: np.asarray(a).item()
    if not hasattr(builtins, "StandardError"):
        builtins.StandardError = Exception
    if not hasattr(builtins, "xrange"):
        builtins.xrange = range
    if "tables" not in sys.modules:
        stub = types.ModuleType("tables")

        class _Col(object):
            def __init__(self, *a, **k):
                pass

        for col in ("StringCol", "Float64Col", "IntCol", "ComplexCol", "Filters",
                    "Float64Atom", "ComplexAtom", "IsDescription"):
            setattr(stub, col, _Col)
        stub.openFile = stub.open_file = None
        sys.modules["tables"] = stub


def _exec_module(qualname, barename, relpath, is_pkg):
    path = os.path.join(REFERENCE_ROOT, "pyflation", relpath)
    with open(path, "r") as fh:
        src = fh.read()
    for pattern, repl in _REWRITES:
        src = pattern.sub(repl, src)
    mod = types.ModuleType(qualname)
    mod.__file__ = path
    if is_pkg:
        mod.__path__ = [os.path.dirname(path)]
        mod.__package__ = qualname
    sys.modules[qualname] = mod
    sys.modules[barename] = mod
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # py2-era docstrings: invalid escape sequences
        code = compile(src, path, "exec")
    exec(code, mod.__dict__)
    return mod


def load():
    """Return the reference's modules as a namespace:
    ``ref.cosmomodels``, ``ref.rk4``, ``ref.cmpotentials``, ``ref.helpers``, ``ref.analysis``."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if "pyflation.cosmomodels" in sys.modules and getattr(
            sys.modules["pyflation.cosmomodels"], "_via_ref_shim", False):
        return _namespace()
    _patch_environment()
    pkg = types.ModuleType("pyflation")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "pyflation")]
    sys.modules["pyflation"] = pkg
    for name, rel, is_pkg in _LOAD_ORDER:
        if name == "analysis":
            # the package __init__ does implicit relative imports of its submodules
            for sub in _ANALYSIS_SUBMODULES:
                m = _exec_module("pyflation.analysis." + sub, sub, "analysis/%s.py" % sub, False)
            apkg = _exec_module("pyflation.analysis", "analysis", rel, True)
            for sub in _ANALYSIS_SUBMODULES:
                setattr(apkg, sub, sys.modules["pyflation.analysis." + sub])
            setattr(pkg, "analysis", apkg)
        else:
            m = _exec_module("pyflation." + name, name, rel, is_pkg)
            setattr(pkg, name, m)
    sys.modules["pyflation.cosmomodels"]._via_ref_shim = True
    return _namespace()


def _namespace():
    ns = types.SimpleNamespace()
    for name in ("configuration", "cmpotentials", "rk4", "helpers", "analysis", "cosmomodels"):
        setattr(ns, name, sys.modules["pyflation." + name])
    return ns
