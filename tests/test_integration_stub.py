"""INTEGRATION.md section B is executable: the ctypes stub a maintainer of the reference would
paste into pyflation/rk4.py is extracted from the document and run against the library, fed
with the reference's own start vectors, and compared with the reference's trajectory."""
import os
import re
import types

import numpy as np
import pytest

from conftest import ROOT, load_fo, rel_err, mode_rel_err

pytestmark = pytest.mark.gpu


def test_documented_ctypes_stub_reproduces_reference():
    from pyflation_b200 import _lib
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n# pyflation/rk4.py  -- additions\n(.*?)```", text, re.S).group(1)
    block = block.replace('"libpyflation_b200.so"', repr(_lib.LIB_PATH))

    class SimRunError(Exception):
        pass

    ns = {"SimRunError": SimRunError}
    exec(compile(block, "INTEGRATION.md", "exec"), ns)
    for name in ("c2_lambdaphi4_2p16", "c3_hybridquadratic_2p20"):
        g = load_fo(name)
        model = types.SimpleNamespace(potential_func=g["potential"], pot_params=dict(g["params"]),
                                      nfields=g["nfields"], ainit=g["ainit"], k=g["k"])
        derivs = types.SimpleNamespace(__self__=model)
        x, y = ns["rkdriver_tsix"](g["foystart"], 0.0, g["fotstartindex"], float(g["fotend"]), 0.01, derivs)
        nf = g["nfields"]
        assert y.shape == (g["T"], 2 * nf * nf + 2 * nf + 1, len(g["k"]))
        assert rel_err(x, g["tresult"]) == 0.0
        assert rel_err(y[g["rows"]][:, :2 * nf + 1], g["yrows"][:, :2 * nf + 1]) < 1e-9
        assert mode_rel_err(y[g["rows"]], g["yrows"], nf) < 1e-9


def test_documented_second_order_stub_reproduces_reference(oracle):
    """Section D continues the stub for the second-order stage (pf_so_run): executed against the
    reference's golden second-order rows, potentials from the drop-in cmpotentials module."""
    from pyflation_b200 import _lib, cmpotentials
    from test_second_order import load_so, fo_trajectory, col_err
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    first = re.search(r"```python\n# pyflation/rk4.py  -- additions\n(.*?)```", text, re.S).group(1)
    second = re.search(r"```python\n# pyflation/rk4.py  -- second-order additions\n(.*?)```", text, re.S).group(1)
    first = first.replace('"libpyflation_b200.so"', repr(_lib.LIB_PATH))

    class SimRunError(Exception):
        pass

    ns = {"SimRunError": SimRunError}
    exec(compile(first + "\n" + second, "INTEGRATION.md", "exec"), ns)
    for name in ("so_msqphisq", "so_lambdaphi4"):
        g = load_so(name)
        fo = types.SimpleNamespace(yresult=fo_trajectory(g), simtstart=float(g["fo_simtstart"]),
                                   tstep_wanted=float(g["fo_tstep"]), bgepsilon=g["bgepsilon"],
                                   fotstartindex=g["fotstartindex"])
        for tag in ("plain", "ramped"):
            so = types.SimpleNamespace(second_stage=fo, potentials=getattr(cmpotentials, g["potential"]),
                                       pot_params=dict(g["params"]), ainit=g["ainit"], k=g["k"],
                                       source=oracle.synthetic_source(g["fo_tresult"], g["k"]),
                                       tstart=g["fotstart"])
            if tag == "ramped":
                so.rampargs = dict(g["ramp"])
            derivs = types.SimpleNamespace(__self__=so)
            x, y = ns["rkdriver_tsix_secondorder"](np.zeros((4, len(g["k"]))), float(g["so_simtstart"]),
                                                   g[tag + "_tstartindex"], float(g["so_tend"]),
                                                   float(g["so_tstep"]), derivs)
            assert np.array_equal(x, g[tag + "_tresult"])
            assert col_err(y[g[tag + "_rows"]], g[tag + "_yrows"]) < 1e-9
