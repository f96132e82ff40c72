"""Pins the CPU oracle (oracle/pyflation_oracle.py) to the unmodified reference.

The golden files were produced by the reference itself (tests/golden/make_golden.py).  The
oracle restates the same NumPy arithmetic, so agreement is expected at rounding level; the
tolerances below are 1e-11 (elementwise) where the north-star budget is 1e-9.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, fo_fixture_names, load_fo, rel_err, mode_rel_err

TOL = 1e-11


def oracle_run(oracle, g, integrate=True):
    kw = {}
    if g["cls"] == "FONoPhase":
        kw["phase"] = False
    if g["cls"] == "FOSuppressOneField":
        kw["suppress_ix"] = g["extra"]["suppress_ix"]
    if g["cls"] == "FixedainitTwoStage":
        kw["fixed_ainit"] = 7.837219134630212e-65          # cosmomodels.py:2043
    return oracle.first_order_run(g["potential"], g["k"], g["nfields"], g["params"],
                                  g["bgystart"].copy(), cq=float(g["cq"]), tend=float(g["tend"]),
                                  h=float(g["h"]), integrate=integrate, **kw)


@pytest.mark.parametrize("name", fo_fixture_names())
def test_first_order_run_matches_reference(oracle, name):
    g = load_fo(name)
    o = oracle_run(oracle, g)
    nf = g["nfields"]
    # background run + initial-condition glue
    assert o["yresult"].shape == (g["T"], oracle.nvar_of(nf), len(g["k"]))
    assert o["fotendindex"] == int(g["fotendindex"])
    assert rel_err(o["fotend"], g["fotend"]) == 0.0
    assert np.array_equal(o["fotstartindex"], g["fotstartindex"])
    assert rel_err(o["fotstart"], g["fotstart"]) == 0.0
    assert rel_err(np.asarray(o["ainit"], dtype=float).reshape(-1), g["ainit"].reshape(-1)) < 1e-15
    assert rel_err(np.asarray(o["etainit"]).reshape(-1), g["etainit"].reshape(-1)) < 1e-15
    assert rel_err(o["bgy"][g["bgrows"]], g["bgyrows"]) < 1e-14
    assert rel_err(o["foystart"], g["foystart"]) < 1e-14
    # trajectory rows, NaN layout included
    assert rel_err(o["tresult"], g["tresult"]) == 0.0
    rows = g["rows"]
    nb = 2 * nf + 1
    assert rel_err(o["yresult"][rows][:, :nb], g["yrows"][:, :nb]) < TOL
    assert mode_rel_err(o["yresult"][rows], g["yrows"], nf) < TOL
    # P_R
    assert rel_err(oracle.pr_of_yresult(o["yresult"], nf, tix=-1), g["Pr_last"]) < TOL
    assert rel_err(oracle.pr_of_yresult(o["yresult"], nf)[rows], g["Pr_rows"]) < TOL
    assert rel_err(oracle.kscaling(g["k"]) * oracle.pr_of_yresult(o["yresult"], nf, tix=-1),
                   g["scaled_Pr_last"]) < TOL


def test_reference_constants(oracle):
    """In-source known answers of the reference: cosmomodels.py:2043 (ainit of the default
    msqphisq run) and solutions/fixtures.py:18 (etainit)."""
    g = load_fo("c1_msqphisq_k4")
    o = oracle_run(oracle, g, integrate=False)
    assert abs(float(o["ainit"][0]) / 7.837219134630212e-65 - 1) < 1e-12
    assert abs(float(o["etainit"][0]) / -2.7559960682873626e+68 - 1) < 1e-10


def test_potentials_match_reference(oracle):
    g = np.load(os.path.join(GOLDEN, "potentials.npz"))
    tags = sorted(set(key.rsplit("_", 1)[0] for key in g.files if key.endswith("_y")))
    assert len(tags) >= 18 + 10
    custom = {"msqphisq__custom": {"mass": 1e-5}, "lambdaphi4__custom": {"lambda": 2e-13},
              "step_potential__custom": {"mass": 6e-6, "c": 0.002, "d": 0.03, "phi_s": 14.8},
              "hybridquadratic__custom": {"m1": 2e-6, "m2": 8e-6},
              "nflation__custom": {"nfields": 2, "mass": 1e-6},
              "quartictwofield__default": {"m1": 5e-6, "m2": 5e-8, "l1": 5e-10, "l2": 5e-14}}
    seen = set()
    for tag in tags:
        name, variant = tag.split("__")
        seen.add(name)
        params = custom.get(tag)
        if name == "nflation" and variant.startswith("nf"):
            params = {"nfields": int(variant[2:])}
        for i, y in enumerate(g[tag + "_y"]):
            U, dU, d2U, d3U = oracle.POTENTIALS[name](y, params)
            assert rel_err(np.asarray(U), g[tag + "_U"][i]) < 1e-14, tag
            assert rel_err(np.asarray(dU), g[tag + "_dU"][i], 1e-200) < 1e-13, tag
            assert rel_err(np.asarray(d2U), g[tag + "_d2U"][i], 1e-200) < 1e-13, tag
            if d3U is not None:
                assert rel_err(np.asarray(d3U, dtype=complex).ravel(), g[tag + "_d3U"][i], 1e-200) < 1e-12, tag
    assert seen == set(oracle.POTENTIALS)


def test_adiabatic_known_answers(oracle):
    """The reference's hand-computed vectors (analysis/test/test_adiabatic.py:81-260)."""
    g = np.load(os.path.join(GOLDEN, "kat_adiabatic.npz"))
    # literal answers from the reference's tests
    np.testing.assert_allclose(g["two_real_Pr"], [[4981 / 16900.0]], rtol=1e-15)
    np.testing.assert_allclose(g["two_complex_Pr"], [[11 / 4.0]], rtol=1e-15)
    np.testing.assert_allclose(g["single_Pr"], [49 / 1.7 ** 2], rtol=1e-15)
    np.testing.assert_allclose(g["two_real_Pphi"].reshape(2, 2), [[10, 17], [17, 29]])
    np.testing.assert_allclose(g["two_complex_Pphi"].reshape(2, 2), [[2, -1 + 4j], [-1 - 4j, 11]])
    for tag in ("two_real", "two_complex", "single", "block"):
        axis = int(g[tag + "_axis"]) if tag != "block" else 1
        got = oracle.pr_spectrum(g[tag + "_phidot"], g[tag + "_modes"], axis)
        assert not np.iscomplexobj(got)
        np.testing.assert_allclose(got, g[tag + "_Pr"], rtol=1e-14)
        np.testing.assert_allclose(oracle.pphi_matrix(g[tag + "_modes"], axis), g[tag + "_Pphi"], rtol=1e-14)
    np.testing.assert_allclose(oracle.kscaling(g["kscaling_k"]), g["kscaling"], rtol=1e-15)


def test_rk4_driver_edge_cases(oracle):
    """rk4.py:63,74 error paths and the NaN / start-row layout on a toy right-hand side."""
    f = lambda y, t: -y
    with pytest.raises(oracle.SimRunError):
        oracle.rk4_drive(np.ones((2, 3)), 0.0, np.array([0, 1, 2]), 1.0, None, f)
    with pytest.raises(oracle.SimRunError):
        oracle.rk4_drive(np.ones((2, 3)), 0.0, np.array([0, 1, 200]), 1.0, 0.01, f)
    y0 = np.arange(1.0, 7.0).reshape(2, 3)
    x, y = oracle.rk4_drive(y0, 0.0, np.array([0, 3, 5]), 0.1, 0.01, f)
    assert y.shape == (11, 2, 3) and x.shape == (11,)
    assert np.isnan(y[:3, :, 1]).all() and np.isnan(y[:5, :, 2]).all()
    assert np.array_equal(y[3, :, 1], y0[:, 1]) and np.array_equal(y[5, :, 2], y0[:, 2])
    np.testing.assert_allclose(y[-1, :, 0], y0[:, 0] * np.exp(-0.1), rtol=1e-9)
    assert oracle.number_of_steps(0.0, 0.025, 0.01) == 3      # around(2.5) = 2 (half to even)
