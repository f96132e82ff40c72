"""Second-order stage: rkdriver_tsix driving CanonicalSecondOrder /
CanonicalRampedSecondOrder.derivs from a first-order result and a given source term.

Golden vectors ``tests/golden/so_*.npz`` are outputs of the UNMODIFIED reference's
``SOCanonicalThreeStage(...).run()`` (through ``oracle/ref_shim.py``, generator
``tests/golden/make_golden.py``) on a reference first-order run, with
``oracle.synthetic_source`` standing in for the source term (computing it is out of scope).

CPU tests pin the oracle's restatement against them; GPU tests compare the CUDA path
(pf_so_run / pf_so_derivs through the drop-in classes) with the goldens and the oracle.
Tolerance 1e-9 of the largest entry of the same variable in the same column (the solution
oscillates through zero); NaN prefixes and start rows exact.
"""
import ctypes
import os
import types

import numpy as np
import pytest

from conftest import GOLDEN

TOL = 1e-9
SO_FIXTURES = ("so_msqphisq", "so_lambdaphi4")


def load_so(name):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    g["potential"] = str(g["potential"])
    g["params"] = {str(a): (int(b) if str(a) == "nfields" else float(b))
                   for a, b in zip(g["param_keys"], g["param_vals"])}
    g["ramp"] = dict(zip([str(x) for x in g["rampargs_keys"]], g["rampargs_vals"]))
    return g


def fo_trajectory(g):
    """The (T, 5, nk) first-order array the second-order stage reads: background rows from the
    golden column 0 (every column of the reference carries the same numbers once it has started,
    NaN before), perturbation rows unused."""
    T, nk = len(g["fo_tresult"]), len(g["k"])
    y = np.full((T, 5, nk), np.nan, dtype=complex)
    for c in range(nk):
        s = int(g["fotstartindex"][c])
        y[s:, 0:3, c] = g["fo_bg0"][s:, :]
    return y


def col_err(got, want):
    assert got.shape == want.shape
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN layout differs"
    scale = np.nanmax(np.abs(want), axis=0, keepdims=True)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.abs(got - want) / scale
    return float(np.nanmax(np.where(scale > 0, r, 0.0)))


# ------------------------------------------------------------------ CPU: oracle vs reference
@pytest.mark.parametrize("name", SO_FIXTURES)
def test_oracle_second_order_matches_reference(oracle, name):
    g = load_so(name)
    fo_y = fo_trajectory(g)
    src = oracle.synthetic_source(g["fo_tresult"], g["k"])
    for tag, ramp in (("plain", None), ("ramped", g["ramp"])):
        t, y, ts = oracle.second_order_run(g["potential"], g["params"], g["k"], g["ainit"],
                                           g["fo_tresult"], fo_y, g["fotstart"], g["fotstartindex"],
                                           g["bgepsilon"], src, float(g["fo_tstep"]),
                                           float(g["fo_simtstart"]), ramp=ramp)
        assert np.array_equal(ts, g[tag + "_tstartindex"])
        assert np.array_equal(t, g[tag + "_tresult"])
        assert col_err(y[g[tag + "_rows"]], g[tag + "_yrows"]) < 1e-12
        f = oracle.make_so_rhs(g["potential"], g["params"], g["k"], g["ainit"], fo_y,
                               float(g["fo_simtstart"]), float(g["fo_tstep"]), g["bgepsilon"], src,
                               ramp=ramp, tstart=g["fotstart"], tstartindex=ts,
                               so_tstep=float(g["so_tstep"]))
        for tt, yy, dd in zip(g[tag + "_derivs_t"], g[tag + "_derivs_y"], g[tag + "_derivs_dydx"]):
            assert col_err(f(yy, tt), dd) < 1e-12
    f = oracle.make_so_rhs(g["potential"], g["params"], g["k"], g["ainit"], fo_y,
                           float(g["fo_simtstart"]), float(g["fo_tstep"]), g["bgepsilon"], None)
    assert col_err(f(g["homog_y"], float(g["homog_t"])), g["homog_dydx"]) < 1e-12


def test_ramp_changes_the_result():
    """The two classes must not be the same computation (the ramp is active for ~1 e-fold)."""
    g = load_so("so_msqphisq")
    assert col_err(g["ramped_yrows"][-4:], g["plain_yrows"][-4:]) > 1e-7


def injected_first_order(g):
    """A FOCanonicalTwoStage carrying the golden first-order data the second-order stage reads."""
    from pyflation_b200 import cosmomodels as c
    import pyflation_oracle as oracle
    m = c.FOCanonicalTwoStage(k=g["k"].copy(), potential_func=g["potential"],
                              pot_params=dict(g["params"]), nfields=1,
                              bgystart=g["bgystart"].copy(), cq=50, solver="rkdriver_tsix")
    m.tresult = g["fo_tresult"]
    m.yresult = fo_trajectory(g)
    m.fotstart, m.fotstartindex = g["fotstart"], g["fotstartindex"]
    m.ainit = g["ainit"]
    m.bgepsilon = g["bgepsilon"]
    m.simtstart = float(g["fo_simtstart"])
    m.source = oracle.synthetic_source(m.tresult, m.k)
    # SOHorizonStart looks the horizon crossings up in the background run
    m.bgmodel = types.SimpleNamespace(tresult=m.tresult, yresult=np.nan_to_num(g["fo_bg0"])[:, :, None])
    return m


def test_driver_setup_matches_reference_without_gpu():
    from pyflation_b200 import cosmomodels as c
    g = load_so("so_msqphisq")
    m = injected_first_order(g)
    so = c.SOCanonicalThreeStage(m, soclass=c.CanonicalRampedSecondOrder)
    assert np.array_equal(so.tstartindex, g["ramped_tstartindex"])
    assert so.tstep_wanted == float(g["so_tstep"]) and so.simtstart == float(g["so_simtstart"])
    assert so.tend == float(g["so_tend"]) and so.solver == "rkdriver_tsix"
    assert so.ystart.shape == (4, len(g["k"])) and not so.ystart.any()
    so.setup_soclass()
    assert isinstance(so.somodel, c.CanonicalRampedSecondOrder)
    assert so.somodel.rampargs == {"a": 15.0, "b": 0.3, "c": 1, "d": 2, "e": 1}
    assert so.somodel.second_stage is m and so.somodel.source is m.source
    assert np.array_equal(so.somodel.tstart, g["fotstart"])
    with pytest.raises(c.ModelError):
        c.SOCanonicalThreeStage(object())
    del m.source
    with pytest.raises(c.ModelError):
        c.SOCanonicalThreeStage(m)
    m.source = None
    with pytest.raises(c.ModelError):
        c.SOCanonicalThreeStage(m)
    # the reference's SOHorizonStart names a solver its rk4 module does not have (:1934)
    m.source = np.zeros((len(m.tresult), len(m.k)), dtype=complex)
    with pytest.raises(ValueError):
        c.SOHorizonStart(m)


@pytest.mark.parametrize("ext", [".npz", ".hf5"])
def test_results_file_carries_the_source_term(tmp_path, ext):
    """The reference's second-order script wraps a first-order results file that carries the
    source term (results.sourceterm -> model.source, bgepsilon recomputed; cosmomodels.py:1607-1661)
    and hands it to SOCanonicalThreeStage: the same round trip through the reference's HDF5 layout
    (hdf5lite) and through the .npz counterpart."""
    from pyflation_b200 import cosmomodels as c
    g = load_so("so_msqphisq")
    m = injected_first_order(g)
    bg = c.CanonicalBackground(ystart=g["bgystart"].copy(), tend=83.0, tstep_wanted=0.01,
                               potential_func=g["potential"], pot_params=dict(g["params"]), nfields=1)
    bg.tresult, bg.yresult = m.tresult, np.nan_to_num(g["fo_bg0"])[:, :, None]
    m.bgmodel = bg
    path = m.saveallresults(filename=str(tmp_path / ("fo" + ext)))
    if ext == ".hf5":
        from pyflation_b200 import hdf5lite
        with hdf5lite.open_file(path) as rf:               # the reference's layout (:279-349)
            assert sorted(rf.root._v_children) == ["bgresults", "results"]
            assert sorted(rf.root.results._v_children) == [
                "bgystart", "fotstart", "fotstartindex", "k", "parameters", "pot_params",
                "sourceterm", "tresult", "yresult", "ystart"]
            assert rf.root.results.yresult.kind == "EARRAY" and rf.root.results.yresult.dtype == np.complex128
            assert rf.root.results.parameters[0]["classname"] == b"FOCanonicalTwoStage"
            assert sorted(rf.root.bgresults._v_children) == ["tresult", "yresult"]
    w = c.make_wrapper_model(path)
    assert np.array_equal(w.source, m.source) and w.bgepsilon.shape == (len(m.tresult),)
    assert np.array_equal(w.yresult.view(np.float64), np.asarray(m.yresult).view(np.float64), equal_nan=True)
    assert w.potential_func == m.potential_func and w.pot_params == m.pot_params and w.cq == m.cq
    so = c.SOCanonicalThreeStage(w, soclass=c.CanonicalRampedSecondOrder)
    assert np.array_equal(so.tstartindex, g["ramped_tstartindex"]) and so.source is w.source
    del m.source
    w2 = c.make_wrapper_model(m.saveallresults(filename=str(tmp_path / ("fo_nosource" + ext))))
    assert w2.source is None
    with pytest.raises(c.ModelError):
        c.SOCanonicalThreeStage(w2)
    so.yresult, so.tresult = g["ramped_yrows"], g["ramped_tresult"][g["ramped_rows"]]
    sopath = so.saveallresults(filename=str(tmp_path / ("so" + ext)))
    with pytest.raises(c.ModelError):
        c.make_wrapper_model(sopath)


@pytest.mark.parametrize("ext,srcext", [(".npz", ".npz"), (".hf5", ".hf5"), (".hf5", ".npz")])
def test_combine_source_and_first_order_files(tmp_path, ext, srcext):
    """sohelpers.combine_source_and_fofile (sohelpers.py:20-86) on HDF5 / .npz files: the merged file is
    what make_wrapper_model + SOCanonicalThreeStage expect (first-order data cut to the source's k
    range, ``sourceterm`` added), with the reference's naming and checks."""
    import pyflation_oracle as oracle
    from pyflation import sohelpers, cosmomodels as c
    g = load_so("so_msqphisq")
    m = injected_first_order(g)
    del m.source
    bg = c.CanonicalBackground(ystart=g["bgystart"].copy(), tend=83.0, tstep_wanted=0.01,
                               potential_func=g["potential"], pot_params=dict(g["params"]), nfields=1)
    bg.tresult, bg.yresult = m.tresult, np.nan_to_num(g["fo_bg0"])[:, :, None]
    m.bgmodel, m.foystart = bg, np.zeros((5, len(m.k)), dtype=complex)
    fofile = m.saveallresults(filename=str(tmp_path / ("fo" + ext)))
    nks = 3                                              # the source covers the first 3 modes
    src = oracle.synthetic_source(m.tresult, m.k[:nks])

    def write_source(path, **data):
        if path.endswith(".npz"):
            np.savez(path, **data)
        else:                                            # the source-term files' layout: /results/...
            from pyflation_b200 import hdf5lite
            with hdf5lite.open_file(path, "w") as sf:
                grp = sf.create_group(sf.root, "results")
                for name in data:
                    sf.create_array(grp, name, data[name])
        return path

    srcfile = write_source(str(tmp_path / ("src-run1" + srcext)), sourceterm=src, k=m.k[:nks],
                           nix=np.arange(len(m.tresult)))
    merged = sohelpers.combine_source_and_fofile(srcfile, fofile)
    assert os.path.basename(merged) == "foandsrc-run1" + srcext and os.path.isfile(merged)
    again = sohelpers.combine_source_and_fofile(srcfile, fofile)
    assert os.path.basename(again) == "foandsrc_-run1" + srcext
    w = c.make_wrapper_model(merged)
    assert w.yresult.shape[2] == nks and len(w.k) == nks and w.foystart.shape == (5, nks)
    assert np.array_equal(w.source, src) and np.array_equal(w.fotstartindex, g["fotstartindex"][:nks])
    so = c.SOCanonicalThreeStage(w)
    assert np.array_equal(so.tstartindex, g["plain_tstartindex"][:nks])
    short = write_source(str(tmp_path / ("src-short" + srcext)), sourceterm=src[:10], k=m.k[:nks], nix=np.arange(10))
    with pytest.raises(ValueError):
        sohelpers.combine_source_and_fofile(short, fofile)
    bad = write_source(str(tmp_path / ("src-bad" + srcext)), k=m.k[:nks])
    with pytest.raises(KeyError):
        sohelpers.combine_source_and_fofile(bad, fofile)
    assert sorted(os.listdir(str(tmp_path))) == sorted(
        ["fo" + ext] + [n + srcext for n in ("src-run1", "foandsrc-run1", "foandsrc_-run1", "src-short", "src-bad")])
    with pytest.raises(IOError):
        sohelpers.combine_source_and_fofile(str(tmp_path / ("src-missing" + srcext)), fofile)


def test_so_entry_points_reject_bad_arguments():
    from pyflation_b200 import _lib
    L = _lib.lib()
    vp, nul = ctypes.c_void_p, ctypes.c_void_p(0)
    one = vp(8)                       # never dereferenced: every call fails its argument checks

    def run(nk=4, T=10, h=0.02, t0=0, t1=10, src=nul, rows=0, pitch=0, tstart=nul, ramp=None,
            state=nul, yout=one, every=1, opitch=0, k=one):
        r = None if ramp is None else (ctypes.c_double * 5)(*ramp)
        return L.pf_so_run(nk, k, one, nul, one, one, src, rows, pitch, tstart, nul,
                           None if r is None else ctypes.cast(r, ctypes.POINTER(ctypes.c_double)),
                           T, h, t0, t1, state, yout, every, opitch, nul)
    assert run(nk=-1) == -1
    assert run(t1=11) == -1
    assert run(h=float("nan")) == -4 and b"Need to specify h" in L.pf_last_error()
    assert run(k=nul) == -1
    assert run(src=one, rows=5, pitch=3) == -1          # pitch narrower than the batch
    assert run(t0=2) == -1                              # a later window needs the carried state
    assert run(yout=nul) == -1
    assert run(opitch=2) == -1
    assert run(ramp=(15.0, 0.3, 1, 2, 1)) == -1         # a ramp needs the start times
    assert run(nk=0) == 0 and run(t0=3, t1=3, state=one) == 0
    assert L.pf_so_derivs(-1, one, nul, nul, one, one, nul, nul, None, one, nul) == -1
    assert L.pf_so_derivs(4, nul, nul, nul, one, one, nul, nul, None, one, nul) == -1
    assert L.pf_so_derivs(0, nul, nul, nul, nul, nul, nul, nul, None, nul, nul) == 0


# ------------------------------------------------------------------ GPU: CUDA path vs reference
@pytest.mark.gpu
@pytest.mark.parametrize("name", SO_FIXTURES)
def test_second_order_stage_matches_reference_golden(name):
    from pyflation_b200 import cosmomodels as c
    g = load_so(name)
    m = injected_first_order(g)
    for tag, cls in (("plain", c.CanonicalSecondOrder), ("ramped", c.CanonicalRampedSecondOrder)):
        so = c.SOCanonicalThreeStage(m, soclass=cls)
        so.run(saveresults=False)
        assert so.yresult.dtype == np.float64
        assert so.yresult.shape == (len(g[tag + "_tresult"]), 4, len(g["k"]))
        assert np.array_equal(so.tresult, g[tag + "_tresult"])
        got, want = so.yresult[g[tag + "_rows"]], g[tag + "_yrows"]
        err = col_err(got, want)
        print("%s %s: max err %.3e of the column scale" % (name, tag, err))
        assert err < TOL
        for col, s in enumerate(g[tag + "_tstartindex"]):     # start rows are ystart itself
            assert np.array_equal(so.yresult[s, :, col], np.zeros(4))
        # single right-hand sides through the model's derivs (pf_so_derivs)
        for tt, yy, dd in zip(g[tag + "_derivs_t"], g[tag + "_derivs_y"], g[tag + "_derivs_dydx"]):
            assert col_err(so.somodel.derivs(yy, tt), dd) < 1e-12
        assert so.dp2modes.shape == (so.yresult.shape[0], 2, len(g["k"]))
        assert np.array_equal(so.a, g["ainit"] * np.exp(so.tresult))
    hm = c.CanonicalHomogeneousSecondOrder(
        k=m.k, ainit=m.ainit, ystart=np.zeros((4, len(m.k))), tstart=m.fotstart,
        tstartindex=so.tstartindex, simtstart=so.simtstart, tend=so.tend,
        tstep_wanted=so.tstep_wanted, potential_func=g["potential"], pot_params=dict(g["params"]))
    hm.second_stage = m
    d = hm.derivs(g["homog_y"], float(g["homog_t"]), tix=int(g["homog_tix"]))
    assert col_err(d, g["homog_dydx"]) < 1e-12
    with pytest.raises(c.ModelError):
        hm.derivs(g["homog_y"], float(g["homog_t"]))


@pytest.mark.gpu
def test_single_step_and_column_subsets(oracle):
    """rk4.rk4stepks on a second-order model (a two-row solve on the device) against the oracle's
    RK4 step, and derivs for a subset of the columns (``k`` / ``kix`` keywords)."""
    from pyflation_b200 import cosmomodels as c, rk4
    g = load_so("so_lambdaphi4")
    m = injected_first_order(g)
    so = c.SOCanonicalThreeStage(m, soclass=c.CanonicalRampedSecondOrder)
    so.setup_soclass()
    sm = so.somodel
    f = oracle.make_so_rhs(g["potential"], g["params"], g["k"], g["ainit"], m.yresult, float(g["fo_simtstart"]),
                           float(g["fo_tstep"]), g["bgepsilon"], m.source, ramp=g["ramp"], tstart=g["fotstart"],
                           tstartindex=so.tstartindex, so_tstep=so.tstep_wanted)
    rng = np.random.RandomState(3)
    # the middle row is a mode's own start row: the ramp is zero there for that mode (:935-937), which
    # the two-row solve must take from the MODEL's start rows, not from the solver's (all zero)
    for row in (int(so.tstartindex.max()) + 7, int(so.tstartindex[1]), int(so.tstartindex.max()) + 400):
        x = so.simtstart + row * so.tstep_wanted
        y = rng.normal(size=(4, len(g["k"]))) * 1e-3
        want = oracle.rk4_step(x, y, so.tstep_wanted, f(y, x), f)
        got = rk4.rk4stepks(x, y, so.tstep_wanted, None, {}, sm.derivs)
        assert col_err(got[None], want[None]) < 1e-12
        kix = np.array([0, 2])
        sub = sm.derivs(y[:, kix], x + 0.5 * so.tstep_wanted, k=g["k"][kix], kix=kix)
        assert col_err(sub[None], f(y, x + 0.5 * so.tstep_wanted)[:, kix][None]) < 1e-12
    with pytest.raises(c.ModelError):
        sm.derivs(y, x, k=g["k"], kix=None)


@pytest.mark.gpu
def test_second_order_launch_properties():
    """Size-independent properties of the kernel on a 1000-mode batch: windows == one launch and
    strided rows == rows of the full run (bit for bit), a source scaled by 2 doubles the result
    exactly, no source and zero start vector stay zero, a mode whose first-order result starts
    after its second-order start is NaN as in the reference (oracle)."""
    import torch
    import pyflation_oracle as oracle
    from pyflation_b200 import engine
    g = load_so("so_msqphisq")
    nk = 1000
    rng = np.random.RandomState(7)
    k = np.sort(rng.uniform(g["k"][0], g["k"][-1], nk))
    k[0] = g["k"][0]
    T_fo, fo_tstep = len(g["fo_tresult"]), float(g["fo_tstep"])
    aH = float(g["ainit"][0]) * np.exp(g["fo_tresult"]) * np.nan_to_num(g["fo_bg0"][:, 2])
    fo_tsix = np.searchsorted(np.maximum.accumulate(50 * aH), k, side="right").astype(np.int64)
    fo_tsix = np.maximum(fo_tsix, int(g["fotstartindex"][0]))
    tsix, h = oracle.so_start_indices(fo_tsix, fo_tstep)
    src = oracle.synthetic_source(g["fo_tresult"], k)
    pm = engine.make_model(g["potential"], g["params"], 1, float(g["ainit"][0]))
    T = int(np.around((g["fo_tresult"][-1] - g["fo_tresult"][0]) / h) + 1)
    ystart = rng.normal(size=(4, nk)) * 1e-3
    common = dict(fo_bg0=g["fo_bg0"], bgepsilon=g["bgepsilon"], fo_simtstart=0.0, fo_tstep=fo_tstep,
                  fo_tsix=fo_tsix, ramp=g["ramp"], tstart=g["fo_tresult"][fo_tsix])

    def solve(source, y0=ystart, every=1, windows=None, **over):
        kw = dict(common, **over)
        prob = engine.SecondOrderProblem(pm, k, tsix, y0, float(g["fo_tresult"][0]), T, h, source=source, **kw)
        if windows is None:
            return engine.second_order_device(prob, out_every=every).cpu().numpy()
        out = torch.empty((T, 4, nk), dtype=torch.float64, device=prob.dev)
        edges = [0] + list(windows) + [T]
        for a, b in zip(edges[:-1], edges[1:]):
            prob.launch(a, b, out[a:b], 1)
        return out.cpu().numpy()

    full = solve(src)
    assert np.isfinite(full[-1]).all()
    for col in (0, 17, nk - 1):
        assert np.isnan(full[: tsix[col], :, col]).all()
        assert np.array_equal(full[tsix[col], :, col], ystart[:, col])
    win = solve(src, windows=(int(tsix.min()) - 5, int(tsix.min()) + 1, int(tsix[nk // 2]), T - 1))
    assert np.array_equal(win, full, equal_nan=True)
    assert np.array_equal(engine.second_order_host(
        engine.SecondOrderProblem(pm, k, tsix, ystart, float(g["fo_tresult"][0]), T, h, source=src, **common),
        slab_bytes=32 * nk * 999), full, equal_nan=True)
    # the same with a page-locked source tensor (streamed without staging) and short windows
    pinned = torch.from_numpy(src).pin_memory()
    prob = engine.SecondOrderProblem(pm, k, tsix, ystart, float(g["fo_tresult"][0]), T, h, source=pinned, **common)
    assert prob.source is None and prob.source_host is not None          # 130 MB: streamed, not resident
    assert np.array_equal(engine.second_order_host(prob, slab_bytes=32 * nk * 57), full, equal_nan=True)
    assert prob.h2d_bytes >= src.nbytes * 0.7 and prob.d2h_bytes == full.nbytes
    strided = solve(src, every=37)
    assert np.array_equal(strided, full[::37], equal_nan=True)
    zero = np.zeros((4, nk))
    base = solve(src, y0=zero)
    assert np.array_equal(solve(2.0 * src, y0=zero), 2.0 * base, equal_nan=True)
    still = solve(None, y0=zero)
    assert np.array_equal(np.nan_to_num(still), np.zeros_like(still))
    # oracle on a few columns (column 0 included: the potential is evaluated there)
    cols = np.array([0, 3, nk // 2, nk - 1])
    fo_y = np.full((T_fo, 5, len(cols)), np.nan, dtype=complex)
    late = fo_tsix[cols].copy()
    late[2] += 9                                  # this column's first-order run "starts" too late
    for i, s in enumerate(late):
        fo_y[s:, 0:3, i] = g["fo_bg0"][s:, :]
    _, want, _ = oracle.second_order_run(g["potential"], g["params"], k[cols], g["ainit"], g["fo_tresult"],
                                         fo_y, g["fo_tresult"][fo_tsix[cols]], fo_tsix[cols], g["bgepsilon"],
                                         src[:, cols], fo_tstep, 0.0, ramp=g["ramp"], ystart=ystart[:, cols],
                                         tstartindex=tsix[cols])
    poisoned = fo_tsix.copy()
    poisoned[cols[2]] = late[2]
    got = solve(src, fo_tsix=poisoned)[:, :, cols]
    assert np.isnan(want[-1, :, 2]).all() and np.isfinite(want[-1, :, [0, 1, 3]]).all()
    err = col_err(got, want)
    print("1000-mode batch vs oracle on 4 columns: %.3e" % err)
    assert err < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("ext", [".npz", ".hf5"])
def test_second_order_script_runs_from_a_results_file(tmp_path, oracle, ext):
    """bin/pyflation_secondorder.py: wrap a first-order file that carries the source term, run
    the three-stage driver with run_config.soargs (ramped class), save; compare with the golden."""
    import importlib.util
    from conftest import ROOT
    from pyflation_b200 import cosmomodels as c
    g = load_so("so_msqphisq")
    m = c.FOCanonicalTwoStage(k=g["k"].copy(), potential_func=g["potential"], pot_params=dict(g["params"]),
                              nfields=1, bgystart=g["bgystart"].copy(), cq=50, solver="rkdriver_tsix")
    m.run(saveresults=False)
    m.source = oracle.synthetic_source(m.tresult, m.k)
    mrg = m.saveallresults(filename=str(tmp_path / ("mrg" + ext)))
    spec = importlib.util.spec_from_file_location("pf_so_script", os.path.join(ROOT, "bin", "pyflation_secondorder.py"))
    script = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(script)
    out = script.runsomodel(mrg, filename=str(tmp_path / "results" / ("so" + ext)))
    if ext == ".npz":
        z = np.load(out)
    else:                                                # the reference's layout: complex yresult (:330)
        from pyflation_b200 import hdf5lite
        with hdf5lite.open_file(out) as rf:
            res = rf.root.results
            z = {"classname": res.parameters[0]["classname"].decode(), "tresult": res.tresult.read()}
            yres = res.yresult.read()
        assert yres.dtype == np.complex128 and not yres.imag[np.isfinite(yres.imag)].any()
        z["yresult"] = np.ascontiguousarray(yres.real)
    assert str(z["classname"]) == "SOCanonicalThreeStage"
    assert col_err(z["yresult"][g["ramped_rows"]], g["ramped_yrows"]) < TOL
    assert np.array_equal(z["tresult"], g["ramped_tresult"])
    with pytest.raises(c.ModelError):
        script.runsomodel(mrg, soargs={"nfields": 2})


@pytest.mark.gpu
def test_full_chain_first_then_second_order(oracle):
    """FOCanonicalTwoStage.run() -> source attached -> SOCanonicalThreeStage.run(), everything on
    the CUDA path, against the reference's golden second-order rows of the same modes."""
    from pyflation_b200 import cosmomodels as c
    g = load_so("so_lambdaphi4")
    m = c.FOCanonicalTwoStage(k=g["k"].copy(), potential_func=g["potential"], pot_params=dict(g["params"]),
                              nfields=1, bgystart=g["bgystart"].copy(), cq=50, solver="rkdriver_tsix")
    m.run(saveresults=False)
    assert np.array_equal(m.fotstartindex, g["fotstartindex"])
    m.source = oracle.synthetic_source(m.tresult, m.k)
    so = c.SOCanonicalThreeStage(m, soclass=c.CanonicalRampedSecondOrder)
    so.run(saveresults=False)
    err = col_err(so.yresult[g["ramped_rows"]], g["ramped_yrows"])
    print("full chain: %.3e" % err)
    assert err < TOL
    assert so.deltaphi.shape == (so.yresult.shape[0], len(g["k"]))
    hs = c.SOHorizonStart(m, solver="rkdriver_tsix")
    assert np.all(hs.tstartindex > so.tstartindex)       # k = aH comes after k = 50 aH
    hs.run(saveresults=False)
    assert np.isfinite(hs.yresult[-1]).all()
