"""Host-side glue of the drop-in package (no GPU): argument checking, vectorised horizon
crossing, Bunch-Davies initial conditions, analysis functions, loud failure without CUDA."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_fo, model_kwargs, rel_err


class _FakeBg(object):
    """A finished background run, injected the way the reference's tests inject yresult
    (analysis/test/test_adiabatic.py:138-144)."""

    def __init__(self, t, y, nf):
        self.tresult, self.yresult = t, y
        self.phidots_ix = slice(1, 2 * nf, 2)

    def getepsilon(self):
        return 0.5 * np.sum(self.yresult[:, self.phidots_ix, 0] ** 2, axis=1)


def make_model(g):
    from pyflation_b200 import cosmomodels as c
    return getattr(c, g["cls"])(**model_kwargs(g))


@pytest.mark.parametrize("name", ["c1_msqphisq_k4", "c3_hybridquadratic_2p20", "ic_nophase",
                                  "ic_suppress", "ic_fixedainit", "hilltopaxion"])
def test_initial_conditions_match_reference(oracle, name):
    g = load_fo(name)
    m = make_model(g)
    nf = g["nfields"]
    bgt, bgy = oracle.run_background(g["potential"], g["params"], nf, g["bgystart"].copy())
    m.bgmodel = _FakeBg(bgt, bgy, nf)
    m.bgmodel.epsilon = m.bgmodel.getepsilon()
    m.fotend, m.fotendindex = oracle.find_inflation_end(bgt, m.bgmodel.epsilon)
    m.setfoics()
    assert np.array_equal(m.fotstartindex, g["fotstartindex"])
    assert rel_err(m.fotstart, g["fotstart"]) == 0.0
    assert rel_err(np.asarray(m.ainit, dtype=float).reshape(-1), g["ainit"].reshape(-1)) < 1e-15
    assert rel_err(np.asarray(m.etainit).reshape(-1), g["etainit"].reshape(-1)) < 1e-15
    assert rel_err(m.foystart, g["foystart"]) < 1e-14
    assert m.foystart.dtype == np.complex128 and m.foystart.shape == g["foystart"].shape


def test_vectorised_crossing_equals_reference_loop(oracle):
    """searchsorted on the running maximum == the reference's per-mode np.where scan
    (cosmomodels.py:1108), including a non-monotone a*H."""
    from pyflation_b200 import cosmomodels as c
    rng = np.random.RandomState(5)
    t = np.arange(600) * 0.01
    H = 1e-5 * (1 + 0.3 * np.sin(40 * t)) * np.exp(-0.02 * t)     # wiggly: a*H not monotone
    m = c.FOCanonicalTwoStage(k=np.array([1.0]), nfields=1)
    m.ainit = 3e-3
    aH = m.ainit * np.exp(t) * H
    assert np.any(np.diff(aH) < 0)
    k = np.sort(rng.uniform(50 * aH[3], 50 * aH.max() * 0.999, 257))
    m.k = k
    want_rows, want_t = oracle.crossing_indices(k, t, H, m.ainit, 50)
    got = m.findallkcrossings(t, H[:, None])
    assert np.array_equal(got[:, 0].astype(int), want_rows)
    assert np.array_equal(got[:, 1], want_t)
    row, efold = m.findkcrossing(k[17], t, H)
    assert row == want_rows[17] and efold == want_t[17]
    m.k = np.array([50 * aH.max() * 1.01])
    with pytest.raises(c.ModelError):
        m.findallkcrossings(t, H)


def test_constructor_validation_matches_reference():
    from pyflation_b200 import cosmomodels as c
    with pytest.raises(ValueError):
        c.CosmologicalModel(tstart=5.0, tend=5.0)
    with pytest.raises(ValueError):
        c.CosmologicalModel(tstart=6.0, tend=5.0)
    with pytest.raises(ValueError):
        c.CosmologicalModel(solver="euler")
    with pytest.raises(ValueError):
        c.CosmologicalModel(nfields=0)
    with pytest.raises(c.ModelError):
        c.CosmologicalModel(pot_params=[1, 2])
    with pytest.raises(AttributeError):
        c.CosmologicalModel(potential_func="nosuchpotential")
    m = c.FOCanonicalTwoStage(nfields=2, potential_func="hybridquadratic")
    assert m.ystart.shape == (2 * 4 + 4 + 1,) and m.cq == 50
    assert (m.H_ix, m.bg_ix, m.dps_ix, m.dpdots_ix) == (4, slice(0, 5), slice(5, None, 2), slice(6, None, 2))
    np.testing.assert_allclose(m.bgystart, [18 / np.sqrt(2), -0.1 / np.sqrt(2)] * 2 + [0.0])
    bg = c.CanonicalBackground(ystart=np.array([18.0, -0.1, 0.0]), potential_func="msqphisq")
    U = 0.5 * 6.3267e-6 ** 2 * 18.0 ** 2
    assert bg.ystart[2] == pytest.approx(np.sqrt(U / (3 - 0.5 * 0.01)), rel=1e-15)   # findH


def test_analysis_known_answers():
    """The reference's own P_R tests (analysis/test/test_adiabatic.py:126-260) against the
    drop-in analysis module, model built and injected exactly as those tests do."""
    from pyflation_b200 import cosmomodels as c
    from pyflation_b200.analysis import adiabatic, utilities
    g = np.load(os.path.join(GOLDEN, "kat_adiabatic.npz"))

    def getmodel(phidot, modes, k, lent, nfields, axis):
        m = c.FOCanonicalTwoStage(k=k, nfields=nfields)
        m.yresult = np.zeros((lent, 2 * nfields + 1 + 2 * nfields ** 2, len(k)), dtype=np.complex128)
        m.yresult[:, m.dps_ix] = utilities.flattenmodematrix(modes, nfields, ix1=axis, ix2=axis + 1)
        m.yresult[:, m.phidots_ix] = phidot
        m.tresult = np.arange(lent) * 0.01
        return m

    k1 = np.array([5.25e-60])
    m = getmodel(np.array([7, 9]).reshape((1, 2, 1)), np.array([[1, 3], [2, 5]]).reshape((1, 2, 2, 1)), k1, 1, 2, 1)
    np.testing.assert_almost_equal(adiabatic.Pr(m), [[4981 / 16900.0]])
    m = getmodel(np.array([1, 1]).reshape((1, 2, 1)), np.array([[1, 1j], [-1j, 3 - 1j]]).reshape((1, 2, 2, 1)), k1, 1, 2, 1)
    arr = adiabatic.Pr(m)
    np.testing.assert_almost_equal(arr, [[11 / 4.0]])
    assert not np.iscomplexobj(arr)
    m = getmodel(1.7, np.array([[[7]]]), k1, 1, 1, 0)
    np.testing.assert_almost_equal(adiabatic.Pr(m), [[49 / 1.7 ** 2]])
    np.testing.assert_almost_equal(adiabatic.scaled_Pr(m), [[49 / 1.7 ** 2 * k1[0] ** 3 / (2 * np.pi ** 2)]])
    # block case: shapes and values against the reference's functions
    modes, phidot = g["block_modes"], g["block_phidot"]
    np.testing.assert_allclose(adiabatic.Pr_spectrum(phidot, modes, 1), g["block_Pr"], rtol=1e-14)
    np.testing.assert_allclose(adiabatic.Pphi_matrix(modes, 1), g["block_Pphi"], rtol=1e-14)
    assert adiabatic.Pphi_matrix(modes, 1).shape == modes.shape
    m = getmodel(np.arange(24.0).reshape((4, 3, 2)), np.arange(72.0).reshape((4, 3, 3, 2)),
                 np.array([1e-60, 5.25e-60]), 4, 3, 1)
    assert adiabatic.Pr(m).shape == (4, 2)
    assert adiabatic.Pr(m, tix=-1, kix=1).shape == (1, 1)
    np.testing.assert_allclose(utilities.kscaling(g["kscaling_k"]), g["kscaling"], rtol=1e-15)
    with pytest.raises(ValueError):
        utilities.getmodematrix(np.zeros((2, 5, 3)), 2)


def test_pzeta_against_reference():
    """analysis.Pzeta / Pzeta_spectrum: the reference's hand-computed vectors
    (test_adiabatic.py:262-327) and its Pzeta of the last rows of every golden run."""
    from pyflation_b200 import cosmomodels as c
    from pyflation_b200.analysis import adiabatic
    from conftest import fo_fixture_names
    g = np.load(os.path.join(GOLDEN, "pzeta.npz"))
    arr = adiabatic.Pzeta_spectrum(3, 0.5, 2, np.array([[7]]), np.array([[5]]), 0)
    np.testing.assert_almost_equal(arr, 29.25 ** 2 / 9.0)
    two = adiabatic.Pzeta_spectrum(np.array([5.5, 2.3]).reshape((2, 1)), np.array([2, 5]).reshape((2, 1)),
                                   np.array([3]).reshape((1, 1)),
                                   np.array([[1 / 3.0, 0.1], [0.1, 0.5]]).reshape((2, 2, 1)),
                                   np.array([[0.1, 0.2], [0.2, 1 / 7.0]]).reshape((2, 2, 1)), 0)
    np.testing.assert_almost_equal(two, [135451.600445493 / 783 ** 2])
    np.testing.assert_allclose(two, g["kat_two"], rtol=1e-14)
    cx = adiabatic.Pzeta_spectrum(np.array([1, 2]).reshape((2, 1)), np.array([1, 1]).reshape((2, 1)),
                                  np.array([1]).reshape((1, 1)),
                                  np.array([[1, 1j], [-1j, 3 - 1j]]).reshape((2, 2, 1)),
                                  np.array([[1, -1j], [1j, 3 + 1j]]).reshape((2, 2, 1)), 0)
    np.testing.assert_almost_equal(cx, [38 / 36.0])
    assert not np.iscomplexobj(cx)
    for name in fo_fixture_names():
        f = load_fo(name)
        m = c.FOCanonicalTwoStage(k=f["k"], nfields=f["nfields"], potential_func=f["potential"],
                                  pot_params=dict(f["params"]))
        m.yresult = f["yrows"][-3:]
        m.tresult = f["tresult"][f["rows"][-3:]]
        assert rel_err(m.Pzeta, g[name + "_Pzeta"]) < 1e-12, name
        assert rel_err(m.scaled_Pzeta[-1:], g[name + "_scaled_Pzeta"]) < 1e-12, name


def test_helpers_default_grid():
    from pyflation_b200 import helpers
    kend = helpers.getkend(0.85e-60, 0.4e-60, 1025)
    assert kend == pytest.approx(8.2165e-58, rel=1e-12)
    k = helpers.seq(0.85e-60, kend, 0.4e-60)
    assert len(k) == 2053 and k[0] == 0.85e-60                 # BASELINE config 1 grid


def test_helpers_utilities_behave_like_the_reference(tmp_path):
    """The rest of pyflation/helpers.py (helpers.py:70-297), behaviour stated by its docstrings."""
    import logging
    from pyflation_b200 import helpers as h
    assert h.eto10(1e-05) == r"1\times 10^{-5}" and h.eto10(2.5e+56) == r"2.5\times 10^{+56}"
    assert h.klegend([1e-60]) == [r"$k=1\times 10^{-60}M_{\mathrm{PL}}$"]
    assert h.klegend([1e-60], mpc=True)[0].endswith(r" M\mathrm{pc}$")
    assert h.invmpc2mpl(2) == 2 * 2.625e-57 and h.mpl2invmpc() == 3.8095e+56
    assert [h.ispower2(n) for n in (1, 2, 3, 1024, 1025, 0)] == [0, 1, 0, 10, 0, 0]
    assert h.removedups([3, 1, 3, 2, 1]).tolist() == [3.0, 1.0, 2.0] and h.removedups([]).size == 0
    f, kw = h.getintfunc(np.linspace(0, 1, 9))
    assert f.__name__ == "romb" and abs(kw["dx"] - 0.125) < 1e-15
    f, kw = h.getintfunc(np.linspace(0, 1, 10))
    assert "simps" in f.__name__ and abs(f(kw["x"] ** 2, **kw) - 1 / 3.0) < 1e-3
    with pytest.raises(ValueError):
        h.getintfunc([])
    assert list(h.cartesian_product([[1, 2], [3]])) == [[1, 3], [2, 3]]
    assert h.cartesian(([1, 2, 3], [4, 5], [6, 7])).tolist()[:3] == [[1, 4, 6], [1, 4, 7], [1, 5, 6]]
    assert h.cartesian(([1, 2], [4, 5])).shape == (4, 2)
    assert h.find_nearest(np.array([1.0, 2.0, 4.0]), 2.9) == 2.0
    target = tmp_path / "a" / "b" / "run.npz"
    h.ensurepath(str(target))
    assert os.path.isdir(str(tmp_path / "a" / "b"))
    log = h.startlogging(logging.getLogger("pf_test_helpers"), str(tmp_path / "run.log"), logging.INFO)
    log.info("hello")
    for handler in list(log.handlers):
        handler.flush(); handler.close(); log.removeHandler(handler)
    assert "hello" in open(str(tmp_path / "run.log")).read()


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from pyflation_b200 import cosmomodels as c, rk4, _lib
    m = c.FOCanonicalTwoStage(k=np.array([1e-60]), potential_func="msqphisq")
    with pytest.raises(_lib.LibraryError):
        m.run(saveresults=False)
    with pytest.raises(NotImplementedError):
        rk4.rkdriver_tsix(np.ones(2), 0.0, np.array([0]), 1.0, 0.01, lambda y, t: -y)
    with pytest.raises(_lib.LibraryError):
        c.CanonicalFirstOrder(k=np.array([1.0]), ystart=np.ones(5)).derivs(np.ones((5, 1)), 0.0)
    with pytest.raises(rk4.SimRunError):
        rk4.rkdriver_tsix(np.ones((3, 1)), 0.0, np.array([0]), 1.0, None,
                          c.CanonicalBackground(ystart=np.array([18.0, -0.1, 1e-5])).derivs)
    with pytest.raises(rk4.SimRunError):
        rk4.rkdriver_tsix(np.ones((3, 1)), 0.0, np.array([500]), 1.0, 0.01,
                          c.CanonicalBackground(ystart=np.array([18.0, -0.1, 1e-5])).derivs)


def test_bench_reference_arm_on_cpu():
    """bench.py's CPU arm (the oracle port, one process per shard) on a tiny sample: the
    reference leg of the benchmark must run without a GPU."""
    import bench
    arm = bench.CpuArm(nk_worker=16, cores=2)
    try:
        wall, work = arm.step()
    finally:
        arm.close()
    assert wall > 0 and work > 2 * 16 * 5000       # ~6000 active steps per mode
    assert "2 processes x 16 consecutive modes" in arm.sample
    cfg = bench.workload_config(1)
    assert cfg["nk_per_gpu"] == 2 ** 16 and cfg["potential"] == "lambdaphi4"


def test_crossing_rows_property(oracle):
    """Property test (hypothesis): for arbitrary positive, not necessarily monotone a*H and any k
    inside its range, the searchsorted-on-running-maximum crossing equals the reference's scan."""
    from hypothesis import given, settings, strategies as st
    from pyflation_b200 import cosmomodels as c

    @settings(max_examples=60, deadline=None)
    @given(st.integers(5, 200), st.integers(1, 40), st.integers(0, 2 ** 31 - 1))
    def check(T, nk, seed):
        rng = np.random.RandomState(seed)
        t = np.arange(T) * 0.01
        H = np.exp(rng.uniform(-12, -9, T))                      # wild, non-monotone
        m = c.FOCanonicalTwoStage(k=np.array([1.0]), nfields=1)
        m.ainit = 10.0 ** rng.uniform(-70, -50)
        aH = 50 * m.ainit * np.exp(t) * H
        k = np.sort(rng.uniform(aH.min() * 0.5, aH.max() * 0.999, nk))
        m.k = k
        want_rows, want_t = oracle.crossing_indices(k, t, H, m.ainit, 50)
        got = m.findallkcrossings(t, H)
        assert np.array_equal(got[:, 0].astype(int), want_rows)
        assert np.array_equal(got[:, 1], want_t)

    check()


def test_partition_property():
    """Property test: contiguous partitions cover the range, are monotone, and no slab exceeds
    the ideal share by more than one item's cost."""
    from hypothesis import given, settings, strategies as st
    from pyflation_b200 import parallel

    @settings(max_examples=100, deadline=None)
    @given(st.integers(1, 300), st.integers(1, 16), st.integers(0, 2 ** 31 - 1))
    def check(n, world, seed):
        cost = np.random.RandomState(seed).uniform(0.5, 2.0, n)
        b = parallel.partition_contiguous(cost, world)
        assert b[0] == 0 and b[-1] == n and len(b) == world + 1 and np.all(np.diff(b) >= 0)
        per = np.array([cost[b[r]:b[r + 1]].sum() for r in range(world)])
        assert per.sum() == pytest.approx(cost.sum())
        assert per.max() <= cost.sum() / world + 2 * cost.max() + 1e-9

    check()


@pytest.mark.parametrize("ext", [".npz", ".hf5"])
def test_save_and_wrapper_model_roundtrip(tmp_path, ext):
    """saveallresults -> make_wrapper_model (cosmomodels.py:236-402, 1552-1677) through the
    reference's HDF5 layout and through the .npz counterpart, with results injected the way the
    reference's tests do."""
    from pyflation_b200 import cosmomodels as c
    g = load_fo("c3_hybridquadratic_2p20")
    m = c.FOCanonicalTwoStage(k=g["k"], nfields=2, potential_func="hybridquadratic",
                              pot_params={"nfields": 2}, bgystart=g["bgystart"].copy(), cq=50)
    m.yresult, m.tresult = g["yrows"], g["tresult"][g["rows"]]
    m.foystart, m.fotstartindex, m.fotstart, m.ainit = g["foystart"], g["fotstartindex"], g["fotstart"], g["ainit"]
    path = m.saveallresults(filename=str(tmp_path / ("run" + ext)))
    w = c.make_wrapper_model(path)
    assert isinstance(w, c.FOCanonicalTwoStage) and w.nfields == 2 and w.potential_func == "hybridquadratic"
    assert np.array_equal(w.k, g["k"]) and np.array_equal(w.fotstartindex, g["fotstartindex"])
    assert np.array_equal(w.yresult.view(np.float64), m.yresult.view(np.float64), equal_nan=True)
    assert rel_err(w.Pr[-1:], g["Pr_last"]) < 1e-12
    with pytest.raises(IOError):
        c.make_wrapper_model(str(tmp_path / ("missing" + ext)))
    with pytest.raises(IOError):
        m.saveallresults(filename=str(tmp_path / "no_such_dir" / ("run" + ext)))
    with pytest.raises(NotImplementedError):
        m.saveallresults(filename=str(tmp_path / "run.txt"), filetype="txt")
    if ext == ".hf5":
        # an existing file is appended to, as in the reference (cosmomodels.py:247-249): two runs
        assert m.saveallresults(filename=path) == path
        w2 = c.make_wrapper_model(path)
        assert w2.yresult.shape[0] == 2 * m.yresult.shape[0] and len(w2.tresult) == 2 * len(m.tresult)
        # zlib-compressed files read back the same
        packed = m.saveallresults(filename=str(tmp_path / "packed.hf5"), hdf5complevel=2)
        assert os.path.getsize(packed) < os.path.getsize(path) // 2
        w3 = c.make_wrapper_model(packed)
        assert np.array_equal(w3.yresult.view(np.float64), m.yresult.view(np.float64), equal_nan=True)
        with pytest.raises(NotImplementedError):
            m.saveallresults(filename=str(tmp_path / "blosc.hf5"), hdf5complevel=2, hdf5complib="blosc")


def test_default_results_file_name_and_type(tmp_path, monkeypatch):
    """No file name: ``run<datetime>.hf5`` in the working directory (cosmomodels.py:239-242);
    a background-only model goes to the ``bgresults`` group with a float64 yresult (:263-267, 346)."""
    from pyflation_b200 import cosmomodels as c, hdf5lite
    monkeypatch.chdir(tmp_path)
    bg = c.CanonicalBackground(ystart=np.array([18.0, -0.1, 0.0]), tend=83.0, tstep_wanted=0.01,
                               potential_func="msqphisq", pot_params={"nfields": 1}, nfields=1)
    bg.tresult = np.arange(5) * 0.01
    bg.yresult = np.arange(15.0).reshape(5, 3, 1)
    bg.lastparams = bg.callingparams()
    path = bg.saveallresults()
    assert os.path.dirname(path) == str(tmp_path) and os.path.basename(path) == "run" + bg.lastparams["datetime"] + ".hf5"
    with hdf5lite.open_file(path) as rf:
        assert list(rf.root._v_children) == ["bgresults"]
        res = rf.root.bgresults
        assert res.yresult.dtype == np.float64 and np.array_equal(res.yresult.read(), bg.yresult)
        assert res.parameters[0]["classname"] == b"CanonicalBackground" and res.parameters[0]["nfields"] == 1
        assert [r["name"] for r in res.pot_params] == [b"nfields"]
    assert c.CosmologicalModel.saveallresults(bg, filename=str(tmp_path / "x.npz")).endswith("x.npz")


def test_committed_bench_record_has_contract_keys():
    """The bench line recorded for this round carries every key of the measurement contract."""
    import json
    from conftest import ROOT
    line = json.loads(open(os.path.join(ROOT, "profiles", "r1_bench_n1.json")).read())
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches",
                "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["dtype"] == "f64" and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
    assert abs(line["roofline"]["frac"] - line["roofline"]["achieved"] / line["roofline"]["peak"]) < 1e-12
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(line["cpu_baseline"])
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(line["clocks"])
    assert line["gpu_launches"] == line["steps"] and line["e2e"]["value"] < line["value"]


def test_pyflation_import_name_resolves_to_the_b200_package():
    """`import pyflation` (the reference's package name, pyflation/__init__.py) gives the module
    objects of pyflation_b200: code written against the reference runs without editing imports."""
    import pyflation
    import pyflation_b200
    from pyflation import cosmomodels, rk4, cmpotentials, helpers, configuration, analysis, run_config  # noqa: F401
    from pyflation.analysis import adiabatic, utilities  # noqa: F401
    from pyflation.analysis.adiabatic import Pr, scaled_Pr  # noqa: F401
    import pyflation.cosmomodels as cm
    assert cm is pyflation_b200.cosmomodels and rk4 is pyflation_b200.rk4
    assert pyflation.__version__ == "0.2.3"
    assert cm.FOCanonicalTwoStage.solverlist == ["rkdriver_tsix"]
    with pytest.raises(ImportError):
        import pyflation.sourceterm  # noqa: F401
