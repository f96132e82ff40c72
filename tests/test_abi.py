"""The C-ABI library builds, loads, and exports exactly what include/pyflation_b200.h declares
(no compute calls here: this file runs without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "pyflation_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from pyflation_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = header_functions()
    assert len(names) >= 17
    for name in names:
        assert hasattr(L, name), "missing export %s" % name
    assert sorted(_lib.EXPORTS) == names
    assert _lib.lib().pf_version() == 2


def test_potential_table_matches_reference_defaults(oracle):
    """Names, field counts and parameter defaults of the device potentials agree with the
    oracle's restatement of cmpotentials.py (itself pinned to the reference)."""
    from pyflation_b200 import _lib, cmpotentials
    table = _lib.potential_table()
    assert set(table) == set(oracle.POTENTIALS)
    assert [n for n, v in sorted(table.items(), key=lambda kv: kv[1][0])][:3] == \
        ["msqphisq", "lambdaphi4", "linde"]
    y1 = np.array([17.3, -0.1, 1e-5])
    y2 = np.array([11.0, -0.1, 0.7, 0.02, 1e-5])
    for name, (pid, nf, plist) in table.items():
        assert nf in (0, 1, 2)
        # defaults: perturbing each library default must be visible through the oracle
        y = y1 if nf == 1 else y2
        base = {"nfields": 2} if name == "nflation" else None
        U0 = oracle.POTENTIALS[name](y, base)[0]
        explicit = dict(base or {})
        explicit.update({p: d for p, d in plist})
        assert oracle.POTENTIALS[name](y, explicit)[0] == U0, name
        # and the host-side drop-in module agrees with the oracle
        got = getattr(cmpotentials, name)(y, base)
        want = oracle.POTENTIALS[name](y, base)
        assert got[0] == pytest.approx(want[0], rel=1e-15)
        np.testing.assert_allclose(got[1], want[1], rtol=1e-15)
        np.testing.assert_allclose(got[2], want[2], rtol=1e-15)
        if want[3] is None:
            assert got[3] is None
        else:
            np.testing.assert_allclose(got[3], want[3], rtol=1e-15, atol=0)
            assert np.shape(got[3]) == np.shape(want[3])


def test_sizes_and_step_count():
    from pyflation_b200 import _lib
    L = _lib.lib()
    assert [L.pf_nvar(n) for n in (1, 2, 8)] == [5, 13, 145]
    assert [L.pf_rec_doubles(n) for n in (1, 2, 8)] == [16, 30, 282]
    # rk4.py:73: int(around((tend - simtstart)/h) + 1), halves to even
    for tend, h in ((83.0, 0.01), (81.63292929292929, 0.01), (0.025, 0.01), (0.035, 0.01), (1.0, 0.3)):
        assert L.pf_number_of_steps(0.0, tend, h) == int(np.around(tend / h) + 1)


def test_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call, so it is testable on a CPU box."""
    from pyflation_b200 import _lib, engine
    L = _lib.lib()
    m = engine.make_model("msqphisq", {}, 1)
    y0 = (ctypes.c_double * 3)(18.0, -0.1, 1e-5)
    assert L.pf_bg_run(ctypes.byref(m), 10, float("nan"), 0, y0, 1, 0, None) == -4   # PF_E_STEP
    assert b"Need to specify h" in L.pf_last_error()
    assert L.pf_bg_run(ctypes.byref(m), 10, 0.01, 11, y0, 1, 0, None) == -5          # PF_E_START
    m.nfields = 2
    assert L.pf_bg_run(ctypes.byref(m), 10, 0.01, 0, y0, 1, 0, None) == -2           # PF_E_POTENTIAL
    m.nfields = 9
    assert L.pf_bg_run(ctypes.byref(m), 10, 0.01, 0, y0, 1, 0, None) == -3           # PF_E_NFIELDS
    with pytest.raises(AttributeError):
        engine.make_model("nosuchpotential", {}, 1)
    with pytest.raises(KeyError):
        engine.make_model("nflation", {}, 2)                                          # cmpotentials.py:828
    with pytest.raises(ValueError):
        engine.make_model("hybridquadratic", {}, 1)


def test_host_fill_narrower_store_widths():
    """The store width of the run-based fill is chosen once per process (AVX-512 / AVX2 / SSE2):
    rerun the sorted cases in child processes that force the narrower ones."""
    import subprocess, sys
    for isa in ("sse2", "avx2"):
        res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", __file__, "-k",
                              "test_host_fill_background_on_cpu and True"],
                             env=dict(os.environ, PF_HOST_FILL_ISA=isa), capture_output=True, text=True,
                             timeout=600)
        assert res.returncode == 0, isa + res.stdout[-2000:] + res.stderr[-2000:]
        assert "5 passed" in res.stdout


@pytest.mark.parametrize("sorted_starts", [False, True])
@pytest.mark.parametrize("nf,pert_nan", [(1, 0), (1, 1), (2, 1), (3, 0), (10, 1)])
def test_host_fill_background_on_cpu(nf, pert_nan, sorted_starts):
    """pf_host_fill_background is host code: checked here without a GPU against a NumPy
    construction (unsorted start rows -> per-element path; sorted -> run-based path with the
    widest non-temporal stores of this CPU; odd sizes, several thread counts, nf = 10 -> the
    any-nfields fallback)."""
    from pyflation_b200 import _lib
    L = _lib.lib()
    rng = np.random.RandomState(3 + nf)
    T, nk = 37, 53
    nb, nvar = 2 * nf + 1, 2 * nf * nf + 2 * nf + 1
    rec_len, bg_off = int(L.pf_rec_doubles(nf)), int(L.pf_rec_bg_offset(nf))
    rec = rng.standard_normal((T, rec_len))
    tsix = rng.randint(0, T + 5, nk).astype(np.int64)          # some modes never start
    if sorted_starts:
        tsix = np.sort(tsix)
    ystart = rng.standard_normal((nvar, nk)) + 1j * rng.standard_normal((nvar, nk))
    want = np.full((T, nvar, nk), 7.0 + 7.0j)
    for t in range(T):
        for k in range(nk):
            if t < tsix[k]:
                want[t, :nb if not pert_nan else nvar, k] = np.nan + 1j * np.nan
            elif t == tsix[k]:
                want[t, :nb, k] = ystart[:nb, k]
            else:
                want[t, :nb, k] = rec[t, bg_off:bg_off + nb]
    for nthreads in (1, 3, 64):
        out = np.full((T, nvar, nk), 7.0 + 7.0j)
        rc = L.pf_host_fill_background(ctypes.c_void_p(out.ctypes.data), nf, nk, 0, T,
                                       ctypes.c_void_p(rec.ctypes.data), rec_len, bg_off,
                                       ctypes.c_void_p(tsix.ctypes.data), ctypes.c_void_p(ystart.ctypes.data),
                                       pert_nan, nthreads)
        assert rc == 0
        assert np.array_equal(out.view(np.float64), want.view(np.float64), equal_nan=True)
    # a sub-range of rows leaves the others untouched
    out = np.full((T, nvar, nk), 7.0 + 7.0j)
    assert L.pf_host_fill_background(ctypes.c_void_p(out.ctypes.data), nf, nk, 10, 20,
                                     ctypes.c_void_p(rec.ctypes.data), rec_len, bg_off,
                                     ctypes.c_void_p(tsix.ctypes.data), ctypes.c_void_p(ystart.ctypes.data),
                                     pert_nan, 2) == 0
    assert np.all(out[:10] == 7.0 + 7.0j) and np.all(out[20:] == 7.0 + 7.0j)
    assert np.array_equal(out[10:20].view(np.float64), want[10:20].view(np.float64), equal_nan=True)
    assert L.pf_host_fill_background(None, nf, nk, 0, T, None, rec_len, bg_off, None, None, 0, 1) == -1


def test_peer_and_probe_entries_reject_bad_arguments():
    """Argument checks of the round-2 entry points that need no GPU to fail."""
    from pyflation_b200 import _lib
    L = _lib.lib()
    buf = ctypes.create_string_buffer(_lib.PF_PEER_HANDLE_BYTES)
    p = ctypes.c_void_p()
    assert L.pf_peer_alloc(0, ctypes.byref(p)) == -1
    assert L.pf_peer_alloc(16, None) == -1
    assert L.pf_peer_export(None, buf) == -1
    assert L.pf_peer_open(None, ctypes.byref(p)) == -1
    assert L.pf_peer_open(buf.raw, None) == -1
    assert L.pf_peer_free(None) == 0 and L.pf_peer_close(None) == 0
    out = ctypes.c_double(0.0)
    assert L.pf_fp64_peak_probe(0, None, 0, ctypes.byref(out)) == -1
    assert b"pf_fp64_peak_probe" in L.pf_last_error()
    # record layout helpers: compiled instantiations hold the dense matrix, runtime-nf the factors
    assert int(L.pf_rec_doubles(1)) == 16 and int(L.pf_rec_bg_offset(1)) == 12
    assert int(L.pf_rec_doubles(8)) == 282 and int(L.pf_rec_bg_offset(8)) == 264
    assert int(L.pf_rec_bg_offset(10)) == 4 * (6 + 20) and int(L.pf_rec_doubles(10)) == 126
    m = _lib.PfModel()
    m.potential_id, m.nfields = 12, 65                       # nflation takes 1..64 fields
    assert L.pf_fo_run(ctypes.byref(m), 1, None, None, None, None, 10, 0.01, 0, 10, None, None, 0, 0, 0,
                       None, None) == -3
    m.potential_id, m.nfields = 10, 9                        # any other potential: 1..8 / its own nf
    assert L.pf_fo_run(ctypes.byref(m), 1, None, None, None, None, 10, 0.01, 0, 10, None, None, 0, 0, 0,
                       None, None) in (-2, -3)
    # out_pitch smaller than the batch is refused
    m.potential_id, m.nfields = 0, 1
    assert L.pf_fo_run(ctypes.byref(m), 8, None, None, None, None, 10, 0.01, 0, 10, None, None, 0, 0, 4,
                       None, None) == -1


def test_field_sums_round_like_numpy():
    """The kernels add the terms of a sum over fields in NumPy's order (np_sum in pf_common.cuh =
    DOUBLE_pairwise_sum: straight for fewer than 8 terms, else eight running sums, a tree, the
    remainder).  pf_np_sum runs the same template on the host: bit-equal to np.sum for every
    length a field count can take, also for the axis-0 reduction of an (n, 1) array
    (cosmomodels.py:569), and different from plain left-to-right addition for n >= 8."""
    import ctypes
    from pyflation_b200 import _lib
    L = _lib.lib()
    rng = np.random.RandomState(12)
    differs = 0
    for n in range(0, 65):
        for _ in range(40):
            a = np.ascontiguousarray(rng.uniform(0.1, 5.0, n) ** 2 * 1e-11)
            got = L.pf_np_sum(a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), n)
            assert got == np.sum(a), (n, got, np.sum(a))
            if n:
                assert got == np.sum(a[:, None], axis=0)[0]
            seq = 0.0
            for x in a:
                seq += x
            differs += int(seq != got)
    assert differs > 100                                # the order matters in practice
    assert np.isnan(L.pf_np_sum(None, 3)) and np.isnan(L.pf_np_sum(a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 129))
