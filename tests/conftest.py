import glob
import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

warnings.filterwarnings("ignore", category=RuntimeWarning)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """The C-ABI library is built in-tree; build it when a fresh checkout runs the tests."""
    lib = os.path.join(ROOT, "pyflation_b200", "lib", "libpyflation_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()


def fo_fixture_names():
    return sorted(os.path.basename(p)[3:-4] for p in glob.glob(os.path.join(GOLDEN, "fo_*.npz")))


def load_fo(name):
    g = np.load(os.path.join(GOLDEN, "fo_%s.npz" % name), allow_pickle=False)
    d = {key: g[key] for key in g.files}
    d["potential"] = str(d["potential"])
    d["cls"] = str(d["cls"])
    d["nfields"] = int(d["nfields"])
    d["T"] = int(d["T"])
    params = {}
    for key, val in zip(d["param_keys"], d["param_vals"]):
        params[str(key)] = int(val) if str(key) == "nfields" else float(val)
    d["params"] = params
    d["extra"] = {str(key): (int(val) if str(key) == "suppress_ix" else float(val))
                  for key, val in zip(d["extra_keys"], d["extra_vals"])}
    return d


def model_kwargs(g):
    kw = dict(k=g["k"].copy(), potential_func=g["potential"], pot_params=dict(g["params"]),
              nfields=g["nfields"], bgystart=g["bgystart"].copy(), cq=50, solver="rkdriver_tsix")
    kw.update(g["extra"])
    return kw


def rel_err(a, b, floor=1e-300):
    """max |a-b|/|b| over entries that are finite in b; NaN patterns must agree."""
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN layout differs"
    ok = ~np.isnan(b)
    if not ok.any():
        return 0.0
    return float((np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), floor)).max())


ENTRY_FLOOR = 1e-6        # entries above this fraction of the largest one are checked one by one


def mode_rel_err(a, b, nf, entry_floor=ENTRY_FLOOR):
    """Perturbation rows.  Returns the larger of
      * the error relative to the largest mode-matrix entry of the same kind (dphi or dphi')
        at that (t, k) -- off-diagonal entries pass through zero, so tiny ones can only be
        held to an absolute bound -- and
      * the per-entry relative error |a-b|/|b| of every entry with |b| > entry_floor * that
        largest entry (off-diagonal entries of multi-field runs are typically 1e-3..1e-1 of the
        diagonal and are held to the full relative tolerance this way)."""
    nb = 2 * nf + 1
    worst = 0.0
    for off in (0, 1):
        x = a[:, nb + off::2, :]
        y = b[:, nb + off::2, :]
        assert np.array_equal(np.isnan(x), np.isnan(y)), "NaN layout differs"
        scale = np.nanmax(np.abs(y), axis=1, keepdims=True)
        with np.errstate(invalid="ignore", divide="ignore"):
            r = np.abs(x - y) / scale
            big = np.abs(y) > entry_floor * scale
            per_entry = np.where(big, np.abs(x - y) / np.abs(y), 0.0)
        if np.isfinite(r).any():
            worst = max(worst, float(np.nanmax(r)), float(np.nanmax(per_entry)))
    return worst


@pytest.fixture(scope="session")
def oracle():
    import pyflation_oracle
    return pyflation_oracle
