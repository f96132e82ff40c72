"""Generate the committed golden vectors from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # all fixtures
    python tests/golden/make_golden.py msqphisq   # a subset, by name

Each ``fo_<name>.npz`` holds a sub-sampled first-order run of the reference
(``FOCanonicalTwoStage(...).run(saveresults=False)`` through ``oracle/ref_shim.py``):
the k sub-grid (always containing column 0 of the full grid), the start indices, the
initial conditions, a set of trajectory rows (every ``row_stride``-th row, every row
around each column's start index, and the last rows) and the final ``P_R`` spectra.
``so_<name>.npz`` holds a second-order run on top of one of them: ``SOCanonicalThreeStage`` with
``CanonicalSecondOrder`` and ``CanonicalRampedSecondOrder``, ``oracle.synthetic_source`` as the source
term, plus single ``derivs`` calls of the three second-order classes.
``potentials.npz`` holds every ``cmpotentials`` function evaluated at fixed points.
``kat_adiabatic.npz`` restates the reference's own known-answer vectors for P_R
(``analysis/test/test_adiabatic.py:81-260``) with the answers recomputed by the reference.
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
warnings.filterwarnings("ignore")

import ref_shim  # noqa: E402

KMIN, KMAX = 0.85e-60, 8.2165e-58      # run_config.py:145-158 (K4 range, getkend)


def k4_grid(ref):
    return ref.helpers.seq(KMIN, KMAX, 0.4e-60)          # 2053 modes (BASELINE config 1)


def subsample(nk_full, n):
    return np.unique(np.linspace(0, nk_full - 1, n).round().astype(np.int64))


def _unequal(nf):
    """nflation start vector with a different value in every field (total amplitude ~ 17)."""
    y = []
    for i in range(nf):
        y += [18.0 / np.sqrt(nf) * (0.8 + 0.2 * i / (nf - 1) + 0.013 * ((i * 5) % 7)),
              -0.1 / np.sqrt(nf) * (1.0 + 0.11 * ((i * 3) % 5))]
    return y + [0.0]


# name -> (potential, nfields, pot_params, bgystart, full-grid size or "K4", #modes kept,
#          row stride, driver class name, extra kwargs)
FIXTURES = {
    # BASELINE configs 1-5
    "c1_msqphisq_k4": ("msqphisq", 1, {"nfields": 1}, [18.0, -0.1, 0], "K4", 33, 40,
                       "FOCanonicalTwoStage", {}),
    "c2_lambdaphi4_2p16": ("lambdaphi4", 1, {"nfields": 1}, [25.0, 0, 0], 2 ** 16, 17, 40,
                           "FOCanonicalTwoStage", {}),
    "c3_hybridquadratic_2p20": ("hybridquadratic", 2, {"nfields": 2},
                                [12.0, 1 / 300.0, 12.0, 49 / 300.0, 0], 2 ** 20, 9, 40,
                                "FOCanonicalTwoStage", {}),
    "c4_nflation8_2p18": ("nflation", 8, {"nfields": 8}, None, 2 ** 18, 3, 100,
                          "FOCanonicalTwoStage", {}),
    "c5_msqphisq_2p20": ("msqphisq", 1, {"nfields": 1}, [18.0, -0.1, 0], 2 ** 20, 9, 40,
                         "FOCanonicalTwoStage", {}),
    # the remaining run_config.py:51-123 fixtures
    "hybrid2and4": ("hybrid2and4", 1, {"nfields": 1}, [25.0, 0, 0], 2 ** 10, 5, 50,
                    "FOCanonicalTwoStage", {}),
    "linde": ("linde", 1, {"nfields": 1}, [25.0, 0, 0], 2 ** 10, 5, 50,
              "FOCanonicalTwoStage", {}),
    "phi2over3": ("phi2over3", 1, {"nfields": 1}, [10.0, 0, 0], 2 ** 10, 5, 50,
                  "FOCanonicalTwoStage", {}),
    # msqphisq_withV0's fixture never ends inflation within tend=83 (the reference raises
    # ModelError at cosmomodels.py:499-500); it is covered by potentials.npz and an error test.
    "step_potential": ("step_potential", 1, {"nfields": 1}, [18.0, -0.1, 0], 2 ** 10, 5, 50,
                       "FOCanonicalTwoStage", {}),
    "bump_potential": ("bump_potential", 1, {"nfields": 1}, [18.0, -0.1, 0], 2 ** 10, 5, 50,
                       "FOCanonicalTwoStage", {}),
    "nflation2": ("nflation", 2, {"nfields": 2}, None, 2 ** 10, 5, 50,
                  "FOCanonicalTwoStage", {}),
    # odd field counts (records only 8-byte aligned at stages 1 and 3) and nfields > 8 (the
    # reference's nflation accepts any nfields, cmpotentials.py:791-846)
    "nflation3": ("nflation", 3, {"nfields": 3}, None, 2 ** 10, 5, 50,
                  "FOCanonicalTwoStage", {}),
    "nflation5": ("nflation", 5, {"nfields": 5}, None, 2 ** 10, 3, 100,
                  "FOCanonicalTwoStage", {}),
    "nflation10": ("nflation", 10, {"nfields": 10}, None, 2 ** 10, 2, 100,
                   "FOCanonicalTwoStage", {}),
    # register-resident instantiations for 9..16 fields (pf_fo_wide.cu): an odd count, and the
    # largest one (whose kernel spills a few registers)
    "nflation13": ("nflation", 13, {"nfields": 13}, None, 2 ** 10, 2, 100,
                   "FOCanonicalTwoStage", {}),
    "nflation16": ("nflation", 16, {"nfields": 16}, None, 2 ** 10, 2, 100,
                   "FOCanonicalTwoStage", {}),
    # unequal fields: the sums over fields (np.sum in cmpotentials.py:837, cosmomodels.py:569) are
    # then sensitive to NumPy's pairwise summation order (8 fields: the tree; 9: tree + remainder)
    "nflation8u": ("nflation", 8, {"nfields": 8}, _unequal(8), 2 ** 10, 2, 100,
                   "FOCanonicalTwoStage", {}),
    "nflation9u": ("nflation", 9, {"nfields": 9}, _unequal(9), 2 ** 10, 2, 100,
                   "FOCanonicalTwoStage", {}),
    "hybridquartic": ("hybridquartic", 2, {"nfields": 2}, [1e-2, 2e-8, 1.63e-9, 3.26e-7, 0],
                      2 ** 10, 5, 50, "FOCanonicalTwoStage", {}),
    "productexponential": ("productexponential", 2, {"nfields": 2}, [18.0, 0.0, 0.001, 0, 0],
                           2 ** 10, 5, 50, "FOCanonicalTwoStage", {}),
    # potentials with no run_config fixture: msqphisq-like starts (they share its mass scale)
    "resonance": ("resonance", 1, {}, [18.0, -0.1, 0], 2 ** 10, 5, 50,
                  "FOCanonicalTwoStage", {}),
    "bump_nothirdderiv": ("bump_nothirdderiv", 1, {}, [18.0, -0.1, 0], 2 ** 10, 5, 50,
                          "FOCanonicalTwoStage", {}),
    # "auto" grid: k chosen so that the modes start 8 % .. 30 % of the way through inflation
    "quartictwofield": ("quartictwofield", 2,
                        {"m1": 5e-6, "m2": 5e-8, "l1": 5e-10, "l2": 5e-14},
                        [20.0, -0.1, 1.0, 0.0, 0], "auto", 5, 50, "FOCanonicalTwoStage", {}),
    "inflection": ("inflection", 2, {}, [0.5, 0.0, 0.0, 0.0, 0], "auto", 5, 50,
                   "FOCanonicalTwoStage", {}),
    "hilltopaxion": ("hilltopaxion", 2, {}, [17.0, -0.1, 0.45, 0.0, 0], 2 ** 10, 5, 50,
                     "FOCanonicalTwoStage", {}),
    # non-default step size / sub-horizon factor / end time (T = around((tend-simtstart)/h)+1)
    "nondefault_steps": ("msqphisq", 1, {"nfields": 1}, [18.0, -0.1, 0], "K4", 5, 25,
                         "FOCanonicalTwoStage", {"tstep_wanted": 0.02, "cq": 40, "tend": 82.0}),
    # initial-condition variants (cosmomodels.py:2017-2197)
    "ic_fixedainit": ("msqphisq", 1, {"nfields": 1}, [18.0, -0.1, 0], "K4", 5, 50,
                      "FixedainitTwoStage", {}),
    "ic_nophase": ("msqphisq", 1, {"nfields": 1}, [18.0, -0.1, 0], "K4", 5, 50,
                   "FONoPhase", {}),
    "ic_suppress": ("hybridquadratic", 2, {"nfields": 2},
                    [12.0, 1 / 300.0, 12.0, 49 / 300.0, 0], 2 ** 10, 5, 50,
                    "FOSuppressOneField", {"suppress_ix": 1}),
}


def keep_rows(T, tsix, stride):
    rows = set(range(0, T, stride))
    for s in np.unique(tsix):
        rows.update(r for r in range(int(s) - 2, int(s) + 6) if 0 <= r < T)
    rows.update(range(max(0, T - 4), T))
    return np.array(sorted(rows), dtype=np.int64)


def make_fo(ref, name):
    pot, nf, params, bgystart, grid, nkeep, stride, cls, extra = FIXTURES[name]
    if grid == "auto":
        probe = getattr(ref.cosmomodels, cls)(
            k=np.array([1e-60]), potential_func=pot, pot_params=dict(params), nfields=nf, cq=50,
            bgystart=None if bgystart is None else np.array(bgystart, dtype=float), **extra)
        probe.runbg()
        Hend = probe.bgmodel.yresult[probe.fotendindex, probe.H_ix]
        ainit = probe.finda_end(Hend) * np.exp(-probe.bgmodel.tresult[probe.fotendindex])
        aH = ainit * np.exp(probe.bgmodel.tresult) * probe.bgmodel.yresult[:, probe.H_ix, 0]
        start_rows = (probe.fotendindex * np.linspace(0.08, 0.3, nkeep)).astype(np.int64)
        assert start_rows[0] > 2, (probe.fotendindex, start_rows)
        k = 50 * aH[start_rows] * (1 - 1e-6)
        kix = np.arange(nkeep)
    else:
        kfull = k4_grid(ref) if grid == "K4" else np.linspace(KMIN, KMAX, grid)
        kix = subsample(len(kfull), nkeep)
        k = kfull[kix]
    kw = dict(k=k, potential_func=pot, pot_params=dict(params), nfields=nf, cq=50,
              solver="rkdriver_tsix")
    kw["bgystart"] = None if bgystart is None else np.array(bgystart, dtype=float)
    kw.update(extra)
    t0 = time.time()
    m = getattr(ref.cosmomodels, cls)(**kw)
    m.run(saveresults=False)
    wall = time.time() - t0
    T = m.yresult.shape[0]
    rows = keep_rows(T, m.fotstartindex, stride)
    bgT = m.bgmodel.yresult.shape[0]
    bgrows = np.array(sorted(set(range(0, bgT, stride)) | set(range(bgT - 3, bgT))
                             | set(int(s) for s in m.fotstartindex)), dtype=np.int64)
    Pr_last = ref.analysis.Pr(m, tix=-1)
    sPr_last = ref.analysis.scaled_Pr(m, tix=-1)
    Pr_rows = ref.analysis.Pr(m)[rows]
    out = dict(
        potential=pot, nfields=nf, cls=cls,
        param_keys=np.array(sorted(params.keys())),
        param_vals=np.array([float(params[q]) for q in sorted(params.keys())]),
        bgystart=np.array(m.bgystart, dtype=float), extra_keys=np.array(sorted(extra.keys())),
        extra_vals=np.array([float(extra[q]) for q in sorted(extra.keys())]),
        grid=str(grid), kix=kix, k=k, cq=float(m.cq), tend=float(m.tend), h=float(m.tstep_wanted),
        simtstart=0.0,
        T=T, rows=rows, tresult=m.tresult, yrows=m.yresult[rows],
        bgT=bgT, bgrows=bgrows, bgyrows=m.bgmodel.yresult[bgrows],
        bgepsilon_rows=m.bgepsilon[bgrows],
        ainit=np.asarray(m.ainit, dtype=float), etainit=np.asarray(m.etainit, dtype=float),
        fotend=m.fotend, fotendindex=m.fotendindex,
        fotstart=m.fotstart, fotstartindex=m.fotstartindex, foystart=m.foystart,
        Pr_last=Pr_last, scaled_Pr_last=sPr_last, Pr_rows=Pr_rows,
        ref_wall_s=wall,
    )
    path = os.path.join(HERE, "fo_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("%-26s nf=%d nk=%d T=%d rows=%d wall=%.1fs -> %.0f KB" % (
        name, nf, len(k), T, len(rows), wall, os.path.getsize(path) / 1024.0))


# second-order stage: name -> (first-order fixture it starts from, #modes kept)
SO_FIXTURES = {
    "so_msqphisq": ("c1_msqphisq_k4", 4),
    "so_lambdaphi4": ("c2_lambdaphi4_2p16", 4),       # d2U depends on phi
}


def make_so(ref, name):
    """SOCanonicalThreeStage on top of a reference first-order run, with oracle.synthetic_source
    as the source term, for CanonicalSecondOrder and CanonicalRampedSecondOrder; plus single
    right-hand-side evaluations of the three second-order classes."""
    import pyflation_oracle as oracle
    fo_name, nkeep = SO_FIXTURES[name]
    pot, nf, params, bgystart, grid, _, _, cls, extra = FIXTURES[fo_name]
    kfull = k4_grid(ref) if grid == "K4" else np.linspace(KMIN, KMAX, grid)
    kix = subsample(len(kfull), nkeep)
    k = kfull[kix]
    c = ref.cosmomodels
    m = getattr(c, cls)(k=k, potential_func=pot, pot_params=dict(params), nfields=nf, cq=50,
                        solver="rkdriver_tsix", bgystart=np.array(bgystart, dtype=float), **extra)
    m.run(saveresults=False)
    m.source = oracle.synthetic_source(m.tresult, k)
    out = dict(potential=pot, nfields=nf, fo_fixture=fo_name,
               param_keys=np.array(sorted(params.keys())),
               param_vals=np.array([float(params[q]) for q in sorted(params.keys())]),
               bgystart=np.array(m.bgystart, dtype=float), k=k, kix=kix,
               ainit=np.asarray(m.ainit, dtype=float), fo_tstep=float(m.tstep_wanted),
               fo_simtstart=float(m.simtstart), fo_tresult=m.tresult,
               fo_bg0=np.real(m.yresult[:, 0:3, 0]), fotstart=m.fotstart,
               fotstartindex=m.fotstartindex, bgepsilon=m.bgepsilon)
    fo_y = m.yresult
    for tag, soclass in (("plain", c.CanonicalSecondOrder), ("ramped", c.CanonicalRampedSecondOrder)):
        t0 = time.time()
        so = c.SOCanonicalThreeStage(m, soclass=soclass)
        so.run(saveresults=False)
        T = so.yresult.shape[0]
        rows = keep_rows(T, so.tstartindex, 40)
        out[tag + "_tstartindex"] = np.asarray(so.tstartindex)
        out[tag + "_tresult"] = so.tresult
        out[tag + "_rows"] = rows
        out[tag + "_yrows"] = so.yresult[rows]
        out[tag + "_wall_s"] = time.time() - t0
        out["so_tstep"] = float(so.tstep_wanted)
        out["so_simtstart"] = float(so.simtstart)
        out["so_tend"] = float(so.tend)
        # single right-hand sides: a running row, a half step, and rows inside the ramp window
        sm = so.somodel
        probe_rows = sorted(set([int(so.tstartindex.max()) + 200, int(so.tstartindex.min()) + 3,
                                 int(so.tstartindex.min()), int(so.tstartindex.max()) + 20]))
        ts, ys, ds = [], [], []
        for r in probe_rows:
            for half in (0.0, 0.5):
                t = so.simtstart + (r + half) * so.tstep_wanted
                y = np.nan_to_num(so.yresult[r]) + 1e-3 * np.arange(1, 5)[:, None]
                ts.append(t); ys.append(y); ds.append(sm.derivs(y, t))
        out[tag + "_derivs_t"] = np.array(ts)
        out[tag + "_derivs_y"] = np.array(ys)
        out[tag + "_derivs_dydx"] = np.array(ds)
        m.yresult = fo_y                     # SOCanonicalThreeStage re-binds it (cosmomodels.py:1749)
    out["rampargs_keys"] = np.array(sorted(so.somodel.rampargs.keys()))
    out["rampargs_vals"] = np.array([float(so.somodel.rampargs[q]) for q in sorted(so.somodel.rampargs.keys())])
    # homogeneous class: derivs only (it needs a tix keyword no solver of the reference passes)
    hm = c.CanonicalHomogeneousSecondOrder(k=k, ainit=m.ainit, ystart=np.zeros((4, len(k))),
                                           tstart=m.fotstart, tstartindex=so.tstartindex,
                                           simtstart=so.simtstart, tend=so.tend,
                                           tstep_wanted=so.tstep_wanted, solver="rkdriver_tsix",
                                           potential_func=pot, pot_params=dict(params), nfields=nf)
    hm.second_stage = m
    tix = int(m.fotstartindex.max()) + 300
    yh = 1e-3 * np.arange(1, 5)[:, None] * np.ones((1, len(k)))
    out["homog_tix"] = tix
    out["homog_t"] = float(m.tresult[tix])
    out["homog_y"] = yh
    out["homog_dydx"] = hm.derivs(yh, m.tresult[tix], tix=tix)
    path = os.path.join(HERE, "%s.npz" % name)
    np.savez_compressed(path, **out)
    print("%-26s nk=%d T_so=%d rows=%d -> %.0f KB" % (name, len(k), T, len(rows), os.path.getsize(path) / 1024.0))


def make_potentials(ref):
    rng = np.random.RandomState(20261019)
    out = {}
    cm = ref.cmpotentials
    single = ["msqphisq", "lambdaphi4", "linde", "hybrid2and4", "phi2over3", "msqphisq_withV0",
              "step_potential", "bump_potential", "resonance", "bump_nothirdderiv"]
    two = ["hybridquadratic", "ridge_twofield", "quartictwofield", "hybridquartic", "inflection",
           "hilltopaxion", "productexponential"]
    pts1 = np.concatenate([[18.0, 14.84, 14.85, 14.80, 0.5, 25.0], rng.uniform(0.1, 26.0, 10)])
    pts2 = np.column_stack([rng.uniform(0.01, 18.0, 12), rng.uniform(-0.2, 0.2, 12),
                            rng.uniform(0.001, 12.0, 12), rng.uniform(-0.2, 0.2, 12),
                            rng.uniform(1e-6, 1e-4, 12)])
    q2params = {"m1": 5e-6, "m2": 5e-8, "l1": 5e-10, "l2": 5e-14}

    def pack(name, ys, params, tag):
        U, dU, d2U, d3U = [], [], [], []
        for y in ys:
            r = getattr(cm, name)(np.array(y), params)
            U.append(r[0]); dU.append(np.asarray(r[1])); d2U.append(np.asarray(r[2]))
            d3U.append(np.zeros(1) if r[3] is None else np.asarray(r[3], dtype=complex).ravel())
        out[tag + "_y"] = np.array(ys)
        out[tag + "_U"] = np.array(U)
        out[tag + "_dU"] = np.array(dU)
        out[tag + "_d2U"] = np.array(d2U)
        out[tag + "_d3U"] = np.array(d3U)

    for name in single:
        ys = [np.array([p, -0.1, 1e-5]) for p in pts1]
        pack(name, ys, None, name + "__default")
        pack(name, [y.astype(complex) for y in ys], None, name + "__complex")
    for name in two:
        params = q2params if name == "quartictwofield" else {}
        pack(name, list(pts2), params, name + "__default")
    for nf in (1, 2, 3, 8):
        ys = [np.concatenate([rng.uniform(1, 18, 2 * nf), [1e-5]]) for _ in range(6)]
        pack("nflation", ys, {"nfields": nf}, "nflation__nf%d" % nf)
    # non-default parameters
    pack("msqphisq", [np.array([17.0, -0.1, 1e-5])], {"mass": 1e-5}, "msqphisq__custom")
    pack("lambdaphi4", [np.array([20.0, -0.1, 1e-5])], {"lambda": 2e-13}, "lambdaphi4__custom")
    pack("step_potential", [np.array([14.83, -0.1, 1e-5])],
         {"mass": 6e-6, "c": 0.002, "d": 0.03, "phi_s": 14.8}, "step_potential__custom")
    pack("hybridquadratic", [pts2[0]], {"m1": 2e-6, "m2": 8e-6}, "hybridquadratic__custom")
    pack("nflation", [np.array([3.0, 0.1, 4.0, 0.2, 1e-5])], {"nfields": 2, "mass": 1e-6},
         "nflation__custom")
    path = os.path.join(HERE, "potentials.npz")
    np.savez_compressed(path, **out)
    print("potentials -> %.0f KB" % (os.path.getsize(path) / 1024.0))


def make_kat(ref):
    """The reference's hand-computed P_R vectors (test_adiabatic.py:81-260), answers
    recomputed here by the reference's own functions."""
    ad = ref.analysis.adiabatic
    out = {}
    cases = {        # shapes exactly as test_adiabatic.py:155-260 builds them
        "two_real": (np.array([[1, 3], [2, 5]], dtype=float).reshape((1, 2, 2, 1)),
                     np.array([7.0, 9.0]).reshape((1, 2, 1)), 1),
        "two_complex": (np.array([[1, 1j], [-1j, 3 - 1j]]).reshape((1, 2, 2, 1)),
                        np.array([1.0, 1.0]).reshape((1, 2, 1)), 1),
        "single": (np.array([[[7.0]]]), np.array([[1.7]]), 0),
    }
    for tag, (modes, phidot, axis) in cases.items():
        out[tag + "_modes"] = modes
        out[tag + "_phidot"] = phidot
        out[tag + "_axis"] = axis
        out[tag + "_Pr"] = np.asarray(ad.Pr_spectrum(phidot, modes, axis))
        out[tag + "_Pphi"] = np.asarray(ad.Pphi_matrix(modes, axis))
    modes = (np.arange(72.0).reshape((4, 3, 3, 2)) + 0j) * (1 + 0.5j)
    phidot = np.arange(1.0, 25.0).reshape((4, 3, 2))
    out["block_modes"] = modes
    out["block_phidot"] = phidot
    out["block_Pr"] = ad.Pr_spectrum(phidot, modes, 1)
    out["block_Pphi"] = ad.Pphi_matrix(modes, 1)
    k = np.array([1e-60, 5.25e-60, 3e-58])
    out["kscaling_k"] = k
    out["kscaling"] = ref.analysis.utilities.kscaling(k)
    path = os.path.join(HERE, "kat_adiabatic.npz")
    np.savez_compressed(path, **out)
    print("kat_adiabatic -> %.0f KB" % (os.path.getsize(path) / 1024.0))


def make_pzeta(ref):
    """P_zeta of the last stored rows of every first-order golden file, computed by the
    reference's analysis.Pzeta on a model with injected yresult (the way the reference's own
    tests build their models, test_adiabatic.py:138-144), plus the reference's hand-computed
    Pzeta_spectrum vectors (test_adiabatic.py:262-327)."""
    import glob
    out = {}
    for path in sorted(glob.glob(os.path.join(HERE, "fo_*.npz"))):
        name = os.path.basename(path)[3:-4]
        g = np.load(path)
        nf = int(g["nfields"])
        params = {str(a): (int(b) if str(a) == "nfields" else float(b)) for a, b in zip(g["param_keys"], g["param_vals"])}
        m = ref.cosmomodels.FOCanonicalTwoStage(k=g["k"], nfields=nf, potential_func=str(g["potential"]),
                                                pot_params=params)
        m.yresult = g["yrows"][-3:]
        m.tresult = g["tresult"][g["rows"][-3:]]
        out[name + "_Pzeta"] = ref.analysis.Pzeta(m)
        out[name + "_scaled_Pzeta"] = ref.analysis.scaled_Pzeta(m, tix=-1)
    ad = ref.analysis.adiabatic
    out["kat_single"] = np.asarray(ad.Pzeta_spectrum(3, 0.5, 2, np.array([[7]]), np.array([[5]]), 0))
    out["kat_two"] = ad.Pzeta_spectrum(np.array([5.5, 2.3]).reshape((2, 1)), np.array([2, 5]).reshape((2, 1)),
                                       np.array([3]).reshape((1, 1)),
                                       np.array([[1 / 3.0, 0.1], [0.1, 0.5]]).reshape((2, 2, 1)),
                                       np.array([[0.1, 0.2], [0.2, 1 / 7.0]]).reshape((2, 2, 1)), 0)
    out["kat_complex"] = ad.Pzeta_spectrum(np.array([1, 2]).reshape((2, 1)), np.array([1, 1]).reshape((2, 1)),
                                           np.array([1]).reshape((1, 1)),
                                           np.array([[1, 1j], [-1j, 3 - 1j]]).reshape((2, 2, 1)),
                                           np.array([[1, -1j], [1j, 3 + 1j]]).reshape((2, 2, 1)), 0)
    path = os.path.join(HERE, "pzeta.npz")
    np.savez_compressed(path, **out)
    print("pzeta -> %.0f KB" % (os.path.getsize(path) / 1024.0))


def main(argv):
    ref = ref_shim.load()
    names = argv[1:] or (["potentials", "kat"] + list(FIXTURES) + ["pzeta"] + list(SO_FIXTURES))
    for name in names:
        if name == "potentials":
            make_potentials(ref)
        elif name == "kat":
            make_kat(ref)
        elif name == "pzeta":
            make_pzeta(ref)
        elif name in SO_FIXTURES:
            make_so(ref, name)
        else:
            make_fo(ref, name)


if __name__ == "__main__":
    main(sys.argv)
