"""cosmomodels - drop-in model classes for the first-order path of ``pyflation.cosmomodels``.

Class names, constructor signatures, ``.run()`` and the result attributes (``yresult``,
``tresult``, ``k``, ``fotstart``, ``fotstartindex``, ``foystart``, ``ainit`` and the index
slices ``H_ix, bg_ix, phis_ix, phidots_ix, pert_ix, dps_ix, dpdots_ix``) follow the reference
(cosmomodels.py:53-206, 481-675, 973-1550, 2017-2197).  The integration itself is not done
here: ``run()`` hands the model to :mod:`pyflation_b200.rk4`, which launches the CUDA kernels.

Second-order stage (cosmomodels.py:678-971, 1680-1946): ``CanonicalSecondOrder``,
``CanonicalRampedSecondOrder``, ``CanonicalHomogeneousSecondOrder`` and the drivers
``SOCanonicalThreeStage`` / ``SOHorizonStart`` integrate on the GPU as well (pf_so_run) from a
first-order model that carries a ``source`` array; computing that source term (the reference's
``sourceterm`` package) is out of scope.  Results files: the reference's HDF5 layout through
:mod:`pyflation_b200.hdf5lite` (no PyTables here), or ``.npz`` archives with the same names.
"""
import datetime
import logging
import os

import numpy as np
from scipy import interpolate

from .configuration import _debug
from . import cmpotentials
from . import rk4
from . import analysis
from . import hdf5lite

module_logger = logging.getLogger(__name__)


class ModelError(Exception):
    """Generic error for model simulating. (cosmomodels.py:49)"""
    pass


class CosmologicalModel(object):
    """Base class: argument checking, potential / solver lookup, ``run()``.
    (cosmomodels.py:53-206)"""
    solverlist = ["rkdriver_tsix"]

    def __init__(self, ystart=None, simtstart=0.0, tstart=0.0, tstartindex=None,
                 tend=83.0, tstep_wanted=0.01, solver="rkdriver_tsix",
                 potential_func=None, pot_params=None, nfields=1, **kwargs):
        self._log = logging.getLogger('%s.%s' % (__name__, self.__class__.__name__))
        self.ystart = ystart
        self.k = getattr(self, "k", None)

        self.tstartindex = np.array([0]) if tstartindex is None else tstartindex

        if np.all(tstart < tend):
            self.tstart, self.tend = tstart, tend
        elif np.all(tstart == tend):
            raise ValueError("End time is the same as start time!")
        else:
            raise ValueError("Ending time is before starting time!")

        if np.all(simtstart < tend):
            self.simtstart = simtstart
        elif simtstart == tend:
            raise ValueError("End time is same a simulation start time!")
        else:
            raise ValueError("End time is before simulation start time!")

        self.tstep_wanted = tstep_wanted

        if solver not in self.solverlist:
            raise ValueError("Solver not recognized!")
        self.solver = solver

        if potential_func is None:
            potential_func = "msqphisq"
        # AttributeError for an unknown name, like cmpotentials.__getattribute__ (:138)
        self.potentials = getattr(cmpotentials, potential_func)
        self.potential_func = potential_func

        if pot_params is None:
            pot_params = {}
        if not isinstance(pot_params, dict):
            raise ModelError("Need to provide pot_params as a dictionary of parameters.")
        self.pot_params = pot_params

        if nfields < 1:
            raise ValueError("Cannot have zero or negative number of fields.")
        self.nfields = nfields

        self.tresult = None
        self.yresult = None

    def derivs(self, yarray, t):
        """Right-hand side of the model's ODE system; subclasses name theirs."""
        pass

    def potentials(self, y, pot_params=None):
        pass

    def findH(self, potential, y):
        pass

    def run(self, saveresults=True):
        """Execute a simulation run using the parameters already provided.
        (cosmomodels.py:173-206)"""
        if self.solver not in self.solverlist:
            raise ModelError("Unknown solver!")
        if not hasattr(self, "tstartindex"):
            raise ModelError("Need to specify initial starting indices!")
        if _debug:
            self._log.debug("Starting simulation with %s.", self.solver)
        solver = getattr(rk4, self.solver)
        extra = {}
        if getattr(self, "row_stride", 1) != 1:      # extension: strided / last-row output
            extra["row_stride"] = self.row_stride
        try:
            self.tresult, self.yresult = solver(ystart=self.ystart,
                                                simtstart=self.simtstart,
                                                tsix=self.tstartindex,
                                                tend=self.tend,
                                                h=self.tstep_wanted,
                                                derivs=self.derivs, **extra)
        except Exception:
            self._log.exception("Error running %s!", self.solver)
            raise
        self.lastparams = self.callingparams()
        if saveresults:
            try:
                fname = self.saveallresults()
                self._log.info("Results saved in " + fname)
            except IOError as er:
                self._log.error("Error trying to save results! Results NOT saved.\n" + str(er))
        return

    def callingparams(self):
        """Dictionary of the parameters the run was called with."""
        return {"tstart": self.tstart,
                "tend": self.tend,
                "tstep_wanted": self.tstep_wanted,
                "solver": self.solver,
                "classname": self.__class__.__name__,
                "datetime": datetime.datetime.now().strftime("%Y%m%d%H%M%S"),
                "nfields": self.nfields}

    def gethf5paramsdict(self):
        """Columns of the ``parameters`` table of a results file (cosmomodels.py:221-234)."""
        return {"solver": hdf5lite.StringCol(50),
                "classname": hdf5lite.StringCol(255),
                "tstart": hdf5lite.Float64Col(np.shape(self.tstart)),
                "simtstart": hdf5lite.Float64Col(),
                "tend": hdf5lite.Float64Col(),
                "tstep_wanted": hdf5lite.Float64Col(),
                "datetime": hdf5lite.Float64Col(),
                "nfields": hdf5lite.IntCol()}

    def saveallresults(self, filename=None, filetype=None, yresultshape=None, **kwargs):
        """Save the results of the last run (cosmomodels.py:236-277).

        ``filetype`` "hf5" writes the reference's HDF5 layout -- groups ``results`` /
        ``bgresults``, extendable ``yresult`` / ``tresult`` arrays, the ``parameters`` and
        ``pot_params`` tables, an existing file is appended to -- through
        :mod:`pyflation_b200.hdf5lite` (PyTables is not available here); "npz" writes the same
        dataset names into a NumPy archive.  With no ``filetype`` the file name's extension
        decides, and a missing file name means ``run<datetime>.hf5`` as in the reference."""
        if self.yresult is None:
            raise ModelError("No results to save yet.")
        if filetype is None:
            ext = os.path.splitext(filename)[1].lower() if filename else ".hf5"
            filetype = {".npz": "npz", ".hf5": "hf5", ".h5": "hf5", ".hdf5": "hf5", ".hdf": "hf5"}.get(ext, "npz")
        if filetype not in ("npz", "hf5"):
            raise NotImplementedError("Saving results in format %s is not implemented." % filetype)
        if not filename:
            now = (getattr(self, "lastparams", None) or self.callingparams())["datetime"]
            filename = os.path.join(os.getcwd(), "run" + now + "." + filetype)
            self._log.info("Filename set to " + filename)
        if not os.path.isdir(os.path.dirname(os.path.abspath(filename))):
            raise IOError("Directory %s does not exist" % os.path.dirname(filename))
        if filetype == "hf5":
            grpname = "bgresults" if getattr(self, "k", None) is None else "results"      # :263-267
            if yresultshape is None:
                yresultshape = [0] + list(np.shape(self.yresult))[1:]
            if os.path.isfile(filename) and hdf5lite.is_hdf5_file(filename):
                rf = hdf5lite.open_file(filename, "a")                                  # append mode (:247-249)
            else:
                rf = self.createhdf5structure(filename, grpname, yresultshape, **kwargs)
            try:
                self.saveresultsinhdf5(rf, grpname)
            finally:
                rf.close()
            return filename
        data = dict(yresult=self.yresult, tresult=self.tresult)
        for name in ("k", "ystart", "bgystart", "foystart", "fotstart", "fotstartindex", "ainit"):
            if getattr(self, name, None) is not None:
                data[name] = np.asarray(getattr(self, name))
        if getattr(self, "bgmodel", None) is not None:
            data["bg_tresult"] = self.bgmodel.tresult
            data["bg_yresult"] = self.bgmodel.yresult
        src = self._source_for_file()
        if src is not None:
            data["sourceterm"] = src
        for name in ("potential_func", "nfields", "tend", "tstep_wanted", "simtstart", "cq"):
            if getattr(self, name, None) is not None:
                data["param_" + name] = np.asarray(getattr(self, name))
        if getattr(self, "pot_params", None):
            data["pot_param_keys"] = np.array(sorted(self.pot_params.keys()))
            data["pot_param_vals"] = np.array([float(self.pot_params[q]) for q in sorted(self.pot_params.keys())])
        data["classname"] = np.asarray(self.__class__.__name__)
        np.savez(filename, **data)
        return filename

    def _source_for_file(self):
        """The second-order source term a first-order model carries, as a host array: the
        reference keeps it next to the first-order results (results.sourceterm, :1607-1612)."""
        src = getattr(self, "source", None)
        if src is None or isinstance(self, SOCanonicalThreeStage):
            return None
        return src.cpu().numpy() if hasattr(src, "cpu") else np.asarray(src)

    def createhdf5structure(self, filename, grpname="results", yresultshape=None, hdf5complevel=0,
                            hdf5complib="zlib"):
        """A new HDF5 file with the reference's structure (cosmomodels.py:279-349); returns the open
        file.  The reference compresses with blosc, level 2, which is a PyTables plug-in and not
        an HDF5 codec: the default here is no compression, ``hdf5complevel`` > 0 uses zlib."""
        rf = hdf5lite.open_file(filename, "w")
        try:
            filters = hdf5lite.Filters(complevel=hdf5complevel, complib=hdf5complib)
            resgroup = rf.create_group(rf.root, grpname, "Results of simulation")
            rf.create_earray(resgroup, "tresult", hdf5lite.Float64Atom(), (0,), filters=filters, expectedrows=8194)
            rf.create_table(resgroup, "parameters", self.gethf5paramsdict(), filters=filters)
            rf.create_table(resgroup, "pot_params", {"name": hdf5lite.StringCol(255),
                                                     "value": hdf5lite.Float64Col()}, filters=filters)
            if grpname == "results":
                if getattr(self, "bgmodel", None) is not None:
                    bggrp = rf.create_group(rf.root, "bgresults", "Background results")
                    rf.create_array(bggrp, "tresult", np.asarray(self.bgmodel.tresult))
                    rf.create_array(bggrp, "yresult", np.asarray(self.bgmodel.yresult))
                rf.create_earray(resgroup, "yresult", hdf5lite.ComplexAtom(itemsize=16), tuple(yresultshape),
                                 filters=filters, expectedrows=8194)
                rf.create_array(resgroup, "k", np.asarray(self.k))
                if getattr(self, "ystart", None) is not None:
                    rf.create_array(resgroup, "ystart", np.asarray(self.ystart))
                if getattr(self, "bgystart", None) is not None:
                    rf.create_array(resgroup, "bgystart", np.asarray(self.bgystart))
                for name in ("foystart", "fotstart", "fotstartindex"):
                    if getattr(self, name, None) is not None:
                        rf.create_array(resgroup, name, np.asarray(getattr(self, name)))
                src = self._source_for_file()
                if src is not None:                     # what the wrapper reads back (:1607-1612)
                    rf.create_array(resgroup, "sourceterm", src)
            else:
                rf.create_earray(resgroup, "yresult", hdf5lite.Float64Atom(), tuple(yresultshape),
                                 filters=filters, expectedrows=8300)
        except Exception:
            rf.isopen = False                           # nothing useful to write
            raise
        return rf

    def saveresultsinhdf5(self, rf, grpname="results"):
        """Append this run to the open results file (cosmomodels.py:351-402)."""
        resgrp = rf.get_node(rf.root, grpname)
        paramstab = resgrp.parameters
        row = paramstab.row
        params = self.callingparams()
        for key in params:
            if key in paramstab.colnames:
                row[key] = params[key]
        for key in ("simtstart", "cq"):                  # columns beyond the calling parameters
            if key in paramstab.colnames and getattr(self, key, None) is not None:
                row[key] = getattr(self, key)
        row.append()
        paramstab.flush()
        potparamstab = resgrp.pot_params
        for key in self.pot_params or {}:
            row = potparamstab.row
            row["name"] = key
            row["value"] = self.pot_params[key]
            row.append()
        potparamstab.flush()
        yres = np.asarray(self.yresult)
        resgrp.yresult.append(yres if yres.dtype == resgrp.yresult.dtype else yres.astype(resgrp.yresult.dtype))
        resgrp.tresult.append(np.asarray(self.tresult, dtype=np.float64))
        rf.flush()


class PhiModels(CosmologicalModel):
    """Models in the e-fold time scheme of Malik 06 (cosmomodels.py:474-522)."""

    def findH(self, U, y):
        """Hubble parameter from the Friedmann constraint, H^2 = U / (3 - phidot^2/2)."""
        phidot = y[self.phidots_ix]
        return np.sqrt(U / (3.0 - 0.5 * (np.sum(phidot ** 2))))

    def getepsilon(self):
        """epsilon = 1/2 sum_i phidot_i^2 for every stored row."""
        if len(self.yresult.shape) == 3:
            phidots = self.yresult[:, self.phidots_ix, 0]
        else:
            phidots = self.yresult[:, self.phidots_ix]
        return 0.5 * np.sum(phidots ** 2, axis=1)

    def findinflend(self):
        """(efold, row index) of the end of inflation, epsilon >= 1; the efold is refined
        with a cubic spline through the rows before it (cosmomodels.py:493-511)."""
        self.epsilon = self.getepsilon()
        if not any(self.epsilon > 1):
            raise ModelError("Inflation did not end during specified number of efoldings. "
                             "Increase tend and try again!")
        endindex = np.where(self.epsilon >= 1)[0][0]
        tck = interpolate.splrep(self.tresult[:endindex], self.epsilon[:endindex])
        t2 = np.linspace(self.tresult[endindex - 1], self.tresult[endindex], 100)
        y2 = interpolate.splev(t2, tck)
        endefold = t2[np.where(y2 > 1)[0][0]]
        return endefold, endindex


class CanonicalBackground(PhiModels):
    """Background model in e-fold time: y = (phi_i, dphi_i/dn, ..., H).
    (cosmomodels.py:525-571)"""

    def __init__(self, *args, **kwargs):
        super(CanonicalBackground, self).__init__(*args, **kwargs)
        self.H_ix = self.nfields * 2
        self.bg_ix = slice(0, self.nfields * 2 + 1)
        self.phis_ix = slice(0, self.nfields * 2, 2)
        self.phidots_ix = slice(1, self.nfields * 2, 2)
        if np.all(self.ystart[self.H_ix] == 0.0):
            U = self.potentials(self.ystart, self.pot_params)[0]
            self.ystart[self.H_ix] = self.findH(U, self.ystart)

    def derivs(self, y, t, **kwargs):
        """Basic background equations of motion, dydx[0] = dy[0]/dn etc.  Evaluated on the GPU
        (pf_derivs); the solver never calls this method, it recognises it and runs the fused
        background kernel instead."""
        from . import engine
        pm = engine.make_model(self.potential_func, self.pot_params, self.nfields)
        return engine.derivs_device(pm, y, t)


class CanonicalFirstOrder(PhiModels):
    """First order model: background rows, H, then the nfields x nfields complex mode matrix
    and its e-fold derivative, interleaved (cosmomodels.py:573-675)."""

    def __init__(self, k=None, ainit=None, *args, **kwargs):
        super(CanonicalFirstOrder, self).__init__(*args, **kwargs)
        self.ainit = 1 if ainit is None else ainit
        self.k = 10 ** (np.arange(10.0) - 8) if k is None else k
        self.setfieldindices()
        if self.ystart is None:
            self.ystart = np.array([18.0, -0.1] * self.nfields + [0.0] + [1.0, 0.0] * self.nfields)
        if np.all(self.ystart[self.H_ix] == 0.0):
            U = self.potentials(self.ystart, self.pot_params)[0]
            self.ystart[self.H_ix] = self.findH(U, self.ystart)

    def setfieldindices(self):
        nf = self.nfields
        self.H_ix = nf * 2
        self.bg_ix = slice(0, nf * 2 + 1)
        self.phis_ix = slice(0, nf * 2, 2)
        self.phidots_ix = slice(1, nf * 2, 2)
        self.pert_ix = slice(nf * 2 + 1, None)
        self.dps_ix = slice(nf * 2 + 1, None, 2)
        self.dpdots_ix = slice(nf * 2 + 2, None, 2)

    def derivs(self, y, t, **kwargs):
        """Return derivatives of fields in y at time t (cosmomodels.py:625-675).  Evaluated on
        the GPU (pf_derivs); the solver never calls this method, it recognises it and runs the
        fused first-order kernels instead."""
        from . import engine
        k = kwargs.get("k")
        if k is None:
            k = self.k
        pm = engine.make_model(self.potential_func, self.pot_params, self.nfields, self.ainit)
        return engine.derivs_device(pm, y, t, k)


class _SecondOrderModel(PhiModels):
    """What the three second-order classes share: y = (Re d2, Re d2', Im d2, Im d2') per mode,
    integrated in e-folds from the rows of a first-order model (``second_stage``) and a source
    term sampled on its time grid (``source``); both are attached by the driver
    (cosmomodels.py:1755-1775)."""
    ramped = False
    with_source = True

    def __init__(self, k=None, ainit=None, *args, **kwargs):
        super(_SecondOrderModel, self).__init__(*args, **kwargs)
        self.ainit = 1 if ainit is None else ainit
        self.k = 10 ** (np.arange(10.0) - 8) if k is None else k
        if self.ystart is None:
            self.ystart = np.array([0.0, 0.0, 0.0, 0.0])

    def _first_order_inputs(self):
        fo = self.second_stage
        y = fo.yresult
        if y is None or y.shape[0] != len(fo.tresult) or y.shape[1] != 5:
            raise ModelError("Second order models need the full single field first order trajectory.")
        if not hasattr(fo, "bgepsilon"):
            fo.bgepsilon = fo.bgmodel.getepsilon()
        return dict(fo_bg0=np.real(y[:, 0:3, 0]), bgepsilon=fo.bgepsilon,
                    fo_simtstart=fo.simtstart, fo_tstep=fo.tstep_wanted,
                    fo_tsix=getattr(fo, "fotstartindex", None))

    def derivs(self, y, t, **kwargs):
        """Equation of motion for second order perturbations (cosmomodels.py:708-766, 797-848,
        893-971), evaluated on the GPU (pf_so_derivs).  The solver never calls this method: it
        recognises it and runs the second-order kernel."""
        from . import engine
        if kwargs.get("k") is None:
            k, kix = self.k, np.arange(len(np.atleast_1d(self.k)))
        else:
            k, kix = kwargs["k"], kwargs.get("kix")
        if kix is None:
            raise ModelError("Need to specify kix in order to calculate 2nd order perturbation!")
        kix = np.atleast_1d(kix)
        fotix = None
        if not self.with_source:
            if kwargs.get("tix") is None:
                raise ModelError("Need to specify tix in order to calculate 2nd order perturbation!")
            fotix = int(kwargs["tix"])
        inp = self._first_order_inputs()
        if inp["fo_tsix"] is not None:
            inp["fo_tsix"] = np.asarray(inp["fo_tsix"])[kix]
        row = None
        if self.with_source:
            row = int(np.around((t - inp["fo_simtstart"]) / inp["fo_tstep"]))
            row = np.asarray(self.source[row])[kix]
        pm = engine.make_model(self.potential_func, self.pot_params, self.nfields, self.ainit)
        extra = {}
        if self.ramped:
            extra = dict(ramp=self.rampargs, tstart=np.broadcast_to(self.tstart, np.shape(self.k))[kix],
                         tsix=np.asarray(self.tstartindex)[kix])
        return engine.so_derivs_device(pm, y, t, np.atleast_1d(k), so_tstep=self.tstep_wanted,
                                       source_row=row, fotix=fotix, **dict(inp, **extra))


class CanonicalSecondOrder(_SecondOrderModel):
    """Second order model using efold as time variable (cosmomodels.py:678-766)."""


class CanonicalHomogeneousSecondOrder(_SecondOrderModel):
    """Second order homogeneous model: no source term; ``derivs`` needs the first-order row as
    the keyword ``tix`` (cosmomodels.py:768-848), which no solver of the reference passes, so it
    offers single right-hand-side evaluations only."""
    with_source = False


class CanonicalRampedSecondOrder(_SecondOrderModel):
    """Second order model whose source term is switched on with a tanh ramp around each mode's
    start time (cosmomodels.py:850-971):
    ramp = (tanh(a*(t - t_ic - b)) + c)/d  while  abs(t - t_ic - b) < e."""
    ramped = True

    def __init__(self, k=None, ainit=None, *args, **kwargs):
        super(CanonicalRampedSecondOrder, self).__init__(k, ainit, *args, **kwargs)
        if "rampargs" not in kwargs:
            self.rampargs = {"a": 15.0, "b": 0.3, "c": 1, "d": 2, "e": 1}
        else:
            self.rampargs = kwargs["rampargs"]


class MultiStageDriver(CosmologicalModel):
    """Parent of the multi-stage drivers: horizon crossing and post-inflation bookkeeping.
    (cosmomodels.py:973-1227)"""

    def __init__(self, *args, **kwargs):
        super(MultiStageDriver, self).__init__(*args, **kwargs)
        self.cq = kwargs["cq"] if "cq" in kwargs else 50

    def find_efolds_after_inflation(self, Hend, Hreh=None):
        """Efolds from the end of inflation until today (instant reheating by default)."""
        if Hreh is None:
            Hreh = Hend
        return 72.3 + 2.0 / 3.0 * np.log(Hend) - 1.0 / 6.0 * np.log(Hreh)

    def finda_end(self, Hend, Hreh=None, a_0=1):
        return a_0 * np.exp(-self.find_efolds_after_inflation(Hend, Hreh))

    def finda_0(self, Hend, Hreh=None, a_end=None):
        if a_end is None:
            try:
                a_end = self.ainit * np.exp(self.tresult[-1])
            except TypeError:
                raise ModelError("Simulation has not been run yet.")
        return a_end * np.exp(self.find_efolds_after_inflation(Hend, Hreh))

    def _crossing_rows(self, k, t, H, factor=None):
        """First row of ``t`` with ``k < factor*a*H`` for every k at once.

        The reference scans ``np.where(np.sign(k - factor*aH) < 0)[0][0]`` per mode in a
        Python loop (cosmomodels.py:1108, 1131).  The first index at which a sequence exceeds k
        is also the first index at which its running maximum does, and the running maximum is
        sorted, so one ``searchsorted`` gives the same rows in O(nk log T)."""
        if factor is None:
            factor = self.cq
        H = np.asarray(H)
        if len(H.shape) > 1:
            H = H[:, 0]
        aH = factor * (self.ainit * np.exp(t) * H)
        rows = np.searchsorted(np.maximum.accumulate(aH), np.atleast_1d(k), side="right")
        late = rows >= len(aH)
        if np.any(late):
            raise ModelError("k mode " + str(np.atleast_1d(k)[late][0])
                             + " crosses horizon after end of inflation!")
        return rows

    def findkcrossing(self, k, t, H, factor=None):
        """(row index, efold) at which one mode satisfies k = factor*a*H."""
        row = int(self._crossing_rows(k, t, H, factor)[0])
        return row, t[row]

    def findallkcrossings(self, t, H):
        """Array of (row index, efold) pairs for every k in ``self.k``."""
        rows = self._crossing_rows(self.k, t, H)
        return np.column_stack((rows.astype(float), t[rows]))

    def findHorizoncrossings(self, factor=1):
        """(row, efold) of k = factor*a*H for every mode, from the first-order results."""
        H = np.real(self.yresult[:, self.H_ix, :])
        out = []
        for onek, oneH in zip(self.k, np.rollaxis(H, -1, 0)):
            aH = self.ainit * np.exp(self.tresult) * oneH
            hit = np.where(np.sign(onek - factor * aH) < 0)[0]
            if hit.size == 0:
                raise ModelError("k mode " + str(onek) + " crosses horizon after end of inflation!")
            out.append((hit[0], self.tresult[hit[0]]))
        return np.array(out)

    @property
    def Pr(self):
        """Unscaled power spectrum of the comoving curvature perturbation, all rows and k."""
        return analysis.Pr(self)

    @property
    def scaled_Pr(self):
        """k^3/(2 pi^2) P_R for all rows and k."""
        return analysis.scaled_Pr(self)

    @property
    def Pzeta(self):
        """Unscaled spectrum of the curvature perturbation on uniform-density hypersurfaces."""
        return analysis.Pzeta(self)

    @property
    def scaled_Pzeta(self):
        """k^3/(2 pi^2) P_zeta for all rows and k."""
        return analysis.scaled_Pzeta(self)

    def getfoystart(self):
        pass

    def callingparams(self):
        try:
            self.k
        except (NameError, AttributeError):
            self.k = None
        return {"tstart": self.tstart,
                "ainit": self.ainit,
                "potential_func": self.potentials.__name__,
                "tend": self.tend,
                "tstep_wanted": self.tstep_wanted,
                "solver": self.solver,
                "classname": self.__class__.__name__,
                "datetime": datetime.datetime.now().strftime("%Y%m%d%H%M%S"),
                "nfields": self.nfields}


    def gethf5paramsdict(self):
        """cosmomodels.py:1213-1227, plus ``cq`` (the reference does not store it and falls back
        to the constructor's default when a file is wrapped)."""
        params = super(MultiStageDriver, self).gethf5paramsdict()
        params.update({"ainit": hdf5lite.Float64Col(),
                       "potential_func": hdf5lite.StringCol(255),
                       "cq": hdf5lite.Float64Col()})
        return params


class FOCanonicalTwoStage(MultiStageDriver):
    """Background run -> initial conditions -> first-order run (cosmomodels.py:1229-1550)."""

    def __init__(self, bgystart=None, tstart=0.0, tstartindex=None, tend=83.0, tstep_wanted=0.01,
                 k=None, ainit=None, solver="rkdriver_tsix", bgclass=None, foclass=None,
                 potential_func=None, pot_params=None, simtstart=0, nfields=1, **kwargs):
        if bgystart is None:
            self.bgystart = np.array([18.0 / np.sqrt(nfields), -0.1 / np.sqrt(nfields)] * nfields
                                     + [0.0])
        else:
            self.bgystart = bgystart
        self.ystart = np.append(self.bgystart, [0.0, 0.0] * nfields ** 2)
        self.tstartindex = tstartindex if tstartindex else np.array([0])
        super(FOCanonicalTwoStage, self).__init__(
            ystart=self.ystart, tstart=tstart, tstartindex=self.tstartindex, tend=tend,
            tstep_wanted=tstep_wanted, solver=solver, potential_func=potential_func,
            pot_params=pot_params, nfields=nfields, **kwargs)
        self.setfieldindices()
        self.ainit = 1 if ainit is None else ainit
        self.k = 10 ** (np.arange(7.0) - 62) if k is None else k
        self.simtstart = simtstart
        self.bgclass = CanonicalBackground if bgclass is None else bgclass
        self.foclass = CanonicalFirstOrder if foclass is None else foclass
        self.bgmodel = self.firstordermodel = None

    def setfieldindices(self):
        nf = self.nfields
        self.H_ix = nf * 2
        self.bg_ix = slice(0, nf * 2 + 1)
        self.phis_ix = slice(0, nf * 2, 2)
        self.phidots_ix = slice(1, nf * 2, 2)
        self.pert_ix = slice(nf * 2 + 1, None)
        self.dps_ix = slice(nf * 2 + 1, None, 2)
        self.dpdots_ix = slice(nf * 2 + 2, None, 2)

    def _choose_ainit(self):
        Hend = self.bgmodel.yresult[self.fotendindex, self.H_ix]
        self.a_end = self.finda_end(Hend)
        return self.a_end * np.exp(-self.bgmodel.tresult[self.fotendindex])

    def setfoics(self):
        """Initial conditions of the first-order run from the finished background run.
        (cosmomodels.py:1307-1341)"""
        self.ainit = self._choose_ainit()
        if not hasattr(self, "bgepsilon"):
            self.bgepsilon = self.bgmodel.getepsilon()
        self.etainit = -1 / (self.ainit * self.bgmodel.yresult[0, self.H_ix] * (1 - self.bgepsilon[0]))
        kcrossings = self.findallkcrossings(self.bgmodel.tresult[:self.fotendindex],
                                            self.bgmodel.yresult[:self.fotendindex, self.H_ix])
        kcrossefolds = kcrossings[:, 1]
        if np.any(kcrossefolds == 0):
            raise ModelError("Some k modes crossed horizon before simulation began and cannot "
                             "be initialized!")
        self.fotstart, self.fotstartindex = kcrossefolds, kcrossings[:, 0].astype(int)
        self.foystart = self.getfoystart()
        return

    def runbg(self):
        """Run the background model and find the end of inflation (cosmomodels.py:1343-1374)."""
        if len(self.ystart.shape) == 1:
            ys = self.ystart[self.bg_ix]
        elif len(self.ystart.shape) == 2:
            ys = self.ystart[self.bg_ix, 0]
        self.bgmodel = self.bgclass(ystart=ys, tstart=self.tstart, tstartindex=np.array([0]),
                                    tend=self.tend, tstep_wanted=self.tstep_wanted,
                                    solver=self.solver, potential_func=self.potential_func,
                                    pot_params=self.pot_params, nfields=self.nfields)
        self._log.info("Running background model...")
        try:
            self.bgmodel.run(saveresults=False)
        except ModelError:
            self._log.exception("Error in background run, aborting!")
        self.fotend, self.fotendindex = self.bgmodel.findinflend()
        self._log.info("Background run complete, inflation ended " + str(self.fotend)
                       + " efoldings after start.")
        return

    def runfo(self):
        """Run the first order model (cosmomodels.py:1376-1404)."""
        self.firstordermodel = self.foclass(ystart=self.foystart, tstart=self.fotstart,
                                            simtstart=self.simtstart,
                                            tstartindex=self.fotstartindex, tend=self.fotend,
                                            tstep_wanted=self.tstep_wanted, solver=self.solver,
                                            k=self.k, ainit=self.ainit,
                                            potential_func=self.potential_func,
                                            pot_params=self.pot_params, nfields=self.nfields)
        if getattr(self, "row_stride", 1) != 1:
            self.firstordermodel.row_stride = self.row_stride
        self._log.info("Beginning first order run...")
        try:
            self.firstordermodel.run(saveresults=False)
        except ModelError as er:
            raise ModelError("Error in first order run, aborting! Message: " + str(er))
        self.tresult, self.yresult = self.firstordermodel.tresult, self.firstordermodel.yresult
        return

    def run(self, saveresults=True):
        """Background run, initial conditions, first-order run (cosmomodels.py:1406-1441)."""
        self.runbg()
        self.setfoics()
        self.runfo()
        self.lastparams = self.callingparams()
        if saveresults:
            try:
                self._log.info("Results saved in " + self.saveallresults())
            except IOError:
                self._log.exception("Error trying to save results! Results NOT saved.")
        return

    # ---- Bunch-Davies initial conditions --------------------------------------------------
    def _ic_inputs(self, ts, tsix):
        if ts is None or tsix is None:
            ts, tsix = self.fotstart, self.fotstartindex
        bgy = self.bgmodel.yresult
        if len(bgy.shape) > 2:
            bgy = bgy[..., 0]
        nvar = 2 * self.nfields ** 2 + self.nfields * 2 + 1
        foystart = np.zeros((nvar, len(self.k)), dtype=np.complex128)
        foystart[self.bg_ix] = bgy[tsix, :].transpose()
        astar = self.ainit * np.exp(ts)
        Hstar = bgy[tsix, self.H_ix]
        return ts, tsix, bgy, foystart, astar, Hstar

    def _phase(self, tsix, bgy, astar, Hstar):
        """exp(-i k (eta_* - eta_init)) with eta = -1/(a H (1 - epsilon))."""
        etastar = -1 / (astar * Hstar * (1 - self.bgepsilon[tsix]))
        try:
            etadiff = etastar - self.etainit
        except AttributeError:
            etadiff = etastar + 1 / (self.ainit * bgy[0, self.H_ix] * (1 - self.bgepsilon[0]))
        return np.exp(-(self.k * etadiff) * 1j)

    def _fill_diagonal(self, foystart, dp, dpdot, skip=None):
        nf = self.nfields
        nb = 2 * nf + 1
        for i in range(nf):
            if skip is not None and i == skip:
                continue
            foystart[nb + 2 * (i * nf + i)] = dp
            foystart[nb + 2 * (i * nf + i) + 1] = dpdot

    def getfoystart(self, ts=None, tsix=None):
        """Start vector of every mode: background rows at its start row, diagonal mode-matrix
        entries in the Bunch-Davies state, off-diagonals zero (cosmomodels.py:1443-1492)."""
        ts, tsix, bgy, foystart, astar, Hstar = self._ic_inputs(ts, tsix)
        arootk = 1 / (astar * (np.sqrt(2 * self.k)))
        osc = self._phase(tsix, bgy, astar, Hstar)
        self._fill_diagonal(foystart, arootk * osc,
                            -arootk * osc * (1 + (self.k / (astar * Hstar)) * 1j))
        return foystart

    def getdeltaphi(self):
        return self.deltaphi

    def deltaphi(self, recompute=False):
        """delta phi mode matrix rows for all rows and k (cosmomodels.py:1497-1520)."""
        if not hasattr(self, "_deltaphi") or recompute:
            self._deltaphi = self.yresult[:, self.dps_ix, :]
        return self._deltaphi

    @property
    def phis(self):
        return self.yresult[:, self.phis_ix]

    @property
    def phidots(self):
        return self.yresult[:, self.phidots_ix]

    @property
    def H(self):
        return self.yresult[:, self.H_ix]

    @property
    def dpmodes(self):
        return self.yresult[:, self.dps_ix]

    @property
    def dpdotmodes(self):
        return self.yresult[:, self.dpdots_ix]

    @property
    def a(self):
        return self.ainit * np.exp(self.tresult)


class FixedainitTwoStage(FOCanonicalTwoStage):
    """First order driver with ainit fixed by the caller (cosmomodels.py:2017-2073)."""

    def __init__(self, *args, **kwargs):
        super(FixedainitTwoStage, self).__init__(*args, **kwargs)
        # default: ainit of the standard msqphisq run (cosmomodels.py:2043)
        self.fixedainit = kwargs["ainit"] if "ainit" in kwargs else 7.837219134630212e-65

    def _choose_ainit(self):
        return self.fixedainit


class FONoPhase(FOCanonicalTwoStage):
    """Initial conditions without the exp(-i k eta) phase (cosmomodels.py:2076-2132)."""

    def getfoystart(self, ts=None, tsix=None):
        ts, tsix, bgy, foystart, astar, Hstar = self._ic_inputs(ts, tsix)
        arootk = 1 / (astar * (np.sqrt(2 * self.k)))
        self._fill_diagonal(foystart, arootk, -arootk * (1 + (self.k / (astar * Hstar)) * 1j))
        return foystart


class FOSuppressOneField(FOCanonicalTwoStage):
    """Standard initial conditions with one field's mode switched off (cosmomodels.py:2135-2197)."""

    def __init__(self, suppress_ix=0, *args, **kwargs):
        self.suppress_ix = suppress_ix
        super(FOSuppressOneField, self).__init__(*args, **kwargs)

    def getfoystart(self, ts=None, tsix=None):
        ts, tsix, bgy, foystart, astar, Hstar = self._ic_inputs(ts, tsix)
        arootk = 1 / (astar * (np.sqrt(2 * self.k)))
        osc = self._phase(tsix, bgy, astar, Hstar)
        self._fill_diagonal(foystart, arootk * osc,
                            -arootk * osc * (1 + (self.k / (astar * Hstar)) * 1j),
                            skip=self.suppress_ix)
        return foystart


class SOCanonicalThreeStage(MultiStageDriver):
    """Runs third stage calculation (typically second order perturbations) using a two stage
    model instance which carries a ``source`` array (cosmomodels.py:1680-1883)."""

    def __init__(self, second_stage, soclass=None, ystart=None, **soclassargs):
        self._adopt(second_stage)
        self.nfields = second_stage.nfields
        if ystart is None:
            ystart = np.zeros((4, len(self.k)))
        super(SOCanonicalThreeStage, self).__init__(**self._driver_kwargs(
            ystart, self.fotstartindex, "rkdriver_tsix", soclassargs, nfields=self.nfields))
        self._finish_init(soclass)

    def _adopt(self, second_stage):
        if not isinstance(second_stage, FOCanonicalTwoStage):
            raise ModelError("Need to provide a FOCanonicalTwoStage instance to get first order results from!")
        self.second_stage = second_stage
        self.k = np.copy(second_stage.k)
        self.simtstart = second_stage.tresult[0]
        self.fotstart = np.copy(second_stage.fotstart)
        self.fotstartindex = np.copy(second_stage.fotstartindex)
        self.ainit = second_stage.ainit
        self.potentials = second_stage.potentials
        self.potential_func = second_stage.potential_func
        self.pot_params = second_stage.pot_params

    def _driver_kwargs(self, ystart, start_rows, solver, soclassargs, **more):
        # start rows on the second-order time grid, whose step is twice the first-order one
        fotstep = self.second_stage.tstep_wanted
        sotstep = fotstep * 2
        sotstartindex = np.around(start_rows * (fotstep / sotstep) + sotstep / 2).astype(int)
        kwargs = dict(ystart=ystart, tstart=self.second_stage.tresult[0], tstartindex=sotstartindex,
                      simtstart=self.simtstart, tend=self.second_stage.tresult[-1],
                      tstep_wanted=sotstep, solver=solver,
                      potential_func=self.second_stage.potential_func,
                      pot_params=self.second_stage.pot_params, **more)
        if soclassargs is not None:
            kwargs.update(soclassargs)
        return kwargs

    def _finish_init(self, soclass):
        self.soclass = CanonicalSecondOrder if soclass is None else soclass
        self.somodel = None
        source = getattr(self.second_stage, "source", None)
        if source is None:
            raise ModelError("First order model does not have a source term!")
        self.source = source
        self.setfieldindices()

    def setfieldindices(self):
        nf = self.nfields
        # indices into self.second_stage.yresult
        self.H_ix = nf * 2
        self.bg_ix = slice(0, nf * 2 + 1)
        self.phis_ix = slice(0, nf * 2, 2)
        self.phidots_ix = slice(1, nf * 2, 2)
        self.pert_ix = slice(nf * 2 + 1, None)
        self.dps_ix = slice(nf * 2 + 1, None, 2)
        self.dpdots_ix = slice(nf * 2 + 2, None, 2)
        # indices of second order quantities, into self.yresult
        self.dp2s_ix = slice(0, None, 2)
        self.dp2dots_ix = slice(1, None, 2)

    def setup_soclass(self):
        """Initialize the second order class that will be used to run simulation."""
        self.somodel = self.soclass(ystart=self.ystart, tstart=self.fotstart,
                                    tstartindex=self.tstartindex, simtstart=self.simtstart,
                                    tend=self.tend, tstep_wanted=self.tstep_wanted,
                                    solver=self.solver, k=self.k, ainit=self.ainit,
                                    potential_func=self.potential_func, pot_params=self.pot_params,
                                    cq=self.cq, nfields=self.nfields)
        self.somodel.source = self.source
        self.somodel.second_stage = self.second_stage

    def runso(self):
        """Run second order model."""
        self.setup_soclass()
        self._log.info("Beginning second order run...")
        try:
            self.somodel.run(saveresults=False)
        except ModelError:
            self._log.exception("Error in second order run, aborting!")
            raise
        self.tresult, self.yresult = self.somodel.tresult, self.somodel.yresult

    def run(self, saveresults=True):
        """Run simulation and save results."""
        self.runso()
        self.lastparams = self.callingparams()
        if saveresults:
            try:
                self._log.info("Results saved in " + self.saveallresults())
            except IOError:
                self._log.exception("Error trying to save results! Results NOT saved.")

    @property
    def deltaphi(self):
        """delta phi = dp1 + dp2/2 for all timesteps and k modes.  The reference's expression
        (cosmomodels.py:1843-1846) reads rows 3 and 5 of the first-order result, i.e. the layout
        of real/imaginary rows it had before the mode functions became complex; with the complex
        layout the first-order mode is row ``dps_ix`` itself, sampled on the second-order grid."""
        if not hasattr(self, "_deltaphi"):
            ratio = int(np.around(self.tstep_wanted / self.second_stage.tstep_wanted))
            dp1 = self.second_stage.yresult[::ratio, self.dps_ix][:self.yresult.shape[0], 0]
            dp2 = self.yresult[:, 0, :] + self.yresult[:, 2, :] * 1j
            self._deltaphi = dp1 + 0.5 * dp2
        return self._deltaphi

    # results of the stage below
    phis = property(lambda self: self.second_stage.yresult[:, self.phis_ix])
    phidots = property(lambda self: self.second_stage.yresult[:, self.phidots_ix])
    H = property(lambda self: self.second_stage.yresult[:, self.H_ix])
    dpmodes = property(lambda self: self.second_stage.yresult[:, self.dps_ix])
    dpdotmodes = property(lambda self: self.second_stage.yresult[:, self.dpdots_ix])
    # second order quantities
    dp2modes = property(lambda self: self.yresult[:, self.dp2s_ix])
    dp2dotmodes = property(lambda self: self.yresult[:, self.dp2dots_ix])
    a = property(lambda self: self.ainit * np.exp(self.tresult))


class SOHorizonStart(SOCanonicalThreeStage):
    """Second order calculation that starts at horizon crossing (k = aH) instead of the first
    order start (cosmomodels.py:1885-1946).  The reference asks for the solver "rkdriver_new",
    which its rk4 module does not have (ValueError "Solver not recognized!", as here); pass
    ``solver="rkdriver_tsix"`` to run it."""

    def __init__(self, second_stage, soclass=None, ystart=None, **soclassargs):
        self._adopt(second_stage)
        self.nfields = second_stage.nfields
        if ystart is None:
            ystart = np.zeros((4, len(self.k)))
        rows = second_stage._crossing_rows(second_stage.k, second_stage.bgmodel.tresult,
                                           second_stage.bgmodel.yresult[:, 2], factor=1)
        super(SOCanonicalThreeStage, self).__init__(**self._driver_kwargs(
            ystart, np.asarray(rows), "rkdriver_new", soclassargs))
        self._finish_init(soclass)


class _ResultsFromHDF5(dict):
    """The datasets of a results file in the reference's HDF5 layout (cosmomodels.py:279-349),
    under the names the ``.npz`` archives use, so that both kinds of file are wrapped by the same
    code.  Read as in the reference's wrapper (:1584-1622): ``results.parameters[0]`` for the
    calling parameters, ``results.pot_params`` row by row, ``bgresults`` for the background."""

    files = property(lambda self: list(self))

    def __init__(self, filename):
        dict.__init__(self)
        text = lambda v: v.decode("utf-8", "replace") if isinstance(v, bytes) else str(v)
        with hdf5lite.open_file(filename, "r") as rf:
            try:
                results = rf.root.results
                params = results.parameters
                for name in ("yresult", "tresult", "k"):
                    self[name] = getattr(results, name).read()
            except hdf5lite.NoSuchNodeError:
                raise ModelError("File does not contain correct model data structure!")
            for name in ("fotstart", "foystart", "ystart", "bgystart", "fotstartindex", "sourceterm"):
                if name in results:
                    self[name] = getattr(results, name).read()
            first = params[0]
            self["classname"] = np.asarray(text(first["classname"]))
            for col in params.colnames:
                if col in ("potential_func",):
                    self["param_" + col] = np.asarray(text(first[col]))
                elif col in ("nfields", "tend", "tstep_wanted", "simtstart", "cq"):
                    self["param_" + col] = np.asarray(first[col])
                elif col == "ainit":
                    self["ainit"] = np.asarray(first[col])
            if "pot_params" in results:
                rows = results.pot_params.read()
                self["pot_param_keys"] = np.array([text(r["name"]) for r in rows])
                self["pot_param_vals"] = np.array([float(r["value"]) for r in rows])
            if "bgresults" in rf.root:
                self["bg_tresult"] = rf.root.bgresults.tresult.read()
                self["bg_yresult"] = rf.root.bgresults.yresult.read()


def make_wrapper_model(modelfile, *args, **kwargs):
    """Return a model instance holding the results stored in ``modelfile`` (the counterpart of
    ``saveallresults``; the reference's version reads its HDF5 files, cosmomodels.py:1552-1677,
    this one reads those too -- through :mod:`pyflation_b200.hdf5lite` -- and the ``.npz`` archives
    written here).  The arrays are read into memory and the file is closed."""
    if not os.path.isfile(modelfile):
        raise IOError("File does not exist!")
    if hdf5lite.is_hdf5_file(modelfile):
        z = _ResultsFromHDF5(modelfile)
    else:
        z = np.load(modelfile, allow_pickle=False)
    classname = str(z["classname"]) if "classname" in z.files else "FOCanonicalTwoStage"
    cls = globals().get(classname, FOCanonicalTwoStage)
    if issubclass(cls, SOCanonicalThreeStage):
        # as in the reference, whose wrapper cannot build a second-order driver without its
        # first-order stage: wrap the first-order file and run the driver again
        raise ModelError("Results of %s cannot be wrapped: it needs its first order model." % classname)
    pot_params = {}
    if "pot_param_keys" in z.files:
        for key, val in zip(z["pot_param_keys"], z["pot_param_vals"]):
            pot_params[str(key)] = int(val) if str(key) == "nfields" else float(val)
    nfields = int(z["param_nfields"]) if "param_nfields" in z.files else 1
    init = dict(potential_func=str(z["param_potential_func"]) if "param_potential_func" in z.files else None,
                pot_params=pot_params, nfields=nfields)
    if issubclass(cls, MultiStageDriver):
        init["k"] = z["k"] if "k" in z.files else None
        if "bgystart" in z.files:
            init["bgystart"] = z["bgystart"]
        if "param_cq" in z.files:
            init["cq"] = float(z["param_cq"])
    elif "ystart" in z.files:
        init["ystart"] = z["ystart"]
        if issubclass(cls, CanonicalFirstOrder) and "k" in z.files:
            init["k"] = z["k"]
    for name in ("tend", "tstep_wanted"):
        if "param_" + name in z.files:
            init[name] = float(z["param_" + name])
    init.update(kwargs)
    m = cls(*args, **init)
    m.yresult, m.tresult = z["yresult"], z["tresult"]
    for name in ("foystart", "fotstart", "fotstartindex", "ainit"):
        if name in z.files:
            setattr(m, name, z[name])
    if "bg_yresult" in z.files and issubclass(cls, MultiStageDriver):
        bg = m.bgclass(ystart=np.array(z["bg_yresult"][0, :, 0]), tend=m.tend, tstep_wanted=m.tstep_wanted,
                       potential_func=m.potential_func, pot_params=m.pot_params, nfields=m.nfields)
        bg.yresult, bg.tresult = z["bg_yresult"], z["bg_tresult"]
        m.bgmodel = bg
        m.bgepsilon = bg.getepsilon()                       # cosmomodels.py:1661
    # the source term of the second-order stage, when the file has one (:1607-1612)
    m.source = z["sourceterm"] if "sourceterm" in z.files else None
    return m
