"""rk4 - drop-in for ``pyflation.rk4`` (reference rk4.py:27-144) backed by the CUDA kernels.

Same names and call signatures as the reference: :func:`rkdriver_tsix`, :func:`rk4stepks`,
:class:`SimRunError`.  A Python callable cannot cross the C ABI, so instead of calling
``derivs`` the driver looks at the model it is bound to (``derivs.__self__``): a
``CanonicalBackground`` goes to the background kernel, a ``CanonicalFirstOrder`` to the fused
first-order kernel, a ``CanonicalSecondOrder`` / ``CanonicalRampedSecondOrder`` to the
second-order kernel.  Any other right-hand side raises ``NotImplementedError``: there is no
CPU fallback.
"""
import logging

import numpy as np

from .configuration import _debug
from . import engine, _lib

rk_log = logging.getLogger(__name__)

# bytes the last first-order rkdriver_tsix call moved over PCIe (host->device, device->host)
last_transfer = {"h2d_bytes": 0, "d2h_bytes": 0}

# P_R(t, k) of recent first-order results, reduced on the device while the trajectory was still
# there (engine.first_order_host, K5 per window) and KEPT on the device (T*nk*8 bytes, at most
# PR_CACHE_BYTES each, two entries); analysis.Pr looks results up here, so ``model.Pr`` neither
# uploads the host trajectory again nor pays for rows nobody asked for: only the requested
# (rows, k) block crosses PCIe.  Keyed by id(yresult); an entry is used only while a fingerprint
# of the array (sampled entries of the last rows) still matches.
_PR_CACHE = {}
_PR_CACHE_MAX = 2
PR_CACHE_BYTES = 1 << 30            # cache P_R when T*nk*8 is at most this


def _fingerprint(y):
    step = max(1, y.shape[2] // 16)
    return y[-2:, :, ::step].copy()


def cached_pr(yresult):
    """P_R rows as a CUDA tensor (T, nk) float64 cached for this very array, or None."""
    e = _PR_CACHE.get(id(yresult))
    if e is None or e[0] != yresult.shape:
        return None
    if not np.array_equal(e[1], _fingerprint(yresult), equal_nan=True):
        _PR_CACHE.pop(id(yresult), None)
        return None
    return e[2]


def _remember_pr(yresult, pr):
    _PR_CACHE[id(yresult)] = (yresult.shape, _fingerprint(yresult), pr)
    while len(_PR_CACHE) > _PR_CACHE_MAX:
        _PR_CACHE.pop(next(iter(_PR_CACHE)))


class SimRunError(Exception):
    """Generic error for model simulating run. (rk4.py:141)"""
    pass


def _number_of_steps(simtstart, tend, h):
    # rk4.py:73 -- np.around rounds halves to even, so 0.5 -> 0 and 1.5 -> 2
    return int(np.around((tend - simtstart) / h) + 1)


def _bound_model(derivs):
    from . import cosmomodels as c
    owner = getattr(derivs, "__self__", None)
    if isinstance(owner, c.CanonicalFirstOrder) and getattr(derivs, "__func__", None) is \
            c.CanonicalFirstOrder.derivs:
        return "fo", owner
    if isinstance(owner, c.CanonicalBackground) and getattr(derivs, "__func__", None) is \
            c.CanonicalBackground.derivs:
        return "bg", owner
    if isinstance(owner, (c.CanonicalSecondOrder, c.CanonicalRampedSecondOrder)) and \
            getattr(derivs, "__func__", None) is c._SecondOrderModel.derivs:
        return "so", owner
    raise NotImplementedError(
        "pyflation_b200.rk4 integrates CanonicalBackground.derivs, CanonicalFirstOrder.derivs and "
        "Canonical(Ramped)SecondOrder.derivs on the GPU; an arbitrary Python `derivs` (%r) cannot be run and there is no CPU "
        "fallback." % (derivs,))


def _check_common(tsix, simtstart, tend, h):
    if h is None:
        raise SimRunError("Need to specify h.")
    tsix = np.atleast_1d(np.asarray(tsix)).astype(np.int64)
    T = _number_of_steps(simtstart, tend, h)
    if np.any(tsix > T):
        raise SimRunError("Start times outside range of steps.")
    if np.any(tsix >= T) or np.any(tsix < 0):
        # the reference lets tsix == T through its check and then fails with an IndexError
        # at rk4.py:107; both are reported as SimRunError here
        raise SimRunError("Start times outside range of steps.")
    return tsix, T


def _xarr(simtstart, T, h):
    return simtstart + np.arange(T) * h                     # rk4.py:78-87,134


def rkdriver_tsix(ystart, simtstart, tsix, tend, h, derivs, yresult_out=None, row_stride=1):
    """Driver function for classical Runge Kutta 4th Order method.
    Uses indexes of starting time values instead of actual times.
    Indexes are number of steps of size h away from initial time simtstart.

    Returns ``(xarr (T,), yarr (T, nvar, nk))`` laid out exactly as the reference's arrays
    (rk4.py:95-100): rows before a mode's start index are NaN, the start row is ``ystart``.

    ``yresult_out`` (extension, optional): a C-contiguous complex128 array or CPU tensor of
    shape (T, nvar, nk) to receive the first-order trajectory -- lets repeated runs reuse one
    page-locked host buffer instead of allocating tens of GB per call.

    ``row_stride`` (extension, optional): keep only rows 0, s, 2s, ... of a first-order run
    (``xarr`` is strided the same way); ``row_stride="last"`` keeps rows 0 and T-1.  Batches
    whose full trajectory does not fit anywhere are run this way.
    """
    kind, model = _bound_model(derivs)
    tsix, T = _check_common(tsix, simtstart, tend, h)
    ystart = np.asarray(ystart)
    if ystart.ndim == 1:
        ystart = ystart[..., np.newaxis]                    # rk4.py:93-94
    if _debug:
        rk_log.debug("rkdriver_tsix: %s run, T=%d, columns=%d", kind, T, ystart.shape[1])
    try:
        if kind == "bg":
            yarr = _run_background(model, ystart, simtstart, tsix, T, h)
        elif kind == "so":
            yarr = _run_second_order(model, ystart, simtstart, tsix, T, h)
        else:
            if row_stride == "last":
                row_stride = max(1, T - 1)
            yarr = _run_first_order(model, ystart, simtstart, tsix, T, h, yresult_out, int(row_stride))
    except _lib.ArgumentError as err:
        raise SimRunError(str(err))
    rk_log.info("Execution of Runge-Kutta method has finished.")
    xarr = _xarr(simtstart, T, h)
    if kind == "fo" and row_stride != 1:
        xarr = xarr[::int(row_stride)]
    return xarr, yarr


def _run_background(model, ystart, simtstart, tsix, T, h):
    nf = model.nfields
    if ystart.shape[1] != 1 or tsix.shape[0] != 1:
        raise NotImplementedError("background runs integrate a single column")
    if ystart.shape[0] != 2 * nf + 1:
        raise SimRunError("background ystart must have 2*nfields+1 rows")
    if np.iscomplexobj(ystart):
        raise NotImplementedError("background runs are real valued")
    pm = engine.make_model(model.potential_func, model.pot_params, nf)
    bg, _, stagey = engine.background_tables(pm, ystart[:, 0], T, h, simtstart, int(tsix[0]),
                                             records=False, return_stage=True)
    bg_host = bg.cpu().numpy()
    # the first-order run that follows starts on this trajectory: keep its stage arguments
    engine.remember_background(pm, h, simtstart, int(tsix[0]), bg_host, stagey, bg.device)
    return bg_host.reshape(T, 2 * nf + 1, 1)


def _run_first_order(model, ystart, simtstart, tsix, T, h, yresult_out=None, row_stride=1):
    import torch
    nf = model.nfields
    nvar = engine.nvar_of(nf)
    k = np.atleast_1d(np.asarray(model.k, dtype=np.float64))
    if ystart.shape != (nvar, k.shape[0]) or tsix.shape[0] != k.shape[0]:
        raise SimRunError("ystart must have shape (2*nfields**2+2*nfields+1, len(k))")
    n0 = int(tsix.min())
    if int(tsix[0]) != n0:
        # cosmomodels.py:641 evaluates the potential on column 0 only: in the reference every
        # column that starts before column 0 is poisoned to NaN for the whole run
        raise SimRunError("column 0 must be the earliest-starting mode (k ascending): the "
                          "potential is evaluated on column 0 (cosmomodels.py:641)")
    pm = engine.make_model(model.potential_func, model.pot_params, nf, model.ainit)
    prob = engine.FirstOrderProblem(pm, k, tsix, ystart.astype(np.complex128, copy=False),
                                    simtstart, T, h)
    if row_stride < 1:
        raise SimRunError("row_stride must be a positive integer")
    y0 = np.real(ystart[0:2 * nf + 1, 0])
    # runbg() already integrated this background: rebuild the records from its stage arguments
    # and skip K1's serial chain -- except for large single-field full-trajectory batches, where
    # the fused launch hides K1 completely and brings the lock-step store pattern
    if not (nf == 1 and row_stride == 1 and k.shape[0] >= 16384):
        hit = engine.cached_stage_args(pm, h, simtstart, n0, y0, T, prob.dev)
        if hit is not None:
            prob.use_stage_args(*hit)
    if row_stride != 1:
        rows = engine.first_order_device(prob, out_every=row_stride, y0=y0, n0=n0)
        prob.check_flags()
        return rows.cpu().numpy()
    out = None
    if yresult_out is not None:
        out = yresult_out if torch.is_tensor(yresult_out) else torch.as_tensor(yresult_out)
    pr_dev = None
    if T * k.shape[0] * 8 <= PR_CACHE_BYTES:
        pr_dev = torch.empty((T, k.shape[0]), dtype=torch.float64, device=prob.dev)
    res = engine.first_order_host(prob, out=out, y0=y0, n0=n0, pr_out=pr_dev)
    prob.check_flags()
    last_transfer.update(h2d_bytes=prob.h2d_bytes, d2h_bytes=prob.d2h_bytes)
    yarr = yresult_out if (yresult_out is not None and not torch.is_tensor(yresult_out)) else res.numpy()
    if pr_dev is not None:
        _remember_pr(yarr, pr_dev)
    return yarr


def _run_second_order(model, ystart, simtstart, tsix, T, h):
    """rkdriver_tsix for CanonicalSecondOrder / CanonicalRampedSecondOrder.derivs: the model
    carries ``second_stage`` (first-order results) and ``source`` (cosmomodels.py:1773-1775)."""
    k = np.atleast_1d(np.asarray(model.k, dtype=np.float64))
    if ystart.shape != (4, k.shape[0]) or tsix.shape[0] != k.shape[0]:
        raise SimRunError("ystart must have shape (4, len(k))")
    if np.iscomplexobj(ystart):
        raise NotImplementedError("second order runs are real valued")
    inp = model._first_order_inputs()
    fo_tsix = inp["fo_tsix"]
    if fo_tsix is not None and int(np.asarray(fo_tsix)[0]) != int(np.min(fo_tsix)):
        # the potential is evaluated on column 0 of the first-order row (cmpotentials.py, single
        # field functions take y[:, 0]): as in the first-order run, column 0 must start first
        raise SimRunError("column 0 must be the earliest-starting mode (k ascending)")
    pm = engine.make_model(model.potential_func, model.pot_params, model.nfields, model.ainit)
    ramp = model.rampargs if model.ramped else None
    prob = engine.SecondOrderProblem(pm, k, tsix, ystart, simtstart, T, h, source=model.source,
                                     ramp=ramp, tstart=model.tstart if model.ramped else None,
                                     ramp_tsix=model.tstartindex if model.ramped else None, **inp)
    yarr = engine.second_order_host(prob)
    last_transfer.update(h2d_bytes=prob.h2d_bytes, d2h_bytes=prob.d2h_bytes)
    return yarr


def rk4stepks(x, y, h, dydx, dargs, derivs):
    '''Do one step of the classical 4th order Runge Kutta method,
    starting from y at x with time step h and derivatives given by derivs.

    Runs as a two-row solve on the device; ``dydx`` is recomputed there, so it is ignored.'''
    y = np.asarray(y)
    single = y.ndim == 1
    ncol = 1 if single else y.shape[1]
    _, yarr = rkdriver_tsix(y, x, np.zeros(ncol, dtype=np.int64), x + h, h, derivs)
    out = yarr[1]
    return out[:, 0] if single else out
