"""analysis.adiabatic - curvature perturbation spectra (pyflation/analysis/adiabatic.py).

``Pr`` / ``scaled_Pr`` accept a model whose ``yresult`` is a host NumPy array (reduced with
NumPy, as post-processing) or a CUDA tensor (reduced on the device by ``pf_pr_spectrum``).
"""
import numpy as np

from . import utilities


def Pphi_matrix(modes, axis):
    """P_IJ = sum_K chi_IK conj(chi_JK) for a mode matrix whose two nfields-long axes sit at
    ``axis`` and ``axis+1``; result has the shape and dtype of ``modes``
    (adiabatic.py:45-81)."""
    modes = np.asarray(modes)
    if axis < 0:
        axis += modes.ndim - 1
    m = np.moveaxis(modes, (axis, axis + 1), (0, 1))
    P = np.einsum("ik...,jk...->ij...", m, m.conj())
    return np.moveaxis(P, (0, 1), (axis, axis + 1)).astype(modes.dtype)


def Pphi_modes(m):
    """P_IJ flattened like ``m.yresult[:, m.dps_ix]`` (adiabatic.py:17-43)."""
    mdp = utilities.getmodematrix(m.yresult, m.nfields, ix=1, ixslice=m.dps_ix)
    return utilities.flattenmodematrix(Pphi_matrix(mdp, axis=1), m.nfields, 1, 2)


def Pr_spectrum(phidot, modes, axis):
    """P_R = sum_IJ phidot_I phidot_J P_IJ / (sum_K phidot_K^2)^2, real valued
    (adiabatic.py:153-199)."""
    phidot = np.atleast_1d(phidot)
    modes = np.asarray(modes)
    norm = (np.sum(phidot ** 2, axis=axis)) ** 2
    weighted = (np.expand_dims(phidot, axis) * np.expand_dims(phidot, axis + 1)
                * Pphi_matrix(modes, axis))
    total = np.sum(weighted, axis=(axis, axis + 1))
    return np.real(total / norm).astype(float)


_DEVICE_REDUCE_BYTES = 64 << 20


def _pr_on_device(y, nfields, chunk_bytes=512 << 20):
    """P_R of a large host trajectory with the K5 kernel: rows are streamed to the device in
    chunks (the reference spends 1.9 s in NumPy on the default 2053-mode run).
    Returns None when no CUDA device is present (small post-processing then stays in NumPy)."""
    import torch
    if not torch.cuda.is_available():
        return None
    from .. import engine
    rows_per = max(1, int(chunk_bytes // max(1, y[0].nbytes)))
    out = np.empty((y.shape[0], y.shape[2]), dtype=np.float64)
    for r0 in range(0, y.shape[0], rows_per):
        r1 = min(y.shape[0], r0 + rows_per)
        block = torch.as_tensor(np.ascontiguousarray(y[r0:r1])).cuda(non_blocking=True)
        out[r0:r1] = engine.pr_spectrum_device(nfields, block).cpu().numpy()
    return out


def _to_host(block):
    """Device block -> fresh NumPy array through a page-locked staging tensor."""
    from ..parallel import to_host
    return to_host(block)


def _slices(m, tix, kix):
    if tix is None:
        tslice = slice(None)
    else:
        if tix < 0:
            tix = len(m.tresult) + tix
        tslice = slice(tix, tix + 1)
    kslice = slice(None) if kix is None else slice(kix, kix + 1)
    return tslice, kslice


def Pr(m, tix=None, kix=None):
    """Unscaled curvature power spectrum of model ``m`` for the requested rows and modes;
    shape (rows, nk) (adiabatic.py:201-258)."""
    tslice, kslice = _slices(m, tix, kix)
    if isinstance(m.yresult, np.ndarray):
        from .. import rk4
        cached = rk4.cached_pr(m.yresult)                 # reduced on the device during the run
        if cached is not None:
            return _to_host(cached[tslice, kslice])
    y = m.yresult[tslice, :, kslice]
    if not isinstance(y, np.ndarray):                     # CUDA tensor -> K5
        from .. import engine
        return engine.pr_spectrum_device(m.nfields, y).cpu().numpy()
    if y.nbytes >= _DEVICE_REDUCE_BYTES and y.dtype == np.complex128:
        out = _pr_on_device(y, m.nfields)                 # whole trajectories: reduce on the GPU
        if out is not None:
            return out
    phidot = np.real(y[:, m.phidots_ix, :]).astype(np.float64)
    modes = utilities.getmodematrix(y, m.nfields, ix=1, ixslice=m.dps_ix)
    return Pr_spectrum(phidot, modes, 1)


def scaled_Pr(m, tix=None, kix=None):
    """k^3/(2 pi^2) P_R (adiabatic.py:261-285)."""
    return utilities.kscaling(m.k, kix) * Pr(m, tix, kix)


def Pzeta_spectrum(Vphi, phidot, H, modes, modesdot, axis):
    """P_zeta = P_{delta rho} / (rho')^2 (adiabatic.py:309-352, with nonadiabatic.py:145-239,
    433-471):  delta rho_I = sum_a [H^2 phi'_a chi'_aI - H^2 phi'_a^2/2 sum_b phi'_b chi_bI
    + V_a chi_aI],  P_{delta rho} = sum_I |delta rho_I|^2,  rho' = -3 H^2 sum_a phi'_a^2."""
    Vphi, phidot, H, modes, modesdot, axis = utilities.correct_shapes(Vphi, phidot, H, modes,
                                                                      modesdot, axis)
    pa = np.expand_dims(phidot, axis + 1)
    Va = np.expand_dims(Vphi, axis + 1)
    Hm = np.expand_dims(H, axis + 1)
    inner = np.sum(pa * modes, axis=axis, keepdims=True)
    drho = np.sum(Hm ** 2 * pa * modesdot - 0.5 * Hm ** 2 * pa ** 2 * inner + Va * modes, axis=axis)
    spectrum = np.real(np.sum(drho * np.conj(drho), axis=axis)).astype(float)
    rhodot = np.sum(-3 * H ** 2 * phidot ** 2, axis=axis)
    return rhodot ** (-2.0) * spectrum


def Pzeta(m, tix=None, kix=None):
    """Unscaled spectrum of the curvature perturbation on uniform-density hypersurfaces
    (adiabatic.py:287-307)."""
    return Pzeta_spectrum(*utilities.components_from_model(m, tix, kix))


def scaled_Pzeta(m, tix=None, kix=None):
    """k^3/(2 pi^2) P_zeta (adiabatic.py:354-378)."""
    return utilities.kscaling(m.k, kix) * Pzeta(m, tix, kix)


def findns(sPr, k, kix, running=False):
    """Spectral index n_s (and running) of a scaled spectrum at ``k[kix]``
    (adiabatic.py:84-131)."""
    return utilities.spectral_index(sPr, k, kix, running)
