"""analysis.utilities - array helpers used by the adiabatic spectra
(pyflation/analysis/utilities.py:12-184)."""
import numpy as np


def kscaling(k, kix=None):
    """k^3/(2 pi^2), for all k or for the single index ``kix`` (utilities.py:12-37)."""
    k = np.atleast_1d(k)
    if kix is not None:
        k = k[kix:kix + 1]
    return k ** 3 / (2 * np.pi ** 2)


def spectral_index(y, k, kix, running=False):
    """1 + dln(y)/dln(k) at ``k[kix]`` from a linear (or, with ``running``, quadratic)
    least-squares fit in log space; with ``running`` also returns the running
    (utilities.py:39-102)."""
    if y.shape != k.shape:
        raise ValueError("y and k arrays must be same shape.")
    if np.any(y == 0):
        raise ValueError("The array y cannot contain any zero values.")
    logy, logk = np.log(y), np.log(k)
    fit = np.poly1d(np.polyfit(logk, logy, deg=2 if running else 1))
    n_s = 1 + np.polyder(fit, m=1)(logk[kix])
    if running:
        return n_s, 2 * np.polyder(fit, m=2)(logk[kix])
    return n_s


def getmodematrix(y, nfields, ix=None, ixslice=None):
    """View of ``y`` with the nfields**2-long axis ``ix`` (after slicing it with ``ixslice``)
    split into two nfields-long axes (utilities.py:104-144)."""
    if ix is None:
        ix = 1
    if ixslice is None:
        ixslice = slice(None)
    index = [slice(None)] * y.ndim
    index[ix] = ixslice
    modes = y[tuple(index)]
    shape = list(modes.shape)
    if shape[ix] != nfields ** 2:
        raise ValueError("Array does not have correct dimensions of nfields**2.")
    shape[ix:ix + 1] = [nfields, nfields]
    return modes.reshape(shape)


def flattenmodematrix(modematrix, nfields, ix1=None, ix2=None):
    """Merge the two nfields-long axes ``ix1, ix2`` back into one nfields**2-long axis
    (utilities.py:146-184)."""
    s = modematrix.shape
    if s.count(nfields) < 2:
        raise ValueError("Mode matrix does not have two nfield long dimensions.")
    try:
        if ix1 is None:
            ix1 = s.index(nfields)
        if ix2 is None:
            ix2 = s.index(nfields, ix1 + 1)
    except ValueError:
        raise ValueError("Cannot determine correct indices for nfield long dimensions!")
    shape = list(s)
    shape.pop(ix2)
    shape[ix1] = nfields ** 2
    return modematrix.reshape(shape)


def makespectrum(modes_I, axis):
    """sum_I |modes_I|^2 along ``axis`` (utilities.py:255-277)."""
    modes_I = np.atleast_1d(modes_I)
    return np.real(np.sum(modes_I * modes_I.conj(), axis=axis)).astype(float)


def correct_shapes(Vphi, phidot, H, modes, modesdot, axis):
    """Bring the inputs of the energy-density spectra to broadcastable shapes
    (utilities.py:279-335): Vphi and phidot (..., nfields, ...), H with a length-1 field axis,
    modes/modesdot with two adjacent nfields-long axes starting at ``axis``."""
    mshape = modes.shape
    if mshape[axis + 1] != mshape[axis]:
        raise ValueError("The mode matrix dimensions are not together.")
    if mshape != modesdot.shape:
        raise ValueError("Mode matrix and its derivative should be the same shape.")
    Vphi, phidot, H = np.atleast_1d(Vphi, phidot, H)
    if Vphi.ndim < phidot.ndim:
        Vphi = np.expand_dims(Vphi, axis=-1)
    if (len(mshape) - 1) != Vphi.ndim != phidot.ndim:
        raise ValueError("Vphi, phidot and modes arrays must have correct shape.")
    if H.ndim < phidot.ndim:
        H = np.expand_dims(H, axis)
    return Vphi, phidot, H, modes, modesdot, axis


def components_from_model(m, tix=None, kix=None):
    """(Vphi, phidot, H, modes, modesdot, axis) of model ``m`` for the requested rows and
    modes (utilities.py:205-253); the potential gradient is evaluated with the host-side
    ``cmpotentials`` function, row by row, as in the reference."""
    if tix is None:
        tslice = slice(None)
    else:
        if tix < 0:
            tix = len(m.tresult) + tix
        tslice = slice(tix, tix + 1)
    kslice = slice(None) if kix is None else slice(kix, kix + 1)
    phidot = m.yresult[tslice, m.phidots_ix, kslice]
    H = m.yresult[tslice, m.H_ix, kslice]
    Vphi = np.array([m.potentials(row, m.pot_params)[1] for row in m.yresult[tslice, m.bg_ix, kslice]])
    modes = getmodematrix(m.yresult[tslice, m.dps_ix, kslice], m.nfields)
    modesdot = getmodematrix(m.yresult[tslice, m.dpdots_ix, kslice], m.nfields)
    return correct_shapes(Vphi, phidot, H, modes, modesdot, 1)
