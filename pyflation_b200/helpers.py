"""helpers - drop-in for pyflation/helpers.py: the k-grid helpers that callers of the first-order
path use (run_config.py:142-158, bin/pyflation_firstorder.py:206) and the small utilities the
reference's scripts import from the same module (paths, logging, unit conversion, labels)."""
import numpy as np


def getkend(kinit, deltak, numsoks):
    """End of the first-order k range so that the second-order range holds ``numsoks`` modes
    (helpers.py:46-68): ``2*(numsoks*deltak + kinit)``."""
    return 2 * ((numsoks) * deltak + kinit)


def seq(min=0.0, max=None, inc=1.0, type=float, return_type='NumPyArray'):
    """Numbers from ``min`` to (and including) ``max`` in steps of ``inc`` (helpers.py:229-247)."""
    if max is None:
        max = min
        min = 0.0
        inc = 1.0
    r = np.arange(min, max + inc / 2.0, inc, type)
    if return_type == 'NumPyArray' or return_type == np.ndarray:
        return r
    elif return_type == 'list':
        return r.tolist()
    elif return_type == 'tuple':
        return tuple(r.tolist())
    return


def nanfillstart(a, l):
    """Pad the start of ``a`` with NaNs up to length ``l`` (helpers.py:18-44)."""
    if len(a) >= l:
        return a
    shape = list(a.shape)
    shape[0] = l - shape[0]
    return np.concatenate((np.full(shape, np.nan), a))


def find_nearest_ix(array, value):
    """Index of the entry of ``array`` closest to ``value``."""
    return int(np.abs(np.asarray(array) - value).argmin())


def find_nearest(array, value):
    """The entry of ``array`` closest to ``value`` (helpers.py:249-254)."""
    array = np.asarray(array)
    return array[find_nearest_ix(array, value)]


# ---- small utilities of the reference's helpers module that its scripts and notebooks import ----
MPC_INVERSE_IN_MPL = 2.625e-57          # helpers.py:92-98: the reference's conversion constants
MPL_IN_MPC_INVERSE = 3.8095e+56


def invmpc2mpl(x=1):
    """Mpc^-1 -> reduced Planck masses."""
    return MPC_INVERSE_IN_MPL * x


def mpl2invmpc(x=1):
    """Reduced Planck masses -> Mpc^-1."""
    return MPL_IN_MPC_INVERSE * x


def eto10(number):
    """'1e-05' -> '1\\times 10^{-5}' for LaTeX labels (helpers.py:70-73)."""
    import re
    return re.sub(r'e(\S)0?(\d+)', r'\\times 10^{\1\2}', str(number))


def klegend(ks, mpc=False):
    """Legend strings for a list of k values, optionally with the Mpc^-1 equivalent."""
    labels = []
    for k in ks:
        text = r"$k=" + eto10(k) + r"M_{\mathrm{PL}}"
        if mpc:
            text += r" = " + eto10(mpl2invmpc(k)) + r" M\mathrm{pc}"
        labels.append(text + "$")
    return labels


def ispower2(n):
    """log2(n) when n is a power of two, else 0 (so n == 1 also gives 0; helpers.py:100-108)."""
    n = int(n)
    return n.bit_length() - 1 if n > 0 and n & (n - 1) == 0 else 0


def removedups(l):
    """Array of the values of ``l`` without duplicates, first occurrences kept, order intact."""
    vals = np.asarray(list(l))
    if vals.size == 0:
        return np.array([])
    _, first = np.unique(vals, return_index=True)
    return vals[np.sort(first)].astype(float)


def getintfunc(x):
    """(integration function, keyword arguments) for samples at ``x``: Romberg when len(x) is a
    power of two plus one, Simpson otherwise (helpers.py:131-158)."""
    from scipy import integrate
    if len(x) == 0:
        raise ValueError("Cannot integrate length 0 array!")
    if ispower2(len(x) - 1):
        return integrate.romb, {"dx": x[1] - x[0]}
    simpson = getattr(integrate, "simpson", None) or integrate.simps
    return simpson, {"x": x}


def cartesian_product(lists, previous_elements=[]):
    """Generator over the cartesian product of ``lists`` (each item a list)."""
    import itertools
    for combo in itertools.product(*lists):
        yield list(previous_elements) + list(combo)


def cartesian(arrays, out=None):
    """All combinations of the entries of the 1-D ``arrays`` as rows of a 2-D array, the first
    array varying slowest (helpers.py:171-215)."""
    arrays = [np.asarray(x) for x in arrays]
    grid = np.stack(np.meshgrid(*arrays, indexing="ij"), axis=-1).reshape(-1, len(arrays))
    grid = grid.astype(arrays[0].dtype)
    if out is None:
        return grid
    out[...] = grid
    return out


def ensurepath(path):
    """Create the directory that will hold the file ``path`` if it is not there yet."""
    import os
    folder = os.path.dirname(os.path.abspath(path))
    if not os.path.isdir(folder):
        try:
            os.makedirs(folder)
        except OSError:
            raise OSError("Error creating results directory!")


def startlogging(log, logfile, loglevel=None, consolelevel=None):
    """Attach a rotating file handler (1 MiB x 50) and a console handler to ``log``
    (helpers.py:263-297)."""
    import logging
    from logging.handlers import RotatingFileHandler
    loglevel = logging.INFO if loglevel is None else loglevel
    consolelevel = consolelevel or loglevel
    log.setLevel(loglevel)
    fmt = logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s")
    for handler, level in ((RotatingFileHandler(filename=logfile, maxBytes=2 ** 20, backupCount=50), loglevel),
                           (logging.StreamHandler(), consolelevel)):
        handler.setLevel(level)
        handler.setFormatter(fmt)
        log.addHandler(handler)
    log.debug("Logging started at level %d", loglevel)
    return log
