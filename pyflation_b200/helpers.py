"""helpers - the k-grid helpers of pyflation/helpers.py that callers of the first-order path
use (run_config.py:142-158, bin/pyflation_firstorder.py:206)."""
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
