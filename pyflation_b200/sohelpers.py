"""sohelpers - drop-in for pyflation/sohelpers.py: merge a source-term file into a first-order
results file ahead of the second-order run (sohelpers.py:20-86).  Files are the ``.npz``
archives of this package (``saveallresults``); the source file holds ``sourceterm`` (T, nk_s),
``k`` (nk_s,) and ``nix`` (the time steps it was calculated for)."""
import logging
import os

import numpy as np

_log = logging.getLogger(__name__)

# first-order datasets whose last axis runs over k: cut to the source term's k range (:67-71)
_PER_K = ("yresult", "fotstart", "fotstartindex", "foystart")


def combine_source_and_fofile(sourcefile, fofile, newfile=None):
    """Copy the source term into a copy of the first-order file; returns the new file's name.
    Default name: the source file's, with "src" replaced by "foandsrc", next to the
    first-order file (an existing file of that name is not overwritten: "foandsrc_")."""
    if not newfile or not os.path.isdir(os.path.dirname(newfile)):
        newfile = os.path.join(os.path.dirname(fofile), os.path.basename(sourcefile).replace("src", "foandsrc"))
        _log.info("Saving combined source and first order results in file %s.", newfile)
    if os.path.isfile(newfile):
        newfile = newfile.replace("foandsrc", "foandsrc_")
    try:
        src = np.load(sourcefile, allow_pickle=False)
        fo = np.load(fofile, allow_pickle=False)
    except IOError:
        _log.exception("Source or first order files not found!")
        raise
    missing = [name for name in ("sourceterm", "k", "nix") if name not in src.files]
    if missing:
        raise KeyError("Source term file not in correct format! Missing: %s" % ", ".join(missing))
    if len(fo["tresult"]) != len(src["nix"]):
        raise ValueError("Not all timesteps have had source term calculated!")
    numks = len(src["k"])
    merged = {name: fo[name] for name in fo.files}
    for name in _PER_K:
        if name in merged:
            merged[name] = merged[name][..., :numks]
    merged["sourceterm"] = src["sourceterm"]
    merged["k"] = src["k"]
    np.savez(newfile, **merged)
    _log.info("Source term successfully copied to new file %s.", newfile)
    return newfile if newfile.endswith(".npz") else newfile + ".npz"
