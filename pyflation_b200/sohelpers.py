"""sohelpers - drop-in for pyflation/sohelpers.py: merge a source-term file into a first-order
results file ahead of the second-order run (sohelpers.py:20-86).  Files are either the reference's
HDF5 results files (read and written by :mod:`pyflation_b200.hdf5lite`) or the ``.npz`` archives
of this package (``saveallresults``); the source file holds ``sourceterm`` (T, nk_s), ``k``
(nk_s,) and ``nix`` (the time steps it was calculated for) -- under ``/results`` in an HDF5 file."""
import logging
import os

import numpy as np

from . import hdf5lite

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
    if os.path.isfile(fofile) and hdf5lite.is_hdf5_file(fofile):
        return _combine_hdf5(sourcefile, fofile, newfile)
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


def _combine_hdf5(sourcefile, fofile, newfile):
    """The HDF5 form of the merge, node for node as sohelpers.py:46-86: the per-k arrays cut to
    the source's k range, ``tresult`` / ``parameters`` / ``pot_params`` and the whole
    ``bgresults`` group copied, then ``sourceterm`` and the source's ``k``."""
    sf = ff = nf = None
    try:
        try:
            if hdf5lite.is_hdf5_file(sourcefile):
                sf = hdf5lite.open_file(sourcefile, "r")
            else:
                np.load(sourcefile, allow_pickle=False).close()         # IOError when it is missing
            ff = hdf5lite.open_file(fofile, "r")
            nf = hdf5lite.open_file(newfile, "w")                        # a new file, never an append
        except IOError:
            _log.exception("Source or first order files not found!")
            raise
        try:
            if sf is not None:
                sres = sf.root.results
                sterm, srck, srcnix = sres.sourceterm, sres.k.read(), sres.nix.read()
            else:
                z = np.load(sourcefile, allow_pickle=False)
                sterm, srck, srcnix = z["sourceterm"], z["k"], z["nix"]
        except (hdf5lite.NoSuchNodeError, KeyError) as err:
            _log.exception("Source term file not in correct format!")
            raise KeyError("Source term file not in correct format! (%s)" % (err,))
        fres = ff.root.results
        if len(fres.tresult) != len(srcnix):
            raise ValueError("Not all timesteps have had source term calculated!")
        nres = nf.create_group(nf.root, "results", fres._v_attrs.get("TITLE", ""))
        numks = len(srck)
        for name in _PER_K:
            nf.create_array(nres, name, getattr(fres, name)[..., :numks])
        for name in ("tresult", "parameters", "pot_params"):
            nf.copy_node(getattr(fres, name), nres)
        if "bgresults" in ff.root:
            nf.copy_node(ff.root.bgresults, nf.root, recursive=True)
        if sf is not None:
            nf.copy_node(sterm, nres)
        else:
            nf.create_array(nres, "sourceterm", sterm)
        nf.create_array(nres, "k", srck)
        nf.close()
        nf = None
        _log.info("Source term successfully copied to new file %s.", newfile)
    finally:
        if nf is not None:                              # failed: leave no half-written file behind
            nf.isopen = False
            if os.path.isfile(newfile):
                os.remove(newfile)
        for fh in (sf, ff):
            if fh is not None:
                fh.close()
    return newfile
