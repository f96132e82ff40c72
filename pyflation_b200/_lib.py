"""ctypes binding of ``libpyflation_b200.so`` (see ``include/pyflation_b200.h``).

There is deliberately no fallback: if the CUDA library has not been built the import of any
compute entry point raises, and every non-zero return code is turned into an exception.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PYFLATION_B200_LIB selects an alternative build of the same ABI (kernel tuning experiments)
LIB_PATH = os.environ.get("PYFLATION_B200_LIB", os.path.join(_HERE, "lib", "libpyflation_b200.so"))

PF_MAX_NFIELDS = 8            # compiled instantiations
PF_MAX_NFIELDS_RT = 64        # runtime-nf kernels (nflation)
PF_MAX_PARAMS = 8
PF_FLAG_BG_MISMATCH = 1
PF_FLAG_TIMEOUT = 2
PF_LAYOUT_FULL, PF_LAYOUT_PERT = 0, 1
PF_PEER_HANDLE_BYTES = 64
PF_SO_COEF = 6                # doubles per stage in pf_so_run's coefficient table

# names every build must export (checked by tests/test_abi.py against the header)
EXPORTS = [
    "pf_version", "pf_last_error", "pf_potential_id", "pf_potential_name",
    "pf_potential_nfields", "pf_potential_nparams", "pf_potential_param_name",
    "pf_potential_param_default", "pf_potential_eval", "pf_number_of_steps", "pf_nvar",
    "pf_rec_doubles", "pf_rec_bg_offset", "pf_bg_run", "pf_stage_table", "pf_fo_run", "pf_solve_scratch_doubles",
    "pf_fo_solve", "pf_pr_spectrum", "pf_pr_spectrum_pert", "pf_scatter_slab", "pf_copy_rows_d2h", "pf_copy_window_d2h",
    "pf_host_fill_background", "pf_crossing_rows", "pf_bunch_davies", "pf_derivs",
    "pf_so_run", "pf_so_derivs", "pf_np_sum",
    "pf_fp64_peak_probe", "pf_peer_alloc", "pf_peer_free", "pf_peer_export", "pf_peer_open", "pf_peer_close",
]


class PfModel(ctypes.Structure):
    _fields_ = [("potential_id", ctypes.c_int32),
                ("nfields", ctypes.c_int32),
                ("params", ctypes.c_double * PF_MAX_PARAMS),
                ("ainit", ctypes.c_double)]


class LibraryError(RuntimeError):
    """Raised for CUDA errors (positive return codes) and a missing library."""


class ArgumentError(ValueError):
    """Raised for negative return codes; carries ``code`` (PF_E_*)."""

    def __init__(self, code, msg):
        ValueError.__init__(self, msg)
        self.code = code


_lib = None


def lib():
    """Load the shared library on first use; raise loudly when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryError(
            "pyflation_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C pyflation_b200/csrc`). There is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    c_i64, c_dbl, c_vp, c_int = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_int
    mp = ctypes.POINTER(PfModel)
    L.pf_version.restype = c_int
    L.pf_last_error.restype = ctypes.c_char_p
    L.pf_potential_id.argtypes = [ctypes.c_char_p]
    L.pf_potential_id.restype = c_int
    L.pf_potential_name.argtypes = [c_int]
    L.pf_potential_name.restype = ctypes.c_char_p
    L.pf_potential_nfields.argtypes = [c_int]
    L.pf_potential_nparams.argtypes = [c_int]
    L.pf_potential_param_name.argtypes = [c_int, c_int]
    L.pf_potential_param_name.restype = ctypes.c_char_p
    L.pf_potential_param_default.argtypes = [c_int, c_int]
    L.pf_potential_param_default.restype = c_dbl
    L.pf_potential_eval.argtypes = [mp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]
    L.pf_np_sum.argtypes = [ctypes.POINTER(c_dbl), c_int]
    L.pf_np_sum.restype = c_dbl
    L.pf_number_of_steps.argtypes = [c_dbl, c_dbl, c_dbl]
    L.pf_number_of_steps.restype = c_i64
    L.pf_nvar.argtypes = [c_int]
    L.pf_nvar.restype = c_i64
    L.pf_rec_doubles.argtypes = [c_int]
    L.pf_rec_doubles.restype = c_i64
    L.pf_rec_bg_offset.argtypes = [c_int]
    L.pf_rec_bg_offset.restype = c_i64
    L.pf_bg_run.argtypes = [mp, c_i64, c_dbl, c_i64, ctypes.POINTER(c_dbl), c_vp, c_vp, c_vp]
    L.pf_stage_table.argtypes = [mp, c_i64, c_dbl, c_dbl, c_i64, c_vp, c_vp, c_vp, c_vp]
    L.pf_fo_run.argtypes = [mp, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_dbl, c_i64, c_i64,
                            c_vp, c_vp, c_i64, ctypes.c_int32, c_i64, c_vp, c_vp]
    L.pf_solve_scratch_doubles.argtypes = [c_int, c_i64]
    L.pf_solve_scratch_doubles.restype = c_i64
    L.pf_fo_solve.argtypes = [mp, c_i64, c_vp, c_vp, c_vp, c_i64, ctypes.POINTER(c_dbl), c_i64, c_dbl,
                              c_dbl, c_i64, c_vp, c_vp, c_vp, c_i64, ctypes.c_int32, c_i64, c_vp, c_vp]
    c_sz = ctypes.c_size_t
    L.pf_copy_rows_d2h.argtypes = [c_vp, c_sz, c_vp, c_sz, c_sz, c_sz, c_vp]
    L.pf_copy_window_d2h.argtypes = [c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp]
    L.pf_host_fill_background.argtypes = [c_vp, c_int, c_i64, c_i64, c_i64, c_vp, c_i64, c_i64,
                                          c_vp, c_vp, c_int, c_int]
    L.pf_pr_spectrum.argtypes = [c_int, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]
    L.pf_pr_spectrum_pert.argtypes = [c_int, c_i64, c_i64, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]
    L.pf_scatter_slab.argtypes = [c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]
    L.pf_crossing_rows.argtypes = [c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]
    L.pf_bunch_davies.argtypes = [c_int, c_i64, c_vp, c_vp, c_vp, c_dbl, c_dbl, c_dbl, c_dbl, c_int,
                                  c_int, c_vp, c_vp]
    L.pf_derivs.argtypes = [mp, c_int, c_i64, c_vp, c_dbl, c_vp, c_vp, c_vp]
    L.pf_so_run.argtypes = [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp,
                            ctypes.POINTER(c_dbl), c_i64, c_dbl, c_i64, c_i64, c_vp, c_vp, c_i64, c_i64, c_vp]
    L.pf_so_derivs.argtypes = [c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.POINTER(c_dbl),
                               c_vp, c_vp]
    L.pf_fp64_peak_probe.argtypes = [c_int, c_vp, c_i64, ctypes.POINTER(c_dbl)]
    L.pf_peer_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(c_vp)]
    L.pf_peer_free.argtypes = [c_vp]
    L.pf_peer_export.argtypes = [c_vp, ctypes.c_char_p]
    L.pf_peer_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(c_vp)]
    L.pf_peer_close.argtypes = [c_vp]
    _lib = L
    return L


def check(rc):
    """Translate a return code of the C ABI into an exception."""
    if rc == 0:
        return
    msg = lib().pf_last_error().decode("utf-8", "replace")
    if rc < 0:
        raise ArgumentError(rc, msg)
    raise LibraryError(msg)


def potential_table():
    """{name: (id, nfields or 0, [(param name, default), ...])} read from the library."""
    L = lib()
    table = {}
    i = 0
    while True:
        name = L.pf_potential_name(i)
        if name is None:
            break
        params = [(L.pf_potential_param_name(i, q).decode(), L.pf_potential_param_default(i, q))
                  for q in range(L.pf_potential_nparams(i))]
        table[name.decode()] = (i, L.pf_potential_nfields(i), params)
        i += 1
    return table
