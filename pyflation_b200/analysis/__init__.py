"""analysis - the adiabatic part of ``pyflation.analysis`` that the first-order path's parity
metric needs (P_R, scaled P_R).  The non-adiabatic spectra are out of scope."""
from .adiabatic import Pr, scaled_Pr, Pr_spectrum, Pphi_matrix, Pphi_modes, findns
from . import adiabatic, utilities

__all__ = ["Pr", "scaled_Pr", "Pr_spectrum", "Pphi_matrix", "Pphi_modes", "findns",
           "adiabatic", "utilities"]
