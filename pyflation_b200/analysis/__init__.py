"""analysis - the adiabatic part of ``pyflation.analysis`` that the first-order path's parity
metric needs (P_R, P_zeta and their scaled forms).  The non-adiabatic spectra are out of scope
(DESIGN.md)."""
from .adiabatic import (Pr, scaled_Pr, Pzeta, scaled_Pzeta, Pr_spectrum, Pzeta_spectrum, Pphi_matrix,
                        Pphi_modes, findns)
from . import adiabatic, utilities

__all__ = ["Pr", "scaled_Pr", "Pzeta", "scaled_Pzeta", "Pr_spectrum", "Pzeta_spectrum", "Pphi_matrix", "Pphi_modes", "findns",
           "adiabatic", "utilities"]
