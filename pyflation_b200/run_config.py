"""run_config - the per-run choices the first-order script reads (reference run_config.py:51-185):
the potential / initial-condition fixtures, the k ranges and the driver class.  Values are the
reference's; a run directory may shadow this module with its own ``run_config.py`` exactly as
with the reference."""
import logging
import os

import numpy as np

from . import cosmomodels as c
from .helpers import getkend

LOGLEVEL = logging.INFO


def _fixture(potential, nfields, bgystart, **extra):
    d = {"potential_func": potential, "pot_params": {"nfields": nfields}, "nfields": nfields,
         "bgystart": None if bgystart is None else np.array(bgystart, dtype=float),
         "cq": 50, "solver": "rkdriver_tsix"}
    d.update(extra)
    return d


# run_config.py:51-123
fixtures = {
    "msqphisq": _fixture("msqphisq", 1, [18.0, -0.1, 0]),
    "lambdaphi4": _fixture("lambdaphi4", 1, [25.0, 0, 0]),
    "hybrid2and4": _fixture("hybrid2and4", 1, [25.0, 0, 0]),
    "linde": _fixture("linde", 1, [25.0, 0, 0]),
    "phi2over3": _fixture("phi2over3", 1, [10.0, 0, 0]),
    "msqphisq_withV0": _fixture("msqphisq_withV0", 1, [18.0, 0, 0]),
    "step_potential": _fixture("step_potential", 1, [18.0, -0.1, 0]),
    "bump_potential": _fixture("bump_potential", 1, [18.0, -0.1, 0]),
    "hybridquadratic": _fixture("hybridquadratic", 2, [12.0, 1 / 300.0, 12.0, 49 / 300.0, 0]),
    "nflation": _fixture("nflation", 2, None),
    "hybridquartic": _fixture("hybridquartic", 2, [1e-2, 2e-8, 1.63e-9, 3.26e-7, 0]),
    "productexponential": _fixture("productexponential", 2, [18.0, 0.0, 0.001, 0, 0]),
}
foargs = fixtures["msqphisq"]

# run_config.py:142-158
K_ranges = {"K1": {"kinit": 0.5e-61, "deltak": 1e-61, "numsoks": 1025},
            "K2": {"kinit": 1.5e-61, "deltak": 3e-61, "numsoks": 1025},
            "K3": {"kinit": 0.25e-60, "deltak": 1e-60, "numsoks": 1025},
            "K4": {"kinit": 0.85e-60, "deltak": 0.4e-60, "numsoks": 1025}}
K_range = K_ranges["K4"]
kinit, deltak, numsoks = K_range["kinit"], K_range["deltak"], K_range["numsoks"]
kend = getkend(kinit, deltak, numsoks)

# run_config.py:170-185
foclass = c.FOCanonicalTwoStage

RESULTSDIR = os.path.join(os.getcwd(), "results")
# results file names (run_config.py:252-261): HDF5 in the reference's layout; a name ending in
# ".npz" makes the scripts write / read NumPy archives instead
foresults = os.path.join(RESULTSDIR, "fo.hf5")
srcresults = os.path.join(RESULTSDIR, "src.hf5")

# second-order stage (run_config.py:178-181, 215-217, 258-260): the ramped class is the default
soclass = c.CanonicalRampedSecondOrder
soargs = {"solver": "rkdriver_tsix",
          "nfields": 1,                 # only single field models can have second order calculated
          "soclass": soclass}
mrgresults = os.path.join(RESULTSDIR, "mrg.hf5")
soresults = os.path.join(RESULTSDIR, "so.hf5")
