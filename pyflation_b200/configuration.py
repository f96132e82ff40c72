"""Configuration constants (mirrors pyflation/configuration.py:20-46 for the names the
first-order path reads)."""
import logging

# debug logging control: 0 for off, 1 for on           (configuration.py:20)
_debug = 0

# default log level, can be overridden by callers     (configuration.py:25)
LOGLEVEL = logging.INFO

# Directory names used by the reference's run scaffolding (configuration.py:30-36)
CODEDIRNAME = "."
RUNDIRNAME = "runs"
RESULTSDIRNAME = "results"
LOGDIRNAME = "applogs"

# On-disk codec of the reference's HDF5 results (configuration.py:45-46); results I/O is
# outside this package's scope, the names are kept for callers that read them.
hdf5complib = "blosc"
hdf5complevel = 2
