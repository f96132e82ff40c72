#!/usr/bin/env python
"""pyflation_firstorder.py - run a first-order simulation and save the results.

Drop-in for the reference's bin/pyflation_firstorder.py (:51-216): same options (-f/--filename,
-p/--potential, -q/-v/--debug/--console), same ``runfomodel(filename, foargs, foclass)`` entry,
fixtures and k range from ``run_config``; the integration runs on the GPU.  Results are written
in the reference's HDF5 layout (pyflation_b200.hdf5lite; a file name ending in ``.npz`` selects a
NumPy archive instead); extra option
``--kstride N`` keeps every N-th k of the range, ``--rowstride N`` every N-th time row.
"""
import logging
import optparse
import os
import sys

try:
    from pyflation_b200 import cli, run_config, helpers, cosmomodels as c
except ImportError:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from pyflation_b200 import cli, run_config, helpers, cosmomodels as c

log = logging.getLogger("pyflation_firstorder")


def runfomodel(filename=None, foargs=None, foclass=None, row_stride=1):
    """Build the first-order driver from ``foargs`` (default: run_config.foargs, k grid from
    run_config's range), run it on the GPU and save the results; returns the file written."""
    settings = dict(run_config.foargs if foargs is None else foargs)
    settings.setdefault("k", helpers.seq(run_config.kinit, run_config.kend, run_config.deltak))
    driver_class = foclass or run_config.foclass
    if not issubclass(driver_class, c.FOCanonicalTwoStage):
        raise ValueError("Must use FOCanonicalTwoStage class for first order run!")
    target = filename or run_config.foresults
    model = driver_class(**settings)
    if row_stride != 1:
        model.row_stride = row_stride
    try:
        model.run(saveresults=False)
    except c.ModelError:
        log.exception("Something went wrong with model, quitting!")
        sys.exit(1)
    try:
        cli.make_parent_directory(target)
        target = model.saveallresults(filename=target)
        log.info("Successfully ran and saved simulation in file %s.", target)
    except Exception:
        log.exception("IO error, nothing saved!")
    return target


def main(argv=None):
    parser = optparse.OptionParser()
    parser.add_option("-f", "--filename", action="store", dest="foresults", default=run_config.foresults,
                      type="string", metavar="FILE", help="file to store results")
    parser.add_option("-p", "--potential", action="store", dest="fixture", default="", type="string",
                      metavar="POTENTIAL", help="the potential to use, choose from the fixtures dict in run_config.py")
    parser.add_option("--kstride", action="store", dest="kstride", default=1, type="int",
                      help="keep every N-th k of the configured range")
    parser.add_option("--rowstride", action="store", dest="rowstride", default=1, type="int",
                      help="keep every N-th time row of the trajectory")
    cli.add_log_options(parser, run_config.LOGLEVEL)
    options, _ = parser.parse_args(args=(argv or sys.argv)[1:])
    cli.configure_logging(options)
    log.info("-----------First order run requested------------------")
    if os.path.isfile(options.foresults):
        raise IOError("First order results file already exists!")
    if options.fixture and options.fixture not in run_config.fixtures:
        log.error("Fixture doesn't exist in run_config.py!")
        return 1
    foargs = dict(run_config.fixtures[options.fixture] if options.fixture else run_config.foargs)
    log.info("Potential %s selected." % foargs["potential_func"])
    kgrid = helpers.seq(run_config.kinit, run_config.kend, run_config.deltak)
    foargs["k"] = kgrid[::max(1, options.kstride)]
    try:
        runfomodel(filename=options.foresults, foargs=foargs, row_stride=options.rowstride)
    except Exception:
        log.exception("Something went wrong during first order run!")
        return 1
    log.info("----------First order run finished-------------------")
    return 0


if __name__ == "__main__":
    sys.exit(main())
