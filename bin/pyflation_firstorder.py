#!/usr/bin/env python
"""pyflation_firstorder.py - run a first-order simulation and save the results.

Drop-in for the reference's bin/pyflation_firstorder.py (:51-216): same options (-f/--filename,
-p/--potential, -q/-v/--debug/--console), same ``runfomodel(filename, foargs, foclass)`` entry,
fixtures and k range from ``run_config``; the integration runs on the GPU.  Results are written
as ``.npz`` (the reference's HDF5 format is out of scope); extra option
``--kstride N`` keeps every N-th k of the range, ``--rowstride N`` every N-th time row.
"""
import logging
import optparse
import os
import sys

try:
    from pyflation_b200 import run_config, helpers, cosmomodels as c
except ImportError:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from pyflation_b200 import run_config, helpers, cosmomodels as c

log = logging.getLogger("pyflation_firstorder")


def runfomodel(filename=None, foargs=None, foclass=None, row_stride=1):
    """Create the first-order driver with ``foargs``, run it and save the results."""
    if foargs is None:
        foargs = dict(run_config.foargs)
    if "k" not in foargs:
        foargs["k"] = helpers.seq(run_config.kinit, run_config.kend, run_config.deltak)
    if filename is None:
        filename = run_config.foresults
    if not foclass:
        foclass = run_config.foclass
    if not issubclass(foclass, c.FOCanonicalTwoStage):
        raise ValueError("Must use FOCanonicalTwoStage class for first order run!")
    model = foclass(**foargs)
    if row_stride != 1:
        model.row_stride = row_stride
    try:
        log.debug("Starting model run...")
        model.run(saveresults=False)
        log.debug("Model run finished.")
    except c.ModelError:
        log.exception("Something went wrong with model, quitting!")
        sys.exit(1)
    try:
        dirname = os.path.dirname(filename)
        if dirname and not os.path.isdir(dirname):
            os.makedirs(dirname)
        filename = model.saveallresults(filename=filename)
        log.info("Successfully ran and saved simulation in file %s.", filename)
    except Exception:
        log.exception("IO error, nothing saved!")
    del model
    return filename


def main(argv=None):
    if not argv:
        argv = sys.argv
    parser = optparse.OptionParser()
    parser.add_option("-f", "--filename", action="store", dest="foresults", default=run_config.foresults,
                      type="string", metavar="FILE", help="file to store results")
    parser.add_option("-p", "--potential", action="store", dest="fixture", default="", type="string",
                      metavar="POTENTIAL", help="the potential to use, choose from the fixtures dict in run_config.py")
    parser.add_option("--kstride", action="store", dest="kstride", default=1, type="int",
                      help="keep every N-th k of the configured range")
    parser.add_option("--rowstride", action="store", dest="rowstride", default=1, type="int",
                      help="keep every N-th time row of the trajectory")
    loggroup = optparse.OptionGroup(parser, "Log Options", "These options affect the verbosity of the log output.")
    loggroup.add_option("-q", "--quiet", action="store_const", const=logging.FATAL, dest="loglevel",
                        help="only print fatal error messages")
    loggroup.add_option("-v", "--verbose", action="store_const", const=logging.INFO, dest="loglevel",
                        help="print informative messages")
    loggroup.add_option("--debug", action="store_const", const=logging.DEBUG, dest="loglevel",
                        help="log lots of debugging information", default=run_config.LOGLEVEL)
    loggroup.add_option("--console", action="store_true", dest="console", default=False,
                        help="if selected matches console log level to selected log level")
    parser.add_option_group(loggroup)
    options, _ = parser.parse_args(args=argv[1:])
    logging.basicConfig(level=options.loglevel if options.console else logging.WARN)
    log.info("-----------First order run requested------------------")
    if os.path.isfile(options.foresults):
        raise IOError("First order results file already exists!")
    if options.fixture != "":
        try:
            foargs = dict(run_config.fixtures[options.fixture])
            log.info("Potential %s selected." % foargs["potential_func"])
        except KeyError:
            log.exception("Fixture doesn't exist in run_config.py!")
            return 1
    else:
        foargs = dict(run_config.foargs)
    foargs["k"] = helpers.seq(run_config.kinit, run_config.kend, run_config.deltak)[::max(1, options.kstride)]
    try:
        runfomodel(filename=options.foresults, foargs=foargs, row_stride=options.rowstride)
    except Exception:
        log.exception("Something went wrong during first order run!")
        return 1
    log.info("----------First order run finished-------------------")
    return 0


if __name__ == "__main__":
    sys.exit(main())
