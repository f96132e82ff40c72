#!/usr/bin/env python
"""pyflation_secondorder.py - run a second-order simulation from a first-order results file.

Drop-in for the reference's bin/pyflation_secondorder.py (:43-170): same options (-f/--filename,
-q/-v/--debug/--console) and the same ``runsomodel(mrgfile, filename, soargs, sodriver)`` entry.
The first-order file must carry the source term (dataset ``sourceterm``; the reference merges it in
with bin/pyflation_source.py + pyflation_combine.py, computing it is outside this package); the
integration runs on the GPU (pf_so_run), results are written as ``.npz``.
"""
import logging
import optparse
import os
import sys

try:
    from pyflation_b200 import run_config, cosmomodels as c
except ImportError:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from pyflation_b200 import run_config, cosmomodels as c

log = logging.getLogger("pyflation_secondorder")


def runsomodel(mrgfile, filename=None, soargs=None, sodriver=None):
    """Wrap the merged first-order + source file, run the three-stage driver, save the results."""
    if soargs is None:
        soargs = dict(run_config.soargs)
    if soargs.get("nfields", 1) > 1:
        log.error("Only single field models can have second order perturbation calculated.")
        raise c.ModelError("Only single field models can have second order perturbation calculated.")
    try:
        fomodel = c.make_wrapper_model(mrgfile)
    except Exception:
        log.exception("Error wrapping model file.")
        raise
    if sodriver is None:
        sodriver = c.SOCanonicalThreeStage
    somodel = sodriver(fomodel, **soargs)
    try:
        log.debug("Starting model run...")
        somodel.run(saveresults=False)
        log.debug("Model run finished.")
    except c.ModelError:
        log.exception("Something went wrong with model, quitting!")
        sys.exit(1)
    if filename is None:
        filename = run_config.soresults
    try:
        dirname = os.path.dirname(filename)
        if dirname and not os.path.isdir(dirname):
            os.makedirs(dirname)
        filename = somodel.saveallresults(filename=filename)
        log.info("Successfully ran and saved simulation in file %s.", filename)
    except Exception:
        log.exception("IO error, nothing saved!")
    del somodel
    return filename


def main(argv=None):
    if not argv:
        argv = sys.argv
    parser = optparse.OptionParser()
    parser.add_option("-f", "--filename", action="store", dest="mrgresults", default=run_config.mrgresults,
                      type="string", metavar="FILE",
                      help="merged first order and source results file, default=%default")
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
    if not os.path.isfile(options.mrgresults):
        raise IOError("Merged results file %s does not exist!" % options.mrgresults)
    if os.path.isfile(run_config.soresults):
        raise IOError("Second order results file already exists!")
    log.info("-----------Second order run requested------------------")
    try:
        runsomodel(mrgfile=options.mrgresults, soargs=dict(run_config.soargs))
    except Exception:
        log.exception("Something went wrong during second order run!")
        return 1
    log.info("----------Second order run finished-------------------")
    return 0


if __name__ == "__main__":
    sys.exit(main())
