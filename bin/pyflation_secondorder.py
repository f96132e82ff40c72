#!/usr/bin/env python
"""pyflation_secondorder.py - run a second-order simulation from a first-order results file.

Drop-in for the reference's bin/pyflation_secondorder.py (:43-170): same options (-f/--filename,
-q/-v/--debug/--console) and the same ``runsomodel(mrgfile, filename, soargs, sodriver)`` entry.
The first-order file must carry the source term (dataset ``sourceterm``; the reference merges it in
with bin/pyflation_source.py + pyflation_combine.py, here ``sohelpers.combine_source_and_fofile``;
computing it is outside this package); the integration runs on the GPU (pf_so_run), results are
written in the reference's HDF5 layout (or as ``.npz`` when the file name ends so).
"""
import logging
import optparse
import os
import sys

try:
    from pyflation_b200 import cli, run_config, cosmomodels as c
except ImportError:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from pyflation_b200 import cli, run_config, cosmomodels as c

log = logging.getLogger("pyflation_secondorder")


def runsomodel(mrgfile, filename=None, soargs=None, sodriver=None):
    """Wrap the merged first-order + source file ``mrgfile``, run the three-stage driver
    (``sodriver``, default SOCanonicalThreeStage, built with ``soargs``) and save its results to
    ``filename`` (default ``run_config.soresults``).  Returns the name of the file written."""
    soargs = dict(run_config.soargs if soargs is None else soargs)
    if soargs.get("nfields", 1) > 1:
        message = "Only single field models can have second order perturbation calculated."
        log.error(message)
        raise c.ModelError(message)
    try:
        first_order = c.make_wrapper_model(mrgfile)
    except Exception:
        log.exception("Error wrapping model file.")
        raise
    driver = (sodriver or c.SOCanonicalThreeStage)(first_order, **soargs)
    try:
        driver.run(saveresults=False)
    except c.ModelError:
        log.exception("Something went wrong with model, quitting!")
        sys.exit(1)
    target = filename or run_config.soresults
    try:
        cli.make_parent_directory(target)
        target = driver.saveallresults(filename=target)
        log.info("Successfully ran and saved simulation in file %s.", target)
    except Exception:
        log.exception("IO error, nothing saved!")
    return target


def main(argv=None):
    parser = optparse.OptionParser()
    parser.add_option("-f", "--filename", action="store", dest="mrgresults", type="string", metavar="FILE",
                      default=run_config.mrgresults,
                      help="merged first order and source results file, default=%default")
    cli.add_log_options(parser, run_config.LOGLEVEL)
    options, _ = parser.parse_args(args=(argv or sys.argv)[1:])
    cli.configure_logging(options)
    if not os.path.isfile(options.mrgresults):
        raise IOError("Merged results file %s does not exist!" % options.mrgresults)
    if os.path.isfile(run_config.soresults):
        raise IOError("Second order results file already exists!")
    log.info("-----------Second order run requested------------------")
    try:
        runsomodel(mrgfile=options.mrgresults)
    except Exception:
        log.exception("Something went wrong during second order run!")
        return 1
    log.info("----------Second order run finished-------------------")
    return 0


if __name__ == "__main__":
    sys.exit(main())
