"""cli - what the command-line callers in bin/ share: the verbosity options of the reference's
scripts (-q / -v / --debug / --console, e.g. bin/pyflation_firstorder.py:171-187) and the
results-directory handling."""
import logging
import optparse
import os

# flag(s), level, help text -- same flags and meaning as the reference's "Log Options" group
_LEVELS = ((("-q", "--quiet"), logging.FATAL, "only print fatal error messages"),
           (("-v", "--verbose"), logging.INFO, "print informative messages"),
           (("--debug",), logging.DEBUG, "log lots of debugging information"))


def add_log_options(parser, default_level):
    group = optparse.OptionGroup(parser, "Log Options",
                                 "These options affect the verbosity of the log output.")
    for flags, level, text in _LEVELS:
        extra = {"default": default_level} if level == logging.DEBUG else {}
        group.add_option(*flags, action="store_const", const=level, dest="loglevel", help=text, **extra)
    group.add_option("--console", action="store_true", dest="console", default=False,
                     help="if selected matches console log level to selected log level")
    parser.add_option_group(group)
    return parser


def configure_logging(options):
    logging.basicConfig(level=options.loglevel if options.console else logging.WARN)


def make_parent_directory(filename):
    folder = os.path.dirname(filename)
    if folder and not os.path.isdir(folder):
        os.makedirs(folder)
