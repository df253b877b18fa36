#!/usr/bin/env python
"""Compare one netCDF file with its baseline — the command the reference's CI scripts call as
`python -m nk_ooc.baseline_cmp` (nk_ooc/baseline_cmp.py:13-49): same options, same default tolerances
(rtol 1e-7, atol 2e-9), exit status 0 when metadata AND values agree, 1 otherwise.

    python -m nk_ooc_b200.baseline_cmp --fname fcn_00.nc --expr_dir W/gen_init_iterate --baseline_dir baselines/ci_short

Files are read with scipy's NETCDF3 reader (nk_ooc_b200/utils.py); `--anom_suffix` is an extension (off by default)."""

import argparse
import logging
import os
import sys

from . import utils

# option -> (help, type, default); the first five are the reference's
OPTIONS = {
    "fname": ("name of file to be compared", str, None),
    "expr_dir": ("directory with file", str, None),
    "baseline_dir": ("directory with baseline file", str, None),
    "rtol": ("relative tolerance", float, 1.0e-7),
    "atol": ("absolute tolerance", float, 2.0e-9),
    "anom_suffix": ("(extension) variables <x><suffix> are anomalies of x: compare them with x's tolerance "
                    "atol + rtol |x| instead of atol + rtol |x - mean(x)|", str, None),
}


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description=__doc__.split("\n")[0],
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    for name, (text, kind, default) in OPTIONS.items():
        parser.add_argument(f"--{name}", help=text, type=kind, default=default)
    return parser.parse_args(list(argv or []))


def compare(fname, expr_dir, baseline_dir, rtol=OPTIONS["rtol"][2], atol=OPTIONS["atol"][2], anom_suffix=None):
    """True when the file in expr_dir has the baseline's metadata and values; both checks always run, so that the
    log names every difference (as the reference's main does)"""
    log = logging.getLogger(__name__)
    ours, theirs = os.path.join(expr_dir, fname), os.path.join(baseline_dir, fname)
    log.info("expr_fname = %s", ours)
    log.info("baseline_fname = %s", theirs)
    checks = [utils.metadata_same(ours, theirs),
              utils.isclose_all_vars(ours, theirs, rtol=rtol, atol=atol, anom_suffix=anom_suffix)]
    return all(checks)


def main(args):
    logging.basicConfig(format="%(filename)s:%(funcName)s:%(message)s", level="INFO", stream=sys.stdout)
    same = compare(args.fname, args.expr_dir, args.baseline_dir, args.rtol, args.atol, args.anom_suffix)
    sys.exit(0 if same else 1)


if __name__ == "__main__":
    main(parse_args(sys.argv[1:]))
