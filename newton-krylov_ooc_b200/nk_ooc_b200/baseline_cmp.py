#!/usr/bin/env python
"""compare a netCDF file to a baseline: the reference's `python -m nk_ooc.baseline_cmp` (nk_ooc/baseline_cmp.py:13-49)
with the same arguments, defaults (rtol 1e-7, atol 2e-9) and exit status, for the files this path writes

    python -m nk_ooc_b200.baseline_cmp --fname fcn_00.nc --expr_dir W/gen_init_iterate --baseline_dir baselines/ci_short
"""

import argparse
import logging
import os
import sys

from .utils import isclose_all_vars, metadata_same


def parse_args(args_list_in=None):
    args_list = [] if args_list_in is None else args_list_in
    parser = argparse.ArgumentParser(description="compare netCDF file to baseline",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--fname", help="name of file to be compared")
    parser.add_argument("--expr_dir", help="directory with file")
    parser.add_argument("--baseline_dir", help="directory with baseline file")
    parser.add_argument("--rtol", help="relative tolerance", type=float, default=1.0e-7)
    parser.add_argument("--atol", help="absolute tolerance", type=float, default=2.0e-9)
    parser.add_argument("--anom_suffix", default=None,
                        help="(extension) compare variables <x><suffix>, anomalies of x, with x's tolerance "
                             "atol + rtol |x| instead of atol + rtol |x - mean(x)|")
    return parser.parse_args(args_list)


def compare(fname, expr_dir, baseline_dir, rtol=1.0e-7, atol=2.0e-9, anom_suffix=None):
    """True when metadata and values agree (both checks always run, as in the reference)"""
    logger = logging.getLogger(__name__)
    baseline_fname = os.path.join(baseline_dir, fname)
    expr_fname = os.path.join(expr_dir, fname)
    logger.info("expr_fname = %s", expr_fname)
    logger.info("baseline_fname = %s", baseline_fname)
    res = True
    if not metadata_same(expr_fname, baseline_fname):
        res = False
    if not isclose_all_vars(expr_fname, baseline_fname, rtol=rtol, atol=atol, anom_suffix=anom_suffix):
        res = False
    return res


def main(args):
    logging.basicConfig(format="%(filename)s:%(funcName)s:%(message)s", level="INFO", stream=sys.stdout)
    sys.exit(0 if compare(args.fname, args.expr_dir, args.baseline_dir, args.rtol, args.atol, args.anom_suffix) else 1)


if __name__ == "__main__":
    main(parse_args(sys.argv[1:]))
