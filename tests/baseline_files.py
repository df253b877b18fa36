"""Puts the reference's committed baselines/ci_* directories back together from the fixtures
tests/golden/baseline_files.npz + baseline_files_meta.json (made by oracle/gen_golden.py:baseline_files from
/root/reference/baselines, which does not exist on the GPU box): real NETCDF3_64BIT_OFFSET files with the
baselines' dimensions, variable order, dtypes, attributes and values, plus the Newton_state.json files."""
import json
import os
import shutil

import numpy as np
from scipy.io import netcdf_file

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def materialise(root):
    """writes root/ci_*/<file>.nc and root/ci_*/Newton_state.json; returns root"""
    vals = np.load(os.path.join(GOLDEN, "baseline_files.npz"))
    with open(os.path.join(GOLDEN, "baseline_files_meta.json")) as fptr:
        meta = json.load(fptr)
    for key, entry in meta.items():
        fname = os.path.join(root, key + ".nc")
        os.makedirs(os.path.dirname(fname), exist_ok=True)
        with netcdf_file(fname, "w", version=2) as nc:
            for name, length in entry["dims"]:
                nc.createDimension(name, length)
            for var in entry["vars"]:
                ncvar = nc.createVariable(var["name"], np.dtype(var["dtype"]), tuple(var["dims"]))
                for akey, aval in var["attrs"].items():
                    setattr(ncvar, akey, aval)
                ncvar[:] = vals[f"{key}/{var['name']}"]
    for cfg in ("ci_long_dye_decay", "ci_long_iage", "ci_py_driver_2d_iage_column_regions"):
        os.makedirs(os.path.join(root, cfg), exist_ok=True)
        shutil.copyfile(os.path.join(GOLDEN, f"Newton_state_{cfg}.json"), os.path.join(root, cfg, "Newton_state.json"))
    return root
