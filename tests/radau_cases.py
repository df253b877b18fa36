"""shared by the Radau-parity GPU tests and scripts/error_vs_steps.py: device models for the cases of
tests/golden/radau_<grid>_<module>.npz (oracle/gen_golden_radau.py)"""
import os
import tempfile

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(grid, module):
    path = os.path.join(GOLDEN, f"radau_{grid}_{module}.npz")
    return np.load(path) if os.path.exists(path) else None


def forcing_record(g, depth, ypos):
    """(times, data[nt, nz, ny]) of the o2_like sink, scalef applied.  Small cases carry the record as the
    reference's gen_forcing_fcn produced it; the large grids rebuild it from the file's native 40 x 50 grid
    (stored with the 40 x 50 case) through the product's own reader — which then is part of what the
    comparison with the reference's F checks (nk_ooc/utils.py:488-537)."""
    if "frc_data" in g.files:
        return g["frc_time"], g["frc_data"]
    from scipy.io import netcdf_file

    from nk_ooc_b200.py_driver_2d.model_state import read_forcing

    src = load(str(g["frc_from"]), "forced")
    de, ye = src["depth_edges"], src["ypos_edges"]
    with tempfile.TemporaryDirectory() as tmp:
        fname = os.path.join(tmp, "po4_sms.nc")
        with netcdf_file(fname, "w", version=2) as f:
            f.createDimension("time", len(src["frc_time"]))
            f.createDimension("depth", len(de) - 1)
            f.createDimension("ypos", len(ye) - 1)
            for name, vals in (("time", src["frc_time"]), ("depth", 0.5 * (de[1:] + de[:-1])),
                               ("ypos", 0.5 * (ye[1:] + ye[:-1]))):
                f.createVariable(name, "f8", (name,))[:] = vals
            f.createVariable("po4_sms", "f8", ("time", "depth", "ypos"))[:] = -3.0 * src["frc_data"]
        return read_forcing(fname, "po4_sms", [depth.mid, ypos.mid], -1.0 / 3.0)


def model(g, module):
    from nk_ooc_b200.py_driver_2d import modules
    from nk_ooc_b200.spatial_axis import SpatialAxis

    depth, ypos = SpatialAxis("depth", g["depth_edges"]), SpatialAxis("ypos", g["ypos_edges"])
    tr = modules.Transport2D(depth, ypos, float(g["params"][3]), float(g["params"][4]))
    if module == "forced":
        # scripts/run_py_driver_2d_forced_o2_like.sh:14-25
        ft, fd = forcing_record(g, depth, ypos)
        return modules.forced_model(tr, "const", 1.0, 1.0 / 3600.0, "file", sms_times=ft, sms_data=fd, sink_thres=0.05)
    if module == "iage":
        return modules.iage_model(tr)
    return modules.phosphorus_model(tr)
