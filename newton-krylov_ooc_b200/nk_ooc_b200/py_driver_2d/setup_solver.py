"""grid set-up of py_driver_2d — mirror of nk_ooc/py_driver_2d/setup_solver.py:134-198
(gen_grid_vars_file: axes, grid_weight, region_mask incl. column regions)."""

from datetime import datetime

import numpy as np
from scipy.io import netcdf_file

from ..spatial_axis import spatial_axis_from_defn


def gen_axis(axisname, modelinfo):
    """axis from the `<axisname>_*` entries of modelinfo (setup_solver.py:185-198)"""
    kw = {"axisname": axisname}
    for key, typ in (("units", str), ("nlevs", int), ("edge_start", float), ("edge_end", float),
                     ("delta_ratio_max", float), ("delta_start", float)):
        name = f"{axisname}_{key}"
        if name in modelinfo:
            kw[key] = typ(modelinfo[name])
    return spatial_axis_from_defn(**kw)


def column_region_mask(shape, max_abs_vvel, horiz_mix_coeff):
    """every ypos column is its own region when there are no lateral processes
    (setup_solver.py:173-182); int32, 1-based"""
    if max_abs_vvel == 0.0 and horiz_mix_coeff == 0.0:
        mask = np.empty(shape, dtype=np.int32)
        for ypos_i in range(shape[1]):
            mask[:, ypos_i] = ypos_i + 1
        return mask
    return np.ones(shape, dtype=np.int32)


def gen_grid_vars_file(modelinfo):
    """write modelinfo["grid_vars_fname"]; returns (depth, ypos)"""
    axes = {name: gen_axis(name, modelinfo) for name in ("depth", "ypos")}
    with netcdf_file(modelinfo["grid_vars_fname"], "w", version=2) as fptr:
        stamp = datetime.now().strftime("%Y-%m-%d %H:%M:%S")
        fptr.history = f"{stamp}: created by nk_ooc_b200.py_driver_2d.setup_solver.gen_grid_vars_file"
        for axis in axes.values():
            axis.define(fptr)
        var = fptr.createVariable("grid_weight", "f8", ("depth", "ypos"))
        var.long_name = "grid-cell area"
        var.units = "m^2"
        var = fptr.createVariable("region_mask", "i4", ("depth", "ypos"))
        var.long_name = "Region Mask"
        var.cell_measures = "area: grid_weight"
        for axis in axes.values():
            axis.write(fptr)
        weight = np.outer(axes["depth"].delta, axes["ypos"].delta)
        fptr.variables["grid_weight"][:] = weight
        fptr.variables["region_mask"][:] = column_region_mask(
            weight.shape, float(modelinfo["max_abs_vvel"]), float(modelinfo["horiz_mix_coeff"]))
    return axes["depth"], axes["ypos"]
