"""Derived variables of the hist files: for every tracer-like variable the reference
writes its time mean, time anomaly, time standard deviation, end-minus-start difference and spatial
integrals / means next to the snapshots (nk_ooc/py_driver_2d/tracer_module_state.py:110-260,
nk_ooc/test_problem/tracer_module_state.py:96-199).  Host-side numpy on member 0's snapshots: this
is file output, not the hot path.  Units strings of the integrals are canonicalised as the reference does
with pint (utils.units_product: "years m", "mmol / m^2 / s"), so that baseline_cmp's metadata check passes."""

import numpy as np

from .utils import units_product


def time_mean_weights(n_time):
    """end points count half: the hist file holds t = 0 and t = end
    (py_driver_2d/tracer_module_state.py:204-212; test_problem/tracer_module_state.py:156-164)"""
    weights = np.full(n_time, 1.0 / (n_time - 1))
    weights[0] *= 0.5
    weights[-1] *= 0.5
    return weights


def derived_specs(name, attrs, depth, ypos=None):
    """[(varname, dimensions, long_name, units)] of the variables derived from tracer-like `name`; the two models
    word the long names differently (test_problem/tracer_module_state.py:115-141: "mean in time";
    py_driver_2d/tracer_module_state.py:131-160: "time mean")"""
    dn = depth.axisname
    cell = (dn,) if ypos is None else (dn, ypos.axisname)
    ln, un = attrs["long_name"], attrs["units"]
    if ypos is None:
        mean_s, anom_s, std_s = "mean in time", "anomaly in time", "std dev in time"
    else:
        mean_s, anom_s, std_s = "time mean", "time anomaly", "time std dev"
    out = [
        (f"{name}_time_mean", cell, f"{ln}, {mean_s}", un),
        (f"{name}_time_anom", ("time",) + cell, f"{ln}, {anom_s}", un),
        (f"{name}_time_std", cell, f"{ln}, {std_s}", un),
        (f"{name}_time_delta", cell, f"{ln}, end state minus start state", un),
    ]
    if ypos is None:
        out.append((f"{name}_{dn}_int", ("time",), f"{ln}, {dn} integral", units_product(un, depth.units)))
    else:
        yn = ypos.axisname
        out += [
            (f"{name}_depth_int", ("time", yn), f"{ln}, depth integral", units_product(un, depth.units)),
            (f"{name}_ypos_mean", ("time", dn), f"{ln}, ypos mean", un),
            (f"{name}_depth_ypos_int", ("time",), f"{ln}, depth-ypos integral",
             units_product(un, depth.units, ypos.units)),
        ]
    return out


def define_derived(fptr, name, attrs, depth, ypos=None):
    for varname, dims, long_name, units in derived_specs(name, attrs, depth, ypos):
        var = fptr.createVariable(varname, "f8", dims)
        var.long_name, var.units = long_name, units
        if "time" in dims:
            var.cell_methods = "time: point"


def derived_values(name, vals, depth, ypos=None):
    """{varname: array} from snapshots vals [time, depth(, ypos)]
    (py_driver_2d/tracer_module_state.py:214-260; test_problem/tracer_module_state.py:166-199)"""
    w = time_mean_weights(vals.shape[0])
    mean = np.einsum("i,i...", w, vals)
    anom = vals - mean
    out = {
        f"{name}_time_mean": mean,
        f"{name}_time_anom": anom,
        f"{name}_time_std": np.sqrt(np.einsum("i,i...", w, anom ** 2)),
        f"{name}_time_delta": vals[-1] - vals[0],
    }
    if ypos is None:
        out[f"{name}_{depth.axisname}_int"] = depth.int_vals_mid(vals, axis=-1)
    else:
        ypos_int = ypos.int_vals_mid(vals, axis=-1)
        out[f"{name}_depth_int"] = depth.int_vals_mid(vals, axis=-2)
        out[f"{name}_ypos_mean"] = ypos_int / (ypos.edges.max() - ypos.edges.min())
        out[f"{name}_depth_ypos_int"] = depth.int_vals_mid(ypos_int, axis=-1)
    return out


def write_derived(fptr, name, vals, depth, ypos=None):
    for varname, arr in derived_values(name, vals, depth, ypos).items():
        fptr.variables[varname][:] = arr
