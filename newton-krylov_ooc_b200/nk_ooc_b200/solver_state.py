"""Persistence of an iterative solver's progress and its statistics file.

`SolverState` keeps the reference's `<name>_state.json` (nk_ooc/solver_state.py:14-146): the
iteration counter, the log of completed steps ("NN:step" for per-iteration steps) and saved values
(numpy arrays tagged `__ndarray__`), rewritten after every change so that an interrupted solve can
be resumed (`resume=True`) or its last step redone (`rewind=True`).

`StatsFile` writes `<name>_stats.nc` (nk_ooc/stats_file.py:13-139; variables defined by
solver_base.py:71-125): NETCDF3 64-bit offset, unlimited `iteration` dimension, `region` dimension,
per tracer module `{iterate,fcn,increment}_{mean,norm}_<module>`, `increment_scalef_<module>`,
`Armijo_factor_<module>`, `Krylov_iterations` (Newton) and `precond_rhs_norm_<module>`,
`precond_resid_norm_<module>` (Krylov), plus the model's own statistics of every iteration's hist file
(`ModelStateBase.def_stats_vars / put_stats_vars`, model_state_base.py:136-180): the time mean of each
tracer-like variable and, on the lat-depth grid, its ypos mean
(py_driver_2d/tracer_module_state.py:281-345, test_problem/tracer_module_state.py:196-240).
The file is rewritten from an in-memory copy at every put
(classic netCDF cannot be grown in place by scipy's writer); a resumed solve reloads it first.
"""

import json
import os
from datetime import datetime

import numpy as np
from scipy.io import netcdf_file

FILL_F8 = 9.969209968386869e36  # netCDF4 default_fillvals["f8"]
FILL_I4 = -2147483647


class _NumpyEncoder(json.JSONEncoder):
    def default(self, o):  # pylint: disable=method-hidden
        if isinstance(o, np.ndarray):
            return {"__ndarray__": o.tolist()}
        if isinstance(o, (np.floating, np.integer)):
            return o.item()
        return json.JSONEncoder.default(self, o)


def _decode(dct):
    if "__ndarray__" in dct:
        return np.asarray(dct["__ndarray__"])
    return dct


class SolverState:
    """step log + saved values of one solver (solver_state.py:14-146)"""

    def __init__(self, name, workdir, resume=False, rewind=False):
        os.makedirs(workdir, exist_ok=True)
        self._name = name
        self._workdir = workdir
        self._state_fname = os.path.join(workdir, f"{name}_state.json")
        self._rewound_step_string = None
        if resume:
            self._read()
            if rewind:
                self._rewound_step_string = self._saved_state["step_log"].pop()
        else:
            if rewind:
                raise RuntimeError(f"rewind cannot be True if resume is False, name={name}")
            self._saved_state = {"iteration": 0, "step_log": []}
            self.log_step("__init__", per_iteration=False)

    def get_workdir(self):
        return self._workdir

    def get_iteration(self):
        return self._saved_state["iteration"]

    def inc_iteration(self):
        self._saved_state["iteration"] += 1
        self.log_step("inc_iteration")
        return self._saved_state["iteration"]

    def _step_log_string(self, stepval, per_iteration):
        return f"{self.get_iteration():02}:{stepval}" if per_iteration else stepval

    def log_step(self, stepval, per_iteration=True):
        if not self.step_logged(stepval, per_iteration):
            self._saved_state["step_log"].append(self._step_log_string(stepval, per_iteration))
            self._write()

    def step_logged(self, stepval, per_iteration=True):
        return self._step_log_string(stepval, per_iteration) in self._saved_state["step_log"]

    def step_was_rewound(self, stepval, per_iteration=True):
        if self._rewound_step_string is None:
            return False
        return self._step_log_string(stepval, per_iteration) == self._rewound_step_string

    def set_value_saved_state(self, key, value):
        """store a value and confirm that it is read back exactly (solver_state.py:107-118)"""
        self._saved_state[key] = value
        self._write()
        self._read()
        back = self._saved_state[key]
        same = np.array_equal(back, value) if isinstance(value, np.ndarray) else back == value
        if not same:
            raise RuntimeError("saved_state value not recovered on reread")

    def get_value_saved_state(self, key):
        return self._saved_state[key]

    def has_value(self, key):
        return key in self._saved_state

    def _write(self):
        with open(self._state_fname, mode="w") as fptr:
            json.dump(self._saved_state, fptr, indent=2, cls=_NumpyEncoder)

    def _read(self):
        with open(self._state_fname, mode="r") as fptr:
            self._saved_state = json.load(fptr, object_hook=_decode)


NEWTON_VARS = {  # newton_solver.py:62-118
    "iterate": ("model_state", "{method} of {name} Newton iterate", None),
    "fcn": ("model_state", "{method} of {name} Newton fcn", None),
    "increment": ("model_state", "{method} of {name} Newton increment", None),
    "increment_scalef": ("per_tracer_module", "factor applied to {name} Newton increment to satisfy bounds", "1"),
    "Armijo_factor": ("per_tracer_module", "factor applied to {name} Newton increment to satisfy Armijo condition", "1"),
    "Krylov_iterations": ("tracer_module_independent", "number of iterations in Krylov solver", "1"),
}
KRYLOV_VARS = {  # krylov_solver.py:50-73
    "precond_rhs_norm": ("per_tracer_module_invariant", "norm of {name} preconditioned rhs", None),
    "precond_resid_norm": ("per_tracer_module", "norm of {name} preconditioned residual", None),
}


class StatsFile:
    """`<name>_stats.nc` of a solver (stats_file.py:13-139)"""

    def __init__(self, name, workdir, region_cnt, tracer_modules, var_table, resume=False):
        """tracer_modules: list of (module name, units or None)"""
        self._name = name
        self._fname = os.path.join(workdir, f"{name}_stats.nc")
        self._region_cnt = region_cnt
        self._dimlen = {"region": region_cnt}  # fixed dimensions (the unlimited one is "iteration")
        self._vars = {}  # varname -> dict(dims, dtype, attrs, data)
        self._n_iter = 0
        self._history = (f"{datetime.now():%Y-%m-%d %H:%M:%S}: created by StatsFile._create_stats_file "
                         f"for {name} solver")
        self._keys = {}  # key -> (category, [varnames] or {"mean": [...], "norm": [...]})
        for key, (category, long_name, units) in var_table.items():
            if category == "model_state":
                names = {"mean": [], "norm": []}
                for method in ("mean", "norm"):
                    for mod, mod_units in tracer_modules:
                        vname = f"{key}_{method}_{mod}"
                        self._define(vname, ("iteration", "region"), "f8",
                                     long_name.format(method=method, name=mod), mod_units)
                        names[method].append(vname)
                self._keys[key] = (category, names)
            elif category in ("per_tracer_module", "per_tracer_module_invariant"):
                names = []
                for mod, mod_units in tracer_modules:
                    vname = f"{key}_{mod}"
                    dims = ("iteration", "region") if category == "per_tracer_module" else ("region",)
                    self._define(vname, dims, "f8", long_name.format(name=mod), units if units else mod_units)
                    names.append(vname)
                self._keys[key] = (category, names)
            else:
                self._define(key, ("iteration",), "i4", long_name, units)
                self._keys[key] = (category, [key])
        if resume and os.path.exists(self._fname):
            self._load()
        else:
            self._flush()

    def _define(self, vname, dims, dtype, long_name, units):
        attrs = {"long_name": long_name}
        if units is not None:
            attrs["units"] = units
        if "iteration" in dims:
            attrs["_FillValue"] = FILL_F8 if dtype == "f8" else FILL_I4
        shape = tuple(0 if d == "iteration" else self._dimlen[d] for d in dims)
        self._vars[vname] = {"dims": dims, "dtype": dtype, "attrs": attrs,
                             "data": np.zeros(shape, dtype=np.float64 if dtype == "f8" else np.int32)}

    def _grow(self, n_iter):
        for var in self._vars.values():
            if "iteration" not in var["dims"]:
                continue
            cur = var["data"]
            if cur.shape[0] >= n_iter:
                continue
            pad = np.full((n_iter - cur.shape[0],) + cur.shape[1:], var["attrs"]["_FillValue"], dtype=cur.dtype)
            var["data"] = np.concatenate([cur, pad], axis=0)
        self._n_iter = max(self._n_iter, n_iter)

    def put(self, iteration, **kwargs):
        """values of one iteration: key -> ModelState (model_state category: its mean and norm are
        written), ndarray [n_modules, region_cnt] (per_tracer_module) or int"""
        self._grow(iteration + 1)
        for key, vals in kwargs.items():
            category, names = self._keys[key]
            if category == "model_state":
                for method in ("mean", "norm"):
                    red = vals.mean() if method == "mean" else vals.norm()
                    for ind, vname in enumerate(names[method]):
                        self._vars[vname]["data"][iteration] = np.asarray(red[ind]).reshape(-1)[: self._region_cnt]
            elif category == "per_tracer_module":
                for ind, vname in enumerate(names):
                    self._vars[vname]["data"][iteration] = vals[ind]
            elif category == "per_tracer_module_invariant":
                raise ValueError(f"{key} has no iteration dimension: use put_invariant")
            else:
                self._vars[names[0]]["data"][iteration] = vals
        self._flush()

    def put_hist_stats(self, iteration, hist_fname, names, mean_weights=None):
        """the model's statistics of one iteration's hist file: time mean of the tracer-like variables
        `names` (end points of the record down-weighted: the hist file holds t = 0 and t = T,
        tracer_module_state.py:206-214 / 156-164) and, for each axis in `mean_weights`
        ({axis name: cell widths}), the weighted mean along it (`<name>_mean_<axis>`).  Dimensions,
        coordinate variables and metadata are taken from the hist file at the first call."""
        self._grow(iteration + 1)
        with netcdf_file(hist_fname, "r", mmap=False) as fptr:
            timelen = fptr.dimensions["time"] or fptr.variables["time"].shape[0]
            weights = np.full(timelen, 1.0 / (timelen - 1))
            weights[0] *= 0.5
            weights[-1] *= 0.5
            for name in names:
                if name not in fptr.variables:
                    continue
                var = fptr.variables[name]
                dims = tuple(var.dimensions[1:])
                for dim, length in zip(dims, var.shape[1:]):
                    if dim not in self._dimlen:
                        self._dimlen[dim] = int(length)
                        if dim in fptr.variables:  # coordinate variable, iteration invariant
                            cvar = fptr.variables[dim]
                            attrs = {k: (v.decode() if isinstance(v, bytes) else v) for k, v in cvar._attributes.items()}
                            self._vars[dim] = {"dims": (dim,), "dtype": "f8", "attrs": attrs,
                                               "data": np.array(cvar.data, dtype=np.float64)}
                attrs = {k: (v.decode() if isinstance(v, bytes) else v) for k, v in var._attributes.items()
                         if k not in ("cell_methods", "_FillValue")}
                mean = np.einsum("i,i...", weights, np.array(var.data, dtype=np.float64))
                targets = [(name, dims, mean)]
                for axis, widths in (mean_weights or {}).items():
                    if axis in dims:
                        w = np.asarray(widths, dtype=np.float64) / np.sum(widths)
                        pos = dims.index(axis)
                        targets.append((f"{name}_mean_{axis}", dims[:pos] + dims[pos + 1:],
                                        np.tensordot(mean, w, axes=([pos], [0]))))
                for vname, vdims, vals in targets:
                    if vname not in self._vars:
                        self._vars[vname] = {"dims": ("iteration",) + vdims, "dtype": "f8",
                                             "attrs": dict(attrs, _FillValue=FILL_F8),
                                             "data": np.full((self._n_iter,) + tuple(self._dimlen[d] for d in vdims),
                                                             FILL_F8)}
                    self._vars[vname]["data"][iteration] = vals
        self._flush()

    def def_hist_stats(self, hist_fname, names, mean_weights=None):
        """define the variables put_hist_stats will fill (model_state_base.py:134-151): dimensions and metadata
        from the hist file; values are fill values until an iteration puts them"""
        with netcdf_file(hist_fname, "r", mmap=False) as fptr:
            for name in names:
                if name not in fptr.variables:
                    continue
                var = fptr.variables[name]
                dims = tuple(var.dimensions[1:])
                for dim, length in zip(dims, var.shape[1:]):
                    self._dimlen.setdefault(dim, int(length))
                attrs = {k: (v.decode() if isinstance(v, bytes) else v) for k, v in var._attributes.items()
                         if k not in ("cell_methods", "_FillValue")}
                targets = [(name, dims)]
                for axis in (mean_weights or {}):
                    if axis in dims:
                        pos = dims.index(axis)
                        targets.append((f"{name}_mean_{axis}", dims[:pos] + dims[pos + 1:]))
                for vname, vdims in targets:
                    if vname not in self._vars:
                        self._vars[vname] = {"dims": ("iteration",) + vdims, "dtype": "f8",
                                             "attrs": dict(attrs, _FillValue=FILL_F8),
                                             "data": np.full((self._n_iter,) + tuple(self._dimlen[d] for d in vdims),
                                                             FILL_F8)}
        self._flush()

    def put_hist_coordinates(self, hist_fname):
        """iteration-invariant coordinate variables of the dimensions in use (model_state_base.py:153-167)"""
        with netcdf_file(hist_fname, "r", mmap=False) as fptr:
            for dim in list(self._dimlen):
                if dim in fptr.variables and dim not in self._vars and dim != "region":
                    cvar = fptr.variables[dim]
                    attrs = {k: (v.decode() if isinstance(v, bytes) else v) for k, v in cvar._attributes.items()}
                    self._vars[dim] = {"dims": (dim,), "dtype": "f8", "attrs": attrs,
                                       "data": np.array(cvar.data, dtype=np.float64)}
        self._flush()

    def put_invariant(self, **kwargs):
        for key, vals in kwargs.items():
            category, names = self._keys[key]
            if category != "per_tracer_module_invariant":
                raise RuntimeError(f"iteration is a dimension for {key}")
            for ind, vname in enumerate(names):
                self._vars[vname]["data"][:] = vals[ind]
        self._flush()

    def _flush(self):
        with netcdf_file(self._fname, "w", version=2) as fptr:
            fptr.history = self._history
            fptr.createDimension("iteration", None)
            for dim, length in self._dimlen.items():
                fptr.createDimension(dim, length)
            it = fptr.createVariable("iteration", "i4", ("iteration",))
            it.long_name = f"{self._name} solver iteration"
            reg = fptr.createVariable("region", "i4", ("region",))
            reg.long_name = "region index (0-based)"
            reg.comment = "axis attribute is a work-around to enable pyferret to read stats files"
            reg.axis = "T"
            reg[:] = np.arange(self._region_cnt, dtype=np.int32)
            handles = {}
            for vname, var in self._vars.items():
                handle = fptr.createVariable(vname, var["dtype"], var["dims"])
                for att, val in var["attrs"].items():
                    setattr(handle, att, val)
                handles[vname] = handle
            if self._n_iter > 0:
                it[: self._n_iter] = np.arange(self._n_iter, dtype=np.int32)
            for vname, var in self._vars.items():
                if "iteration" in var["dims"]:
                    if self._n_iter > 0:
                        handles[vname][: self._n_iter] = var["data"][: self._n_iter]
                else:
                    handles[vname][:] = var["data"]

    def _load(self):
        with netcdf_file(self._fname, "r", mmap=False) as fptr:
            self._history = fptr.history.decode() if isinstance(fptr.history, bytes) else str(fptr.history)
            self._n_iter = int(fptr.variables["iteration"].shape[0])
            for dim, length in fptr.dimensions.items():
                if dim != "iteration" and dim not in self._dimlen:
                    self._dimlen[dim] = int(length)
            for vname, fvar in fptr.variables.items():
                if vname in ("iteration", "region"):
                    continue
                if vname in self._vars:
                    var = self._vars[vname]
                    var["data"] = np.array(fvar.data, dtype=var["data"].dtype)
                else:  # statistics of the hist files, defined on the fly by put_hist_stats
                    attrs = {k: (v.decode() if isinstance(v, bytes) else v) for k, v in fvar._attributes.items()}
                    self._vars[vname] = {"dims": tuple(fvar.dimensions), "dtype": "f8", "attrs": attrs,
                                         "data": np.array(fvar.data, dtype=np.float64)}
        self._grow(self._n_iter)


class RefStatsFileAdapter:
    """Lets the model-state stats hooks (ModelStateBase.def_stats_vars / put_stats_vars_iteration_invariant /
    put_stats_vars) write into the REFERENCE's own StatsFile (nk_ooc/stats_file.py:13-139: def_dimensions, def_vars,
    put_vars_iteration_invariant, put_vars) when this package's ModelState runs under the reference's NewtonSolver
    (newton_solver.py:52-58,330 hands its own stats file to those hooks).  Same variables as StatsFile's methods
    of the same names: per tracer-like hist variable its time mean with down-weighted end points
    (tracer_module_state.py:325-341) and the weighted mean along the axes in `mean_weights`, plus the
    coordinate variables of the axes."""

    def __init__(self, ref_stats_file):
        self._ref = ref_stats_file

    @staticmethod
    def _attrs(var, drop=()):
        return {k: (v.decode() if isinstance(v, bytes) else v) for k, v in var._attributes.items() if k not in drop}

    def _targets(self, fptr, names, mean_weights):
        for name in names:
            if name not in fptr.variables:
                continue
            var = fptr.variables[name]
            dims = tuple(var.dimensions[1:])
            yield name, name, dims, None
            for axis in (mean_weights or {}):
                if axis in dims:
                    pos = dims.index(axis)
                    yield name, f"{name}_mean_{axis}", dims[:pos] + dims[pos + 1:], (axis, pos)

    def def_hist_stats(self, hist_fname, names, mean_weights=None):
        dimensions, vars_metadata = {}, {}
        with netcdf_file(hist_fname, "r", mmap=False) as fptr:
            for name, vname, dims, _ in self._targets(fptr, names, mean_weights):
                var = fptr.variables[name]
                for dim, length in zip(var.dimensions[1:], var.shape[1:]):
                    if dim not in dimensions:
                        dimensions[dim] = int(length)
                        if dim in fptr.variables:
                            vars_metadata[dim] = {"datatype": "f8", "dimensions": (dim,),
                                                  "attrs": self._attrs(fptr.variables[dim])}
                vars_metadata[vname] = {"datatype": "f8", "dimensions": ("iteration",) + dims,
                                        "attrs": self._attrs(var, drop=("cell_methods", "_FillValue"))}
        self._ref.def_dimensions(dimensions)
        self._ref.def_vars(vars_metadata)

    def put_hist_coordinates(self, hist_fname):
        with netcdf_file(hist_fname, "r", mmap=False) as fptr:
            vals = {dim: np.array(fptr.variables[dim].data, dtype=np.float64) for dim in fptr.dimensions
                    if dim in fptr.variables and dim != "time"}
        self._ref.put_vars_iteration_invariant(vals)

    def put_hist_stats(self, iteration, hist_fname, names, mean_weights=None):
        vals = {}
        with netcdf_file(hist_fname, "r", mmap=False) as fptr:
            timelen = fptr.dimensions["time"] or fptr.variables["time"].shape[0]
            weights = np.full(timelen, 1.0 / (timelen - 1))
            weights[0] *= 0.5
            weights[-1] *= 0.5
            means = {}
            for name, vname, _, axis in self._targets(fptr, names, mean_weights):
                if name not in means:
                    means[name] = np.einsum("i,i...", weights, np.array(fptr.variables[name].data, dtype=np.float64))
                if axis is None:
                    vals[vname] = means[name]
                else:
                    w = np.asarray(mean_weights[axis[0]], dtype=np.float64)
                    vals[vname] = np.tensordot(means[name], w / w.sum(), axes=([axis[1]], [0]))
        self._ref.put_vars(iteration, vals)


def as_hist_stats(stats_file):
    """the object the model-state stats hooks write through: a StatsFile of this package as it is, the
    reference's StatsFile behind RefStatsFileAdapter"""
    return stats_file if hasattr(stats_file, "def_hist_stats") else RefStatsFileAdapter(stats_file)
