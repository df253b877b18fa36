"""Host mirror of the reference's operator surface (nk_ooc/model_state_base.py,
nk_ooc/tracer_module_state_base.py, nk_ooc/model_config.py) over the CUDA library.

Same method names, argument meaning and error behaviour as the reference for the hot-path
operators (comp_fcn, apply_precond_jacobian, comp_jacobian_fcn_state_prod, dot_prod, norm,
mean, mod_gram_schmidt, lin_comb, arithmetic with per-(module, region) scalars, dump).  The
values of a state live in HBM, one float64 tensor [tracer, depth(, ypos), member] per tracer
module with the member index fastest; a state may carry B >= 1 independent members
(B == 1 is the reference's single state).  Scalars are then arrays [n_modules, region_cnt]
(B == 1, as in the reference) or [n_modules, region_cnt, B].

State files are NETCDF3_64BIT_OFFSET with the reference's layout,
read/written with scipy.io.netcdf_file; with B > 1 member 0 is written.
"""

import copy
import logging
import os
from datetime import datetime

import numpy as np
import torch
from scipy.io import netcdf_file

from . import engine
from .solver_state import as_hist_stats


def _expand_defs(defs, names):
    """tracer-module defs with `{suff}` templating: a name `root:suff1:suff2...` gives one module
    per suffix (model_config.py:80-125)"""

    def subst(obj, suff):
        if isinstance(obj, str):
            return obj.replace("{suff}", suff)
        if isinstance(obj, dict):
            return {subst(k, suff): subst(v, suff) for k, v in obj.items()}
        if isinstance(obj, list):
            return [subst(v, suff) for v in obj]
        return obj

    out = {}
    for full in names:
        root, _, suffs = full.partition(":")
        if root not in defs:
            raise ValueError(f"unknown tracer module name {root}")
        if not suffs:
            out[root] = copy.deepcopy(defs[root])
            continue
        for suff in suffs.split(":"):
            out[subst(root, suff)] = subst(copy.deepcopy(defs[root]), suff)
    return out



def _expand_matrix_defs(defs, names):
    """precond matrix definitions with `{suff}` templating, expanded with the suffixes of the tracer modules in use
    (input/py_driver_2d/tracer_module_defs.yaml:60-62: `forced_{suff}` -> `forced_o2_like`)"""
    suffs = [suff for full in names for suff in full.split(":")[1:]]
    out = {}
    for name, mdef in defs.items():
        if "{suff}" not in name:
            out[name] = copy.deepcopy(mdef)
            continue
        for suff in suffs:
            text = lambda obj: obj.replace("{suff}", suff) if isinstance(obj, str) else obj  # noqa: E731
            out[text(name)] = {key: [text(v) for v in val] if isinstance(val, list) else copy.deepcopy(val)
                               for key, val in mdef.items()}
    return out


def copy_hist_attrs(src, dst, long_name_suffix=None, drop_time_cell_methods=False):
    """attributes of a hist-file variable carried over to the precond file (model_state_base.py:449-470): all of
    them, `cell_methods` dropped when it refers to a time dimension the result does not have, the long name
    extended by the time reduction"""
    for key, val in src._attributes.items():  # pylint: disable=protected-access
        if isinstance(val, bytes):
            val = val.decode()
        if key == "cell_methods" and drop_time_cell_methods and "time:" in val:
            continue
        if key == "long_name" and long_name_suffix:
            val = val + long_name_suffix
        setattr(dst, key, val)

def propagate_base_matrix_defs_to_all(matrix_defs):
    """the settings of the matrix definition "base" become defaults of every other one, in place
    (model_config.py:197-228, tests/test_model_config.py:25-57): a key the matrix lacks is taken over as a deep copy;
    list settings are merged by OPTION NAME (the first word of an entry), so an option the matrix sets itself keeps its
    own sub-option and nothing is added twice; dict settings gain the base's missing keys"""
    base = matrix_defs.get("base")
    if base is None:
        return
    for name, mdef in matrix_defs.items():
        if name == "base":
            continue
        for key, dflt in base.items():
            if key not in mdef:
                mdef[key] = copy.deepcopy(dflt)
            elif isinstance(dflt, list):
                have = {entry.split()[0] for entry in mdef[key]}
                mdef[key].extend(entry for entry in dflt if entry.split()[0] not in have)
            elif isinstance(dflt, dict):
                for sub, val in dflt.items():
                    mdef[key].setdefault(sub, val)
            else:
                raise TypeError(f"base defn type {type(dflt)} not supported")


def check_precond_matrix_defs(matrix_defs):
    """hist_to_precond_varnames entries are `name`, `name:mean` or `name:log_mean` (model_config.py:231-246)"""
    for name, mdef in matrix_defs.items():
        for hist_var in mdef.get("hist_to_precond_varnames", []):
            time_op = hist_var.partition(":")[2]
            if time_op not in ("", "mean", "log_mean"):
                raise ValueError(f"unknown time_op={time_op} in {hist_var} from {name}")


class ModelConfig:
    """modelinfo + tracer module definitions + region weights (nk_ooc/model_config.py:17-78,
    249-315).  tracer_module_defs: dict as in input/<model>/tracer_module_defs.yaml; precond_matrix_defs: the
    section of that name of the same file (None: the definitions of the modules this package ships)."""

    def __init__(self, modelinfo, tracer_module_defs, grid_vars=None, precond_matrix_defs=None):
        self.modelinfo = modelinfo
        names = modelinfo["tracer_module_names"].split(",")
        self.tracer_module_defs = _expand_defs(tracer_module_defs, names)
        # expanded names in the order given ("forced_{suff}:o2_like" -> "forced_o2_like")
        self.tracer_module_names = list(self.tracer_module_defs)
        if grid_vars is None:
            grid_vars = read_grid_vars(modelinfo["grid_vars_fname"], "region_mask")
        self.region_mask = grid_vars["region_mask"]
        self.grid_weight = grid_vars["grid_weight"]
        self.weights = engine.RegionWeights(self.region_mask, self.grid_weight)
        self.region_cnt = self.weights.region_cnt
        if precond_matrix_defs is None:
            self.precond_matrix_defs = self._precond_matrix_defs()
        else:
            self.precond_matrix_defs = _expand_matrix_defs(precond_matrix_defs, names)
        propagate_base_matrix_defs_to_all(self.precond_matrix_defs)
        check_precond_matrix_defs(self.precond_matrix_defs)

    def _precond_matrix_defs(self):
        """{matrix name: {"hist_to_precond_varnames": [...]}} (input/<model>/tracer_module_defs.yaml
        precond_matrix_defs; model_config.py:197-228: the settings of "base" are propagated to all matrices).
        The models of this package use: base = time (py_driver_2d) or the two mixing-coefficient reductions
        (test_problem); one matrix per tracer that names one, fed by that tracer's hist variable."""
        column = self.modelinfo.get("model_name", "") == "test_problem"
        base = ["mixing_coeff:mean", "mixing_coeff:log_mean"] if column else ["time"]
        defs = {"base": {"hist_to_precond_varnames": list(base)}}
        for mdef in self.tracer_module_defs.values():
            for tname, meta in mdef["tracers"].items():
                pm = meta.get("precond_matrix")
                if pm is None or pm in defs:
                    continue
                own = ["po4_s_restore_tau_r:mean"] if (column and pm == "phosphorus") else [tname]
                defs[pm] = {"hist_to_precond_varnames": list(own)}
        return defs


def read_grid_vars(fname, region_mask_varname):
    """region_mask and the weight variable named by its cell_measures attribute
    (nk_ooc/model_config.py:249-289)"""
    with netcdf_file(fname, "r", mmap=False) as fptr:
        var = fptr.variables[region_mask_varname]
        mask = np.array(var.data, dtype=np.int32)
        cell_measures = var.cell_measures
        if isinstance(cell_measures, bytes):
            cell_measures = cell_measures.decode()
        parts = cell_measures.split(":")
        if len(parts) != 2:
            raise RuntimeError(f"unexpected number of words in {region_mask_varname}:cell_measures")
        wname = parts[-1].split()[0]
        weight = np.array(fptr.variables[wname].data, dtype=np.float64)
    return {"region_mask": mask, "grid_weight": weight}


class TracerModuleStateBase:
    """values of one tracer module for B members (nk_ooc/tracer_module_state_base.py:13-515)"""

    def __init__(self, name, tracer_module_def, cell_shape, config, vals=None, members=1):
        self.name = name
        self._def = tracer_module_def
        self.tracer_names = list(tracer_module_def["tracers"])
        self.tracer_cnt = len(self.tracer_names)
        self.cell_shape = tuple(cell_shape)
        self.config = config
        self.members = members
        ldb = engine.padded_members(members)
        if vals is None:
            vals = torch.zeros((self.tracer_cnt,) + self.cell_shape + (ldb,), dtype=torch.float64, device="cuda")
        self.vals = vals

    # ---- host <-> device --------------------------------------------------------------
    def tracer_index(self, tracer_name):
        try:
            return self.tracer_names.index(tracer_name)
        except ValueError:
            raise KeyError(f"unknown tracer_name={tracer_name}") from None

    def get_tracer_vals(self, tracer_name, member=0):
        return self.vals[self.tracer_index(tracer_name), ..., member].cpu().numpy()

    def set_tracer_vals(self, tracer_name, vals, member=None):
        t = torch.as_tensor(np.ascontiguousarray(vals), dtype=torch.float64).cuda()
        if member is None:
            self.vals[self.tracer_index(tracer_name), ..., : self.members] = t.unsqueeze(-1)
        else:
            self.vals[self.tracer_index(tracer_name), ..., member] = t

    def get_tracer_vals_all(self, member=0):
        return self.vals[..., member].cpu().numpy()

    def set_tracer_vals_all(self, vals, member=None):
        t = torch.as_tensor(np.ascontiguousarray(vals), dtype=torch.float64).cuda()
        if member is None:
            self.vals[..., : self.members] = t.unsqueeze(-1)
        else:
            self.vals[..., member] = t

    def clone(self):
        res = copy.copy(self)
        res.vals = self.vals.clone()
        return res

    # ---- reductions (K5) --------------------------------------------------------------
    def _flat(self, t):
        return t.reshape(self.tracer_cnt, -1, t.shape[-1])

    def dot_prod(self, other):
        """[region_cnt, B] (tracer_module_state_base.py:379-388)"""
        return self.config.weights.dot(self._flat(self.vals), self._flat(other.vals), self.members)

    def mean(self):
        """[region_cnt, B] (tracer_module_state_base.py:371-377)"""
        return self.config.weights.dot(self._flat(self.vals), None, self.members)

    # ---- elementwise with per-(region, member) scalars (K6) --------------------------------
    def axpby(self, alpha, x, beta):
        """self <- alpha*x + beta*self; alpha/beta float or device [region_cnt, B]"""
        xv = None if x is None else self._flat(x.vals)
        self.config.weights.axpby(alpha, xv, beta, self._flat(self.vals), self.members)
        return self

    # ---- limiter (tracer_module_state_base.py:115-176) ----------------------------------------
    def has_bounds(self):
        if "bounds" in self._def:
            return True
        return any("bounds" in meta for meta in self._def["tracers"].values())

    def get_bounds(self, tracer_name):
        lob, upb = None, None
        for metadata in (self._def, self._def["tracers"][tracer_name]):
            if "bounds" in metadata:
                lob = metadata["bounds"].get("lob", lob)
                upb = metadata["bounds"].get("upb", upb)
        return lob, upb

    def apply_limiter(self, base):
        """scale self so that base + scalef*self is within bounds; returns scalef [region_cnt, B]
        on the host (1.0 without bounds), as tracer_module_state_base.py:115-151"""
        R, B = self.config.region_cnt, self.members
        if not self.has_bounds():
            return np.ones((R, B))
        scalef = np.ones((R, B))
        for tname in self.tracer_names:
            lob, upb = self.get_bounds(tname)
            if lob is None and upb is None:
                continue
            t = self.tracer_index(tname)
            got = self.config.weights.limiter_scalef(self._flat(base.vals[t : t + 1]), self._flat(self.vals[t : t + 1]),
                                                     lob, upb, B).cpu().numpy()
            # regions without cells come back as +inf (min_by_region, utils.py:557); the running
            # minimum starts from ones (tracer_module_state_base.py:123-124), so they end at 1.0
            scalef = np.minimum(scalef, got)
        if (scalef < 1.0).any():
            self.axpby(None, None, torch.from_numpy(np.ascontiguousarray(scalef)).cuda())
        return scalef

    def apply_region_mask(self):
        """zero where region_mask == 0 (tracer_module_state_base.py:153-176)"""
        R, B = self.config.region_cnt, self.members
        ones = torch.ones((R, B), dtype=torch.float64, device="cuda")
        self.config.weights.axpby(None, None, ones, self._flat(self.vals), B, fill_beta=0.0)
        return self

    def zero_extra_tracers(self):
        """tracers not being solved for; none in the models of this package besides shadows
        handled by the model classes (tracer_module_state_base.py:483-500)"""
        return self

    def shadow_pairs(self):
        out = []
        for name, meta in self._def["tracers"].items():
            if "shadows" in meta:
                out.append((self.tracer_index(name), self.tracer_index(meta["shadows"])))
        return out

    def precond_matrix_list(self):
        res = []
        for meta in self._def["tracers"].values():
            pm = meta.get("precond_matrix")
            if pm is not None and pm not in res:
                res.append(pm)
        return res

    def append_tracer_names_per_precond_matrix(self, res):
        """{matrix name: [tracer names]} (tracer_module_state_base.py:86-98)"""
        for tname, meta in self._def["tracers"].items():
            pm = meta.get("precond_matrix")
            if pm is not None:
                res.setdefault(pm, []).append(tname)

    def log_vals(self, msg, vals):
        """per-tracer-module values to the log, in the reference's format (tracer_module_state_base.py:178-198)"""
        logger = logging.getLogger(__name__)
        vals = np.asarray(vals)
        if vals.ndim >= 1 and vals.shape[-1] == 1:
            self.log_vals(msg, vals[..., 0])
            return
        if vals.ndim == 0:
            logger.info("%s[%s]=%e", msg, self.name, vals)
        elif vals.ndim == 1:
            for j in range(vals.shape[0]):
                logger.info("%s[%s,%d]=%e", msg, self.name, j, vals[j])
        elif vals.ndim == 2:
            for i in range(vals.shape[0]):
                for j in range(vals.shape[1]):
                    logger.info("%s[%s,%d,%d]=%e", msg, self.name, i, j, vals[i, j])
        else:
            raise ValueError(f"vals.ndim={vals.ndim} not handled")

    # ---- statistics of a hist file (tracer_module_state.py of the models: stats_vars_*) -------------
    def stats_vars_tracer_like(self):
        """tracer-like hist variables that go into the stats file (tracer_module_state_base.py:100-104)"""
        return list(self.tracer_names)


class ModelStateBase:
    """state space of a model (nk_ooc/model_state_base.py:24-577)"""

    __array_priority__ = 100
    model_config_obj = None

    # ---- construction -----------------------------------------------------------------
    def __init__(self, fname, members=1):
        if self.model_config_obj is None:
            raise RuntimeError("self.model_config_obj is None, it should be set in derived class")
        engine.require_cuda()
        cfg = self.model_config_obj
        self.members = members
        self.tracer_modules = np.empty(len(cfg.tracer_module_names), dtype=object)
        for ind, name in enumerate(cfg.tracer_module_names):
            tms = self._new_tracer_module(name, cfg.tracer_module_defs[name], members)
            self._load(tms, fname)
            self.tracer_modules[ind] = tms

    def _new_tracer_module(self, name, tracer_module_def, members):
        raise NotImplementedError("Method must be implemented in derived class")

    def _gen_init_iterate(self, tms):
        raise NotImplementedError("Method must be implemented in derived class")

    def _load(self, tms, fname):
        if isinstance(fname, dict):  # {tracer_name: ndarray} (convenience, not in the reference)
            for tname in tms.tracer_names:
                tms.set_tracer_vals(tname, np.asarray(fname[tname]).reshape(tms.cell_shape))
            return
        if fname == "zeros":
            return
        if fname == "gen_init_iterate":
            self._gen_init_iterate(tms)
            return
        with netcdf_file(fname, "r", mmap=False) as fptr:
            for tname in tms.tracer_names:
                if tname not in fptr.variables:
                    raise KeyError(f"{tname} not found in {fname}")
                vals = np.array(fptr.variables[tname].data, dtype=np.float64)
                if vals.size != int(np.prod(tms.cell_shape)):
                    raise ValueError(f"unexpected dimension lengths for {tname} in {fname}")
                tms.set_tracer_vals(tname, vals.reshape(tms.cell_shape))

    def _like(self, clone_vals=True):
        res = copy.copy(self)
        res.tracer_modules = np.empty(len(self.tracer_modules), dtype=object)
        for i, tms in enumerate(self.tracer_modules):
            res.tracer_modules[i] = tms.clone() if clone_vals else copy.copy(tms)
        return res

    @classmethod
    def from_members(cls, states):
        """stack single-member states into one batched state (B = len(states))"""
        res = cls("zeros", members=len(states))
        for i, tms in enumerate(res.tracer_modules):
            for b, st in enumerate(states):
                tms.vals[..., b] = st.tracer_modules[i].vals[..., 0]
        return res

    def member(self, b):
        """single-member state holding member b"""
        res = type(self)("zeros", members=1)
        for i, tms in enumerate(res.tracer_modules):
            tms.vals[..., 0] = self.tracer_modules[i].vals[..., b]
        return res

    def member_slice(self, lo, hi):
        """batched state holding the members [lo, hi) (the shard of one rank, distributed.py)"""
        res = type(self)("zeros", members=hi - lo)
        for i, tms in enumerate(res.tracer_modules):
            tms.vals[..., : hi - lo] = self.tracer_modules[i].vals[..., lo:hi]
        return res

    # ---- file output --------------------------------------------------------------------
    def _axes(self):
        raise NotImplementedError("Method must be implemented in derived class")

    def dump(self, fname, caller=None):
        """write the state (member 0) in the reference's state-file layout
        (model_state_base.py:91-112; py_driver_2d/tracer_module_state.py:71-96)"""
        if fname is None:
            return self
        if caller is None:
            raise ValueError("caller unknown")
        os.makedirs(os.path.dirname(os.path.abspath(fname)), exist_ok=True)
        with netcdf_file(fname, "w", version=2) as fptr:
            stamp = datetime.now().strftime("%Y-%m-%d %H:%M:%S")
            fptr.history = f"{stamp}: created by {type(self).__module__}.{type(self).__name__}.dump called from {caller}"
            axes = self._axes()
            for axis in axes:
                axis.define(fptr)
            dims = tuple(axis.axisname for axis in axes)
            for tms in self.tracer_modules:
                for tname in tms.tracer_names:
                    fptr.createVariable(tname, "f8", dims)
            for axis in axes:
                axis.write(fptr)
            for tms in self.tracer_modules:
                for tname in tms.tracer_names:
                    fptr.variables[tname][:] = tms.get_tracer_vals(tname).reshape([len(a) for a in axes])
        return self

    def log_vals(self, msg, vals):
        """write per-tracer module values to the log (model_state_base.py:113-122)"""
        for ind, tms in enumerate(self.tracer_modules):
            if isinstance(msg, list):
                for msg_ind, submsg in enumerate(msg):
                    tms.log_vals(submsg, vals[msg_ind, ind, ...])
            else:
                tms.log_vals(msg, vals[ind, ...])

    def log(self, msg=None):
        """mean and norm of the instance to the log (model_state_base.py:124-132)"""
        msg_full = ["mean", "norm"] if msg is None else [f"{msg},mean", f"{msg},norm"]
        self.log_vals(msg_full, np.stack((self.mean(), self.norm())))

    # ---- model-specific statistics of an iteration's hist file (model_state_base.py:134-180) ----------
    def _stats_names_and_weights(self):
        names = []
        for tms in self.tracer_modules:
            names += tms.stats_vars_tracer_like()
        ypos = getattr(type(self), "ypos", None)
        weights = {ypos.axisname: ypos.delta} if ypos is not None and hasattr(ypos, "axisname") else None
        return names, weights

    def def_stats_vars(self, stats_file, hist_fname, solver_state):
        """define the model specific stats variables: dimensions, coordinate variables and one variable per
        tracer-like hist variable (+ its ypos mean), with the hist file's metadata"""
        step = "ModelStateBase.def_stats_vars"
        if solver_state is not None and solver_state.step_logged(step, per_iteration=False):
            return
        names, weights = self._stats_names_and_weights()
        as_hist_stats(stats_file).def_hist_stats(hist_fname, names, weights)
        if solver_state is not None:
            solver_state.log_step(step, per_iteration=False)

    def put_stats_vars_iteration_invariant(self, stats_file, hist_fname, solver_state):
        """values of the iteration-invariant stats variables (the coordinate variables of the axes)"""
        step = "ModelStateBase.put_stats_vars_iteration_invariant"
        if solver_state is not None and solver_state.step_logged(step, per_iteration=False):
            return
        as_hist_stats(stats_file).put_hist_coordinates(hist_fname)
        if solver_state is not None:
            solver_state.log_step(step, per_iteration=False)

    def put_stats_vars(self, stats_file, hist_fname, solver_state):
        """stats variables of the current iteration: time mean of every tracer-like hist variable with the end
        points of the record down-weighted, and its ypos mean"""
        step = "ModelStateBase.put_stats_vars"
        if solver_state is not None and solver_state.step_logged(step):
            return
        names, weights = self._stats_names_and_weights()
        iteration = solver_state.get_iteration() if solver_state is not None else 0
        as_hist_stats(stats_file).put_hist_stats(iteration, hist_fname, names, weights)
        if solver_state is not None:
            solver_state.log_step(step)

    def gen_precond_jacobian(self, hist_fname, precond_fname, solver_state=None):
        """hist file -> precond file (model_state_base.py:404-481, 583-615): for every entry `name[:mean|:log_mean]` of
        hist_vars_for_precond_list(), in that order, the hist variable reduced over time as the entry says; first all
        dimensions the results have with their coordinate variables, then the results.  Layout (order of dimensions
        and variables, names, attributes) as the reference writes it — the CI flows compare these files with the
        baselines' metadata"""
        os.makedirs(os.path.dirname(os.path.abspath(precond_fname)), exist_ok=True)
        entries = [entry.partition(":")[::2] for entry in self.hist_vars_for_precond_list()]
        with netcdf_file(hist_fname, "r", mmap=False) as fin, netcdf_file(precond_fname, "w", version=2) as fout:
            stamp = datetime.now().strftime("%Y-%m-%d %H:%M:%S")
            msg = f"{stamp}: created by {type(self).__name__}.gen_precond_jacobian"
            prior = getattr(fin, "history", None)
            fout.history = msg if prior is None else msg + "\n" + (prior.decode() if isinstance(prior, bytes) else prior)

            def result_dims(name, time_op):
                var = fin.variables[name]
                dims = list(zip(var.dimensions, var.shape))
                if time_op in ("mean", "log_mean"):
                    dims = [d for d in dims if d[0] != "time"]
                return [d for d in dims if not (d[0] == "time" and d[1] == 1)]

            coords = []
            for name, time_op in entries:
                for dim, length in result_dims(name, time_op):
                    if dim not in fout.dimensions:
                        fout.createDimension(dim, length)
                    elif fout.dimensions[dim] != length:
                        raise ValueError(f"dimension {dim} of {name} has length {length}, not {fout.dimensions[dim]}")
                    if dim in fin.variables and dim not in coords:
                        coords.append(dim)
            for dim in coords:
                src = fin.variables[dim]
                dst = fout.createVariable(dim, src.data.dtype.newbyteorder("="), (dim,))
                copy_hist_attrs(src, dst)
                dst[:] = np.array(src.data)
            for name, time_op in entries:
                if name in fout.dimensions:
                    continue  # a coordinate variable: written above
                src = fin.variables[name]
                dims = tuple(d[0] for d in result_dims(name, time_op))
                vals = np.array(src.data, dtype=np.float64)
                if time_op == "mean":
                    out_name, suffix, vals = f"{name}_mean", ", mean over time dim", vals.mean(axis=0)
                elif time_op == "log_mean":
                    out_name, suffix, vals = f"{name}_log_mean", ", log mean over time dim", np.exp(np.log(vals).mean(axis=0))
                else:
                    out_name, suffix = name, None
                    vals = vals.reshape([fout.dimensions[d] for d in dims])
                dst = fout.createVariable(out_name, "f8", dims)
                copy_hist_attrs(src, dst, suffix, drop_time_cell_methods="time" not in dims)
                dst[:] = vals

    def hist_vars_for_precond_list(self):
        """hist variables needed for the preconditioner (model_state_base.py:379-388)"""
        res = []
        defs = self.model_config_obj.precond_matrix_defs
        for matrix_name in self.precond_matrix_list() + ["base"]:
            for varname in defs[matrix_name]["hist_to_precond_varnames"]:
                if varname not in res:
                    res.append(varname)
        return res

    def tracer_names_per_precond_matrix(self):
        """{matrix name: [tracer names]} (model_state_base.py:397-402)"""
        res = {}
        for tms in self.tracer_modules:
            tms.append_tracer_names_per_precond_matrix(res)
        return res

    # ---- scalars ------------------------------------------------------------------------
    def _scalar_shape(self):
        n, R = len(self.tracer_modules), self.model_config_obj.region_cnt
        return (n, R) if self.members == 1 else (n, R, self.members)

    def _to_host(self, per_module):
        """list of device [R, B] -> ndarray [n_modules, R] (B == 1) or [n_modules, R, B]"""
        arr = torch.stack(per_module).cpu().numpy()
        return arr[..., 0] if self.members == 1 else arr

    def _module_scalars(self, arr, ind):
        """host scalars (float | [n] | [n, R] | [n, R, B]) -> float or device [R, B] for module ind"""
        if isinstance(arr, (int, float)):
            return float(arr)
        arr = np.asarray(arr, dtype=np.float64)
        n, R, B = len(self.tracer_modules), self.model_config_obj.region_cnt, self.members
        if arr.shape == (n,):
            sel = np.full((R, B), arr[ind])
        elif arr.shape == (n, R):
            sel = np.repeat(arr[ind][:, None], B, axis=1)
        elif arr.shape == (n, R, B):
            sel = arr[ind]
        else:
            raise ValueError(f"unsupported scalar shape {arr.shape}")
        return torch.from_numpy(np.ascontiguousarray(sel)).cuda()

    def mean(self):
        return self._to_host([tms.mean() for tms in self.tracer_modules])

    def dot_prod(self, other):
        return self._to_host([tms.dot_prod(o) for tms, o in zip(self.tracer_modules, other.tracer_modules)])

    def norm(self):
        return np.sqrt(self.dot_prod(self))

    # ---- arithmetic (model_state_base.py:180-345) -------------------------------------------
    def __neg__(self):
        res = self._like()
        for tms in res.tracer_modules:
            tms.axpby(None, None, -1.0)
        return res

    def _iop_state(self, other, sign):
        for tms, o in zip(self.tracer_modules, other.tracer_modules):
            tms.axpby(sign, o, 1.0)
        return self

    def __iadd__(self, other):
        if isinstance(other, ModelStateBase):
            return self._iop_state(other, 1.0)
        return NotImplemented

    def __isub__(self, other):
        if isinstance(other, ModelStateBase):
            return self._iop_state(other, -1.0)
        return NotImplemented

    def __add__(self, other):
        if isinstance(other, ModelStateBase):
            return self._like().__iadd__(other)
        return NotImplemented

    def __radd__(self, other):
        """res = other + self (model_state_base.py:201-206)"""
        return self + other

    def __sub__(self, other):
        if isinstance(other, ModelStateBase):
            return self._like().__isub__(other)
        return NotImplemented

    def _scale(self, other, reciprocal):
        if isinstance(other, ModelStateBase):
            return NotImplemented
        if not isinstance(other, (int, float, np.ndarray)):
            return NotImplemented
        if reciprocal:
            other = 1.0 / np.asarray(other, dtype=np.float64) if not isinstance(other, (int, float)) else 1.0 / other
        for ind, tms in enumerate(self.tracer_modules):
            tms.axpby(None, None, self._module_scalars(other, ind))
        return self

    def __imul__(self, other):
        return self._scale(other, False)

    def __itruediv__(self, other):
        return self._scale(other, True)

    def __mul__(self, other):
        if isinstance(other, ModelStateBase):
            return NotImplemented
        return self._like().__imul__(other)

    def __rmul__(self, other):
        return self * other

    def __truediv__(self, other):
        if isinstance(other, ModelStateBase):
            return NotImplemented
        return self._like().__itruediv__(other)

    def __rtruediv__(self, other):
        """res = other / self, other a float or ndarray [n_modules(, region_cnt)] (model_state_base.py:310-328).
        Not used by the solvers; the elementwise reciprocal is a torch op (cold path), the scaling the library's."""
        if not isinstance(other, (int, float, np.ndarray)):
            return NotImplemented
        res = self._like()
        for tms in res.tracer_modules:
            tms.vals = torch.reciprocal(tms.vals)
        return res.__imul__(other).apply_region_mask()

    def apply_limiter(self, base):
        """scale self so that base + scalef*self is within bounds (model_state_base.py:76-84);
        returns scalef [n_modules, region_cnt(, B)]"""
        per = [tms.apply_limiter(b) for tms, b in zip(self.tracer_modules, base.tracer_modules)]
        arr = np.stack(per)
        return arr[..., 0] if self.members == 1 else arr

    # ---- Krylov building blocks -----------------------------------------------------------
    def mod_gram_schmidt(self, basis_cnt, fname_fcn, quantity):
        """in-place modified Gram-Schmidt against basis files (model_state_base.py:365-377).
        fname_fcn may also return an in-memory ModelState (HBM-resident basis).  Per tracer module ONE
        library call (nkb_mgs: w resident on the chip, every basis vector read once) and, at the end, one
        device-to-host copy of all scalars."""
        basis = []
        for i_val in range(basis_cnt):
            basis_i = fname_fcn(quantity, i_val)
            basis.append(basis_i if isinstance(basis_i, ModelStateBase) else type(self)(basis_i))
        per_module = []
        for ind, tms in enumerate(self.tracer_modules):
            vecs = [tms._flat(b.tracer_modules[ind].vals) for b in basis]
            per_module.append(self.model_config_obj.weights.mgs(tms._flat(tms.vals), vecs, self.members))
        h_val = torch.stack(per_module).cpu().numpy()  # [n_modules, k, R, B]
        return h_val[..., 0] if self.members == 1 else h_val

    def comp_jacobian_fcn_state_prod(self, fcn, direction, res_fname, solver_state):
        """finite-difference Jacobian-vector product (model_state_base.py:492-527)"""
        step = f"comp_jacobian_fcn_state_prod complete for {res_fname}"
        if solver_state is not None and solver_state.step_logged(step):
            return type(self)(res_fname)
        sigma = 1.0e-4 * self.norm()
        sigma = np.where(sigma == 0.0, 1.0, sigma)
        perturb_ms = self + sigma * direction
        perturb_fname = None
        if res_fname is not None:
            workdir = solver_state.get_workdir() if solver_state is not None else os.path.dirname(res_fname)
            perturb_fname = os.path.join(workdir, f"perturb_fcn_{os.path.basename(res_fname)}")
        perturb_fcn = perturb_ms.comp_fcn(perturb_fname, solver_state)
        caller = f"{type(self).__name__}.comp_jacobian_fcn_state_prod"
        res = ((perturb_fcn - fcn) / sigma).dump(res_fname, caller)
        if solver_state is not None:
            solver_state.log_step(step)
        return res

    def comp_fcn_postprocess(self, res_fname, caller):
        """model_state_base.py:483-490"""
        return self.zero_extra_tracers().apply_region_mask().dump(res_fname, f"comp_fcn_postprocess called from {caller}")

    def get_tracer_vals(self, tracer_name):
        for tms in self.tracer_modules:
            try:
                return tms.get_tracer_vals(tracer_name)
            except KeyError:
                pass
        raise KeyError(f"unknown tracer_name={tracer_name}")

    def set_tracer_vals(self, tracer_name, vals):
        for tms in self.tracer_modules:
            try:
                tms.set_tracer_vals(tracer_name, vals)
                return
            except KeyError:
                pass
        raise KeyError(f"unknown tracer_name={tracer_name}")

    def shadow_tracers_on(self):
        return any(tms.shadow_pairs() for tms in self.tracer_modules)

    def copy_shadow_tracers_to_real_tracers(self):
        for tms in self.tracer_modules:
            for shadow, real in tms.shadow_pairs():
                tms.vals[real] = tms.vals[shadow]
        return self

    def copy_real_tracers_to_shadow_tracers(self):
        for tms in self.tracer_modules:
            for shadow, real in tms.shadow_pairs():
                tms.vals[shadow] = tms.vals[real]
        return self

    def zero_extra_tracers(self):
        for tms in self.tracer_modules:
            tms.zero_extra_tracers()
        return self

    def apply_region_mask(self):
        for tms in self.tracer_modules:
            tms.apply_region_mask()
        return self

    def precond_matrix_list(self):
        res = []
        for tms in self.tracer_modules:
            res.extend(tms.precond_matrix_list())
        return res


def lin_comb(res_type, coeff, fname_fcn, quantity):
    """linear combination of model states in files (or in HBM) (model_state_base.py:619-624): one pass over
    the k vectors per tracer module (nkb_lin_comb); coeff [n_modules, k, region_cnt(, B)]"""
    coeff = np.asarray(coeff, dtype=np.float64)
    states = []
    for ind in range(coeff.shape[1]):
        st = fname_fcn(quantity, ind)
        states.append(st if isinstance(st, ModelStateBase) else res_type(st))
    res = states[0]._like(clone_vals=False)
    B, R = res.members, res.model_config_obj.region_cnt
    for ind, tms in enumerate(res.tracer_modules):
        c = coeff[ind]
        if c.ndim == 2:  # [k, R] -> [k, R, B]
            c = np.repeat(c[:, :, None], B, axis=2)
        cdev = torch.from_numpy(np.ascontiguousarray(c.reshape(len(states), R, B))).cuda()
        vecs = [tms._flat(s.tracer_modules[ind].vals) for s in states]
        out = res.model_config_obj.weights.lin_comb(cdev, vecs, B)
        tms.vals = out.reshape(states[0].tracer_modules[ind].vals.shape)
    return res


def get_subclasses(mod_name, base_class):
    """subclasses of base_class defined in module mod_name, [] when the module does not exist
    (nk_ooc/utils.py:84-96)"""
    import importlib
    import inspect

    try:
        mod = importlib.import_module(mod_name)
    except ModuleNotFoundError as err:
        if err.name and mod_name.startswith(err.name):
            return []
        raise
    return [obj for _, obj in inspect.getmembers(mod, inspect.isclass)
            if issubclass(obj, base_class) and obj is not base_class and obj.__module__ == mod.__name__]


def get_model_state_class(model_name, lvl=logging.DEBUG):
    """model state class of model_name: the first ModelStateBase subclass of
    nk_ooc_b200.<model_name>.model_state (model_state_base.py:627-646)"""
    logger = logging.getLogger(__name__)
    cls = ModelStateBase
    subclasses = get_subclasses(".".join([__package__, model_name, "model_state"]), ModelStateBase)
    if subclasses:
        cls = subclasses[0]
    logger.log(lvl, "using class %s from %s for model state", cls.__name__, cls.__module__)
    return cls


def get_tracer_module_state_class(model_name, tracer_module_name, tracer_module_def):
    """tracer module state class: the model's TracerModuleState (nk_ooc_b200.<model>.tracer_module_state),
    overridden by the tracer module specific class of nk_ooc_b200.<model>.<py_mod_name or module name>
    (model_state_base.py:649-667)"""
    cls = TracerModuleStateBase
    subclasses = get_subclasses(".".join([__package__, model_name, "tracer_module_state"]), cls)
    if subclasses:
        cls = subclasses[0]
    py_mod_name = tracer_module_def.get("py_mod_name", tracer_module_name)
    subclasses = get_subclasses(".".join([__package__, model_name, py_mod_name]), cls)
    if subclasses:
        cls = subclasses[0]
    return cls
