"""test_problem model specifics for ModelStateBase — mirror of nk_ooc/test_problem/model_state.py
on the device.  The model year is integrated by the persistent column kernel (csrc/nkb_column.cu)
on a schedule whose step boundaries sit on the kinks of the mixing coefficient; a Richardson
pair (h, h/2) of the second-order scheme gives the accuracy the reference's CI tolerance
(rtol 1e-7, atol 2e-9 against a Radau solve at 1e-12) asks for."""

import os
from datetime import datetime

import numpy as np
import torch
from scipy.io import netcdf_file

from .. import engine
from .. import hist as hist_mod
from ..model_state_base import ModelConfig, ModelStateBase, get_tracer_module_state_class
from ..spatial_axis import spatial_axis_from_file
from . import modules
from .modules import SEC_PER_YEAR

# input/test_problem/tracer_module_defs.yaml of the reference, restated
TRACER_MODULE_DEFS = {
    "iage": {
        "region_mask_varname": "region_mask",
        "tracers": {"iage": {"attrs": {"long_name": "ideal age", "units": "years"},
                             "init_iterate_val_depths": [125.0, 650.0], "init_iterate_vals": [0.0, 1000]}},
    },
    "phosphorus": {
        "region_mask_varname": "region_mask",
        "tracers": {
            "po4": {"attrs": {"long_name": "phosphate", "units": "mmol / m^3"},
                    "init_iterate_val_depths": [125.0, 375.0], "init_iterate_vals": [0.0, 4.1],
                    "precond_matrix": "phosphorus"},
            "dop": {"attrs": {"long_name": "dissolved organic phosphorus", "units": "mmol / m^3"},
                    "init_iterate_val_depths": [100.0, 250.0], "init_iterate_vals": [7.3e-2, 0]},
            "pop": {"attrs": {"long_name": "particulate organic phosphorus", "units": "mmol / m^3"},
                    "init_iterate_val_depths": [175.0, 425.0], "init_iterate_vals": [1.8e-2, 0.0]},
            "po4_s": {"attrs": {"long_name": "shadow phosphate", "units": "mmol / m^3"}, "shadows": "po4"},
            "dop_s": {"attrs": {"long_name": "shadow dissolved organic phosphorus", "units": "mmol / m^3"},
                      "shadows": "dop"},
            "pop_s": {"attrs": {"long_name": "shadow particulate organic phosphorus", "units": "mmol / m^3"},
                      "shadows": "pop"},
        },
    },
    "dye_decay_{suff}": {
        "region_mask_varname": "region_mask",
        "py_mod_name": "dye_decay",
        "tracers": {"dye_decay_{suff}": {"attrs": {"long_name": "dye decay {suff}", "units": "mol / m^3"},
                                         "init_iterate_val_depths": [150.0], "init_iterate_vals": [0.0]}},
    },
}

# coarse leg of the Richardson pair, steps per year (the fine leg has twice as many)
DEFAULT_STEPS_PER_YEAR = {"iage": 8000, "dye_decay": 8000, "phosphorus": 16000}


class ModelState(ModelStateBase):
    """test_problem model specifics for ModelStateBase"""

    __array_priority__ = 100
    time_range = (0.0, SEC_PER_YEAR)
    depth = None
    steps_per_year = None
    richardson = True
    _side_stream = None  # second CUDA stream for the coarse leg of the Richardson pair
    _models = {}
    _precond_cache = {}

    @classmethod
    def configure(cls, modelinfo, tracer_module_defs=None, steps_per_year=None, richardson=True):
        cls.reset()
        cls.model_config_obj = ModelConfig(modelinfo, tracer_module_defs or TRACER_MODULE_DEFS)
        cls.depth = spatial_axis_from_file(modelinfo["grid_vars_fname"], modelinfo.get("depth_axisname", "depth"))
        cls.steps_per_year = steps_per_year
        cls.richardson = richardson

    @classmethod
    def reset(cls):
        cls.model_config_obj = None
        cls.depth = None
        cls._models = {}
        cls._precond_cache = {}

    def __init__(self, fname, members=1):
        if ModelState.model_config_obj is None:
            raise RuntimeError("ModelState.model_config_obj is None")
        if self.model_config_obj.region_cnt != 1:
            raise NotImplementedError("region_cnt > 1 is not supported by test_problem")
        super().__init__(fname, members)

    def _new_tracer_module(self, name, tracer_module_def, members):
        """the tracer module's own class (iage, dye_decay, phosphorus: model_state_base.py:649-667)"""
        cls = get_tracer_module_state_class("test_problem", name, tracer_module_def)
        return cls(name, tracer_module_def, (len(self.depth),), self.model_config_obj, members=members)

    def _gen_init_iterate(self, tms):
        """test_problem/tracer_module_state.py:41-68"""
        metas = tms._def["tracers"]
        for tname, meta in metas.items():
            if "init_iterate_vals" not in meta and "shadows" in meta:
                meta = metas[meta["shadows"]]
            if "init_iterate_vals" not in meta:
                raise ValueError(f"gen_init_iterate failure for {tname}")
            tms.set_tracer_vals(tname, np.interp(self.depth.mid, meta["init_iterate_val_depths"],
                                                 meta["init_iterate_vals"]))

    def _axes(self):
        return [self.depth]

    # ---- device models: (coarse, fine) pair per tracer module ---------------------------------
    @classmethod
    def models_for(cls, tms):
        if tms.name in cls._models:
            return cls._models[tms.name]
        kind = tms._def.get("py_mod_name", tms.name)
        info = cls.model_config_obj.modelinfo

        def build():
            if kind == "iage":
                return modules.iage_model(cls.depth)
            if kind == "dye_decay":
                return modules.dye_decay_model(cls.depth, tms.name[10:])
            if kind == "phosphorus":
                return modules.phosphorus_model(cls.depth, int(info.get("po4_s_restoring_opt", 1)))
            raise NotImplementedError(f"tracer module {tms.name} is not available in test_problem")

        spy = cls.steps_per_year or DEFAULT_STEPS_PER_YEAR[kind]
        extra = (0.1, 0.2, 0.6, 0.7) if kind == "dye_decay" else ()
        sched = modules.aligned_schedule(cls.depth, spy, extra)
        coarse = build()
        coarse.set_schedule(*sched)
        fine = None
        if cls.richardson:
            fine = build()
            fine.set_schedule(*modules.halved(sched))
        cls._models[tms.name] = (coarse, fine)
        return cls._models[tms.name]

    def _eval_module(self, tms, hist_times=None):
        """F for one tracer module: Richardson combination (4 F_{h/2} - F_h)/3 of the two legs"""
        coarse, fine = self.models_for(tms)
        x = tms.vals.reshape(tms.tracer_cnt, len(self.depth), 1, -1)
        lead = fine if fine is not None else coarse
        snaps = None
        f_c = snaps_c = None
        if fine is not None:
            # the coarse leg runs on a side stream, concurrently with the fine leg (a column-year launch of a
            # small batch occupies a few SMs only)
            cur = torch.cuda.current_stream()
            side = type(self)._side_stream
            if side is None:
                side = type(self)._side_stream = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                if hist_times is not None:
                    f_c, snaps_c = coarse.eval(x, self.members, hist_steps=coarse.step_index_of_times(hist_times))
                else:
                    f_c = coarse.eval(x, self.members)
        if hist_times is not None:
            f_lead, snaps = lead.eval(x, self.members, hist_steps=lead.step_index_of_times(hist_times))
        else:
            f_lead = lead.eval(x, self.members)
        if fine is not None:
            cur.wait_stream(side)
            f_c.record_stream(cur)
            x.record_stream(side)
            if snaps_c is not None:
                snaps_c.record_stream(cur)
                snaps.mul_(4.0 / 3.0).add_(snaps_c, alpha=-1.0 / 3.0)
            # F <- 4/3 F_fine - 1/3 F_coarse with the library's axpby kernel (K6)
            flat = (tms.tracer_cnt, len(self.depth), f_lead.shape[-1])
            self.model_config_obj.weights.axpby(-1.0 / 3.0, f_c.reshape(flat), 4.0 / 3.0, f_lead.reshape(flat),
                                                self.members)
        return f_lead.reshape(tms.vals.shape), snaps

    # ---- operators ------------------------------------------------------------------------
    def comp_fcn(self, res_fname, solver_state, hist_fname=None):
        """test_problem/model_state.py:52-117"""
        step = f"comp_fcn complete for {res_fname}"
        if solver_state is not None and solver_state.step_logged(step):
            return ModelState(res_fname)
        res_ms = self._like(clone_vals=False)
        hist = {}
        times = np.linspace(self.time_range[0], self.time_range[1], 101) if hist_fname is not None else None
        for ind, tms in enumerate(self.tracer_modules):
            res_ms.tracer_modules[ind].vals, snaps = self._eval_module(tms, times)
            hist[tms.name] = snaps
        if hist_fname is not None:
            self._write_hist(hist_fname, times, hist)
        res_ms.comp_fcn_postprocess(res_fname, f"{type(self).__name__}.comp_fcn")
        if solver_state is not None:
            solver_state.log_step(step)
        return res_ms

    def zero_extra_tracers(self):
        """when shadow tracers run, the real tracers they shadow are not solved for
        (tracer_module_state_base.py:483-500)"""
        for tms in self.tracer_modules:
            for _, real in tms.shadow_pairs():
                tms.vals[real] = 0.0
        return self

    @staticmethod
    def _hist_tracer_like(tms):
        """{name: attrs} of the tracer-like hist variables (test_problem/tracer_module_state.py:149-154;
        phosphorus adds po4_uptake and po4_s_restore_tau_r, phosphorus.py:122-135)"""
        res = {tname: meta["attrs"] for tname, meta in tms._def["tracers"].items()}
        if tms._def.get("py_mod_name", tms.name) == "phosphorus":
            res["po4_uptake"] = {"long_name": "uptake of po4", "units": f"{res['po4']['units']} / s"}
            res["po4_s_restore_tau_r"] = {"long_name": "inverse timescale for po4_s restoring", "units": "1 / s"}
        return res

    def _write_hist(self, hist_fname, times, hist):
        """time, depth axis, bldepth, mixing_coeff, tracer snapshots of member 0 and their derived
        variables (test_problem/model_state.py:119-225; tracer_module_state.py:96-199)"""
        os.makedirs(os.path.dirname(os.path.abspath(hist_fname)), exist_ok=True)
        model = self.models_for(self.tracer_modules[0])[0]
        dn, de = self.depth.axisname, self.depth.dump_names["edges"]
        with netcdf_file(hist_fname, "w", version=2) as fptr:
            stamp = datetime.now().strftime("%Y-%m-%d %H:%M:%S")
            fptr.history = f"{stamp}: created by {__name__}._gen_hist"
            fptr.createDimension("time", None)
            self.depth.define(fptr)
            var = fptr.createVariable("time", "f8", ("time",))
            var.long_name, var.units, var.calendar = "time", "seconds since 0001-01-01", "noleap"
            var = fptr.createVariable("bldepth", "f8", ("time",))
            var.long_name, var.units, var.cell_methods = "boundary layer depth", "m", "time: point"
            var = fptr.createVariable("mixing_coeff", "f8", ("time", de))
            var.long_name, var.units, var.cell_methods = "vertical mixing coefficient", "m^2 / s", "time: point"
            for tms in self.tracer_modules:
                for tname, attrs in self._hist_tracer_like(tms).items():
                    var = fptr.createVariable(tname, "f8", ("time", dn))
                    var.long_name, var.units = attrs["long_name"], attrs["units"]
                    var.cell_methods = "time: point"
                    hist_mod.define_derived(fptr, tname, attrs, self.depth)
            self.depth.write(fptr)
            for ti, t in enumerate(times):
                fptr.variables["time"][ti] = t
                fptr.variables["bldepth"][ti] = modules.bldepth(t)
                mc = np.empty(len(self.depth) + 1)
                mc[1:-1] = model.mixing_coeff(t).cpu().numpy()[:, 0] * self.depth.delta_mid
                mc[0], mc[-1] = mc[1], mc[-2]
                fptr.variables["mixing_coeff"][ti, :] = mc
            for tms in self.tracer_modules:
                snaps = hist[tms.name].cpu().numpy()  # [n_time, T, nz, 1]
                like = {tname: snaps[:, ind, :, 0] for ind, tname in enumerate(tms.tracer_names)}
                if tms._def.get("py_mod_name", tms.name) == "phosphorus":
                    po4 = like["po4"]
                    like["po4_uptake"] = modules.po4_uptake(self.depth, po4)
                    opt = int(self.model_config_obj.modelinfo.get("po4_s_restoring_opt", 1))
                    like["po4_s_restore_tau_r"] = modules.po4_s_restore_tau_r(self.depth, po4, like["po4_uptake"], opt)
                for tname, vals in like.items():
                    fptr.variables[tname][:] = vals
                    hist_mod.write_derived(fptr, tname, vals, self.depth)

    def apply_precond_jacobian(self, precond_fname, res_fname, solver_state):
        """res = A^-1 (self / T) - self, A the tridiagonal Jacobian with the log-mean mixing
        coefficient (test_problem/model_state.py:227-270; iage.py:31-52; dye_decay.py:49-73)"""
        step = f"apply_precond_jacobian complete for {res_fname}"
        if solver_state is not None and solver_state.step_logged(step):
            return ModelState(res_fname)
        with netcdf_file(precond_fname, "r", mmap=False) as fptr:
            mca = np.array(fptr.variables["mixing_coeff_log_mean"].data)[1:-1]
        res_ms = self._like(clone_vals=False)
        t0, t1 = self.time_range
        for ind, tms in enumerate(self.tracer_modules):
            kind = tms._def.get("py_mod_name", tms.name)
            if kind == "phosphorus":
                res_ms.tracer_modules[ind].vals = self._apply_precond_phosphorus(tms, precond_fname, mca)
                continue
            if kind not in ("iage", "dye_decay"):
                raise NotImplementedError(f"preconditioner of {tms.name} is not on the B200 path yet")
            key = (tms.name, precond_fname)
            if key not in self._precond_cache:
                d = self.depth
                ab = np.zeros((3, len(d)))
                ab[0, 1:] = mca * d.delta_mid_r * d.delta_r[:-1]
                ab[1, :-1] -= mca * d.delta_mid_r * d.delta_r[:-1]
                ab[1, 1:] -= mca * d.delta_mid_r * d.delta_r[1:]
                ab[2, :-1] = mca * d.delta_mid_r * d.delta_r[1:]
                if kind == "iage":
                    ab[1, 0] -= 24.0 * (1.0 / 86400.0) * 10.0 * d.delta_r[0]
                else:
                    ab[1, :] -= int(tms.name[10:]) * 0.001 * (1.0 / SEC_PER_YEAR)
                self._precond_cache[key] = engine.BandedFactor(ab, 1, 1)
            y = tms.vals[0].reshape(len(self.depth), -1)
            out = self._precond_cache[key].solve(y, self.members, 1.0 / (t1 - t0), subtract_rhs=True)
            res_ms.tracer_modules[ind].vals = out.reshape(tms.vals.shape)
        if solver_state is not None:
            solver_state.log_step(step)
        return res_ms.dump(res_fname, f"{type(self).__name__}.apply_precond_jacobian")

    def _apply_precond_phosphorus(self, tms, precond_fname, mca):
        """preconditioner of the shadow phosphorus tracers (test_problem/phosphorus.py:169-211):
        two regularised solves (shifts 1e-11, 0.5e-11) + Richardson extrapolation, removal of the
        null vector weighted by layer thickness, minus the input.  The member-independent pieces
        (band matrices, SVD null vector) are set up on the host once per precond file; the solves,
        the weighted sums and the updates of all members run in the library's kernels (K4-K6)."""
        nz, B = len(self.depth), self.members
        weights = self.model_config_obj.weights
        key = (tms.name, precond_fname)
        if key not in self._precond_cache:
            with netcdf_file(precond_fname, "r", mmap=False) as fptr:
                tau_r = np.array(fptr.variables["po4_s_restore_tau_r_mean"].data)
            ab_a, kl, ku, _ = _phosphorus_precond_band(self.depth, mca, tau_r, 1.0e-11)
            ab_b, _, _, _ = _phosphorus_precond_band(self.depth, mca, tau_r, 0.5e-11)
            _, _, _, dense = _phosphorus_precond_band(self.depth, mca, tau_r, 0.0)
            _, sing_vals, r_sing_vects = np.linalg.svd(dense)
            null_vect = r_sing_vects[sing_vals.argmin(), :]
            dz3 = np.concatenate((self.depth.delta,) * 3)
            # region-weighted "mean" of K5 = sum_k dz_k/sum(dz) * (.) summed over the 3 tracers
            denom_mean = (null_vect * dz3).sum() / self.depth.delta.sum()
            self._precond_cache[key] = (engine.BandedFactor(ab_a, kl, ku), engine.BandedFactor(ab_b, kl, ku),
                                        null_vect.reshape(3, nz), denom_mean)
        fac_a, fac_b, null_vect, denom_mean = self._precond_cache[key]
        t0, t1 = self.time_range
        ldb = tms.vals.shape[-1]
        y3 = tms.vals[3:6].reshape(3 * nz, ldb)
        res_a = fac_a.solve(y3, B, 1.0 / (t1 - t0))
        res = fac_b.solve(y3, B, 1.0 / (t1 - t0))
        weights.axpby(-1.0, res_a.reshape(3, nz, ldb), 2.0, res.reshape(3, nz, ldb), B)  # 2 b - a
        numer_mean = weights.dot(res.reshape(3, nz, ldb), None, B)  # [1, B]
        nv = torch.zeros((3, nz, ldb), dtype=torch.float64, device="cuda")
        nv[..., :B] = torch.from_numpy(null_vect).cuda().unsqueeze(-1)
        weights.axpby(-numer_mean / denom_mean, nv, 1.0, res.reshape(3, nz, ldb), B)
        weights.axpby(-1.0, y3.reshape(3, nz, ldb), 1.0, res.reshape(3, nz, ldb), B)
        out = tms.vals.clone()  # real tracers keep their values (res_ms = deepcopy(self), :237)
        out[3:6] = res.reshape(out[3:6].shape)
        return out


def _phosphorus_precond_band(depth, mca, tau_r, shift):
    """banded storage [kl+ku+1, 3nz] (kl = ku = 2nz) of the 7-diagonal preconditioner matrix of the
    shadow phosphorus tracers minus shift*I (test_problem/phosphorus.py:177-290)"""
    nz = len(depth)
    day_per_sec = 1.0 / 86400.0
    single = np.zeros(nz)
    single[:-1] -= mca * depth.delta_mid_r * depth.delta_r[:-1]
    single[1:] -= mca * depth.delta_mid_r * depth.delta_r[1:]
    d0 = np.concatenate((single - tau_r, single - 0.01 * day_per_sec, single - 0.01 * day_per_sec))
    d0[2 * nz:3 * nz - 1] -= day_per_sec * depth.delta_r[:-1]  # pop_s sinking loss to the layer below
    up = mca * depth.delta_mid_r * depth.delta_r[:-1]
    lo = mca * depth.delta_mid_r * depth.delta_r[1:]
    zero = np.zeros(1)
    diags = {
        0: d0 - shift,
        1: np.concatenate((up, zero, up, zero, up)),
        -1: np.concatenate((lo, zero, lo, zero, lo + day_per_sec * depth.delta_r[1:])),
        nz: np.concatenate((0.01 * day_per_sec * np.ones(nz), np.zeros(nz))),
        -nz: np.concatenate((0.67 * tau_r, np.zeros(nz))),
        2 * nz: 0.01 * day_per_sec * np.ones(nz),
        -2 * nz: 0.33 * tau_r,
    }
    n, kl, ku = 3 * nz, 2 * nz, 2 * nz
    ab = np.zeros((kl + ku + 1, n))
    dense = np.zeros((n, n))
    for off, vals in diags.items():
        rows = np.arange(len(vals)) + max(0, -off)
        cols = rows + off
        ab[ku + rows - cols, cols] = vals
        dense[rows, cols] = vals
    return ab, kl, ku, dense


def gen_depth_axis_file(modelinfo, depth):
    """depth_axis.nc = axis dump + region_mask(depth) with cell_measures "thickness: depth_delta"
    (test_problem/setup_solver.py:101-117)"""
    fname = modelinfo["grid_vars_fname"]
    with netcdf_file(fname, "w", version=2) as fptr:
        stamp = datetime.now().strftime("%Y-%m-%d %H:%M:%S")
        fptr.history = f"{stamp}: generated by SpatialAxis.dump called from test_problem.setup_solver"
        if depth.defn_dict_values is not None:
            fptr.defn_dict_values = depth.defn_dict_values
        depth.define(fptr)
        var = fptr.createVariable("region_mask", "i4", (depth.axisname,))
        var.long_name = "Region Mask"
        var.cell_measures = "thickness: depth_delta"
        depth.write(fptr)
        var[:] = np.ones(len(depth), dtype=np.int32)
