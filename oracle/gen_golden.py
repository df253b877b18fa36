"""TEST INFRASTRUCTURE: generate tests/golden/*.npz from the REFERENCE itself.

Runs only in the build container (needs /root/reference).  It imports the reference's own
numerical modules through oracle/ref_harness.py (stubbed netCDF4/xarray/pint), evaluates them
on seeded inputs, and copies the values of the committed baselines
(/root/reference/baselines/ci_*) that the hot path is compared against.  The fixtures are
small (values only, float64/int32) and are committed together with this script.

    python -m oracle.gen_golden            # from the repo root
"""

import os
import sys

import numpy as np
from scipy.io import netcdf_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
YEAR = 365.0 * 86400.0


def baselines():
    """values of the committed CI baselines that involve the hot path"""
    base = os.path.join(rh.REF_ROOT, "baselines")
    out = {}
    want = {
        "ci_short": ["depth_axis", "fcn_00", "init_iterate", "init_iterate_00"],
        "ci_long_iage": ["basis_00", "increment_00", "iterate_01", "krylov_res_00", "perturb_fcn_w_raw_00",
                         "precond_00", "precond_fcn_00", "w_00", "w_raw_00"],
        "ci_py_driver_2d_iage": ["fcn_0000", "grid_vars", "init_iterate", "init_iterate_0000"],
        "ci_py_driver_2d_iage_column_regions": ["basis_00", "fcn_0000", "grid_vars", "increment_00", "init_iterate",
                                                "init_iterate_0000", "iterate_01", "krylov_res_00",
                                                "perturb_fcn_w_raw_00", "precond_fcn_00"],
    }
    for cfg, files in want.items():
        for f in files:
            for var, vals in rh.read_nc(os.path.join(base, cfg, f + ".nc")).items():
                out[f"{cfg}/{f}/{var}"] = vals
    # a few hist variables (process fields, pinned at rtol 1e-3/atol 1e-6 by the CI script)
    hist = rh.read_nc(os.path.join(base, "ci_py_driver_2d_iage", "hist_0000.nc"))
    derived = ["time_mean", "time_std", "time_delta", "depth_int"]  # (time_anom is large and redundant)
    for var in ["time", "stream", "vvel", "wvel", "horiz_mixing_coeff", "bldepth", "vert_mixing_coeff", "iage",
                "iage_slow_rest"] + [f"iage_{d}" for d in derived + ["ypos_mean", "depth_ypos_int"]]:
        out[f"ci_py_driver_2d_iage/hist_0000/{var}"] = hist[var]
    hist = rh.read_nc(os.path.join(base, "ci_short", "hist_00.nc"))
    for var in (["time", "bldepth", "mixing_coeff", "iage", "po4", "po4_uptake", "po4_s_restore_tau_r"]
                + [f"{v}_{d}" for v in ("iage", "po4", "po4_uptake", "po4_s_restore_tau_r") for d in derived]):
        out[f"ci_short/hist_00/{var}"] = hist[var]
    np.savez_compressed(os.path.join(OUT, "baselines.npz"), **out)
    print("baselines.npz:", len(out), "arrays")
    # the step log of the reference's own ci_long_iage run (solver persistence tests)
    import shutil

    shutil.copyfile(os.path.join(base, "ci_long_iage", "Newton_state.json"),
                    os.path.join(OUT, "Newton_state_ci_long_iage.json"))


def baseline_files():
    """every committed CI baseline file in full — variable values in baseline_files.npz, dimensions / variable
    order / attributes in baseline_files_meta.json — so that tests/baseline_files.py can put the reference's
    baselines/ directory back together on the GPU box (where /root/reference does not exist) and the ports of
    scripts/ci_*.sh can run `baseline_cmp` against real files, metadata check included"""
    import glob
    import json
    import shutil

    base = os.path.join(rh.REF_ROOT, "baselines")
    vals, meta = {}, {}
    for fname in sorted(glob.glob(os.path.join(base, "ci_*", "*.nc"))):
        key = os.path.relpath(fname, base)[:-3]
        nc = netcdf_file(fname, "r", mmap=False)
        dims = [[name, None if length is None else int(length)] for name, length in nc.dimensions.items()]
        entry = {"dims": dims, "vars": []}
        for name, var in nc.variables.items():
            attrs = {}
            for akey, aval in var._attributes.items():
                if isinstance(aval, bytes):
                    aval = aval.decode()
                elif isinstance(aval, np.ndarray):
                    aval = aval.tolist()
                elif isinstance(aval, np.generic):
                    aval = aval.item()
                attrs[akey] = aval
            entry["vars"].append({"name": name, "dims": list(var.dimensions), "dtype": var.data.dtype.str[1:],
                                  "attrs": attrs})
            vals[f"{key}/{name}"] = np.array(var.data, dtype=var.data.dtype.newbyteorder("="))
        nc.close()
        meta[key] = entry
    np.savez_compressed(os.path.join(OUT, "baseline_files.npz"), **vals)
    with open(os.path.join(OUT, "baseline_files_meta.json"), "w") as fptr:
        json.dump(meta, fptr, indent=1, sort_keys=True)
    for cfg in ("ci_long_dye_decay", "ci_long_iage", "ci_py_driver_2d_iage_column_regions"):
        shutil.copyfile(os.path.join(base, cfg, "Newton_state.json"), os.path.join(OUT, f"Newton_state_{cfg}.json"))
    print("baseline_files.npz:", len(vals), "arrays of", len(meta), "files")


def remap_cases():
    """SpatialAxis.remap_linear_interpolant of the reference on seeded inputs + the known
    answers of the reference's unit tests (tests/test_spatial_axis.py:138-198)"""
    rng = np.random.default_rng(100)
    out = {}
    ax5 = rh.make_axis("depth", edge_end=50.0, nlevs=5, delta_ratio_max=1.0)
    out["known/edges"] = ax5.edges
    known = [
        ([-15.0, -5.0], [1.0, 2.0], [2.0] * 5),
        ([-15.0, 25.0], [0.0, 8.0], [4.0, 6.0, 7.75, 8.0, 8.0]),
        ([5.0, 25.0], [0.0, 8.0], [0.5, 4.0, 7.5, 8.0, 8.0]),
        ([22.5, 27.5], [0.0, 8.0], [0.0, 0.0, 4.0, 8.0, 8.0]),
        ([42.5, 47.5], [0.0, 8.0], [0.0, 0.0, 0.0, 0.0, 4.0]),
        ([45.0, 55.0], [0.0, 8.0], [0.0, 0.0, 0.0, 0.0, 1.0]),
    ]
    for i, (x, y, e) in enumerate(known):
        assert (ax5.remap_linear_interpolant(np.array(x), np.array(y)) == np.array(e)).all()
        out[f"known/{i}/x"], out[f"known/{i}/y"], out[f"known/{i}/expected"] = map(np.array, (x, y, e))
    out["known/count"] = np.array(len(known))
    ax = rh.make_axis("depth", nlevs=30)
    out["rand/edges"] = ax.edges
    n_case = 60
    for i in range(n_case):
        n = int(rng.integers(1, 6))
        x = np.sort(rng.uniform(-150.0, 1050.0, n))
        y = rng.normal(size=n)
        out[f"rand/{i}/x"], out[f"rand/{i}/y"] = x, y
        out[f"rand/{i}/res"] = ax.remap_linear_interpolant(x, y)
    out["rand/count"] = np.array(n_case)
    np.savez_compressed(os.path.join(OUT, "remap.npz"), **out)
    print("remap.npz:", len(out), "arrays")


def synthetic_forcing(nz, ny, seed):
    """small synthetic forcing record standing in for input/py_driver_2d/po4_sms.nc"""
    rng = np.random.default_rng(seed)
    nt = 13
    times = np.linspace(0.0, YEAR, nt)
    data = 3.0e-8 * rng.normal(size=(nt, nz, ny))
    return times, data


def write_forcing_nc(fname, times, data, depth_mid, ypos_mid):
    from scipy.io import netcdf_file

    with netcdf_file(fname, "w", version=2) as f:
        f.createDimension("time", len(times))
        f.createDimension("depth", len(depth_mid))
        f.createDimension("ypos", len(ypos_mid))
        for name, vals in (("time", times), ("depth", depth_mid), ("ypos", ypos_mid)):
            v = f.createVariable(name, "f8", (name,))
            v[:] = vals
        v = f.createVariable("po4_sms", "f8", ("time", "depth", "ypos"))
        v[:] = data


def py_driver_2d_cases():
    """reference process fields + comp_tend / comp_jacobian / apply_precond_jacobian of the three
    py_driver_2d tracer modules on seeded states"""
    rng = np.random.default_rng(200)
    out = {}
    for tag, (nz, ny, ratio, vvel, kh) in {
        "g14x11": (14, 11, 19.0, 0.1, 1000.0),
        "g20x3cr": (20, 3, 19.0, 0.0, 0.0),
        "g30x30": (30, 30, 19.0, 0.1, 1000.0),
    }.items():
        depth, ypos, procs = rh.make_py_driver_2d(nz, ny, ratio, vvel, kh)
        from nk_ooc.py_driver_2d.advection import Advection

        out[f"{tag}/params"] = np.array([nz, ny, ratio, vvel, kh])
        out[f"{tag}/depth_edges"], out[f"{tag}/ypos_edges"] = depth.edges, ypos.edges
        out[f"{tag}/stream"], out[f"{tag}/vvel"], out[f"{tag}/wvel"] = Advection.stream, Advection.vvel, Advection.wvel
        out[f"{tag}/hmix"] = procs["horiz_mix"]._mixing_coeff
        times = YEAR * np.array([0.0, 0.1, 0.26, 0.3, 0.349, 0.5, 0.66, 0.7, 0.99])
        out[f"{tag}/times"] = times
        out[f"{tag}/bldepth"] = np.stack([procs["vert_mix"].bldepth(t) for t in times])
        out[f"{tag}/mixing_coeff"] = np.stack([procs["vert_mix"].mixing_coeff(t).copy() for t in times])
        if tag == "g30x30":
            continue
        # tracer modules
        iage = rh.make_2d_iage(depth, ypos)
        x = rng.normal(size=(2, nz, ny))
        out[f"{tag}/iage/x"] = x
        out[f"{tag}/iage/tend"] = np.stack([iage.comp_tend(t, x.reshape(-1), procs).reshape(x.shape) for t in times[:4]])
        out[f"{tag}/iage/jac_dense_t3"] = iage.comp_jacobian(times[3], x.reshape(-1), procs).toarray()

        class _Res:  # receives apply_precond_jacobian's result
            def set_tracer_vals_all(self, vals, reseat_vals=False):
                self.vals = vals

        iage.get_tracer_vals_all = lambda x=x: x
        res = _Res()
        iage.apply_precond_jacobian((0.0, YEAR), res, procs)
        out[f"{tag}/iage/precond"] = res.vals

        phos = rh.make_2d_phosphorus(depth, ypos)
        x = np.abs(rng.normal(size=(3, nz, ny))) * np.array([2.0, 0.05, 0.01])[:, None, None]
        out[f"{tag}/phosphorus/x"] = x
        out[f"{tag}/phosphorus/tend"] = np.stack(
            [phos.comp_tend(t, x.reshape(-1), procs).reshape(x.shape) for t in times[:4]]
        )

        out[f"{tag}/phosphorus/jac_dense_t3"] = phos.comp_jacobian(times[3], x.reshape(-1), procs).toarray()
        out.update(phosphorus_2d_precond_case(tag, depth, ypos, procs, rng))

        ft, fd = synthetic_forcing(nz, ny, 300)
        fname = f"/tmp/golden_forcing_{tag}.nc"
        write_forcing_nc(fname, ft, fd, depth.mid, ypos.mid)
        modelinfo = {
            "forced_surf_restore_opt": "const", "forced_surf_restore_const": "1.0",
            "forced_surf_restore_rate_10m": "1.0 / 3600.0", "forced_sms_opt": "file",
            "forced_sms_fname": fname, "forced_sms_varname": "po4_sms", "forced_sms_scalef": "-1.0 / 3.0",
            "forced_sink_thres": "0.05",
        }
        forced = rh.make_2d_forced(depth, ypos, modelinfo)
        x = np.abs(rng.normal(size=(1, nz, ny))) * 0.06
        out[f"{tag}/forced/frc_time"], out[f"{tag}/forced/frc_data"] = ft, fd
        out[f"{tag}/forced/x"] = x
        tt = YEAR * np.array([0.0, 0.1234, 0.5, 0.987])
        out[f"{tag}/forced/times"] = tt
        out[f"{tag}/forced/tend"] = np.stack([forced.comp_tend(t, x.reshape(-1), procs).reshape(x.shape) for t in tt])
    np.savez_compressed(os.path.join(OUT, "py_driver_2d.npz"), **out)
    print("py_driver_2d.npz:", len(out), "arrays")


def phosphorus_2d_precond_case(tag, depth, ypos, procs, rng):
    """the reference's py_driver_2d phosphorus.apply_precond_jacobian (phosphorus.py:197-274) run
    unmodified; only the xarray-backed base-class plumbing it touches (value access, region mean,
    the three in-place operators, the null-space file dump) is replaced by numpy equivalents with
    the semantics of tracer_module_state_base.py:255-388 for ONE region covering the grid"""
    import nk_ooc.py_driver_2d.phosphorus as ref_mod

    nz, ny = len(depth), len(ypos)
    weight = np.outer(depth.delta, ypos.delta)
    weight = weight / weight.sum()  # region_comp_mean_matrix row (model_config.py:292-315)

    class _P(ref_mod.phosphorus):  # pylint: disable=invalid-name
        def get_tracer_vals_all(self):
            return self._v

        def set_tracer_vals_all(self, vals, reseat_vals=False):
            self._v = vals if reseat_vals else np.array(vals)

        def mean(self):  # tracer_module_state_base.py:371-377: sum over tracers of the region means
            return np.array([(weight[None] * self._v).sum()])

        def __itruediv__(self, other):
            self._v /= float(np.asarray(other).reshape(-1)[0])
            return self

        def __rmul__(self, other):
            res = copy.copy(self)
            res._v = float(np.asarray(other).reshape(-1)[0]) * self._v
            return res

        def __isub__(self, other):
            self._v -= other._v
            return self

        def dump(self, fptr, action):
            pass

    import copy

    class _NullDataset:
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    class _Var:
        def __init__(self, arr):
            self._a = arr

        def __getitem__(self, key):
            return self._a[key]

    class _Precond:
        def __init__(self, times, po4):
            self.variables = {"time": _Var(times), "po4": _Var(po4)}

        def filepath(self):
            return "/tmp/golden_precond/precond_00.nc"

    ref_mod.Dataset = _NullDataset
    base = rh.make_2d_phosphorus(depth, ypos)
    tm = object.__new__(_P)
    tm.__dict__.update(base.__dict__)
    y = rng.normal(size=(3, nz, ny)) * np.array([1.0, 0.05, 0.01])[:, None, None]
    tm._v = y.copy()
    n_t = 4
    ptimes = YEAR * np.linspace(0.0, 1.0, n_t)
    po4 = np.abs(rng.normal(size=(n_t, nz, ny))) * 2.0

    class _Res:
        def set_tracer_vals_all(self, vals, reseat_vals=False):
            self.vals = np.array(vals)

    res = _Res()
    tm.apply_precond_jacobian((0.0, YEAR), res, procs, _Precond(ptimes, po4))
    return {f"{tag}/phosphorus/precond_y": y, f"{tag}/phosphorus/precond_times": ptimes,
            f"{tag}/phosphorus/precond_po4": po4, f"{tag}/phosphorus/precond": res.vals}


def test_problem_cases():
    """reference test_problem vert_mix + tracer-module tendencies + tridiagonal preconditioners"""
    rng = np.random.default_rng(400)
    out = {}
    nz = 20
    depth, vert_mix = rh.make_test_problem(nz)
    out["depth_edges"] = depth.edges
    times = YEAR * np.array([0.0, 0.05, 0.15, 0.3, 0.5, 0.65, 0.9])
    out["times"] = times
    out["bldepth"] = np.array([vert_mix.bldepth(t) for t in times])
    out["mixing_coeff"] = np.stack([vert_mix.mixing_coeff(t).copy() for t in times])
    for kind, name in (("iage", "iage"), ("dye_decay", "dye_decay_010"), ("phosphorus", "phosphorus")):
        tm = rh.make_tp_module(kind, depth, name)
        x = np.abs(rng.normal(size=(tm.tracer_cnt, nz)))
        out[f"{name}/x"] = x
        out[f"{name}/tend"] = np.stack([tm.comp_tend(t, x.reshape(-1), vert_mix).reshape(x.shape) for t in times])
        if kind != "phosphorus":
            mca = np.abs(rng.normal(size=nz - 1)) * 1.0e-3

            class _Res:
                def set_tracer_vals_all(self, vals, reseat_vals=False):
                    self.vals = vals

            tm.get_tracer_vals_all = lambda x=x: x
            res = _Res()
            tm.apply_precond_jacobian((0.0, YEAR), res, mca)
            out[f"{name}/mca"] = mca
            out[f"{name}/precond"] = res.vals
    # phosphorus: derived hist variables and the 3nz x 3nz preconditioner of the shadow tracers
    # (test_problem/phosphorus.py:60-120,169-290)
    tm = rh.make_tp_module("phosphorus", depth, "phosphorus")
    x = out["phosphorus/x"]
    uptake = tm.po4_uptake(x[0])
    tau_r = tm.po4_s_restore_tau_r(x[0], uptake)
    out["phosphorus/po4_uptake"] = uptake
    out["phosphorus/po4_s_restore_tau_r"] = tau_r
    mca = np.abs(rng.normal(size=nz - 1)) * 1.0e-3
    y = rng.normal(size=(6, nz))

    class _ResP:
        def __init__(self):
            self.vals = {}

        def set_tracer_vals(self, name, vals):
            self.vals[name] = np.array(vals)

    tm.get_tracer_vals_all = lambda y=y: y
    res = _ResP()
    tm.apply_precond_jacobian((0.0, YEAR), res, mca, tau_r)
    out["phosphorus/precond_mca"] = mca
    out["phosphorus/precond_y"] = y
    out["phosphorus/precond"] = np.stack([res.vals[n] for n in ("po4_s", "dop_s", "pop_s")])
    np.savez_compressed(os.path.join(OUT, "test_problem.npz"), **out)
    print("test_problem.npz:", len(out), "arrays")


def forced_restore_file_case():
    """the reference's forced module with forced_surf_restore_opt = file (forced.py:46-51,124-130): the
    restoring record [time, ypos] is given on a COARSER ypos axis than the model's, so the reference's
    gen_forcing_fcn also interpolates it in space (utils.py:518-531); written to its own fixture so that
    the seeded states of py_driver_2d.npz stay what they are"""
    from scipy.io import netcdf_file

    rng = np.random.default_rng(410)
    nz, ny = 12, 9
    depth, ypos, procs = rh.make_py_driver_2d(nz, ny, 19.0, 0.1, 1000.0)
    nt, ny_file = 7, 5
    rtimes = np.linspace(0.0, YEAR, nt)
    ypos_file = np.linspace(ypos.mid[0], ypos.mid[-1], ny_file)
    rdata = 1.0 + 0.3 * rng.normal(size=(nt, ny_file))
    fname = "/tmp/golden_surf_restore.nc"
    with netcdf_file(fname, "w", version=2) as f:
        f.createDimension("time", nt)
        f.createDimension("ypos", ny_file)
        for name, vals in (("time", rtimes), ("ypos", ypos_file)):
            v = f.createVariable(name, "f8", (name,))
            v[:] = vals
        v = f.createVariable("surf_vals", "f8", ("time", "ypos"))
        v[:] = rdata
    modelinfo = {
        "forced_surf_restore_opt": "file", "forced_surf_restore_fname": fname,
        "forced_surf_restore_varname": "surf_vals", "forced_surf_restore_rate_10m": "1.0 / 7200.0",
        "forced_sms_opt": "const", "forced_sms_const": "-1.0e-9",
    }
    forced = rh.make_2d_forced(depth, ypos, modelinfo, suff="restore_file")
    x = np.abs(rng.normal(size=(1, nz, ny)))
    tt = YEAR * np.array([0.0, 0.21, 0.5, 0.93, 1.0])
    out = {
        "params": np.array([nz, ny, 19.0, 0.1, 1000.0]), "depth_edges": depth.edges, "ypos_edges": ypos.edges,
        "rtimes": rtimes, "ypos_file": ypos_file, "rdata": rdata,
        "x": x, "times": tt,
        "restore_to": np.stack([forced.surf_restore_fcn(t) for t in tt]),
        "tend": np.stack([forced.comp_tend(t, x.reshape(-1), procs).reshape(x.shape) for t in tt]),
        "jac_dense_t1": forced.comp_jacobian(tt[1], x.reshape(-1), procs).toarray(),
    }
    np.savez_compressed(os.path.join(OUT, "forced_restore_file.npz"), **out)
    print("forced_restore_file.npz:", len(out), "arrays")


def precond_2d_cases():
    """the reference's general 2-D preconditioners WITH lateral processes, run unmodified:
    iage.apply_precond_jacobian (py_driver_2d/iage.py:66-93) on 14x11 and 30x30, and
    forced.apply_precond_jacobian (forced.py:204-241) for the o2_like configuration
    (scripts/run_py_driver_2d_forced_o2_like.sh:14-25; sink record = input/py_driver_2d/po4_sms.nc through the
    reference's own gen_forcing_fcn) whose Jacobian depends on the precond file's tracer snapshots through
    the sink_thres term (forced.py:190-202,221-229).  Also the reference's comp_jacobian of both modules at
    the three interval mid-points, so that the host-side assembly can be compared entry by entry.
    Own fixture (precond_2d.npz) so that the seeded states of py_driver_2d.npz stay what they are."""
    rng = np.random.default_rng(500)
    out = {}

    class _Res:
        def set_tracer_vals_all(self, vals, reseat_vals=False):
            self.vals = np.array(vals)

    class _Var:
        def __init__(self, arr):
            self._a = arr

        def __getitem__(self, key):
            return self._a[key]

    class _Precond:
        def __init__(self, variables):
            self.variables = {k: _Var(v) for k, v in variables.items()}

    for tag, (nz, ny) in {"g14x11": (14, 11), "g30x30": (30, 30)}.items():
        depth, ypos, procs = rh.make_py_driver_2d(nz, ny, 19.0, 0.1, 1000.0)
        out[f"{tag}/params"] = np.array([nz, ny, 19.0, 0.1, 1000.0])
        out[f"{tag}/depth_edges"], out[f"{tag}/ypos_edges"] = depth.edges, ypos.edges
        mids = YEAR * (np.arange(3) + 0.5) / 3.0
        # iage
        iage = rh.make_2d_iage(depth, ypos)
        y = rng.normal(size=(2, nz, ny))
        iage.get_tracer_vals_all = lambda y=y: y
        res = _Res()
        iage.apply_precond_jacobian((0.0, YEAR), res, procs)
        out[f"{tag}/iage/y"], out[f"{tag}/iage/precond"] = y, res.vals
        if tag == "g14x11":
            out[f"{tag}/iage/jac_dense_mids"] = np.stack(
                [iage.comp_jacobian(t, y.reshape(-1), procs).toarray() for t in mids])
        # forced o2_like
        from oracle.gen_golden_radau import O2_LIKE

        forced = rh.make_2d_forced(depth, ypos, dict(O2_LIKE))
        forced._tracer_module_def = {"tracers": {"o2_like": {}}}
        ftimes = np.linspace(0.0, YEAR, 61)
        out[f"{tag}/forced/frc_time"] = ftimes
        out[f"{tag}/forced/frc_data"] = np.stack([forced.sms_fcn(t) for t in ftimes])
        y = rng.normal(size=(1, nz, ny))
        # precond file: 61 hist times, tracer snapshots with values below, inside and above (0, sink_thres)
        ptimes = ftimes
        snaps = np.abs(rng.normal(size=(61, nz, ny))) * 0.06 - 0.005
        forced.get_tracer_vals_all = lambda y=y: y
        res = _Res()
        forced.apply_precond_jacobian((0.0, YEAR), res, procs, _Precond({"time": ptimes, "o2_like": snaps}))
        out[f"{tag}/forced/y"], out[f"{tag}/forced/precond"] = y, res.vals
        out[f"{tag}/forced/precond_times"], out[f"{tag}/forced/precond_snaps"] = ptimes, snaps
        if tag == "g14x11":
            jacs = []
            for i, t in enumerate(mids):
                tv = snaps[np.argmin(abs(YEAR * (i + 1.0) / 3.0 - ptimes))].reshape(-1)
                jacs.append(forced.comp_jacobian(t, tv, procs).toarray())
            out[f"{tag}/forced/jac_dense_mids"] = np.stack(jacs)
    np.savez_compressed(os.path.join(OUT, "precond_2d.npz"), **out)
    print("precond_2d.npz:", len(out), "arrays")


def main():
    if not rh.available():
        raise SystemExit("reference tree not found: golden vectors can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    rh.install_stubs()
    baselines()
    baseline_files()
    remap_cases()
    py_driver_2d_cases()
    test_problem_cases()
    forced_restore_file_case()
    precond_2d_cases()


if __name__ == "__main__":
    main()
