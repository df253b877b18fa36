"""TEST INFRASTRUCTURE: F(x) = x(T) - x(0) golden vectors from the REFERENCE's own integrator call.

Runs only in the build container (needs /root/reference).  Every case builds the reference's own
tracer-module object through oracle/ref_harness.py and integrates it EXACTLY as the reference does
(py_driver_2d/model_state.py:102-114: scipy solve_ivp "Radau", max_step = T/100, analytic
comp_jacobian + comp_jacobian_sparsity; test_problem/model_state.py:83-92: no Jacobian, rtol = atol =
1e-12), twice: at the reference's own tolerance (rtol = atol = 1e-6 for py_driver_2d) so that the
reference's own integration error is on record, and at rtol = atol = 1e-9 as the truth the GPU path is
compared with.  The o2_like sink record is the reference's input/py_driver_2d/po4_sms.nc read through
the reference's own utils.gen_forcing_fcn (which interpolates it onto the case's grid) with the options
of scripts/run_py_driver_2d_forced_o2_like.sh:14-25.

    python -m oracle.gen_golden_radau [case ...]      # from the repo root; cases run in parallel

Writes tests/golden/radau_<case>.npz (one small file per case so that cases can be regenerated
independently).  Wall times of the reference integrations are stored too (cpu seconds, this container).
"""

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
YEAR = 365.0 * 86400.0

GRIDS = {  # nz, ny, depth delta_ratio_max (input/py_driver_2d/model_params.cfg; CI overrides)
    "g14x11": (14, 11, 19.0),
    "g30x30": (30, 30, 19.0),
    "g40x50": (40, 50, 19.0),
    "g80x100": (80, 100, 9.0),
    "g125x150": (125, 150, 11.8),
}

O2_LIKE = {  # scripts/run_py_driver_2d_forced_o2_like.sh:14-25
    "forced_surf_restore_opt": "const", "forced_surf_restore_const": "1.0",
    "forced_surf_restore_rate_10m": "1.0 / 3600.0", "forced_sms_opt": "file",
    "forced_sms_fname": os.path.join(rh.REF_ROOT, "input", "py_driver_2d", "po4_sms.nc"),
    "forced_sms_varname": "po4_sms", "forced_sms_scalef": "-1.0 / 3.0", "forced_sink_thres": "0.05",
}


def init_state(module, depth, ypos, seed):
    """gen_init_iterate profiles (input/py_driver_2d/tracer_module_defs.yaml:7-48) made 2-D and
    perturbed smoothly so that no symmetry hides an indexing error.  forced: the o2_like field is
    given an oxygen-minimum-like interior (values down to ~0.01, below sink_thres = 0.05) so that the
    sink limiter of forced.py:140-152 switches on AND off during the year."""
    rng = np.random.default_rng(seed)
    nz, ny = len(depth), len(ypos)
    zz = depth.mid[:, None]
    yy = (ypos.mid / ypos.edges[-1])[None, :]
    wob = 1.0 + 0.1 * np.sin(2 * np.pi * yy) * np.cos(np.pi * zz / 4000.0) + 0.01 * rng.standard_normal((nz, ny))
    if module == "forced":
        omz = np.exp(-(((zz - 600.0) / 500.0) ** 2)) * (0.3 + 0.7 * np.sin(np.pi * yy) ** 2)
        x = (1.0 - 0.99 * omz) * wob
        return np.maximum(x, 0.004)[None]
    if module == "iage":
        prof = np.interp(depth.mid, [55.0, 200.0], [0.0, 2.0])[:, None]
        return np.stack([prof * wob, prof * wob * 1.1])
    cols = [
        np.interp(depth.mid, [1.3e2, 2.6e2], [5.5e-3, 4.1]),
        np.interp(depth.mid, [9.5e1, 1.4e2], [7.1e-2, 1.5e-4]),
        np.interp(depth.mid, [1.7e2, 2.5e2], [1.8e-2, 7.9e-4]),
    ]
    return np.stack([c[:, None] * wob for c in cols])


def make_module(module, depth, ypos):
    if module == "forced":
        return rh.make_2d_forced(depth, ypos, dict(O2_LIKE))
    if module == "iage":
        return rh.make_2d_iage(depth, ypos)
    return rh.make_2d_phosphorus(depth, ypos)


def radau_2d(tm, procs, x0, tol, t_eval=None):
    """the reference's call, py_driver_2d/model_state.py:95-121"""
    from scipy import integrate

    flat0 = x0.reshape(-1).copy()
    sparsity = tm.comp_jacobian_sparsity(0.0, flat0, procs)
    t0 = time.process_time()
    sol = integrate.solve_ivp(
        tm.comp_tend, (0.0, YEAR), flat0, "Radau", t_eval, max_step=YEAR * 0.01, atol=tol, rtol=tol,
        args=(procs,), jac=tm.comp_jacobian, jac_sparsity=sparsity,
    )
    cpu = time.process_time() - t0
    assert sol.status == 0, sol.message
    return sol, cpu


def case_2d(grid, module, tols=(1.0e-6, 1.0e-9)):
    rh.install_stubs()
    nz, ny, ratio = GRIDS[grid]
    depth, ypos, procs = rh.make_py_driver_2d(nz, ny, ratio, 0.1, 1000.0)
    tm = make_module(module, depth, ypos)
    x0 = init_state(module, depth, ypos, seed=nz * 131 + ny * 7 + len(module))
    out = {"params": np.array([nz, ny, ratio, 0.1, 1000.0]), "depth_edges": depth.edges, "ypos_edges": ypos.edges,
           "x0": x0}
    if module == "forced":
        # the record as the reference's gen_forcing_fcn hands it to comp_tend: on the model grid, scalef applied
        ftimes = np.linspace(0.0, YEAR, 61)
        out["frc_time"] = ftimes
        if nz * ny <= 2000:
            out["frc_data"] = np.stack([tm.sms_fcn(t) for t in ftimes])
        else:
            # too large for a fixture: the test rebuilds the record from the file's native 40 x 50 grid (stored in
            # radau_g40x50_forced.npz) with the product's own forcing reader (the mirror of utils.gen_forcing_fcn)
            out["frc_from"] = np.array("g40x50")
        # the file's own time axis must be those 61 points for the record above to be the whole forcing
        nc = rh.read_nc(O2_LIKE["forced_sms_fname"])
        assert np.allclose(nc["time"], ftimes, rtol=0, atol=1e-6 * YEAR), "po4_sms.nc time axis is not linspace(0, T, 61)"
    for tol in tols:
        tag = f"tol{tol:.0e}"
        t_eval = np.linspace(0.0, YEAR, 61) if tol == tols[-1] else None
        sol, cpu = radau_2d(tm, procs, x0, tol, t_eval)
        out[f"{tag}/fcn"] = (sol.y[:, -1] - x0.reshape(-1)).reshape(x0.shape)
        out[f"{tag}/cpu_s"] = np.array(cpu)
        out[f"{tag}/nfev_njev_nlu"] = np.array([sol.nfev, sol.njev, sol.nlu])
        if t_eval is not None:  # a few snapshots of the truth run (61 hist times: 0, 15, 18, 21, 30, 42, 60)
            keep = [15, 18, 21, 30, 42] if nz * ny <= 2000 else [18, 42]
            out[f"{tag}/snap_idx"] = np.array(keep)
            out[f"{tag}/snaps"] = sol.y[:, keep].T.reshape((len(keep),) + x0.shape)
        print(f"{grid}/{module} tol {tol:.0e}: {cpu:.1f} cpu-s, nfev {sol.nfev} njev {sol.njev} nlu {sol.nlu}, "
              f"max|F| {np.abs(out[f'{tag}/fcn']).max():.3e}", flush=True)
    if len(tols) > 1:
        d = np.abs(out[f"tol{tols[0]:.0e}/fcn"] - out[f"tol{tols[-1]:.0e}/fcn"])
        print(f"{grid}/{module}: reference's own error at its tolerance: max |dF| {d.max():.3e}", flush=True)
    np.savez_compressed(os.path.join(OUT, f"radau_{grid}_{module}.npz"), **out)


def case_tp(name):
    """test_problem module F(x) with the reference's call (test_problem/model_state.py:83-92): Radau,
    rtol = atol = 1e-12, no Jacobian; x0 = the reference's gen_init_iterate profile for the module"""
    from scipy import integrate

    rh.install_stubs()
    kind = "dye_decay" if name.startswith("dye_decay") else name
    nz = 20
    depth, vert_mix = rh.make_test_problem(nz)
    tm = rh.make_tp_module(kind, depth, name)
    if kind == "dye_decay":
        tm.suff = name.split("_")[-1]
    rng = np.random.default_rng(77)
    if kind == "phosphorus":
        raise SystemExit("test_problem phosphorus F(x) is pinned by baselines/ci_short")
    # input/test_problem/tracer_module_defs.yaml: iage [55, 200] -> [0, 2]; dye_decay: zeros -> use a
    # smooth positive profile so that decay and mixing both act on it
    x0 = (np.interp(depth.mid, [55.0, 200.0], [0.0, 2.0]) + 0.05 * rng.random(nz))[None]
    out = {"depth_edges": depth.edges, "x0": x0}
    for tol in (1.0e-12,):
        t0 = time.process_time()
        sol = integrate.solve_ivp(tm.comp_tend, (0.0, YEAR), x0.reshape(-1), "Radau", np.linspace(0.0, YEAR, 101),
                                  atol=tol, rtol=tol, args=(vert_mix,))
        cpu = time.process_time() - t0
        assert sol.status == 0
        out["fcn"] = (sol.y[:, -1] - x0.reshape(-1)).reshape(x0.shape)
        out["hist"] = sol.y.T.reshape((101,) + x0.shape)[::10]
        out["cpu_s"] = np.array(cpu)
        print(f"tp20/{name}: {cpu:.1f} cpu-s, nfev {sol.nfev}, max|F| {np.abs(out['fcn']).max():.3e}", flush=True)
    np.savez_compressed(os.path.join(OUT, f"radau_tp20_{name}.npz"), **out)


DEFAULT = ["g14x11/iage", "g14x11/forced", "g14x11/phosphorus", "g30x30/forced", "g30x30/phosphorus",
           "tp20/dye_decay_010", "tp20/dye_decay_001"]


def run_case(spec):
    os.environ["OMP_NUM_THREADS"] = "1"
    grid, module = spec.split("/")
    if grid == "tp20":
        case_tp(module)
    else:
        tols = (1.0e-6, 1.0e-9)
        if ":" in module:  # e.g. g80x100/forced:1e-6  (one tolerance only)
            module, t = module.split(":")
            tols = (float(t),)
        case_2d(grid, module, tols)
    return spec


def main():
    if not rh.available():
        raise SystemExit("reference tree not found: golden vectors can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    cases = sys.argv[1:] or DEFAULT
    if len(cases) == 1:
        run_case(cases[0])
        return
    import multiprocessing as mp

    with mp.get_context("spawn").Pool(min(len(cases), os.cpu_count() or 1)) as pool:
        for spec in pool.imap_unordered(run_case, cases):
            print("done", spec, flush=True)


if __name__ == "__main__":
    main()
