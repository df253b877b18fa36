"""File-to-file command line of the B200 path — the reference's three entry
scripts in one module:

    python -m nk_ooc_b200.cli setup_solver --model_name py_driver_2d --workdir W [--fp_cnt N] [overrides]
    python -m nk_ooc_b200.cli nk_driver    --model_name py_driver_2d --workdir W [overrides]
    python -m nk_ooc_b200.cli comp_fcn | gen_precond_jacobian | apply_precond_jacobian
                                           --model_name ... --in_fname ... --res_fname ... [--hist_fname ...]

* setup_solver = nk_ooc/<model>/setup_solver.py:64-161 (grid file, init iterate from
  `gen_init_iterate` or a file, `fp_cnt` fixed-point iterations x += F(x), init_iterate.nc);
* nk_driver    = nk_ooc/nk_driver.py (Newton's method to convergence) on the device-resident
  solver of nk_ooc_b200/solver.py, writing the reference's file names into the work directory;
* the three model-state commands = nk_ooc/run_cmd.py:45-87.

Configuration: the reference's cfg files can be passed as they are (`--cfg_fnames a.cfg,b.cfg`,
configparser with the HOME / USER / repo_root defaults of nk_ooc/share.py:99-125; sections
DEFAULT / solverinfo / modelinfo).  Without them the defaults below restate
input/<model>/newton_krylov.cfg and model_params.cfg.  The command-line overrides are those of
share.py:11-31 (`--workdir`, `--tracer_module_names`, `--newton_rel_tol`, `--newton_max_iter`,
`--init_iterate_fname`) plus the grid sizes the reference's setup scripts take.
State, hist, precond and grid files are the reference's NETCDF3_64BIT_OFFSET layouts.
"""

import argparse
import configparser
import logging
import os
import sys

# input/<model>/newton_krylov.cfg [solverinfo] (lines 32-43) and [modelinfo], model_params.cfg
DEFAULTS = {
    "test_problem": {
        "solverinfo": {"newton_rel_tol": "1.0e-8", "newton_max_iter": "5", "post_newton_fp_iter": "1",
                       "krylov_rel_tol": "0.01"},
        "modelinfo": {"tracer_module_names": "iage,phosphorus", "po4_s_restoring_opt": "1", "depth_axisname": "depth",
                      "depth_units": "m", "depth_nlevs": "30", "depth_edge_start": "0.0", "depth_edge_end": "900.0",
                      "depth_delta_ratio_max": "5.0", "reinvoke": "False"},
        "grid_vars": "depth_axis.nc",
    },
    "py_driver_2d": {
        "solverinfo": {"newton_rel_tol": "1.0e-5", "newton_max_iter": "5", "post_newton_fp_iter": "1",
                       "krylov_rel_tol": "0.01"},
        "modelinfo": {"tracer_module_names": "iage", "depth_axisname": "depth", "depth_units": "m",
                      "depth_edge_start": "0.0", "depth_edge_end": "4000.0", "depth_nlevs": "40",
                      "depth_delta_ratio_max": "19.0", "ypos_axisname": "ypos", "ypos_units": "m",
                      "ypos_edge_start": "0.0", "ypos_edge_end": "50.0e5", "ypos_delta_ratio_max": "1.0",
                      "ypos_nlevs": "50", "max_abs_vvel": "0.1", "horiz_mix_coeff": "1000.0", "reinvoke": "False"},
        "grid_vars": "grid_vars.nc",
    },
}
SOLVER_OVERRIDES = ("newton_rel_tol", "newton_max_iter", "krylov_rel_tol", "post_newton_fp_iter", "init_iterate_fname")
MODEL_OVERRIDES = ("tracer_module_names", "depth_nlevs", "depth_delta_ratio_max", "ypos_nlevs", "max_abs_vvel",
                   "horiz_mix_coeff", "po4_s_restoring_opt")


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description=__doc__.split("\n")[0],
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("cmd", choices=["setup_solver", "nk_driver", "comp_fcn", "gen_precond_jacobian",
                                        "apply_precond_jacobian"])
    parser.add_argument("--model_name", default="test_problem", choices=sorted(DEFAULTS))
    parser.add_argument("--cfg_fnames", default=None, help="comma-separated cfg files of the reference's format")
    parser.add_argument("--workdir", default=None)
    for name in SOLVER_OVERRIDES + MODEL_OVERRIDES:
        parser.add_argument(f"--{name}", default=None, help=f"override {name} from the cfg / defaults")
    parser.add_argument("--resume", action="store_true", help="nk_driver: resume from the saved solver state")
    parser.add_argument("--rewind", action="store_true", help="nk_driver: with --resume, redo the last logged step")
    parser.add_argument("--persist", action="store_true", help="accepted for compatibility (there is no re-invocation)")
    parser.add_argument("--fp_cnt", type=int, default=2, help="fixed-point iterations applied to the init iterate")
    parser.add_argument("--init_iterate_opt", default="gen_init_iterate",
                        help="setup_solver: initial iterate (gen_init_iterate, zeros or a file name)")
    parser.add_argument("--deprecation_warning_to_error", action="store_true",
                        help="treat DeprecationWarning warnings as errors")
    parser.add_argument("--armijo_batch", type=int, default=1, help="speculative Armijo candidates per evaluation")
    parser.add_argument("--fname_dir", default=".", help="directory that relative fname arguments are relative to")
    parser.add_argument("--hist_fname", default=None)
    parser.add_argument("--precond_fname", default=None)
    parser.add_argument("--in_fname", default=None)
    parser.add_argument("--res_fname", default=None)
    return parser.parse_args(argv)


def read_config(args):
    """{"workdir", "solverinfo", "modelinfo"} from cfg files (share.py:99-125) or the defaults"""
    solverinfo = dict(DEFAULTS[args.model_name]["solverinfo"])
    modelinfo = dict(DEFAULTS[args.model_name]["modelinfo"])
    workdir = args.workdir
    if args.cfg_fnames:
        defaults = {key: os.environ.get(key, "") for key in ("HOME", "USER")}
        defaults["repo_root"] = os.environ.get("NK_REF_ROOT", os.getcwd())
        config = configparser.ConfigParser(defaults, allow_no_value=True)
        if not config.read(args.cfg_fnames.split(",")):
            raise RuntimeError(f"cfg_fnames not read: {args.cfg_fnames}")
        if args.workdir is not None:
            config["DEFAULT"]["workdir"] = args.workdir
        workdir = config["DEFAULT"].get("workdir", workdir)
        for section, target in (("solverinfo", solverinfo), ("modelinfo", modelinfo)):
            if config.has_section(section):
                target.update({k: v for k, v in config[section].items() if v is not None})
    if workdir is None:
        workdir = os.path.join(os.environ.get("HOME", "."), f"{args.model_name}_work")
    for name in SOLVER_OVERRIDES:
        if getattr(args, name) is not None:
            solverinfo[name] = getattr(args, name)
    for name in MODEL_OVERRIDES:
        if getattr(args, name) is not None:
            modelinfo[name] = getattr(args, name)
    modelinfo["model_name"] = args.model_name
    modelinfo.setdefault("grid_vars_fname", os.path.join(workdir, DEFAULTS[args.model_name]["grid_vars"]))
    solverinfo.setdefault("init_iterate_fname", os.path.join(workdir, "gen_init_iterate", "init_iterate.nc"))
    return {"workdir": workdir, "solverinfo": solverinfo, "modelinfo": modelinfo}


def _model_state_class(model_name):
    """model_state_base.get_model_state_class (nk_ooc/model_state_base.py:627-646)"""
    if model_name == "py_driver_2d":
        from .py_driver_2d.model_state import ModelState
    else:
        from .test_problem.model_state import ModelState
    return ModelState


def _gen_grid_file(config):
    modelinfo = config["modelinfo"]
    os.makedirs(os.path.dirname(os.path.abspath(modelinfo["grid_vars_fname"])), exist_ok=True)
    if modelinfo["model_name"] == "py_driver_2d":
        from .py_driver_2d.setup_solver import gen_grid_vars_file

        gen_grid_vars_file(modelinfo)
    else:
        from .spatial_axis import spatial_axis_from_defn
        from .test_problem.model_state import gen_depth_axis_file

        kw = {"axisname": modelinfo.get("depth_axisname", "depth"), "units": modelinfo.get("depth_units", "m"),
              "nlevs": int(modelinfo["depth_nlevs"]), "edge_start": float(modelinfo["depth_edge_start"]),
              "edge_end": float(modelinfo["depth_edge_end"]),
              "delta_ratio_max": float(modelinfo["depth_delta_ratio_max"])}
        gen_depth_axis_file(modelinfo, spatial_axis_from_defn(**kw))


def setup_solver(config, fp_cnt, init_iterate_src="gen_init_iterate"):
    """nk_ooc/<model>/setup_solver.py:64-161"""
    logger = logging.getLogger(__name__)
    _gen_grid_file(config)
    cls = _model_state_class(config["modelinfo"]["model_name"])
    cls.configure(config["modelinfo"])
    caller = "nk_ooc_b200.cli.setup_solver"
    init_iterate = cls(init_iterate_src)
    init_dir = os.path.join(config["workdir"], "gen_init_iterate")
    # py_driver_2d/setup_solver.py:116-122 numbers these files with four digits, test_problem/setup_solver.py:146-152 with two
    width = 4 if config["modelinfo"]["model_name"] == "py_driver_2d" else 2
    for fp_iter in range(fp_cnt):
        logger.info("fp_iter=%d", fp_iter)
        init_iterate.dump(os.path.join(init_dir, f"init_iterate_{fp_iter:0{width}}.nc"), caller)
        fcn = init_iterate.comp_fcn(os.path.join(init_dir, f"fcn_{fp_iter:0{width}}.nc"), None,
                                    os.path.join(init_dir, f"hist_{fp_iter:0{width}}.nc"))
        init_iterate += fcn
        init_iterate.copy_shadow_tracers_to_real_tracers()
    init_iterate.dump(config["solverinfo"]["init_iterate_fname"], caller)
    return init_iterate


def nk_driver(config, armijo_batch=1, resume=False, rewind=False):
    """nk_ooc/nk_driver.py: Newton's method from solverinfo["init_iterate_fname"] to convergence;
    resume continues an interrupted solve at the step where it stopped (Newton_state.json / Krylov_state.json step
    logs), rewind redoes the last logged step first (nk_driver.py --resume / --rewind)"""
    from .solver import NewtonSolver

    cls = _model_state_class(config["modelinfo"]["model_name"])
    if cls.model_config_obj is None:
        cls.configure(config["modelinfo"])
    iterate = cls(config["solverinfo"]["init_iterate_fname"])
    solver = NewtonSolver(iterate, config["solverinfo"], workdir=config["workdir"], armijo_batch=armijo_batch,
                          resume=resume, rewind=rewind)
    solver.solve()
    return solver


def run_cmd(config, args):
    """nk_ooc/run_cmd.py:45-87"""

    def resolve(fname):
        return fname if fname is None or os.path.isabs(fname) else os.path.join(args.fname_dir, fname)

    cls = _model_state_class(config["modelinfo"]["model_name"])
    if cls.model_config_obj is None:
        cls.configure(config["modelinfo"])
    if args.cmd == "gen_precond_jacobian":
        ms_in = cls(resolve(args.in_fname)) if args.in_fname else cls("zeros")
        ms_in.gen_precond_jacobian(resolve(args.hist_fname), resolve(args.precond_fname), solver_state=None)
        return None
    ms_in = cls(resolve(args.in_fname))
    ms_in.log("state_in")
    if args.cmd == "comp_fcn":
        res = ms_in.comp_fcn(resolve(args.res_fname), None, hist_fname=resolve(args.hist_fname))
        res.log("fcn")
    else:
        res = ms_in.apply_precond_jacobian(resolve(args.precond_fname), resolve(args.res_fname), None)
        res.log("precond_res")
    return res


def main(argv=None):
    args = parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s:%(name)s:%(message)s")
    config = read_config(args)
    os.makedirs(config["workdir"], exist_ok=True)
    if args.deprecation_warning_to_error:
        import warnings

        warnings.filterwarnings("error", category=DeprecationWarning)
    if args.cmd == "setup_solver":
        setup_solver(config, args.fp_cnt, args.init_iterate_opt)
    elif args.cmd == "nk_driver":
        solver = nk_driver(config, args.armijo_batch, resume=args.resume, rewind=args.rewind)
        logging.getLogger(__name__).info("converged after %d Newton iterations", solver.iteration)
    else:
        run_cmd(config, args)
    return 0


if __name__ == "__main__":
    sys.exit(main())
