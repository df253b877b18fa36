"""CPU test of the command line's configuration handling (nk_ooc_b200/cli.py): cfg files in the
reference's format (configparser, %(workdir)s interpolation, sections solverinfo / modelinfo),
defaults restating input/<model>/*.cfg, command-line overrides (share.py:11-31)."""
import os

from nk_ooc_b200 import cli


def test_defaults_and_overrides(tmp_path):
    args = cli.parse_args(["setup_solver", "--model_name", "py_driver_2d", "--workdir", str(tmp_path),
                           "--depth_nlevs", "20", "--ypos_nlevs", "3", "--max_abs_vvel", "0.0",
                           "--horiz_mix_coeff", "0.0", "--newton_rel_tol", "1.0e-6"])
    cfg = cli.read_config(args)
    assert cfg["workdir"] == str(tmp_path)
    assert cfg["modelinfo"]["depth_nlevs"] == "20" and cfg["modelinfo"]["ypos_nlevs"] == "3"
    assert cfg["modelinfo"]["depth_delta_ratio_max"] == "19.0"  # input/py_driver_2d/model_params.cfg:15-16
    assert cfg["modelinfo"]["grid_vars_fname"] == os.path.join(str(tmp_path), "grid_vars.nc")
    assert cfg["solverinfo"]["newton_rel_tol"] == "1.0e-6" and cfg["solverinfo"]["krylov_rel_tol"] == "0.01"
    assert cfg["solverinfo"]["init_iterate_fname"].endswith(os.path.join("gen_init_iterate", "init_iterate.nc"))
    cfg = cli.read_config(cli.parse_args(["nk_driver"]))
    assert cfg["modelinfo"]["tracer_module_names"] == "iage,phosphorus"  # input/test_problem/newton_krylov.cfg:58
    assert cfg["solverinfo"]["newton_rel_tol"] == "1.0e-8"


def test_reference_style_cfg_files(tmp_path):
    cfg_a = tmp_path / "newton_krylov.cfg"
    cfg_a.write_text(
        "[DEFAULT]\nmodel_name=test_problem\nworkdir=%(HOME)s/some_work\nno_value_allowed=cfg_fname_out\n"
        "[solverinfo]\nnewton_rel_tol=1.0e-8\nnewton_max_iter=5\npost_newton_fp_iter=1\nkrylov_rel_tol=0.01\n"
        "init_iterate_fname=%(workdir)s/gen_init_iterate/init_iterate.nc\n"
        "[modelinfo]\nreinvoke=True\ngrid_vars_fname=%(workdir)s/depth_axis.nc\ntracer_module_names=iage\n")
    cfg_b = tmp_path / "model_params.cfg"
    cfg_b.write_text("[modelinfo]\npo4_s_restoring_opt=1\ndepth_axisname=depth\ndepth_units=m\ndepth_edge_start=0.0\n"
                     "depth_edge_end=900.0\ndepth_delta_ratio_max=5.0\n")
    args = cli.parse_args(["comp_fcn", "--cfg_fnames", f"{cfg_a},{cfg_b}", "--workdir", str(tmp_path / "w"),
                           "--tracer_module_names", "dye_decay_{suff}:001:010", "--depth_nlevs", "20"])
    cfg = cli.read_config(args)
    assert cfg["workdir"] == str(tmp_path / "w")
    assert cfg["modelinfo"]["grid_vars_fname"] == str(tmp_path / "w" / "depth_axis.nc")
    assert cfg["solverinfo"]["init_iterate_fname"] == str(tmp_path / "w" / "gen_init_iterate" / "init_iterate.nc")
    assert cfg["modelinfo"]["tracer_module_names"] == "dye_decay_{suff}:001:010"
    assert cfg["modelinfo"]["depth_edge_end"] == "900.0" and cfg["modelinfo"]["depth_nlevs"] == "20"
