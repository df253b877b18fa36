"""CPU tests of the baseline_cmp port (nk_ooc_b200/baseline_cmp.py, nk_ooc_b200/utils.py; reference
nk_ooc/baseline_cmp.py:30-49, nk_ooc/utils.py:186-324) and of the baseline fixture that the CI-script
tests compare against on the GPU box."""
import importlib.util
import os
import shutil
import sys

import numpy as np
import pytest
from scipy.io import netcdf_file

from baseline_files import materialise

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "newton-krylov_ooc_b200", "nk_ooc_b200")
REF_BASELINES = "/root/reference/baselines"


def _load(name):
    """utils.py / baseline_cmp.py are numpy + scipy only: load them without importing the package
    (whose __init__ needs torch and the CUDA library)"""
    import types

    if "nkb_host" not in sys.modules:
        pkg = types.ModuleType("nkb_host")
        pkg.__path__ = [PKG]
        sys.modules["nkb_host"] = pkg
    spec = importlib.util.spec_from_file_location(f"nkb_host.{name}", os.path.join(PKG, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[f"nkb_host.{name}"] = mod
    spec.loader.exec_module(mod)
    return mod


utils = _load("utils")
baseline_cmp = _load("baseline_cmp")


def test_units_strings_equal_the_baselines_spelling():
    """the unit strings pint gives the reference (hist files of baselines/ci_short, ci_py_driver_2d_iage)"""
    assert utils.units_product("years", "m") == "years m"
    assert utils.units_product("years", "m", "m") == "years m^2"
    assert utils.units_product("mmol / m^3", "m") == "mmol / m^2"
    assert utils.units_product("mmol / m^3 / s", "m") == "mmol / m^2 / s"
    assert utils.units_product("1 / s", "m") == "m / s"
    assert utils.units_str_format("m^2 / s") == "m^2 / s"
    assert utils.units_str_format("mmol / s / m^2") == "mmol / m^2 / s"  # utils.py:201-204: time unit last
    assert utils.units_str_format("(years) (m)") == "years m"


def test_materialised_baselines_equal_the_reference_files(tmp_path):
    """every variable, dimension and attribute of every baseline file survives the fixture round trip
    (checked with the port's own metadata_same / isclose_all_vars at zero tolerance)"""
    if not os.path.isdir(REF_BASELINES):
        pytest.skip("reference baselines not available here")
    root = materialise(str(tmp_path / "baselines"))
    n = 0
    for cfg in sorted(os.listdir(REF_BASELINES)):
        for fname in sorted(os.listdir(os.path.join(REF_BASELINES, cfg))):
            if fname.endswith(".nc"):
                assert baseline_cmp.compare(fname, os.path.join(root, cfg), os.path.join(REF_BASELINES, cfg), 0.0, 0.0), \
                    (cfg, fname)
                n += 1
            else:
                assert open(os.path.join(root, cfg, fname)).read() == open(os.path.join(REF_BASELINES, cfg, fname)).read()
    assert n == 31


def test_baseline_cmp_detects_value_metadata_and_fill_differences(tmp_path):
    root = materialise(str(tmp_path / "baselines"))
    base = os.path.join(root, "ci_short")
    expr = str(tmp_path / "expr")
    os.makedirs(expr)

    def rewrite(fname, edit):
        """copy of the baseline with one edit applied"""
        with netcdf_file(os.path.join(base, fname), "r", mmap=False) as src, \
                netcdf_file(os.path.join(expr, fname), "w", version=2) as dst:
            for name, length in src.dimensions.items():
                dst.createDimension(name, length)
            for name, var in src.variables.items():
                dims, data = var.dimensions, np.array(var.data, dtype=var.data.dtype.newbyteorder("="))
                attrs = dict(var._attributes)
                name, dims, data, attrs = edit(name, dims, data, attrs)
                if name is None:
                    continue
                out = dst.createVariable(name, data.dtype, dims)
                for key, val in attrs.items():
                    setattr(out, key, val)
                out[:] = data

    same = lambda n, d, v, a: (n, d, v, a)  # noqa: E731
    rewrite("fcn_00.nc", same)
    assert baseline_cmp.compare("fcn_00.nc", expr, base)
    # a value off by more than rtol 1e-7 / atol 2e-9, and inside a looser tolerance
    rewrite("fcn_00.nc", lambda n, d, v, a: (n, d, v * (1.0 + 1.0e-4) if n == "iage" else v, a))
    assert not baseline_cmp.compare("fcn_00.nc", expr, base)
    assert baseline_cmp.compare("fcn_00.nc", expr, base, rtol=1.0e-3)
    # an attribute, a missing variable, a changed dimension
    rewrite("fcn_00.nc", lambda n, d, v, a: (n, d, v, dict(a, units="cm") if n == "depth" else a))
    assert not utils.metadata_same(os.path.join(expr, "fcn_00.nc"), os.path.join(base, "fcn_00.nc"))
    rewrite("fcn_00.nc", lambda n, d, v, a: (None if n == "pop_s" else n, d, v, a))
    assert not utils.metadata_same(os.path.join(expr, "fcn_00.nc"), os.path.join(base, "fcn_00.nc"))
    assert utils.isclose_all_vars(os.path.join(expr, "fcn_00.nc"), os.path.join(base, "fcn_00.nc"), 1e-7, 2e-9)
    # _FillValue pattern (utils.py:282-287)
    fill = 9.969209968386869e36

    def with_fill(where):
        def edit(n, d, v, a):
            if n == "iage":
                v = v.copy()
                v[where] = fill
                a = dict(a, _FillValue=fill)
            return n, d, v, a
        return edit

    shutil.rmtree(expr)
    os.makedirs(expr)
    base2 = str(tmp_path / "base2")
    os.makedirs(base2)
    rewrite("fcn_00.nc", with_fill(3))
    shutil.move(os.path.join(expr, "fcn_00.nc"), os.path.join(base2, "fcn_00.nc"))
    rewrite("fcn_00.nc", with_fill(3))
    assert baseline_cmp.compare("fcn_00.nc", expr, base2)
    rewrite("fcn_00.nc", with_fill(4))
    assert not baseline_cmp.compare("fcn_00.nc", expr, base2)


def test_baseline_cmp_command_line_exit_status(tmp_path):
    """`python -m nk_ooc_b200.baseline_cmp` exits 0 / 1 like the reference's module (baseline_cmp.py:49)"""
    root = materialise(str(tmp_path / "baselines"))
    args = baseline_cmp.parse_args(["--fname", "depth_axis.nc", "--expr_dir", os.path.join(root, "ci_short"),
                                    "--baseline_dir", os.path.join(root, "ci_short")])
    assert (args.rtol, args.atol) == (1.0e-7, 2.0e-9)
    with pytest.raises(SystemExit) as exc:
        baseline_cmp.main(args)
    assert exc.value.code == 0
    args = baseline_cmp.parse_args(["--fname", "init_iterate.nc", "--expr_dir", os.path.join(root, "ci_short"),
                                    "--baseline_dir", os.path.join(root, "ci_py_driver_2d_iage")])
    with pytest.raises(SystemExit) as exc:
        baseline_cmp.main(args)
    assert exc.value.code == 1


def test_anomaly_variables_compared_at_the_parent_variables_tolerance(tmp_path):
    """--anom_suffix (extension used by the py_driver_2d script ports): x_time_anom = x - mean(x) is compared with
    atol + rtol |x|; without the option the plain comparison fails where the anomaly is small"""
    def write(fname, x):
        with netcdf_file(fname, "w", version=2) as nc:
            nc.createDimension("time", x.shape[0])
            nc.createDimension("depth", x.shape[1])
            var = nc.createVariable("x", "f8", ("time", "depth"))
            var[:] = x
            var = nc.createVariable("x_time_anom", "f8", ("time", "depth"))
            var[:] = x - x.mean(axis=0)

    rng = np.random.default_rng(0)
    x = 5.0 + 0.01 * rng.normal(size=(7, 4))
    base, expr = str(tmp_path / "b.nc"), str(tmp_path / "e.nc")
    write(base, x)
    write(expr, x * (1.0 + 2.0e-4 * rng.normal(size=x.shape)))  # x agrees to rtol 1e-3, its small anomaly does not
    assert not utils.isclose_all_vars(expr, base, rtol=1.0e-3, atol=1.0e-6)
    assert utils.isclose_all_vars(expr, base, rtol=1.0e-3, atol=1.0e-6, anom_suffix="_time_anom")
    write(expr, x * (1.0 + 5.0e-3))  # x itself off: fails either way
    assert not utils.isclose_all_vars(expr, base, rtol=1.0e-3, atol=1.0e-6, anom_suffix="_time_anom")
