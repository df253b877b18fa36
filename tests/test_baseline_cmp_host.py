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


class _AttrVar:
    """netCDF4.Variable look-alike whose __dict__ holds the attributes and nothing else (utils.metadata_same compares
    the two variables' __dict__)"""

    __slots__ = ("_v", "_name", "__dict__")

    def __init__(self, name, var):
        object.__setattr__(self, "_v", var)
        object.__setattr__(self, "_name", name)
        for key, val in var._attributes.items():  # noqa: SLF001
            if isinstance(val, bytes):
                val = val.decode()
            elif isinstance(val, np.ndarray) and val.size == 1:
                val = val.reshape(-1)[0].item()
            self.__dict__[key] = val

    name = property(lambda self: self._name)
    dimensions = property(lambda self: self._v.dimensions)
    shape = property(lambda self: self._v.shape)

    def __getitem__(self, key):
        return np.array(self._v.data, dtype=self._v.data.dtype.newbyteorder("="))[key]


class _CmpDataset:
    """read-only netCDF4.Dataset look-alike over scipy's reader"""

    def __init__(self, fname, mode="r", **kwargs):
        assert mode == "r"
        self._nc = netcdf_file(fname, "r", mmap=False)
        self.variables = {name: _AttrVar(name, var) for name, var in self._nc.variables.items()}
        self.dimensions = {name: range(length or next(v.shape[0] for v in self._nc.variables.values()
                                                     if v.dimensions and v.dimensions[0] == name))
                           for name, length in self._nc.dimensions.items()}

    def set_auto_mask(self, flag):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self._nc.close()
        return False


class _SameUnits:
    """pint.UnitRegistry stand-in: a unit is its string (files whose units agree) or, for the metre-based units of the
    reference's fixtures input/tests/isclose_*.nc, its size in metres — enough for `ureg(u1) != ureg(u2)` and
    `ureg.Quantity(vals, u1).to(u2).magnitude` (utils.py:306-310)"""

    METRES = {"m": 1.0, "cm": 0.01, "km": 1000.0}

    def __call__(self, units):
        return ("length", self.METRES[units]) if units in self.METRES else units

    def Quantity(self, vals, units):  # noqa: N802
        outer = self

        class _Quantity:
            def to(self, other):
                from types import SimpleNamespace

                return SimpleNamespace(magnitude=vals * (outer.METRES[units] / outer.METRES[other]))

        return _Quantity()


def test_verdicts_equal_the_references_own_comparison_functions(tmp_path):
    """the reference's `metadata_same` and `isclose_all_vars` (nk_ooc/utils.py:212-324, imported unmodified; build
    container only) and this package's give the same verdict on every pair: the reference's own fixtures
    input/tests/isclose_{base,same,diff}.nc and edited copies of a baseline file (values, attributes, variables,
    dimensions, fill patterns)"""
    from oracle import ref_harness

    if not ref_harness.available():
        pytest.skip("the reference is not mounted here")
    ref_harness.install_stubs()
    import nk_ooc.utils as ref_utils

    saved = ref_utils.Dataset, ref_utils.UnitRegistry
    ref_utils.Dataset, ref_utils.UnitRegistry = _CmpDataset, _SameUnits
    try:
        pairs = []
        fix = os.path.join(ref_harness.REF_ROOT, "input", "tests")
        for other in ("isclose_same.nc", "isclose_diff.nc", "isclose_base.nc"):
            pairs.append((os.path.join(fix, "isclose_base.nc"), os.path.join(fix, other)))
        root = materialise(str(tmp_path / "baselines"))
        base = os.path.join(root, "ci_short", "fcn_00.nc")
        fill = 9.969209968386869e36

        def edited(tag, edit):
            out = str(tmp_path / f"{tag}.nc")
            with netcdf_file(base, "r", mmap=False) as src, netcdf_file(out, "w", version=2) as dst:
                dims = edit("dims", dict(src.dimensions))
                for name, length in dims.items():
                    dst.createDimension(name, length)
                for name, var in src.variables.items():
                    data = np.array(var.data, dtype=var.data.dtype.newbyteorder("="))
                    name, vdims, data, attrs = edit("var", (name, var.dimensions, data, dict(var._attributes)))  # noqa: SLF001
                    if name is None:
                        continue
                    handle = dst.createVariable(name, data.dtype, vdims)
                    for key, val in attrs.items():
                        setattr(handle, key, val)
                    handle[:] = data
            return out

        def var_edit(fn):
            return lambda kind, arg: fn(*arg) if kind == "var" else arg

        def with_fill(where):
            def fn(n, d, v, a):
                if n == "iage":
                    v = v.copy()
                    v[where] = fill
                    a = dict(a, _FillValue=fill)
                return n, d, v, a
            return var_edit(fn)

        variants = {
            "same": var_edit(lambda n, d, v, a: (n, d, v, a)),
            "value_off": var_edit(lambda n, d, v, a: (n, d, v * (1.0 + 1.0e-4) if n == "iage" else v, a)),
            "value_close": var_edit(lambda n, d, v, a: (n, d, v * (1.0 + 1.0e-9) if n == "po4" else v, a)),
            "attr": var_edit(lambda n, d, v, a: (n, d, v, dict(a, long_name="x") if n == "depth" else a)),
            "extra_attr": var_edit(lambda n, d, v, a: (n, d, v, dict(a, note="x") if n == "iage" else a)),
            "missing_var": var_edit(lambda n, d, v, a: (None if n == "pop_s" else n, d, v, a)),
            "renamed_var": var_edit(lambda n, d, v, a: ("pop_t" if n == "pop_s" else n, d, v, a)),
            "fill3": with_fill(3),
            "fill4": with_fill(4),
            "extra_dim": lambda kind, arg: dict(arg, extra=3) if kind == "dims" else arg,
        }
        files = {tag: edited(tag, edit) for tag, edit in variants.items()}
        for a in ("same", "fill3"):
            for b in files:
                pairs.append((files[a], files[b]))
        verdicts = set()
        for f1, f2 in pairs:
            # (the last three are the tolerances of the reference's tests/test_utils.py:53-75 on its fixtures)
            for rtol, atol in ((1.0e-7, 2.0e-9), (1.0e-3, 1.0e-6), (0.0, 0.0), (1.0e-8, 1.0e-8), (1.0e-5, 1.0e-5)):
                want = (bool(ref_utils.metadata_same(f1, f2)), bool(ref_utils.isclose_all_vars(f1, f2, rtol, atol)))
                got = (bool(utils.metadata_same(f1, f2)), bool(utils.isclose_all_vars(f1, f2, rtol=rtol, atol=atol)))
                assert got == want, (os.path.basename(f1), os.path.basename(f2), rtol, got, want)
                verdicts.add(want)
        assert verdicts == {(True, True), (True, False), (False, True), (False, False)}
    finally:
        ref_utils.Dataset, ref_utils.UnitRegistry = saved


def test_isclose_all_vars_known_answers_of_the_reference_tests(tmp_path):
    """tests/test_utils.py:53-75 on files with the content of input/tests/isclose_{base,same,diff}.nc (restated: two
    variables of three values in m; `same` holds the second in cm, `diff` moves the first by 1e-7), plus the conversion
    factors the comparison rests on"""
    def write(name, var1, var2, units2):
        path = str(tmp_path / name)
        with netcdf_file(path, "w", version=2) as nc:
            nc.createDimension("dim", 3)
            for vname, vals, units in (("var1", var1, "m"), ("var2", var2, units2)):
                var = nc.createVariable(vname, "f8", ("dim",))
                var.units = units
                var[:] = vals
        return path

    base = write("base.nc", [1.0, 2.0, 3.0], [1.0, 2.0, 3.0], "m")
    same = write("same.nc", [1.0, 2.0, 3.0], [100.0, 200.0, 300.0], "cm")
    diff = write("diff.nc", [1.0000001, 2.0000001, 3.0000001], [100.0, 200.0, 300.0], "cm")
    assert utils.isclose_all_vars(base, base, rtol=0.0, atol=0.0)
    assert utils.isclose_all_vars(base, base, rtol=1.0e-5, atol=1.0e-5)
    assert utils.isclose_all_vars(base, same, rtol=0.0, atol=0.0)
    assert utils.isclose_all_vars(base, same, rtol=1.0e-5, atol=1.0e-5)
    assert not utils.isclose_all_vars(base, diff, rtol=0.0, atol=0.0)
    assert not utils.isclose_all_vars(base, diff, rtol=1.0e-8, atol=1.0e-8)
    assert utils.isclose_all_vars(base, diff, rtol=1.0e-5, atol=1.0e-5)
    assert utils.units_conversion_factor("m", "cm") == 100.0
    assert utils.units_conversion_factor("mmol / m^3", "mol / m^3") == 1.0e-3
    assert utils.units_conversion_factor("m / d", "m / s") == pytest.approx(1.0 / 86400.0, rel=1e-15)
    assert utils.units_conversion_factor("years", "d") == 365.25
    assert utils.units_conversion_factor("mmol / m^2 / s", "mol / m^2 / d") == pytest.approx(86.4, rel=1e-15)
    with pytest.raises(ValueError, match="cannot convert"):
        utils.units_conversion_factor("m", "s")
    with pytest.raises(ValueError, match="unknown unit"):
        utils.units_conversion_factor("furlong", "m")
    other = write("other.nc", [1.0, 2.0, 3.0], [1.0, 2.0, 3.0], "s")
    assert not utils.isclose_all_vars(base, other, rtol=1.0e-5, atol=1.0e-5)
