"""TEST INFRASTRUCTURE (build-container only): import the reference's numerical modules.

The reference (klindsay28/Newton-Krylov_OOC, mounted read-only at /root/reference) needs
netCDF4, xarray and pint, none of which are installed here.  Its numerical modules
(spatial_axis, py_driver_2d/{advection,horiz_mix,vert_mix,iage,forced,phosphorus},
test_problem/{vert_mix,iage,dye_decay,phosphorus}) run unmodified once those three names
are stubbed.  This harness installs the stubs and builds reference objects with
``object.__new__`` plus the handful of attributes their numeric methods read.

It is used ONLY by ``oracle/gen_golden.py`` (which writes tests/golden/*.npz in this
container) and by tests that skip when /root/reference is absent.  Nothing on the GPU box
may import it: /root/reference does not exist there.
"""

import os
import sys
import types

import numpy as np
from scipy.io import netcdf_file

REF_ROOT = os.environ.get("NK_REF_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "nk_ooc"))


class _Var:
    """minimal netCDF4.Variable look-alike over scipy.io.netcdf_file"""

    def __init__(self, var):
        self._var = var
        self.dimensions = var.dimensions
        self.shape = var.shape
        self.ndim = len(var.shape)
        for key, val in var._attributes.items():
            if isinstance(val, bytes):
                val = val.decode()
            setattr(self, key, val)

    def __getitem__(self, key):
        return np.array(self._var.data, dtype=self._var.data.dtype.newbyteorder("="))[key]

    def __len__(self):
        return self.shape[0]


class _Dataset:
    """read-only netCDF4.Dataset look-alike (enough for gen_forcing_fcn & friends)"""

    def __init__(self, fname, mode="r", **kwargs):
        if mode != "r":
            raise NotImplementedError("stub Dataset is read-only")
        self._nc = netcdf_file(fname, "r", mmap=False)
        self.variables = {k: _Var(v) for k, v in self._nc.variables.items()}
        self.dimensions = dict(self._nc.dimensions)

    def set_auto_mask(self, flag):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self._nc.close()
        return False


def install_stubs():
    if "netCDF4" not in sys.modules:
        mod = types.ModuleType("netCDF4")
        mod.Dataset = _Dataset
        mod.default_fillvals = {"f8": 9.969209968386869e36, "i4": -2147483647}
        sys.modules["netCDF4"] = mod
    if "xarray" not in sys.modules:
        mod = types.ModuleType("xarray")
        mod.Dataset = object
        mod.DataArray = object
        sys.modules["xarray"] = mod
    if "pint" not in sys.modules:
        mod = types.ModuleType("pint")

        class UnitRegistry:  # pylint: disable=too-few-public-methods
            def __init__(self, *a, **k):
                pass

        mod.UnitRegistry = UnitRegistry
        sys.modules["pint"] = mod
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def make_axis(axisname, **defn):
    """reference SpatialAxis from a defn dict (spatial_axis.py:225-250)"""
    install_stubs()
    from nk_ooc.spatial_axis import spatial_axis_defn_dict, spatial_axis_from_defn_dict

    return spatial_axis_from_defn_dict(
        defn_dict=spatial_axis_defn_dict(axisname=axisname, **defn)
    )


def make_py_driver_2d(nz, ny, depth_ratio, max_abs_vvel=0.1, horiz_mix_coeff=1000.0,
                      depth_end=4000.0, ypos_end=50.0e5):
    """reference py_driver_2d process objects on the grid of model_params.cfg"""
    install_stubs()
    from nk_ooc.py_driver_2d.advection import Advection
    from nk_ooc.py_driver_2d.horiz_mix import HorizMix
    from nk_ooc.py_driver_2d.vert_mix import VertMix

    depth = make_axis("depth", units="m", nlevs=nz, edge_start=0.0, edge_end=depth_end,
                      delta_ratio_max=depth_ratio)
    ypos = make_axis("ypos", units="m", nlevs=ny, edge_start=0.0, edge_end=ypos_end,
                     delta_ratio_max=1.0)
    modelinfo = {"max_abs_vvel": repr(max_abs_vvel), "horiz_mix_coeff": repr(horiz_mix_coeff)}
    processes = {}
    processes["advection"] = Advection(depth, ypos, modelinfo)
    processes["horiz_mix"] = HorizMix(depth, ypos, modelinfo)
    processes["vert_mix"] = VertMix(depth, ypos)
    return depth, ypos, processes


def make_2d_iage(depth, ypos):
    install_stubs()
    from nk_ooc.py_driver_2d.iage import iage

    tm = object.__new__(iage)
    tm.name = "iage"
    tm.tracer_cnt = 2
    tm.depth = depth
    tm.ypos = ypos
    tm.surf_restore_rate = 24.0 / 86400.0 * 10.0 / depth.delta[0]
    tm.surf_slow_factor = 0.01
    return tm


def make_2d_phosphorus(depth, ypos):
    install_stubs()
    from nk_ooc.py_driver_2d.phosphorus import phosphorus

    tm = object.__new__(phosphorus)
    tm.name = "phosphorus"
    tm.tracer_cnt = 3
    tm.depth = depth
    tm.ypos = ypos
    tm.light_lim = np.outer(
        np.exp((-1.0 / 25.0) * depth.mid),
        np.exp(-1.0 * ((ypos.mid - 2.5e6) / 1.5e6) ** 2),
    )
    tm.po4_ind, tm.dop_ind, tm.pop_ind = 0, 1, 2
    tm.params = phosphorus.gen_params({})
    tm.pop_sink_work = np.zeros((len(depth) + 1, len(ypos)))
    return tm


def make_2d_forced(depth, ypos, modelinfo, suff="o2_like"):
    """modelinfo: dict of forced_* cfg strings (scripts/run_py_driver_2d_forced_*.sh)"""
    install_stubs()
    from nk_ooc.py_driver_2d.forced import forced

    forced.forced_class_vars_set = False
    tm = object.__new__(forced)
    tm.name = f"forced_{suff}"
    tm.tracer_cnt = 1
    tm.depth = depth
    tm.ypos = ypos
    tm._set_forced_class_vars(modelinfo)
    return tm


def make_test_problem(nz, depth_end=900.0, ratio=5.0):
    install_stubs()
    from nk_ooc.test_problem.vert_mix import VertMix

    depth = make_axis("depth", units="m", nlevs=nz, edge_start=0.0, edge_end=depth_end,
                      delta_ratio_max=ratio)
    return depth, VertMix(depth)


def make_tp_module(kind, depth, name=None):
    install_stubs()
    import importlib

    mod = importlib.import_module(f"nk_ooc.test_problem.{kind}")
    cls = getattr(mod, kind)
    tm = object.__new__(cls)
    tm.name = name or kind
    tm.depth = depth
    if kind == "iage":
        tm.tracer_cnt = 1
        tm.pist_vel = 24.0 * (1.0 / 86400.0) * 10.0
    elif kind == "dye_decay":
        from nk_ooc.test_problem import constants

        tm.tracer_cnt = 1
        tm._dye_decay_surf_flux_times = constants.sec_per_year * np.array([0.1, 0.2, 0.6, 0.7])
        tm._dye_decay_surf_flux_vals = constants.year_per_sec * np.array([0.0, 2.0, 2.0, 0.0])
        tm._dye_decay_surf_flux_time = None
        tm._dye_decay_surf_flux_val = 0.0
    elif kind == "phosphorus":
        tm.tracer_cnt = 6
        tm.light_lim = np.exp((-1.0 / 25.0) * depth.mid)
        tm.po4_s_restoring_opt = 1  # input/test_problem/model_params.cfg:6
        tm._sinking_tend_work = np.zeros(1 + len(depth))
    return tm


def read_nc(fname):
    """all variables of a NETCDF3 file as native-endian numpy arrays"""
    nc = netcdf_file(fname, "r", mmap=False)
    out = {}
    for name, var in nc.variables.items():
        out[name] = np.array(var.data, dtype=var.data.dtype.newbyteorder("="))
    nc.close()
    return out
