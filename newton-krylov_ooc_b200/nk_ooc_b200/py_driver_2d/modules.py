"""Tracer modules of py_driver_2d -> device models (engine.Model).

Each builder fills the nkb_model_desc that describes how the module's tendency
(nk_ooc/py_driver_2d/{iage,forced,phosphorus}.py comp_tend) is split between the explicit
sources and the implicit vertical operator of the fused stage kernel.
"""

import numpy as np

from .. import _lib
from ..engine import Model
from .processes import SEC_PER_YEAR, Advection, HorizMix, VertMix, explicit_stencil


class Transport2D:
    """grid + processes shared by all tracer modules (py_driver_2d/model_state.py:42-65)"""

    def __init__(self, depth, ypos, max_abs_vvel=0.1, horiz_mix_coeff=1000.0):
        self.depth, self.ypos = depth, ypos
        self.max_abs_vvel = float(max_abs_vvel)
        self.horiz_mix_coeff = float(horiz_mix_coeff)
        self.time_range = (0.0, SEC_PER_YEAR)
        self.advection = Advection(depth, ypos, self.max_abs_vvel)
        self.horiz_mix = HorizMix(depth, ypos, self.horiz_mix_coeff, self.advection)
        self.vert_mix = VertMix(depth, ypos)
        self.estencil = explicit_stencil(depth, ypos, self.advection, self.horiz_mix)

    def base_desc(self, keep):
        d = _lib.ModelDesc()
        d.nz, d.ny = len(self.depth), len(self.ypos)
        d.column_model = 0
        d.t0, d.t1 = self.time_range
        keep["edges"] = np.ascontiguousarray(self.depth.edges, dtype=np.float64)
        keep["ymid"] = np.ascontiguousarray(self.ypos.mid, dtype=np.float64)
        keep["wvel"] = np.ascontiguousarray(self.advection.wvel, dtype=np.float64)
        keep["est"] = self.estencil
        keep["bld"] = np.ascontiguousarray(self.vert_mix.bldepth_max, dtype=np.float64)
        d.h_depth_edges = _lib.dptr(keep["edges"])
        d.h_ypos_mid = _lib.dptr(keep["ymid"])
        d.h_wvel = _lib.dptr(keep["wvel"])
        d.h_estencil = _lib.dptr(keep["est"])
        d.h_bld_max = _lib.dptr(keep["bld"])
        return d


def iage_model(tr):
    """iage + iage_slow_rest (py_driver_2d/iage.py:13-41): surface restoring to 0 at
    24/day over 10 m (x0.01 for the slow tracer) is implicit, ageing 1/yr explicit"""
    keep = {}
    d = tr.base_desc(keep)
    d.n_tracers, d.kind, d.n_classes = 2, _lib.MOD_LINEAR, 2
    d.class_of[0], d.class_of[1] = 0, 1
    rate = 24.0 / 86400.0 * 10.0 / tr.depth.delta[0]
    d.surf_diag[0] = -rate
    d.surf_diag[1] = -0.01 * rate
    d.src_const[0] = d.src_const[1] = 1.0 / SEC_PER_YEAR
    return Model(d, keep)


def forced_model(tr, surf_restore_opt="const", surf_restore_const=0.0, surf_restore_rate_10m=24.0 / 86400.0,
                 sms_opt="none", sms_const=0.0, sms_decay_rate=0.0, sms_times=None, sms_data=None,
                 sink_thres=None, surf_restore_times=None, surf_restore_data=None):
    """forced_{suff} (py_driver_2d/forced.py:57-154).  sms_data [nt, nz, ny] must already be
    on the model grid with scalef applied (utils.gen_forcing_fcn does both when reading);
    surf_restore_data [nt, ny] (forced_surf_restore_opt = file) on the model's ypos axis."""
    if surf_restore_opt not in ("none", "const", "file"):
        raise ValueError(f"unknown forced_surf_restore_opt={surf_restore_opt}")
    if sms_opt not in ("none", "const", "decay", "file"):
        raise ValueError(f"unknown forced_sms_opt={sms_opt}")
    if surf_restore_opt == "none" and sms_opt != "decay":
        raise ValueError("forced_sms_opt must be decay if forced_surf_restore_opt == none")
    keep = {}
    d = tr.base_desc(keep)
    d.n_tracers, d.n_classes = 1, 1
    d.class_of[0] = 0
    d.kind = _lib.MOD_FORCED_FILE if sms_opt == "file" else _lib.MOD_LINEAR
    if surf_restore_opt == "const":
        rate = 10.0 / tr.depth.delta[0] * surf_restore_rate_10m
        d.surf_diag[0] = -rate
        d.surf_aff[0] = rate * surf_restore_const
    elif surf_restore_opt == "file":
        rate = 10.0 / tr.depth.delta[0] * surf_restore_rate_10m
        keep["st"] = np.ascontiguousarray(surf_restore_times, dtype=np.float64)
        keep["sd"] = np.ascontiguousarray(surf_restore_data, dtype=np.float64)
        if keep["sd"].shape != (len(keep["st"]), len(tr.ypos)) or len(keep["st"]) < 2:
            raise ValueError("surf_restore_data must be [nt >= 2, ny] on the model grid")
        d.surf_diag[0] = -rate
        d.srf_rate[0] = rate
        d.n_srf = len(keep["st"])
        d.h_srf_time = _lib.dptr(keep["st"])
        d.h_srf_data = _lib.dptr(keep["sd"])
    if sms_opt == "const":
        d.src_const[0] = sms_const
    elif sms_opt == "decay":
        d.decay[0] = -sms_decay_rate
    elif sms_opt == "file":
        keep["ft"] = np.ascontiguousarray(sms_times, dtype=np.float64)
        keep["fd"] = np.ascontiguousarray(sms_data, dtype=np.float64)
        if keep["fd"].shape != (len(keep["ft"]), len(tr.depth), len(tr.ypos)):
            raise ValueError("sms_data must be [nt, nz, ny] on the model grid")
        d.n_frc = len(keep["ft"])
        d.h_frc_time = _lib.dptr(keep["ft"])
        d.h_frc_data = _lib.dptr(keep["fd"])
        d.sink_thres = float(sink_thres) if sink_thres is not None else 0.0
    return Model(d, keep)


def phosphorus_model(tr, params=None):
    """po4/dop/pop (py_driver_2d/phosphorus.py:15-95): Michaelis-Menten uptake and
    remineralisation explicit, pop sinking (upwind, 2 m/day) in the implicit operator"""
    p = {
        "po4_halfsat": 0.5,
        "max_uptake_rate": 1.0 / (3.0 * 86400.0),
        "sigma": 0.67,
        "dop_remin_rate": 1.0 / (0.5 * 365.0 * 86400.0),
        "pop_remin_rate": 1.0 / (0.5 * 365.0 * 86400.0),
        "pop_sink_vel": 2.0 / 86400.0,
    }
    p.update(params or {})
    keep = {}
    d = tr.base_desc(keep)
    d.n_tracers, d.kind, d.n_classes = 3, _lib.MOD_PHOSPHORUS, 2
    d.class_of[0], d.class_of[1], d.class_of[2] = 0, 0, 1
    d.sink_vel[1] = p["pop_sink_vel"]
    keep["light"] = np.ascontiguousarray(
        np.outer(np.exp((-1.0 / 25.0) * tr.depth.mid), np.exp(-1.0 * ((tr.ypos.mid - 2.5e6) / 1.5e6) ** 2))
    )
    d.h_light = _lib.dptr(keep["light"])
    d.po4_halfsat, d.max_uptake_rate, d.sigma = p["po4_halfsat"], p["max_uptake_rate"], p["sigma"]
    d.dop_remin_rate, d.pop_remin_rate = p["dop_remin_rate"], p["pop_remin_rate"]
    return Model(d, keep)
