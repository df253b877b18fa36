"""Tracer modules of test_problem (1-D column) -> device models (engine.Model).

The vertical operator (time-varying mixing, piston velocity / decay on the diagonal, particle
sinking) is implicit; the surface flux of dye_decay is an affine k=0 source; the phosphorus
sources are explicit (nk_ooc/test_problem/{vert_mix,iage,dye_decay,phosphorus}.py)."""

import numpy as np

from .. import _lib
from ..engine import Model

SEC_PER_DAY = 86400.0
SEC_PER_YEAR = SEC_PER_DAY * 365.0  # nk_ooc/test_problem/constants.py:3-5


def _base_desc(depth, keep):
    d = _lib.ModelDesc()
    d.nz, d.ny = len(depth), 1
    d.column_model = 1
    d.t0, d.t1 = 0.0, SEC_PER_YEAR
    keep["edges"] = np.ascontiguousarray(depth.edges, dtype=np.float64)
    d.h_depth_edges = _lib.dptr(keep["edges"])
    return d


def iage_model(depth):
    """test_problem/iage.py:11-29: piston velocity 240 m/day as a surface flux, ageing 1/yr"""
    keep = {}
    d = _base_desc(depth, keep)
    d.n_tracers, d.kind, d.n_classes = 1, _lib.MOD_LINEAR, 1
    d.class_of[0] = 0
    pist_vel = 24.0 * (1.0 / SEC_PER_DAY) * 10.0
    d.surf_diag[0] = -pist_vel * depth.delta_r[0]
    d.src_const[0] = 1.0 / SEC_PER_YEAR
    return Model(d, keep)


def dye_decay_model(depth, suff):
    """test_problem/dye_decay.py:11-47: trapezoid-in-time surface flux (1 mol/m^2 per year),
    decay at suff/1000 per year"""
    keep = {}
    d = _base_desc(depth, keep)
    d.n_tracers, d.kind, d.n_classes = 1, _lib.MOD_LINEAR, 1
    d.class_of[0] = 0
    d.decay[0] = -int(suff) * 0.001 * (1.0 / SEC_PER_YEAR)
    d.n_flux_pts = 4
    for i, (t, v) in enumerate(zip([0.1, 0.2, 0.6, 0.7], [0.0, 2.0, 2.0, 0.0])):
        d.flux_t[i] = SEC_PER_YEAR * t
        d.flux_v[i] = (1.0 / SEC_PER_YEAR) * v
    return Model(d, keep)


def phosphorus_model(depth, po4_s_restoring_opt=1):
    """test_problem/phosphorus.py:11-120: po4, dop, pop and their shadows; pop / pop_s sink at
    1 m/day (implicit, class 1)"""
    keep = {}
    d = _base_desc(depth, keep)
    d.n_tracers, d.kind, d.n_classes = 6, _lib.MOD_PHOSPHORUS_1D, 2
    for t, c in enumerate([0, 0, 1, 0, 0, 1]):
        d.class_of[t] = c
    d.sink_vel[1] = 1.0 / SEC_PER_DAY
    keep["light"] = np.ascontiguousarray(np.exp((-1.0 / 25.0) * depth.mid))
    d.h_light = _lib.dptr(keep["light"])
    d.po4_s_restoring_opt = int(po4_s_restoring_opt)
    return Model(d, keep)


def po4_uptake(depth, po4):
    """test_problem/phosphorus.py:78-85: po4 [..., nz]; light e-folding depth 25 m, half saturation 0.5,
    maximum rate 1/day"""
    light = np.exp((-1.0 / 25.0) * depth.mid)
    return (1.0 / SEC_PER_DAY) * light * (po4 / (po4 + 0.5))


def po4_s_restore_tau_r(depth, po4, uptake, opt=1):
    """inverse time scale of the po4_s restoring (test_problem/phosphorus.py:58-76): opt 0: 1/day in
    the surface layer; opt 1: finite-difference d uptake / d po4"""
    if opt == 0:
        res = np.zeros(np.shape(po4))
        res[..., 0] = 1.0 / SEC_PER_DAY
        return res
    delta = 1.0e-3 * np.abs(po4)
    delta = np.where(delta < 1.0e-8, 1.0e-8, delta)
    return (po4_uptake(depth, po4 + delta) - uptake) / delta


def bldepth(time):
    """test_problem/vert_mix.py:50-57 (host copy for hist output)"""
    frac = 0.5 + 0.5 * np.cos((2 * np.pi) * ((1.0 / SEC_PER_YEAR) * time - 0.25))
    return 50.0 + (150.0 - 50.0) * frac


def kink_times(depth):
    """fractions of the year at which the mixing coefficient of some interior edge has a kink:
    bldepth(t) -/+ 20 m crosses the edge depth (test_problem/vert_mix.py:38-47)"""
    out = set()
    for e in depth.edges[1:-1]:
        for off in (-20.0, 20.0):
            c = (e - off - 100.0) / 50.0
            if abs(c) < 1.0:
                a = np.arccos(c) / (2 * np.pi)
                for ph in (a, -a):
                    out.add(float(np.round((ph + 0.25) % 1.0, 15)))
    return sorted(out)


def aligned_schedule(depth, steps_per_year, extra_breaks=(), n_hist=101):
    """step boundaries on every kink of the mixing coefficient, on the hist times k/100 yr
    (test_problem/model_state.py:66) and on `extra_breaks` (e.g. the dye flux kinks);
    uniform steps of at most T/steps_per_year in between"""
    br = set(kink_times(depth)) | {0.0, 1.0} | set(np.arange(n_hist) / float(n_hist - 1)) | set(extra_breaks)
    br = sorted(br)
    merged = [br[0]]
    for x in br[1:]:
        if x - merged[-1] > 1.0e-12:
            merged.append(x)
    ts, hs = [], []
    for a, b in zip(merged[:-1], merged[1:]):
        n = max(1, int(np.ceil((b - a) * steps_per_year - 1.0e-9)))
        h = (b - a) * SEC_PER_YEAR / n
        ts.append(a * SEC_PER_YEAR + h * np.arange(n))
        hs.append(np.full(n, h))
    return np.concatenate(ts), np.concatenate(hs)


def halved(schedule):
    """every step split in two (the fine leg of the Richardson pair)"""
    ts, hs = schedule
    tf = np.ravel(np.column_stack([ts, ts + 0.5 * hs]))
    return tf, np.repeat(0.5 * hs, 2)
