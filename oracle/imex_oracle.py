"""TEST INFRASTRUCTURE — numpy restatement of the *product's* fixed-schedule integrator.

The reference integrates with SciPy's adaptive Radau IIA(5) (oracle/nk_oracle.py restates
that path).  The B200 path uses a fixed-schedule IMEX Runge-Kutta scheme, ARS(2,2,2)
(Ascher, Ruuth & Spiteri 1997): vertical mixing + vertical advection + surface restoring
(+ particle sinking) implicit through one tridiagonal solve per (member, tracer, column)
and stage, horizontal advection/mixing and the tracer sources explicit.

This file states that scheme in plain numpy so that the CUDA kernels can be checked to
rounding error (independently of the discretisation-error comparison against the
reference's Radau solution).  It uses the SAME building blocks as nk_oracle.py
(Grid2D/Column1D fields) for every model coefficient.  Only tests/ may import it.
"""

import numpy as np

from .nk_oracle import SEC_PER_YEAR

GAMMA = 1.0 - 1.0 / np.sqrt(2.0)
DELTA = 1.0 - 1.0 / (2.0 * GAMMA)


def uniform_schedule(nsteps, t0=0.0, t1=SEC_PER_YEAR):
    """step start times and step sizes; step k covers [k*h, (k+1)*h] with times computed as
    t0 + k*h (not accumulated) so that k/60-year hist times are hit exactly"""
    h = (t1 - t0) / nsteps
    return t0 + h * np.arange(nsteps), np.full(nsteps, h)


def explicit_stencil_2d(g):
    """E_transport(c)[k,j] = eL*c[k,j-1] + eC*c[k,j] + eR*c[k,j+1]: horizontal advection
    (advection.py:58-65) + horizontal mixing (horiz_mix.py:59-65) in coefficient form"""
    nz, ny = g.nz, g.ny
    dyr = g.ypos.delta_r[None, :]
    v = g.vvel.copy()
    v[:, 0] = 0.0
    v[:, -1] = 0.0
    kh = np.zeros((nz, ny + 1))
    kh[:, 1:-1] = g.hmix
    eL = dyr * (0.5 * v[:, :-1] + kh[:, :-1])
    eR = dyr * (-0.5 * v[:, 1:] + kh[:, 1:])
    eC = dyr * (0.5 * v[:, :-1] - 0.5 * v[:, 1:] - kh[:, :-1] - kh[:, 1:])
    eL[:, 0] = 0.0
    eR[:, -1] = 0.0
    return eL, eC, eR


def implicit_tridiag_2d(g, time):
    """vertical advection + vertical mixing as sub/diag/super coefficients [nz, ny]
    (advection.py:67-74, vert_mix.py:34-39)"""
    nz, ny = g.nz, g.ny
    dzr = g.depth.delta_r[:, None]
    w = g.wvel.copy()
    w[0, :] = 0.0
    w[-1, :] = 0.0
    mc = np.zeros((nz + 1, ny))
    mc[1:-1] = g.vert_mixing_coeff(time)
    sub = dzr * (-0.5 * w[:-1] + mc[:-1])
    sup = dzr * (0.5 * w[1:] + mc[1:])
    diag = dzr * (0.5 * w[1:] - 0.5 * w[:-1] - mc[1:] - mc[:-1])
    sub[0] = 0.0
    sup[-1] = 0.0
    return sub, diag, sup


def thomas_factor(sub, diag, sup, hg):
    """tables (m, ib, g) of the LU factorisation of I - hg*tridiag(sub, diag, sup) along axis 0:
       forward  y[k] = r[k] - m[k]*y[k-1];  backward x[k] = y[k]*ib[k] - g[k]*x[k+1]"""
    a = -hg * sub
    b = 1.0 - hg * diag
    c = -hg * sup
    nz = a.shape[0]
    m = np.zeros_like(a)
    ib = np.zeros_like(a)
    beta = b[0].copy()
    ib[0] = 1.0 / beta
    for k in range(1, nz):
        m[k] = a[k] * ib[k - 1]
        beta = b[k] - m[k] * c[k - 1]
        ib[k] = 1.0 / beta
    gg = c * ib
    gg[-1] = 0.0
    return m, ib, gg


def thomas_solve(m, ib, gg, r):
    """r: [nz, ..., B]; tables broadcast over trailing member axis"""
    nz = r.shape[0]
    y = np.empty_like(r)
    y[0] = r[0]
    for k in range(1, nz):
        y[k] = r[k] - m[k][..., None] * y[k - 1]
    x = np.empty_like(r)
    x[nz - 1] = y[nz - 1] * ib[nz - 1][..., None]
    for k in range(nz - 2, -1, -1):
        x[k] = y[k] * ib[k][..., None] - gg[k][..., None] * x[k + 1]
    return x


class Module2D:
    """module-specific split of the py_driver_2d tendency into explicit sources and extra
    implicit (vertical) terms; `kind` in {"iage", "forced", "phosphorus"}"""

    def __init__(self, kind, g, forced=None, phos=None):
        self.kind = kind
        self.g = g
        self.T = {"iage": 2, "forced": 1, "phosphorus": 3}[kind]
        self.forced = forced  # nk_oracle.Forced2D (parameters + forcing record)
        self.phos = phos  # nk_oracle.Phosphorus2D

    def tracer_class(self):
        return {"iage": [0, 1], "forced": [0], "phosphorus": [0, 0, 1]}[self.kind]

    def implicit_extra(self, cls, sub, diag, sup, time=None):
        """add module terms to the class's tridiagonal (in place); returns affine surface
        source rate for k=0 (0.0, a scalar, or [ny] when the surface is restored to a record)"""
        g = self.g
        aff = 0.0
        if self.kind == "iage":
            r = 24.0 / 86400.0 * 10.0 / g.depth.delta[0]
            diag[0] -= r if cls == 0 else 0.01 * r
        elif self.kind == "forced":
            f = self.forced
            if f.restore_const is not None:
                diag[0] -= f.rate
                aff = f.rate * f.restore_const
            if f.restore_times is not None:  # forced_surf_restore_opt = file (forced.py:124-130)
                diag[0] -= f.rate
                aff = aff + f.rate * f.restore_to(time)
            if f.sms_opt == "decay":
                diag -= f.sms_decay_rate
        elif self.kind == "phosphorus" and cls == 1:
            p = self.phos
            dzr = g.depth.delta_r[:, None]
            sub[1:] += p.sink * dzr[1:]
            diag[:-1] -= p.sink * dzr[:-1]
        return aff

    def explicit_sources(self, time, c):
        """c: [T, nz, ny, B] -> explicit source tendency (same shape)"""
        out = np.zeros_like(c)
        if self.kind == "iage":
            out += 1.0 / SEC_PER_YEAR
        elif self.kind == "forced":
            f = self.forced
            if f.sms_opt == "const":
                out[0] += f.sms_const
            elif f.sms_opt == "file":
                s = np.broadcast_to(f.sms(time)[..., None], c[0].shape)
                if f.sink_thres is not None:
                    q = (1.0 / f.sink_thres) * c[0]
                    s = s * np.where((s < 0.0) & (q > 0.0) & (q < 1.0), q, 1.0)
                out[0] += s
        elif self.kind == "phosphorus":
            p = self.phos
            u = p.umax * p.light[..., None] * (c[0] / (c[0] + p.halfsat))
            rd = p.dop_remin * c[1]
            rp = p.pop_remin * c[2]
            out[0] = -u + rd + rp
            out[1] = p.sigma * u - rd
            out[2] = (1.0 - p.sigma) * u - rp
        return out


def explicit_tend_2d(mod, stencil, time, c):
    eL, eC, eR = stencil
    out = eC[None, :, :, None] * c
    out[:, :, 1:] += eL[None, :, 1:, None] * c[:, :, :-1]
    out[:, :, :-1] += eR[None, :, :-1, None] * c[:, :, 1:]
    return out + mod.explicit_sources(time, c)


def stage_tables_2d(mod, time, hg):
    """per tracer class: (m, ib, g, aff) for I - hg*L(time)"""
    tabs = []
    for cls in sorted(set(mod.tracer_class())):
        sub, diag, sup = implicit_tridiag_2d(mod.g, time)
        aff = mod.implicit_extra(cls, sub, diag, sup, time)
        tabs.append(thomas_factor(sub, diag, sup, hg) + (np.asarray(aff, dtype=np.float64)[..., None],))
    return tabs


def piecewise_schedule(breaks, steps_per_piece, t0=0.0, t1=SEC_PER_YEAR):
    """uniform steps inside each [breaks[i], breaks[i+1]] (fractions of the year)"""
    ts, hs = [], []
    for (a, b), n in zip(zip(breaks[:-1], breaks[1:]), steps_per_piece):
        h = (b - a) * (t1 - t0) / n
        ts.append(t0 + a * (t1 - t0) + h * np.arange(n))
        hs.append(np.full(n, h))
    return np.concatenate(ts), np.concatenate(hs)


def model_year_2d(mod, x0, nsteps=2400, snapshots=None, schedule=None):
    """ARS(2,2,2) over one year for x0 [T, nz, ny, B]; returns x(T) - x(0).
    snapshots: optional list that receives (step_index, state copy) after every step."""
    g = mod.g
    stencil = explicit_stencil_2d(g)
    tstart, hs = schedule if schedule is not None else uniform_schedule(nsteps, *g.time_range)
    nsteps = len(hs)
    cls_of = mod.tracer_class()
    u = np.array(x0, dtype=np.float64)
    a1 = (1.0 - GAMMA) / GAMMA
    a0 = 1.0 - a1
    for n in range(nsteps):
        t, h = tstart[n], hs[n]
        hg = h * GAMMA
        t1 = t + GAMMA * h
        t2 = tstart[n + 1] if n + 1 < nsteps else g.time_range[1]
        e_n = explicit_tend_2d(mod, stencil, t, u)
        rhs = u + hg * e_n
        tabs = stage_tables_2d(mod, t1, hg)
        u1 = np.empty_like(u)
        for tr, cls in enumerate(cls_of):
            m, ib, gg, aff = tabs[cls]
            r = rhs[tr].copy()
            r[0] += hg * aff
            u1[tr] = thomas_solve(m, ib, gg, r)
        e_1 = explicit_tend_2d(mod, stencil, t1, u1)
        rhs = a0 * u + a1 * u1 + h * (DELTA - 1.0 + GAMMA) * e_n + h * (1.0 - DELTA) * e_1
        tabs = stage_tables_2d(mod, t2, hg)
        u2 = np.empty_like(u)
        for tr, cls in enumerate(cls_of):
            m, ib, gg, aff = tabs[cls]
            r = rhs[tr].copy()
            r[0] += hg * aff
            u2[tr] = thomas_solve(m, ib, gg, r)
        u = u2
        if snapshots is not None:
            snapshots.append((n + 1, u.copy()))
    return u - x0


# --------------------------------------------------------------------------------------
# test_problem (1-D column): everything is vertical, so the scheme reduces to the L-stable
# SDIRK2 (gamma = 1 - 1/sqrt 2) with one tridiagonal solve per stage
# --------------------------------------------------------------------------------------


class Module1D:
    """split of the test_problem tendencies (test_problem/{iage,dye_decay}.py) for the column
    model; `kind` in {"iage", "dye_decay"}; the vertical operator carries mixing, the piston
    velocity (iage) or the decay; the surface flux of dye_decay is an affine k=0 source"""

    def __init__(self, kind, col, suff=None):
        self.kind = kind
        self.g = col
        self.T = 1
        self.suff = int(suff) if suff is not None else 0

    def tridiag(self, time):
        d = self.g.depth
        nz = self.g.nz
        mc = np.zeros(nz + 1)
        mc[1:-1] = self.g.mixing_coeff(time)
        sub = d.delta_r * mc[:-1]
        sup = d.delta_r * mc[1:]
        diag = -d.delta_r * (mc[1:] + mc[:-1])
        aff = 0.0
        if self.kind == "iage":
            diag[0] -= 24.0 * (1.0 / 86400.0) * 10.0 * d.delta_r[0]
        else:
            diag -= self.suff * 0.001 * (1.0 / SEC_PER_YEAR)
            flux = np.interp(time, SEC_PER_YEAR * np.array([0.1, 0.2, 0.6, 0.7]),
                             (1.0 / SEC_PER_YEAR) * np.array([0.0, 2.0, 2.0, 0.0]))
            aff = flux * d.delta_r[0]
        return sub, diag, sup, aff

    def src_const(self):
        return 1.0 / SEC_PER_YEAR if self.kind == "iage" else 0.0


def model_year_1d(mod, x0, nsteps=None, schedule=None):
    """x0 [nz, B] -> x(T) - x(0) with SDIRK2 (ARS(2,2,2) without explicit transport)"""
    g = mod.g
    tstart, hs = schedule if schedule is not None else uniform_schedule(nsteps, *g.time_range)
    nsteps = len(hs)
    u = np.array(x0, dtype=np.float64)
    a1 = (1.0 - GAMMA) / GAMMA
    a0 = 1.0 - a1
    sc = mod.src_const()
    for n in range(nsteps):
        t, h = tstart[n], hs[n]
        hg = h * GAMMA
        t1 = t + GAMMA * h
        t2 = tstart[n + 1] if n + 1 < nsteps else g.time_range[1]
        sub, diag, sup, aff = mod.tridiag(t1)
        m, ib, gg = thomas_factor(sub, diag, sup, hg)
        r = u + hg * sc
        r[0] += hg * aff
        u1 = thomas_solve(m, ib, gg, r)
        sub, diag, sup, aff = mod.tridiag(t2)
        m, ib, gg = thomas_factor(sub, diag, sup, hg)
        r = a0 * u + a1 * u1 + h * (DELTA - 1.0 + GAMMA) * sc + h * (1.0 - DELTA) * sc
        r[0] += hg * aff
        u = thomas_solve(m, ib, gg, r)
    return u - x0


class Phosphorus1DSplit:
    """test_problem phosphorus (test_problem/phosphorus.py:28-120) in IMEX form: mixing (all six
    tracers) and pop / pop_s sinking (1 m/day, upwind) implicit; uptake, remineralisation and
    the shadow restoring explicit"""

    T = 6

    def __init__(self, col, phos):
        self.g = col
        self.p = phos  # nk_oracle.Phosphorus1D

    def tridiag(self, time, sinking):
        d = self.g.depth
        nz = self.g.nz
        mc = np.zeros(nz + 1)
        mc[1:-1] = self.g.mixing_coeff(time)
        sub = d.delta_r * mc[:-1]
        sup = d.delta_r * mc[1:]
        diag = -d.delta_r * (mc[1:] + mc[:-1])
        if sinking:
            v = 1.0 / 86400.0
            sub[1:] += v * d.delta_r[1:]
            diag[:-1] -= v * d.delta_r[:-1]
        return sub, diag, sup

    def explicit(self, c):
        """c [6, nz, B]"""
        p = self.p
        out = np.empty_like(c)
        po4 = c[0]
        u = (1.0 / 86400.0) * p.light[:, None] * (po4 / (po4 + 0.5))
        rem = 0.01 * (1.0 / 86400.0)
        for o3 in (0, 3):
            out[o3 + 0] = -u + rem * c[o3 + 1] + rem * c[o3 + 2]
            out[o3 + 1] = 0.67 * u - rem * c[o3 + 1]
            out[o3 + 2] = (1.0 - 0.67) * u - rem * c[o3 + 2]
        if p.opt == 0:
            tau = np.zeros_like(po4)
            tau[0] = 1.0 / 86400.0
        else:
            delta = 1.0e-3 * np.abs(po4)
            delta[delta < 1.0e-8] = 1.0e-8
            pd = po4 + delta
            tau = ((1.0 / 86400.0) * p.light[:, None] * (pd / (pd + 0.5)) - u) / delta
        rest = tau * (c[0] - c[3])
        out[3] += rest
        out[4] -= 0.67 * rest
        out[5] -= 0.33 * rest
        return out


def model_year_1d_phosphorus(mod, x0, nsteps):
    """x0 [6, nz, B]"""
    g = mod.g
    tstart, hs = uniform_schedule(nsteps, *g.time_range)
    u = np.array(x0, dtype=np.float64)
    a1 = (1.0 - GAMMA) / GAMMA
    a0 = 1.0 - a1
    sink = [False, False, True, False, False, True]

    def solve(rhs, time, hg):
        out = np.empty_like(rhs)
        fac = {s: thomas_factor(*mod.tridiag(time, s), hg) for s in (False, True)}
        for t in range(6):
            out[t] = thomas_solve(*fac[sink[t]], rhs[t])
        return out

    for n in range(nsteps):
        t, h = tstart[n], hs[n]
        hg = h * GAMMA
        t1 = t + GAMMA * h
        t2 = tstart[n + 1] if n + 1 < nsteps else g.time_range[1]
        e_n = mod.explicit(u)
        u1 = solve(u + hg * e_n, t1, hg)
        e_1 = mod.explicit(u1)
        u = solve(a0 * u + a1 * u1 + h * (DELTA - 1.0 + GAMMA) * e_n + h * (1.0 - DELTA) * e_1, t2, hg)
    return u - x0
