"""TEST INFRASTRUCTURE — CPU oracle for the nk_ooc hot path.  NOT part of the product.

A numpy/scipy restatement of the reference's algorithm for the batched function
evaluation F(x) = x(T) - x(0) (klindsay28/Newton-Krylov_OOC).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module; the product path (``nk_ooc_b200``) never does.

Parity pinning: every function below is checked in ``tests/test_oracle.py`` against
(i) the known answers held by the reference's own unit tests
(``tests/test_spatial_axis.py:138-198``), (ii) golden vectors produced by importing the
reference's own modules in the build container (``oracle/gen_golden.py`` ->
``tests/golden/*.npz``) and (iii) the reference's committed ``baselines/ci_*`` files
(copied as values into the same npz fixtures).

The time integrator of the reference is third party: ``scipy.integrate.solve_ivp(...,
"Radau")`` (SciPy; the reference pins scipy 1.9.1 in ``environments/fixed.yaml:23-31``,
this image has 1.18).  ``comp_fcn_*`` below call it exactly as the reference's call sites
do (``nk_ooc/py_driver_2d/model_state.py:102-114``,
``nk_ooc/test_problem/model_state.py:83-92``).

All citations are ``file:line`` under the reference root.
"""

import numpy as np
from scipy import integrate, linalg, sparse
from scipy.sparse import linalg as sp_linalg

SEC_PER_YEAR = 365.0 * 86400.0  # nk_ooc/test_problem/constants.py:3-5; py_driver_2d/model_state.py:49


# --------------------------------------------------------------------------------------
# spatial axis  (nk_ooc/spatial_axis.py)
# --------------------------------------------------------------------------------------


def stretched_edges(nlevs, edge_start, edge_end, delta_ratio_max):
    """layer edges of the reference's stretched grid (spatial_axis.py:253-290)"""
    xi = np.linspace(-1.0, 1.0, nlevs)
    fcn = 0.125 * xi * (15 + xi * xi * (3 * xi * xi - 10))
    delta_avg = (1.0 / nlevs) * (edge_end - edge_start)
    stretch = delta_avg * (delta_ratio_max - 1) / (delta_ratio_max + 1)
    delta = delta_avg + stretch * fcn
    edges = np.empty(nlevs + 1)
    edges[0] = edge_start
    edges[1:] = edge_start + delta.cumsum()
    return edges


class Axis:
    """grid metrics derived from edges (spatial_axis.py:35-39)"""

    def __init__(self, edges):
        self.edges = np.asarray(edges, dtype=np.float64)
        self.mid = 0.5 * (self.edges[:-1] + self.edges[1:])
        self.delta = self.edges[1:] - self.edges[:-1]
        self.delta_r = 1.0 / self.delta
        self.delta_mid = self.mid[1:] - self.mid[:-1]
        self.delta_mid_r = 1.0 / self.delta_mid

    def __len__(self):
        return len(self.mid)

    def remap_linear_interpolant(self, xvals, yvals):
        """layer averages of the piecewise-linear interpolant (spatial_axis.py:136-187).

        Restated as: integrate the clamped interpolant between consecutive breakpoints
        (layer edges and the x values that fall strictly inside the axis) with the
        trapezoid rule, then divide by the layer thickness.  The reference walks the same
        trapezoids with explicit while loops; layers holding no x value get the plain
        two-edge average (:143-144).
        """
        xvals = np.asarray(xvals, dtype=np.float64)
        yvals = np.asarray(yvals, dtype=np.float64)
        y_edge = np.interp(self.edges, xvals, yvals)
        res = 0.5 * (y_edge[:-1] + y_edge[1:])
        inside = (xvals >= self.edges[0]) & (xvals < self.edges[-1])
        layer_of = np.searchsorted(self.edges, xvals, side="right") - 1
        for layer in np.unique(layer_of[inside]):
            sel = inside & (layer_of == layer)
            xs = np.concatenate(([self.edges[layer]], xvals[sel], [self.edges[layer + 1]]))
            ys = np.concatenate(([y_edge[layer]], yvals[sel], [y_edge[layer + 1]]))
            acc = 0.0
            for i in range(len(xs) - 1):
                acc += (xs[i + 1] - xs[i]) * (0.5 * (ys[i] + ys[i + 1]))
            res[layer] = acc * self.delta_r[layer]
        return res


    def remap_ramp(self, x0, x1, y0, y1):
        """remap_linear_interpolant([x0, x1], [y0, y1]) vectorised over layers: the only form
        the hot path uses (vert_mix.py:61-70).  Same trapezoids as the generic routine."""
        e = self.edges
        f = np.interp(e, [x0, x1], [y0, y1])
        a, b, fa, fb = e[:-1], e[1:], f[:-1], f[1:]
        res = 0.5 * (fa + fb)
        in0 = (x0 >= a) & (x0 < b)
        in1 = (x1 >= a) & (x1 < b)
        both = in0 & in1
        only0 = in0 & ~in1
        only1 = in1 & ~in0
        if both.any():
            acc = (x0 - a) * (0.5 * (fa + y0)) + (x1 - x0) * (0.5 * (y0 + y1)) + (b - x1) * (0.5 * (y1 + fb))
            res = np.where(both, acc * self.delta_r, res)
        if only0.any():
            acc = (x0 - a) * (0.5 * (fa + y0)) + (b - x0) * (0.5 * (y0 + fb))
            res = np.where(only0, acc * self.delta_r, res)
        if only1.any():
            acc = (x1 - a) * (0.5 * (fa + y1)) + (b - x1) * (0.5 * (y1 + fb))
            res = np.where(only1, acc * self.delta_r, res)
        return res


# --------------------------------------------------------------------------------------
# py_driver_2d transport  (nk_ooc/py_driver_2d/{advection,horiz_mix,vert_mix}.py)
# --------------------------------------------------------------------------------------


class Grid2D:
    """time-invariant fields of the py_driver_2d model"""

    def __init__(self, depth_edges, ypos_edges, max_abs_vvel=0.1, horiz_mix_coeff=1000.0):
        self.depth = Axis(depth_edges)
        self.ypos = Axis(ypos_edges)
        self.nz = len(self.depth)
        self.ny = len(self.ypos)
        self.max_abs_vvel = float(max_abs_vvel)
        self.horiz_mix_coeff = float(horiz_mix_coeff)
        self.time_range = (0.0, SEC_PER_YEAR)
        self._vel_field()
        self._horiz_mix()
        # axis whose layer edges are the depth midpoints (vert_mix.py:17)
        self.depth_edges_axis = Axis(self.depth.mid)

    def _vel_field(self):
        """advection.py:22-49"""
        ze, ye = self.depth.edges, self.ypos.edges
        zn = (ze - ze.min()) / (ze.max() - ze.min())
        zn = 2.0 * zn / (1 + (2.0 - 1) * zn)
        zf = (27.0 / 4.0) * zn * (1.0 - zn) ** 2
        yn = (ye - ye.min()) / (ye.max() - ye.min())
        yf = 4.0 * yn * (1.0 - yn)
        stream = np.outer(zf, yf)
        vraw = (stream[1:, :] - stream[:-1, :]) * self.depth.delta_r[:, None]
        with np.errstate(invalid="ignore", divide="ignore"):
            stream = stream * self.max_abs_vvel / abs(vraw).max()
        self.stream = stream
        self.vvel = (stream[1:, :] - stream[:-1, :]) * self.depth.delta_r[:, None]
        self.wvel = (stream[:, 1:] - stream[:, :-1]) * self.ypos.delta_r

    def _horiz_mix(self):
        """horiz_mix.py:25-46  (includes the 1/dy_mid factor)"""
        K = self.horiz_mix_coeff
        vin = np.abs(self.vvel[:, 1:-1])
        if K > 0.0:
            pe = (0.5 / K) * self.ypos.delta_mid * vin
            self.hmix = K * np.where(pe > 1.0, pe, 1.0) * self.ypos.delta_mid_r
        else:
            self.hmix = 0.5 * vin

    def bldepth(self, time):
        """vert_mix.py:89-101"""
        bld_min = 35.0
        bld_max = np.interp(
            self.ypos.mid,
            [0.4e6, 0.8e6, 1.0e6, 1.2e6, 1.4e6, 1.5e6],
            [3000.0, 800.0, 415.0, 325.0, 280.0, bld_min],
        )
        tv = SEC_PER_YEAR * np.array([0.25, 0.35, 0.65, 0.75])
        frac = np.interp(time, tv, [0.0, 1.0, 1.0, 0.0])
        return bld_min + (bld_max - bld_min) * frac

    def vert_mixing_coeff(self, time):
        """kappa/dz_mid at interior depth edges, [nz-1, ny]  (vert_mix.py:43-87).
        Like the reference, the last result is cached by time (vert_mix.py:50-54) and columns
        with equal bldepth share one remap (vert_mix.py:60-72)."""
        if time == getattr(self, "_mc_time", None):
            return self._mc_vals
        bld = self.bldepth(time)
        log_sh, log_dp = np.log(1.0e1), np.log(5.0e-4)
        out = np.empty((self.nz - 1, self.ny))
        cache_val, cache_j = None, 0
        for j in range(self.ny):
            if bld[j] != cache_val:
                out[:, j] = self.depth_edges_axis.remap_ramp(bld[j] - 20.0, bld[j] + 20.0, log_sh, log_dp)
                cache_val, cache_j = bld[j], j
            else:
                out[:, j] = out[:, cache_j]
        out = np.exp(out)
        pe = 0.5 * self.depth.delta_mid[:, None] * np.abs(self.wvel[1:-1, :]) / out
        out = out * np.where(pe > 1.0, pe, 1.0)
        self._mc_time = time
        self._mc_vals = out * self.depth.delta_mid_r[:, None]
        return self._mc_vals

    def transport_tend(self, time, c):
        """advection + horizontal mixing + vertical mixing for c[T, nz, ny]
        (advection.py:51-76, horiz_mix.py:48-67, vert_mix.py:24-41)"""
        T, nz, ny = c.shape
        dzr = self.depth.delta_r[None, :, None]
        dyr = self.ypos.delta_r[None, None, :]
        fy = np.zeros((T, nz, ny + 1))
        fy[:, :, 1:-1] = 0.5 * (c[:, :, 1:] + c[:, :, :-1]) * self.vvel[None, :, 1:-1]
        fz = np.zeros((T, nz + 1, ny))
        fz[:, 1:-1, :] = 0.5 * (c[:, 1:, :] + c[:, :-1, :]) * self.wvel[None, 1:-1, :]
        tend = dyr * (fy[:, :, :-1] - fy[:, :, 1:]) + dzr * (fz[:, 1:, :] - fz[:, :-1, :])
        gy = np.zeros((T, nz, ny + 1))
        gy[:, :, 1:-1] = self.hmix[None] * (c[:, :, 1:] - c[:, :, :-1])
        tend = tend + dyr * (gy[:, :, 1:] - gy[:, :, :-1])
        gz = np.zeros((T, nz + 1, ny))
        gz[:, 1:-1, :] = self.vert_mixing_coeff(time)[None] * (c[:, 1:, :] - c[:, :-1, :])
        tend = tend + dzr * (gz[:, 1:, :] - gz[:, :-1, :])
        return tend

    def transport_jacobian(self, time, tracer_cnt):
        """CSR Jacobian of transport_tend, cell = j + ny*k
        (advection.py:111-179, horiz_mix.py:100-149, vert_mix.py:140-188)"""
        nz, ny = self.nz, self.ny
        n = nz * ny
        idx = np.arange(n).reshape(nz, ny)
        dzr = self.depth.delta_r[:, None]
        dyr = self.ypos.delta_r[None, :]
        mc = self.vert_mixing_coeff(time)
        rows, cols, vals = [], [], []

        def add(r, c, v):
            rows.append(r.ravel())
            cols.append(c.ravel())
            vals.append(np.broadcast_to(v, r.shape).ravel())

        diag = np.zeros((nz, ny))
        # shallower neighbour (k-1)
        v = (-0.5 * self.wvel[1:-1, :] + mc) * dzr[1:]
        add(idx[1:], idx[:-1], v)
        diag[1:] += -0.5 * self.wvel[1:-1, :] * dzr[1:] - mc * dzr[1:]
        # deeper neighbour (k+1)
        v = (0.5 * self.wvel[1:-1, :] + mc) * dzr[:-1]
        add(idx[:-1], idx[1:], v)
        diag[:-1] += 0.5 * self.wvel[1:-1, :] * dzr[:-1] - mc * dzr[:-1]
        # south neighbour (j-1)
        v = (0.5 * self.vvel[:, 1:-1] + self.hmix) * dyr[:, 1:]
        add(idx[:, 1:], idx[:, :-1], v)
        diag[:, 1:] += 0.5 * self.vvel[:, 1:-1] * dyr[:, 1:] - self.hmix * dyr[:, 1:]
        # north neighbour (j+1)
        v = (-0.5 * self.vvel[:, 1:-1] + self.hmix) * dyr[:, :-1]
        add(idx[:, :-1], idx[:, 1:], v)
        diag[:, :-1] += -0.5 * self.vvel[:, 1:-1] * dyr[:, :-1] - self.hmix * dyr[:, :-1]
        add(idx, idx, diag)
        one = sparse.csr_matrix(
            (np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)
        )
        return sparse.block_diag(tracer_cnt * [one], "csr")


# --------------------------------------------------------------------------------------
# py_driver_2d tracer modules  (nk_ooc/py_driver_2d/{iage,forced,phosphorus}.py)
# --------------------------------------------------------------------------------------


class Iage2D:
    """iage.py:13-93"""

    tracer_cnt = 2

    def __init__(self, grid):
        self.g = grid
        self.rate = 24.0 / 86400.0 * 10.0 / grid.depth.delta[0]
        self.slow = 0.01

    def comp_tend(self, time, flat):
        g = self.g
        c = flat.reshape(2, g.nz, g.ny)
        tend = g.transport_tend(time, c)
        tend[0, 0, :] -= self.rate * c[0, 0, :]
        tend[1, 0, :] -= self.slow * self.rate * c[1, 0, :]
        tend += 1.0 / SEC_PER_YEAR
        return tend.reshape(-1)

    def comp_jacobian(self, time, flat=None):
        g = self.g
        n = g.nz * g.ny
        jac = g.transport_jacobian(time, 2)
        d = np.zeros(2 * n)
        d[: g.ny] = -self.rate
        d[n : n + g.ny] = -self.slow * self.rate
        return (jac + sparse.diags(d)).tocsr()

    def apply_precond_jacobian(self, y):
        """res = M^-1 y - y,  M = I - prod_i (I - dt J((i+1/2) dt)), dt = T/3 (iage.py:66-93)"""
        shape = y.shape
        yv = y.reshape(-1)
        t0, t1 = self.g.time_range
        dt = (t1 - t0) / 3
        ident = sparse.identity(yv.size, format="csr")
        mat = ident.copy()
        for i in range(3):
            mat = mat @ (ident - dt * self.comp_jacobian(t0 + (i + 0.5) * dt))
        mat = (ident - mat).tocsc()
        return (sp_linalg.spsolve(mat, yv) - yv).reshape(shape)


class Phosphorus2D:
    """phosphorus.py:15-172"""

    tracer_cnt = 3

    def __init__(self, grid):
        self.g = grid
        self.light = np.outer(
            np.exp((-1.0 / 25.0) * grid.depth.mid),
            np.exp(-1.0 * ((grid.ypos.mid - 2.5e6) / 1.5e6) ** 2),
        )
        self.halfsat = 0.5
        self.umax = 1.0 / (3.0 * 86400.0)
        self.sigma = 0.67
        self.dop_remin = 1.0 / (0.5 * 365.0 * 86400.0)
        self.pop_remin = 1.0 / (0.5 * 365.0 * 86400.0)
        self.sink = 2.0 / 86400.0

    def uptake(self, po4):
        return self.umax * self.light * (po4 / (po4 + self.halfsat))

    def comp_tend(self, time, flat):
        g = self.g
        c = flat.reshape(3, g.nz, g.ny)
        tend = g.transport_tend(time, c)
        u = self.uptake(c[0])
        tend[0] -= u
        tend[1] += self.sigma * u
        tend[2] += (1.0 - self.sigma) * u
        rd = self.dop_remin * c[1]
        rp = self.pop_remin * c[2]
        tend[0] += rd + rp
        tend[1] -= rd
        tend[2] -= rp
        s = np.zeros((g.nz + 1, g.ny))
        s[1:-1] = self.sink * c[2, :-1]
        tend[2] += g.depth.delta_r[:, None] * (s[:-1] - s[1:])
        return tend.reshape(-1)

    def comp_jacobian(self, time, flat):
        g = self.g
        n = g.nz * g.ny
        c = flat.reshape(3, g.nz, g.ny)
        jac = g.transport_jacobian(time, 3).tolil()
        du = (self.umax * self.light * self.halfsat / (c[0] + self.halfsat) ** 2).reshape(-1)
        blocks = [[None] * 3 for _ in range(3)]
        ident = sparse.identity(n, format="csr")
        dub = sparse.diags(du)
        d0 = np.broadcast_to(-self.sink * g.depth.delta_r[:, None], (g.nz, g.ny)).copy()
        d0[-1, :] = 0.0
        dm1 = np.broadcast_to(self.sink * g.depth.delta_r[1:, None], (g.nz - 1, g.ny))
        sinkb = sparse.diags((d0.reshape(-1), dm1.reshape(-1)), (0, -g.ny))
        blocks[0][0] = -dub
        blocks[1][0] = self.sigma * dub
        blocks[2][0] = (1.0 - self.sigma) * dub
        blocks[0][1] = self.dop_remin * ident
        blocks[0][2] = self.pop_remin * ident
        blocks[1][1] = -self.dop_remin * ident
        blocks[2][2] = -self.pop_remin * ident + sinkb
        return (jac.tocsr() + sparse.bmat(blocks, format="csr")).tocsr()


    def apply_precond_jacobian(self, y, precond_times, precond_po4, weight=None):
        """res [3, nz, ny] (phosphorus.py:197-274): one interval of length T, Jacobian at T/2 with po4
        from the precond snapshot nearest T; null vector and shift from ARPACK shift-invert
        (sigma = 0); two shifted sparse solves + Richardson extrapolation; the multiple of the
        null vector that makes the region mean of the result vanish is removed; minus the input.
        weight: normalised cell weights of the (single) region; default layer thickness x width."""
        g = self.g
        shape = (3, g.nz, g.ny)
        if weight is None:
            weight = np.outer(g.depth.delta, g.ypos.delta)
            weight = weight / weight.sum()
        self_vals = np.asarray(y, dtype=np.float64).reshape(-1)
        t0, t1 = g.time_range
        time_delta = t1 - t0
        tracer_vals = np.zeros(shape)
        tracer_vals[0] = precond_po4[np.argmin(abs(t1 - precond_times))]
        mat_id = sparse.identity(self_vals.size)
        mat = mat_id - (mat_id - time_delta * self.comp_jacobian(t0 + 0.5 * time_delta, tracer_vals.reshape(-1)))
        e_vals, e_vects = sp_linalg.eigs(mat, k=5, sigma=0.0)
        null_comp = e_vects[:, 0]
        if max(abs(null_comp.imag)) > 1.0e-10 * max(abs(null_comp.real)):
            raise RuntimeError("1st eigenvector has non-trivial imaginary part")
        null_vect = null_comp.real.reshape(shape)
        shift = 0.5 * e_vals[1].real
        solve_tmp = sp_linalg.spsolve(mat - shift * mat_id, self_vals)
        solve_vals = sp_linalg.spsolve(mat - (0.5 * shift) * mat_id, self_vals)
        solve_vals = (2.0 * solve_vals - solve_tmp).reshape(shape)
        e_vect = null_vect / (weight[None] * null_vect).sum()
        solve_vals = solve_vals - (weight[None] * solve_vals).sum() * e_vect
        return solve_vals - self_vals.reshape(shape)


class Forced2D:
    """forced.py:11-202; forcing interpolation utils.py:488-537.

    sms_times/sms_data: forcing record [nt], [nt, nz, ny] already on the model grid and
    already multiplied by scalef (the reference applies scalef when reading, utils.py:515).
    """

    tracer_cnt = 1

    def __init__(self, grid, restore_rate_10m=24.0 / 86400.0, restore_const=None,
                 sms_opt="none", sms_const=0.0, sms_decay_rate=0.0, sms_times=None,
                 sms_data=None, sink_thres=None, restore_times=None, restore_data=None):
        self.g = grid
        self.restore_const = restore_const
        self.restore_times = restore_times  # forced_surf_restore_opt = file (forced.py:46-51): [nt], [nt, ny]
        self.restore_data = restore_data
        self.rate = 10.0 / grid.depth.delta[0] * restore_rate_10m
        self.sms_opt = sms_opt
        self.sms_const = sms_const
        self.sms_decay_rate = sms_decay_rate
        self.sms_times = sms_times
        self.sms_data = sms_data
        self.sink_thres = sink_thres

    def sms(self, time):
        """linear interpolation in time with linear extrapolation (interp1d
        fill_value="extrapolate", utils.py:533-535)"""
        t = self.sms_times
        i = int(np.clip(np.searchsorted(t, time, side="right") - 1, 0, len(t) - 2))
        w = (time - t[i]) / (t[i + 1] - t[i])
        return self.sms_data[i] + w * (self.sms_data[i + 1] - self.sms_data[i])

    def restore_to(self, time):
        """surface restoring value at `time`: interp1d(linear, extrapolate) over the record"""
        t = self.restore_times
        i = int(np.clip(np.searchsorted(t, time, side="right") - 1, 0, len(t) - 2))
        w = (time - t[i]) / (t[i + 1] - t[i])
        return self.restore_data[i] + w * (self.restore_data[i + 1] - self.restore_data[i])

    def comp_tend(self, time, flat):
        g = self.g
        c = flat.reshape(1, g.nz, g.ny)
        tend = g.transport_tend(time, c)
        if self.restore_const is not None:
            tend[0, 0, :] += self.rate * (self.restore_const - c[0, 0, :])
        if self.restore_times is not None:  # forced.py:124-130
            tend[0, 0, :] += self.rate * (self.restore_to(time) - c[0, 0, :])
        if self.sms_opt == "const":
            tend[0] += self.sms_const
        elif self.sms_opt == "decay":
            tend[0] += -self.sms_decay_rate * c[0]
        elif self.sms_opt == "file":
            s = self.sms(time)
            if self.sink_thres is not None:
                q = (1.0 / self.sink_thres) * c[0]
                s = s * np.where((s < 0.0) & (q > 0.0) & (q < 1.0), q, 1.0)
            tend[0] += s
        return tend.reshape(-1)

    def comp_jacobian(self, time, flat):
        g = self.g
        n = g.nz * g.ny
        jac = g.transport_jacobian(time, 1)
        d = np.zeros(n)
        if self.restore_const is not None or self.restore_times is not None:
            d[: g.ny] -= self.rate
        if self.sms_opt == "decay":
            d -= self.sms_decay_rate
        if self.sms_opt == "file" and self.sink_thres is not None:
            s = self.sms(time)
            q = (1.0 / self.sink_thres) * flat.reshape(g.nz, g.ny)
            d += np.where((s < 0.0) & (q > 0.0) & (q < 1.0), s / self.sink_thres, 0.0).reshape(-1)
        return (jac + sparse.diags(d)).tocsr()

    def apply_precond_jacobian(self, y, precond_times, precond_snaps):
        """res = M^-1 y - y,  M = I - prod_i (I - dt J_i), dt = T/3, J_i = J((i+1/2) dt) evaluated at the
        precond file's tracer snapshot nearest (i+1) dt (forced.py:204-241)"""
        shape = y.shape
        yv = y.reshape(-1)
        t0, t1 = self.g.time_range
        dt = (t1 - t0) / 3
        ident = sparse.identity(yv.size, format="csr")
        mat = ident.copy()
        for i in range(3):
            snap = precond_snaps[np.argmin(abs(t0 + (i + 1.0) * dt - precond_times))]
            mat = mat @ (ident - dt * self.comp_jacobian(t0 + (i + 0.5) * dt, snap.reshape(-1)))
        mat = (ident - mat).tocsc()
        return (sp_linalg.spsolve(mat, yv) - yv).reshape(shape)


def comp_fcn_2d(module, x0, t_eval=None, rtol=1.0e-6, atol=1.0e-6, return_sol=False):
    """F(x) = x(T) - x(0) with the reference's solve_ivp call (py_driver_2d/model_state.py:102-121)"""
    g = module.g
    flat0 = np.asarray(x0, dtype=np.float64).reshape(-1)
    jac0 = module.comp_jacobian(g.time_range[0], flat0)
    r, c, _ = sparse.find(jac0)
    sparsity = sparse.csr_matrix((np.ones(r.shape), (r, c)))
    sol = integrate.solve_ivp(
        module.comp_tend,
        g.time_range,
        flat0,
        "Radau",
        t_eval,
        max_step=(g.time_range[1] - g.time_range[0]) * 0.01,
        atol=atol,
        rtol=rtol,
        jac=module.comp_jacobian,
        jac_sparsity=sparsity,
    )
    res = (sol.y[:, -1] - flat0).reshape(np.shape(x0))
    return (res, sol) if return_sol else res


# --------------------------------------------------------------------------------------
# test_problem (1-D column)  (nk_ooc/test_problem/*.py)
# --------------------------------------------------------------------------------------


class Column1D:
    """vert_mix.py:8-57"""

    def __init__(self, depth_edges):
        self.depth = Axis(depth_edges)
        self.nz = len(self.depth)
        self.time_range = (0.0, SEC_PER_YEAR)

    @staticmethod
    def bldepth(time):
        frac = 0.5 + 0.5 * np.cos((2 * np.pi) * ((1.0 / SEC_PER_YEAR) * time - 0.25))
        return 50.0 + (150.0 - 50.0) * frac

    def mixing_coeff(self, time):
        bld = self.bldepth(time)
        lg = np.interp(self.depth.edges[1:-1], [bld - 20.0, bld + 20.0], [0.0, -5.0])
        return 10.0**lg * self.depth.delta_mid_r

    def mix_tend(self, time, c, surf_flux=0.0):
        w = np.zeros(self.nz + 1)
        w[0] = -surf_flux
        w[1:-1] = self.mixing_coeff(time) * (c[1:] - c[:-1])
        return (w[1:] - w[:-1]) * self.depth.delta_r


class Iage1D:
    """test_problem/iage.py:11-52"""

    tracer_cnt = 1

    def __init__(self, col):
        self.g = col
        self.pist_vel = 24.0 * (1.0 / 86400.0) * 10.0

    def comp_tend(self, time, c):
        return self.g.mix_tend(time, c, -self.pist_vel * c[0]) + 1.0 / SEC_PER_YEAR

    def precond_diagonals(self, mca):
        d = self.g.depth
        m = np.zeros((3, self.g.nz))
        m[0, 1:] = mca * d.delta_mid_r * d.delta_r[:-1]
        m[1, :-1] -= mca * d.delta_mid_r * d.delta_r[:-1]
        m[1, 1:] -= mca * d.delta_mid_r * d.delta_r[1:]
        m[1, 0] -= self.pist_vel * d.delta_r[0]
        m[2, :-1] = mca * d.delta_mid_r * d.delta_r[1:]
        return m

    def apply_precond_jacobian(self, y, mca):
        """iage.py:31-52; mca = mixing_coeff_log_mean[1:-1] of the precond file (m^2/s)"""
        t0, t1 = self.g.time_range
        rhs = (1.0 / (t1 - t0)) * y
        return linalg.solve_banded((1, 1), self.precond_diagonals(mca), rhs) - y


class DyeDecay1D:
    """test_problem/dye_decay.py:11-73"""

    tracer_cnt = 1

    def __init__(self, col, suff):
        self.g = col
        self.suff = int(suff)
        self.flux_t = SEC_PER_YEAR * np.array([0.1, 0.2, 0.6, 0.7])
        self.flux_v = (1.0 / SEC_PER_YEAR) * np.array([0.0, 2.0, 2.0, 0.0])

    def comp_tend(self, time, c):
        flux = np.interp(time, self.flux_t, self.flux_v)
        return self.g.mix_tend(time, c, flux) - self.suff * 0.001 * (1.0 / SEC_PER_YEAR) * c

    def apply_precond_jacobian(self, y, mca):
        d = self.g.depth
        m = np.zeros((3, self.g.nz))
        m[0, 1:] = mca * d.delta_mid_r * d.delta_r[:-1]
        m[1, :-1] -= mca * d.delta_mid_r * d.delta_r[:-1]
        m[1, 1:] -= mca * d.delta_mid_r * d.delta_r[1:]
        m[2, :-1] = mca * d.delta_mid_r * d.delta_r[1:]
        m[1, :] -= self.suff * 0.001 * (1.0 / SEC_PER_YEAR)
        t0, t1 = self.g.time_range
        return linalg.solve_banded((1, 1), m, (1.0 / (t1 - t0)) * y) - y


class Phosphorus1D:
    """test_problem/phosphorus.py:11-120 (po4,dop,pop + shadows po4_s,dop_s,pop_s)"""

    tracer_cnt = 6

    def __init__(self, col, restoring_opt=1):
        self.g = col
        self.light = np.exp((-1.0 / 25.0) * col.depth.mid)
        self.opt = restoring_opt

    def uptake(self, po4):
        return (1.0 / 86400.0) * self.light * (po4 / (po4 + 0.5))

    def tau_r(self, po4, uptake):
        if self.opt == 0:
            res = np.zeros(po4.shape)
            res[0] = 1.0 / 86400.0
            return res
        delta = 1.0e-3 * np.abs(po4)
        delta[delta < 1.0e-8] = 1.0e-8
        return (self.uptake(po4 + delta) - uptake) / delta

    def _core(self, time, u, c):
        g = self.g
        rd = 0.01 * (1.0 / 86400.0) * c[1]
        rp = 0.01 * (1.0 / 86400.0) * c[2]
        out = np.empty((3, g.nz))
        out[0] = -u + rd + rp + g.mix_tend(time, c[0])
        out[1] = 0.67 * u - rd + g.mix_tend(time, c[1])
        s = np.zeros(g.nz + 1)
        s[1:-1] = -(1.0 / 86400.0) * c[2, :-1]
        out[2] = (1.0 - 0.67) * u - rp + g.mix_tend(time, c[2]) + g.depth.delta_r * (s[1:] - s[:-1])
        return out

    def comp_tend(self, time, flat):
        c = flat.reshape(6, -1)
        out = np.empty(c.shape)
        u = self.uptake(c[0])
        out[0:3] = self._core(time, u, c[0:3])
        out[3:6] = self._core(time, u, c[3:6])
        rest = self.tau_r(c[0], u) * (c[0] - c[3])
        out[3] += rest
        out[4] -= 0.67 * rest
        out[5] -= 0.33 * rest
        return out.reshape(-1)


    def precond_matrix(self, mca, tau_r):
        """3nz x 3nz, 7 diagonals, unknowns [po4_s, dop_s, pop_s] (test_problem/phosphorus.py:213-290)"""
        d = self.g.depth
        nz = self.g.nz
        day_per_sec = 1.0 / 86400.0
        single = np.zeros(nz)
        single[:-1] -= mca * d.delta_mid_r * d.delta_r[:-1]
        single[1:] -= mca * d.delta_mid_r * d.delta_r[1:]
        d0_po4 = single - tau_r
        d0_dop = single - 0.01 * day_per_sec
        d0_pop = single - 0.01 * day_per_sec
        d0_pop[:-1] -= day_per_sec * d.delta_r[:-1]
        zero = np.zeros(1)
        up = mca * d.delta_mid_r * d.delta_r[:-1]
        lo = mca * d.delta_mid_r * d.delta_r[1:]
        lo_pop = lo + day_per_sec * d.delta_r[1:]
        return sparse.diags(
            [
                np.concatenate((d0_po4, d0_dop, d0_pop)),
                np.concatenate((up, zero, up, zero, up)),
                np.concatenate((lo, zero, lo, zero, lo_pop)),
                np.concatenate((0.01 * day_per_sec * np.ones(nz), np.zeros(nz))),
                np.concatenate((0.67 * tau_r, np.zeros(nz))),
                0.01 * day_per_sec * np.ones(nz),
                0.33 * tau_r,
            ],
            [0, 1, -1, nz, -nz, 2 * nz, -2 * nz],
            format="csr",
        )

    def apply_precond_jacobian(self, y_shadow, mca, tau_r):
        """res for the shadow tracers [3, nz] (test_problem/phosphorus.py:169-211): two regularised
        sparse solves + Richardson extrapolation, removal of the null vector (smallest singular
        value) weighted by layer thickness, minus the input"""
        from scipy import linalg
        from scipy.sparse import linalg as sp_linalg

        nz = self.g.nz
        self_vals = np.asarray(y_shadow, dtype=np.float64).reshape(-1)
        t0, t1 = self.g.time_range
        rhs_vals = (1.0 / (t1 - t0)) * self_vals
        matrix = self.precond_matrix(mca, tau_r)
        res_a = sp_linalg.spsolve(matrix - 1.0e-11 * sparse.eye(3 * nz), rhs_vals)
        res_b = sp_linalg.spsolve(matrix - 0.5e-11 * sparse.eye(3 * nz), rhs_vals)
        res_vals = 2.0 * res_b - res_a
        _, sing_vals, r_sing_vects = linalg.svd(matrix.todense())
        min_ind = sing_vals.argmin()
        dz3 = np.concatenate((self.g.depth.delta,) * 3)
        numer = (res_vals * dz3).sum()
        denom = (r_sing_vects[min_ind, :] * dz3).sum()
        res_vals -= numer / denom * r_sing_vects[min_ind, :]
        return (res_vals - self_vals).reshape(3, nz)


def comp_fcn_1d(module, x0, t_eval=None, rtol=1.0e-12, atol=1.0e-12, return_sol=False):
    """test_problem/model_state.py:83-103"""
    flat0 = np.asarray(x0, dtype=np.float64).reshape(-1)
    sol = integrate.solve_ivp(
        module.comp_tend, module.g.time_range, flat0, "Radau", t_eval, atol=atol, rtol=rtol
    )
    res = (sol.y[:, -1] - flat0).reshape(np.shape(x0))
    return (res, sol) if return_sol else res


def log_mean_mixing_coeff_1d(col, n_t=101):
    """mixing_coeff_log_mean of the precond file: exp(mean_t log(mixing_coeff*dz_mid)), plain mean
    over the 101 hist times (model_state_base.py:463-466), end rows copied from neighbours
    (test_problem/model_state.py:200-225; model_state_base.py:580-616;
    tracer_module_state_base.py hist_time_mean_weights)"""
    times = np.linspace(col.time_range[0], col.time_range[1], n_t)
    vals = np.empty((n_t, col.nz + 1))
    for i, t in enumerate(times):
        vals[i, 1:-1] = col.mixing_coeff(t) * col.depth.delta_mid
    vals[:, 0] = vals[:, 1]
    vals[:, -1] = vals[:, -2]
    return np.exp(np.log(vals).mean(axis=0))


# --------------------------------------------------------------------------------------
# solver scalars  (model_config.py:292-315; tracer_module_state_base.py:371-388;
#                  model_state_base.py:492-527)
# --------------------------------------------------------------------------------------


def region_weights(region_mask, grid_weight):
    """w[r, cell] = grid_weight/sum_region grid_weight, 0 outside (model_config.py:280-315)"""
    mask = np.where(grid_weight == 0.0, 0, region_mask)
    wgt = np.where(mask == 0, 0.0, grid_weight)
    rcnt = int(mask.max())
    out = np.zeros((rcnt,) + mask.shape)
    for r in range(rcnt):
        sel = mask == r + 1
        out[r][sel] = (1.0 / sum(wgt[sel])) * wgt[sel]
    return out


def dot_prod(w, a, b):
    """dot[r] = sum_tracer sum_cell w[r]*a*b for a, b [T, ...] (tracer_module_state_base.py:379-388)"""
    T = a.shape[0]
    return np.array([sum((w[r] * a[t] * b[t]).sum() for t in range(T)) for r in range(w.shape[0])])


def norm(w, a):
    return np.sqrt(dot_prod(w, a, a))


def mean(w, a):
    T = a.shape[0]
    return np.array([sum((w[r] * a[t]).sum() for t in range(T)) for r in range(w.shape[0])])


def broadcast_region_vals(region_mask, vals, fill=1.0):
    """tracer_module_state_base.py:502-515"""
    out = np.full(region_mask.shape, fill)
    for r, v in enumerate(vals):
        out = np.where(region_mask == r + 1, v, out)
    return out


def fd_sigma(w, x):
    """sigma = 1e-4*||x||, 1 where 0 (model_state_base.py:509-511)"""
    s = 1.0e-4 * norm(w, x)
    return np.where(s == 0.0, 1.0, s)


# --------------------------------------------------------------------------------------
# column regions, colouring and index maps  (py_driver_2d/setup_solver.py:170-182;
#     notebooks/IRF_coloring_dev.ipynb:73-259,294-305,572-617)
# --------------------------------------------------------------------------------------


def min_by_region(region_cnt, region_mask, vals):
    """utils.py:544-558"""
    out = np.empty(region_cnt)
    for region_ind in range(region_cnt):
        out[region_ind] = np.amin(vals, initial=np.inf, where=region_mask == region_ind + 1)
    return out


def comp_scalef_lob(region_cnt, region_mask, base, increment, lob):
    """largest 0 <= scalef <= 1 by region with base + scalef*increment >= lob (utils.py:561-579)"""
    if lob is None or (base + increment >= lob).all():
        return np.ones(region_cnt)
    if (base < lob).any():
        raise ValueError("base < lob")
    scalef_all = np.ones(base.shape)
    np.divide(lob - base, increment, out=scalef_all, where=base + increment < lob)
    return min_by_region(region_cnt, region_mask, scalef_all)


def comp_scalef_upb(region_cnt, region_mask, base, increment, upb):
    """utils.py:582-600"""
    if upb is None or (base + increment <= upb).all():
        return np.ones(region_cnt)
    if (base > upb).any():
        raise ValueError("base > upb")
    scalef_all = np.ones(base.shape)
    np.divide(upb - base, increment, out=scalef_all, where=base + increment > upb)
    return min_by_region(region_cnt, region_mask, scalef_all)


def apply_limiter(region_cnt, region_mask, base, increment, lob, upb):
    """scalef of one tracer module, tracers stacked in front (tracer_module_state_base.py:115-151)"""
    scalef = np.ones(region_cnt)
    for t in range(base.shape[0]):
        if lob is not None:
            scalef = np.minimum(scalef, comp_scalef_lob(region_cnt, region_mask, base[t], increment[t], lob))
        if upb is not None:
            scalef = np.minimum(scalef, comp_scalef_upb(region_cnt, region_mask, base[t], increment[t], upb))
    return scalef


def column_region_mask(nz, ny, max_abs_vvel, horiz_mix_coeff):
    if max_abs_vvel == 0.0 and horiz_mix_coeff == 0.0:
        mask = np.empty((nz, ny), dtype=np.int32)
        for j in range(ny):
            mask[:, j] = j + 1
        return mask
    return np.ones((nz, ny), dtype=np.int32)


def index_maps(mask):
    """nd_to_flat (int32, -1 where masked) and flat_to_nd, C-order over mask != 0
    (IRF_coloring_dev.ipynb:572-580)"""
    nd_to_flat = np.full(mask.shape, -1, dtype=np.int32)
    cells = np.argwhere(mask != 0).astype(np.int32)
    for flat, idx in enumerate(cells):
        nd_to_flat[tuple(idx)] = flat
    return nd_to_flat, cells


def stencil_neighbours(mask, idx, offsets):
    out = []
    for off in offsets:
        nb = tuple(i + o for i, o in zip(idx, off))
        if all(0 <= n < s for n, s in zip(nb, mask.shape)) and mask[nb] != 0:
            out.append(nb)
    return out


def distance2_adjacency(mask, offsets):
    """cells that share a stencil neighbour (or are neighbours) conflict
    (IRF_coloring_dev.ipynb:73-259: conn_nd -> conn2_nd)"""
    cells = [tuple(i) for i in np.argwhere(mask != 0)]
    conn = {c: set(stencil_neighbours(mask, c, offsets)) for c in cells}
    conn2 = {}
    for c in cells:
        s = set(conn[c])
        for nb in conn[c]:
            s |= conn[nb]
        s.discard(c)
        conn2[c] = s
    return cells, conn2


def greedy_colouring(mask, offsets):
    """first-fit colouring in C-order of the distance-2 graph (IRF_coloring_dev.ipynb:294-305).
    Returns int32 colours (1-based; 0 where masked) and the colour count."""
    cells, conn2 = distance2_adjacency(mask, offsets)
    colour = np.zeros(mask.shape, dtype=np.int32)
    for c in cells:
        used = {int(colour[nb]) for nb in conn2[c]}
        k = 1
        while k in used:
            k += 1
        colour[c] = k
    return colour, int(colour.max())


def dimacs_edges(mask, offsets):
    """DIMACS 'p edge n m' / 'e i j' lines, 1-based flat ids, i<j (IRF_coloring_dev.ipynb:606-617)"""
    nd_to_flat, _ = index_maps(mask)
    cells, conn2 = distance2_adjacency(mask, offsets)
    edges = []
    for c in cells:
        i = int(nd_to_flat[c]) + 1
        for nb in sorted(conn2[c]):
            j = int(nd_to_flat[nb]) + 1
            if i < j:
                edges.append((i, j))
    lines = [f"p edge {len(cells)} {len(edges)}"] + [f"e {i} {j}" for i, j in edges]
    return lines


# ---- notebook-faithful connectivity (IRF_coloring_dev.ipynb cells 4-13, 23) --------------------
def ind_wrap(mask_shape, ind):
    """cell 4: periodic in the last dimension, tripole fold beyond the end of the second-to-last"""
    ind_list = list(ind)
    if ind_list[-1] < 0:
        ind_list[-1] += mask_shape[-1]
    if ind_list[-1] >= mask_shape[-1]:
        ind_list[-1] -= mask_shape[-1]
    if ind_list[-2] >= mask_shape[-2]:
        ind_list[-1] = mask_shape[-1] - 1 - ind_list[-1]
        ind_list[-2] = 2 * mask_shape[-2] - 1 - ind_list[-2]
    return tuple(ind_list)


def apply_ind_offsets(mask, inds, ind_offsets):
    """cell 4: wet cells reached from `inds` by `ind_offsets` (after ind_wrap)"""
    ret_val = set()
    for ind in inds:
        for ind_offset in ind_offsets:
            nb = ind_wrap(mask.shape, tuple(i + o for i, o in zip(ind, ind_offset[-mask.ndim:])))
            if all(0 <= nb[n] < mask.shape[n] for n in range(mask.ndim)) and mask[nb]:
                ret_val.add(nb)
    return ret_val


def gen_conn_nd(mask, ind):
    """cell 4: MOM6's computational stencil around ind"""
    if mask[ind] == 0:
        return set()
    ret_val = set()
    if mask.ndim == 3 and ind[-3] > 0:
        ret_val.update(apply_ind_offsets(mask, [ind], [(-1, -1, 0), (-1, 0, -1), (-1, 0, 0), (-1, 0, 1), (-1, 1, 0)]))
    offs_x = [(0, 0, -1), (0, 0, 0), (0, 0, 1)]
    offs_y = [(0, -1, 0), (0, 0, 0), (0, 1, 0)]
    ret_val.update(apply_ind_offsets(mask, apply_ind_offsets(mask, [ind], offs_x), offs_y))
    ret_val.update(apply_ind_offsets(mask, apply_ind_offsets(mask, [ind], offs_y), offs_x))
    if mask.ndim == 3 and ind[-3] < mask.shape[-3] - 1:
        ret_val.update(apply_ind_offsets(mask, [ind], [(1, -1, 0), (1, 0, -1), (1, 0, 0), (1, 0, 1), (1, 1, 0)]))
    return ret_val


def conn_mom6(mask):
    """cells 5-7: conn_nd and conn2_nd as dicts of sets over the wet cells"""
    cells = [ind for ind in np.ndindex(mask.shape) if mask[ind]]
    conn = {c: gen_conn_nd(mask, c) for c in cells}
    conn2 = {}
    for c in cells:
        s2 = set()
        for nb in conn[c]:
            s2.update(conn[nb])
        conn2[c] = s2
    return cells, conn, conn2


def greedy_colouring_sets(mask, cells, conn2, order=None):
    """cells 9-13: first-fit over `order` (default: C order); 0-based colours, -1 where masked"""
    color = np.full(mask.shape, -1)
    for ind in (cells if order is None else order):
        used = [color[nb] for nb in conn2[ind]]
        val = 0
        while val in used:
            val += 1
        color[ind] = val
    return color

