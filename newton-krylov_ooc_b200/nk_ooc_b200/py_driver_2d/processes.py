"""Time-invariant fields of the py_driver_2d processes (one-off set-up, host side).

Mirrors the set-up halves of the reference's process classes; the per-RHS halves
(comp_tend / comp_jacobian / mixing_coeff) run on the device (csrc/nkb_tables.cu,
csrc/nkb_stage.cu).
"""

import numpy as np

SEC_PER_YEAR = 365.0 * 86400.0


class Advection:
    """velocity field from a fixed overturning streamfunction
    (nk_ooc/py_driver_2d/advection.py:22-49)"""

    def __init__(self, depth, ypos, max_abs_vvel):
        self.depth, self.ypos = depth, ypos
        zn = (depth.edges - depth.edges.min()) / (depth.edges.max() - depth.edges.min())
        stretch = 2.0
        zn = stretch * zn / (1 + (stretch - 1) * zn)
        zf = (27.0 / 4.0) * zn * (1.0 - zn) ** 2
        yn = (ypos.edges - ypos.edges.min()) / (ypos.edges.max() - ypos.edges.min())
        yf = 4.0 * yn * (1.0 - yn)
        stream = np.outer(zf, yf)
        vraw = (stream[1:, :] - stream[:-1, :]) * depth.delta_r[:, np.newaxis]
        stream = stream * max_abs_vvel / abs(vraw).max()
        self.stream = stream
        self.vvel = (stream[1:, :] - stream[:-1, :]) * depth.delta_r[:, np.newaxis]
        self.wvel = (stream[:, 1:] - stream[:, :-1]) * ypos.delta_r


class HorizMix:
    """Peclet-limited horizontal mixing coefficient, including the 1/dy_mid factor
    (nk_ooc/py_driver_2d/horiz_mix.py:25-46)"""

    def __init__(self, depth, ypos, horiz_mix_coeff, advection):
        self.depth, self.ypos = depth, ypos
        vin = abs(advection.vvel[:, 1:-1])
        if horiz_mix_coeff > 0.0:
            res = np.full((len(depth), len(ypos) - 1), horiz_mix_coeff)
            peclet_p5 = (0.5 / horiz_mix_coeff) * ypos.delta_mid[:] * vin
            res *= np.where(peclet_p5 > 1.0, peclet_p5, 1.0)
            res *= ypos.delta_mid_r
        else:
            res = 0.5 * vin
        self.mixing_coeff = res


class VertMix:
    """static inputs of the time-varying vertical mixing (nk_ooc/py_driver_2d/vert_mix.py:89-97);
    bldepth(t), the remap and the Peclet limiter are evaluated on the device"""

    bldepth_min = 35.0

    def __init__(self, depth, ypos):
        self.depth, self.ypos = depth, ypos
        self.bldepth_max = np.interp(
            ypos.mid,
            [0.4e6, 0.8e6, 1.0e6, 1.2e6, 1.4e6, 1.5e6],
            [3000.0, 800.0, 415.0, 325.0, 280.0, self.bldepth_min],
        )

    def bldepth(self, time):
        """host copy, used only for hist output (vert_mix.py:89-101)"""
        tvals = SEC_PER_YEAR * np.array([0.25, 0.35, 0.65, 0.75])
        frac = np.interp(time, tvals, [0.0, 1.0, 1.0, 0.0])
        return self.bldepth_min + (self.bldepth_max - self.bldepth_min) * frac


def explicit_stencil(depth, ypos, advection, horiz_mix):
    """horizontal advection + mixing tendency in coefficient form,
         E(c)[k, j] = eL[k, j] c[k, j-1] + eC[k, j] c[k, j] + eR[k, j] c[k, j+1],
    algebraically the flux form of advection.py:58-65 and horiz_mix.py:59-65 with zero
    flux through the side walls.  Returns [3, nz, ny]."""
    nz, ny = len(depth), len(ypos)
    dyr = ypos.delta_r[np.newaxis, :]
    v = advection.vvel.copy()
    v[:, 0] = 0.0
    v[:, -1] = 0.0
    kh = np.zeros((nz, ny + 1))
    kh[:, 1:-1] = horiz_mix.mixing_coeff
    e_l = dyr * (0.5 * v[:, :-1] + kh[:, :-1])
    e_r = dyr * (-0.5 * v[:, 1:] + kh[:, 1:])
    e_c = dyr * (0.5 * v[:, :-1] - 0.5 * v[:, 1:] - kh[:, :-1] - kh[:, 1:])
    e_l[:, 0] = 0.0
    e_r[:, -1] = 0.0
    return np.ascontiguousarray(np.stack([e_l, e_c, e_r]))
