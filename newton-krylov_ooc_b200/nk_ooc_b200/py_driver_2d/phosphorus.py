"""phosphorus tracer module of py_driver_2d (nk_ooc/py_driver_2d/phosphorus.py): po4, dop, pop with
Michaelis-Menten uptake of po4 under a light field, first-order remineralisation and sinking of pop."""

import numpy as np
from scipy import sparse

from .tracer_module_state import TracerModuleState


class phosphorus(TracerModuleState):  # pylint: disable=invalid-name
    """phosphorus tracer module specifics for TracerModuleState"""

    def stats_vars_tracer_like(self):
        return list(self.tracer_names)

    def comp_po4_uptake(self, po4):
        """phosphorus.py:91-95"""
        model = self._model()
        desc = model.desc
        return desc.max_uptake_rate * model._keepalive["light"] * po4 / (po4 + desc.po4_halfsat)

    def comp_po4_uptake_jacobian(self, po4):
        """d uptake / d po4 (phosphorus.py:97-103)"""
        model = self._model()
        desc = model.desc
        return desc.max_uptake_rate * model._keepalive["light"] * desc.po4_halfsat / (po4 + desc.po4_halfsat) ** 2

    def comp_jacobian(self, time, tracer_vals, processes=None):
        """transport blocks + the uptake / remineralisation coupling + pop sinking (phosphorus.py:105-172)"""
        model = self._model()
        desc = model.desc
        nz, ny = self.cell_shape
        ncell = nz * ny
        depth = self._transport().depth
        po4 = np.asarray(tracer_vals, dtype=np.float64).reshape((3,) + self.cell_shape)[0]
        blocks = [[None] * 3 for _ in range(3)]
        for t in range(3):
            blocks[t][t] = self.comp_jacobian_transport(time, t)
        du = sparse.diags(self.comp_po4_uptake_jacobian(po4).reshape(-1))
        ident = sparse.identity(ncell, format="csr")
        sink = desc.sink_vel[desc.class_of[2]]
        d0 = np.broadcast_to(-sink * depth.delta_r[:, np.newaxis], (nz, ny)).copy()
        d0[-1, :] = 0.0
        dm1 = np.broadcast_to(sink * depth.delta_r[1:, np.newaxis], (nz - 1, ny))
        sinkb = sparse.diags((d0.reshape(-1), dm1.reshape(-1)), (0, -ny))
        blocks[0][0] = blocks[0][0] - du
        blocks[1][0] = desc.sigma * du
        blocks[2][0] = (1.0 - desc.sigma) * du
        blocks[0][1] = desc.dop_remin_rate * ident
        blocks[0][2] = desc.pop_remin_rate * ident
        blocks[1][1] = blocks[1][1] - desc.dop_remin_rate * ident
        blocks[2][2] = blocks[2][2] - desc.pop_remin_rate * ident + sinkb
        return sparse.bmat(blocks, format="csr")
