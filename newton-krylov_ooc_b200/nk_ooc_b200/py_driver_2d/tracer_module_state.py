"""py_driver_2d tracer-module hooks — mirror of nk_ooc/py_driver_2d/tracer_module_state.py:98-108,262-279
(comp_tend, comp_jacobian, comp_jacobian_sparsity with the reference's solve_ivp-callback signatures) over
the CUDA library: the tendency is kernel nkb_model_tend; the Jacobian is assembled on the host from the
device's vertical mixing coefficients (K3), member independent."""

import numpy as np
import torch
from scipy import sparse

from .. import engine
from ..model_state_base import TracerModuleStateBase


class TracerModuleState(TracerModuleStateBase):
    """py_driver_2d specifics of TracerModuleStateBase"""

    def _model(self):
        from .model_state import ModelState

        return ModelState.model_for(self)

    def _transport(self):
        from .model_state import ModelState

        return ModelState.transport

    # ---- tendency (tracer_module_state.py:98-108 + the module's own sources) ------------------------------
    def comp_tend(self, time, tracer_vals, processes=None):
        """d tracer / dt at `time`.  tracer_vals: flat ndarray [tracer_cnt * nz * ny] (the reference's
        solve_ivp callback layout; a flat ndarray comes back) or a member-fastest device tensor
        [tracer, nz, ny, ldb] (a device tensor comes back).  `processes` is accepted for signature parity:
        the process coefficients are class-level device tables (Transport2D)."""
        model = self._model()
        if isinstance(tracer_vals, torch.Tensor):
            return model.tend(time, tracer_vals, self.members)
        flat = np.ascontiguousarray(tracer_vals, dtype=np.float64).reshape(-1)
        shape = (self.tracer_cnt,) + self.cell_shape
        if flat.size != int(np.prod(shape)):
            raise ValueError(f"tracer_vals has {flat.size} values, expected {int(np.prod(shape))}")
        x = torch.from_numpy(flat.reshape(shape + (1,))).cuda()
        return model.tend(time, x, 1)[..., 0].cpu().numpy().reshape(-1)

    # ---- Jacobian -------------------------------------------------------------------------------------------
    def comp_jacobian_transport(self, time, tracer_ind):
        """CSR Jacobian of one tracer's transport + linear module terms, cell = j + ny*k (advection.py:111-179,
        horiz_mix.py:100-149, vert_mix.py:140-188, iage.py:55-64, forced.py:170-186)"""
        model = self._model()
        tr = self._transport()
        nz, ny = self.cell_shape
        n = nz * ny
        idx = np.arange(n).reshape(nz, ny)
        dzr = tr.depth.delta_r[:, np.newaxis]
        mc = model.mixing_coeff(time).cpu().numpy()
        w = tr.advection.wvel
        e_l, e_c, e_r = tr.estencil
        rows, cols, vals = [], [], []

        def add(r, c, v):
            rows.append(r.ravel())
            cols.append(c.ravel())
            vals.append(np.broadcast_to(v, r.shape).ravel())

        diag = e_c.copy()
        add(idx[:, 1:], idx[:, :-1], e_l[:, 1:])
        add(idx[:, :-1], idx[:, 1:], e_r[:, :-1])
        add(idx[1:], idx[:-1], (-0.5 * w[1:-1] + mc) * dzr[1:])
        diag[1:] += (-0.5 * w[1:-1] - mc) * dzr[1:]
        add(idx[:-1], idx[1:], (0.5 * w[1:-1] + mc) * dzr[:-1])
        diag[:-1] += (0.5 * w[1:-1] - mc) * dzr[:-1]
        desc = model.desc
        cls_ind = desc.class_of[tracer_ind]
        diag[0] += desc.surf_diag[cls_ind]
        diag += desc.decay[cls_ind]
        add(idx, idx, diag)
        return sparse.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))

    def comp_jacobian(self, time, tracer_vals, processes=None):
        """sparse Jacobian of comp_tend, block diagonal over the tracers (tracer_module_state.py:262-270);
        modules with state-dependent or coupling terms override and add them"""
        return sparse.block_diag([self.comp_jacobian_transport(time, t) for t in range(self.tracer_cnt)], format="csr")

    def comp_jacobian_sparsity(self, time, tracer_vals, processes=None):
        """sparsity pattern of comp_jacobian (tracer_module_state.py:272-279)"""
        row_ind, col_ind, _ = sparse.find(self.comp_jacobian(time, tracer_vals, processes))
        return sparse.csr_matrix((np.ones(row_ind.shape), (row_ind, col_ind)))

    # ---- preconditioner hook (the per-module implementations live with ModelState, which owns the factor cache)
    def apply_precond_jacobian(self, time_range, res_tms, processes=None, fptr_precond=None):
        """res_tms <- M^-1 self - self (iage.py:66-93, forced.py:204-241, phosphorus.py:197-274); fptr_precond:
        the precond file's name (forced and phosphorus read tracer snapshots from it)"""
        from .model_state import ModelState

        res_tms.vals = ModelState.apply_precond_module(self, fptr_precond)
        return res_tms


def tracer_snapshot_nearest(precond_fname, tracer_name, time_end):
    """tracer snapshot of the precond file nearest time_end (forced.py:221-224, phosphorus.py:215-218)"""
    from scipy.io import netcdf_file

    with netcdf_file(precond_fname, "r", mmap=False) as fptr:
        ptimes = np.array(fptr.variables["time"].data)
        return np.array(fptr.variables[tracer_name].data)[np.argmin(abs(time_end - ptimes))]


__all__ = ["TracerModuleState", "tracer_snapshot_nearest", "engine"]
