"""test_problem tracer-module hooks — mirror of the comp_tend callbacks of nk_ooc/test_problem/{iage,dye_decay,
phosphorus}.py (signature comp_tend(time, tracer_vals_flat, vert_mix), the solve_ivp right-hand side) over
the CUDA library (kernel nkb_model_tend on the column model)."""

import numpy as np
import torch

from ..model_state_base import TracerModuleStateBase


class TracerModuleState(TracerModuleStateBase):
    """test_problem specifics of TracerModuleStateBase"""

    def _model(self):
        from .model_state import ModelState

        return ModelState.models_for(self)[0]

    def comp_tend(self, time, tracer_vals_flat, vert_mix=None):
        """d tracer / dt at `time` for a flat ndarray [tracer_cnt * nz] (a flat ndarray comes back) or a
        member-fastest device tensor [tracer, nz, 1, ldb].  `vert_mix` is accepted for signature parity: the
        mixing coefficient is a device table of the model (K3)."""
        model = self._model()
        if isinstance(tracer_vals_flat, torch.Tensor):
            return model.tend(time, tracer_vals_flat, self.members)
        flat = np.ascontiguousarray(tracer_vals_flat, dtype=np.float64).reshape(-1)
        nz = self.cell_shape[0]
        if flat.size != self.tracer_cnt * nz:
            raise ValueError(f"tracer_vals_flat has {flat.size} values, expected {self.tracer_cnt * nz}")
        x = torch.from_numpy(flat.reshape(self.tracer_cnt, nz, 1, 1)).cuda()
        return model.tend(time, x, 1)[:, :, 0, 0].cpu().numpy().reshape(-1)
