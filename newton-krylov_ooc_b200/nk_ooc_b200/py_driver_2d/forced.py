"""forced_{suff} tracer module of py_driver_2d (nk_ooc/py_driver_2d/forced.py): one tracer, surface restoring to
a constant or a record, interior source / sink: constant, first-order decay, or a record with the sink_thres
limiter (the sink is scaled by tracer / thres where 0 < tracer < thres)."""

import numpy as np
from scipy import sparse

from .. import _lib
from .tracer_module_state import TracerModuleState


class forced(TracerModuleState):  # pylint: disable=invalid-name
    """forced tracer module specifics for TracerModuleState"""

    def sms(self, time):
        """forcing record at `time`: linear in time, linearly extrapolated (utils.py:533-535)"""
        keep = self._model()._keepalive
        ft, fd = keep["ft"], keep["fd"]
        i = int(np.clip(np.searchsorted(ft, time, side="right") - 1, 0, len(ft) - 2))
        return fd[i] + (time - ft[i]) / (ft[i + 1] - ft[i]) * (fd[i + 1] - fd[i])

    def comp_jacobian_sms_file(self, time, tracer_vals):
        """d sms / d tracer of the sink_thres limiter (forced.py:188-202)"""
        desc = self._model().desc
        sms = self.sms(time)
        q = np.asarray(tracer_vals, dtype=np.float64).reshape(self.cell_shape) / desc.sink_thres
        return sparse.diags(np.where((sms < 0.0) & (q > 0.0) & (q < 1.0), sms / desc.sink_thres, 0.0).reshape(-1))

    def comp_jacobian(self, time, tracer_vals, processes=None):
        """forced.py:156-168 (surface restoring and decay are in the transport block's diagonal)"""
        jac = super().comp_jacobian(time, tracer_vals, processes)
        desc = self._model().desc
        if desc.kind == _lib.MOD_FORCED_FILE and desc.sink_thres > 0.0:
            jac = jac + self.comp_jacobian_sms_file(time, tracer_vals)
        return jac.tocsr()
