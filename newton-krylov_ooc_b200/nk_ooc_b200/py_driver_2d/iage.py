"""iage tracer module of py_driver_2d (nk_ooc/py_driver_2d/iage.py): two ideal-age tracers, surface restoring
to zero (the second one 100 times slower), ageing at 1 / year.  The restoring sits on the diagonal of the
implicit operator (Model.desc.surf_diag), the ageing is the constant explicit source."""

from .tracer_module_state import TracerModuleState


class iage(TracerModuleState):  # pylint: disable=invalid-name
    """iage tracer module specifics for TracerModuleState: everything is linear, so the generic hooks apply"""
