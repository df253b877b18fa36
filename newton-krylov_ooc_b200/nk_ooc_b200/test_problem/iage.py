"""iage tracer module of test_problem (nk_ooc/test_problem/iage.py): piston-velocity surface flux, ageing 1 / year"""

from .tracer_module_state import TracerModuleState


class iage(TracerModuleState):  # pylint: disable=invalid-name
    """iage tracer module specifics for TracerModuleState"""
