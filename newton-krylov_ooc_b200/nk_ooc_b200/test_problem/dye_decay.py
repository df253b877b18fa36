"""dye_decay_{suff} tracer module of test_problem (nk_ooc/test_problem/dye_decay.py): trapezoid-in-time surface
flux (1 mol / m^2 per year), decay at suff / 1000 per year"""

from .tracer_module_state import TracerModuleState


class dye_decay(TracerModuleState):  # pylint: disable=invalid-name
    """dye_decay tracer module specifics for TracerModuleState"""
