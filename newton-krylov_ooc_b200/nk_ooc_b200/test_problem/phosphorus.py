"""phosphorus tracer module of test_problem (nk_ooc/test_problem/phosphorus.py): po4, dop, pop and their shadows"""

from . import modules
from .tracer_module_state import TracerModuleState


class phosphorus(TracerModuleState):  # pylint: disable=invalid-name
    """phosphorus tracer module specifics for TracerModuleState"""

    def stats_vars_tracer_like(self):
        """tracers + po4_uptake (test_problem/phosphorus.py:161-167)"""
        return list(self.tracer_names) + ["po4_uptake"]

    def po4_uptake(self, po4):
        """test_problem/phosphorus.py:73-79"""
        from .model_state import ModelState

        return modules.po4_uptake(ModelState.depth, po4)

    def po4_s_restore_tau_r(self, po4, po4_uptake):
        """test_problem/phosphorus.py:58-71"""
        from .model_state import ModelState

        opt = int(ModelState.model_config_obj.modelinfo.get("po4_s_restoring_opt", 1))
        return modules.po4_s_restore_tau_r(ModelState.depth, po4, po4_uptake, opt)
