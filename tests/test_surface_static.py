"""CPU: the operator surface the reference's solvers and drivers actually use, read from the reference's own
sources with `ast` (build container only: needs /root/reference), checked against the host mirror:

* every ModelStateBase method that nk_ooc/{newton_solver,krylov_solver,solver_base,nk_driver,run_cmd}.py and the
  models' setup_solver.py call exists on nk_ooc_b200's ModelState classes and accepts the arguments of every
  call site (positional count and keyword names);
* every operator (dunder) the reference's ModelStateBase defines exists;
* module-level entry points (lin_comb, get_model_state_class) exist with compatible signatures;
* the tracer-module hooks (comp_tend, comp_jacobian, comp_jacobian_sparsity, apply_precond_jacobian) of the
  reference's per-module classes exist on the classes that the mirror discovers by the same module-path rule.

This does not run the reference's solvers over the mirror (the mirror needs a GPU, the reference's sources are
not on the GPU box); it removes the failure mode "AttributeError at the first call the solver makes"."""
import ast
import inspect
import os

import pytest

REF = os.environ.get("NK_REF_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "nk_ooc")), reason="reference tree not present")

CALLERS = ["newton_solver.py", "krylov_solver.py", "solver_base.py", "nk_driver.py", "run_cmd.py",
           "py_driver_2d/setup_solver.py", "test_problem/setup_solver.py"]


def _parse(rel):
    with open(os.path.join(REF, "nk_ooc", rel)) as f:
        return ast.parse(f.read())


def _class_methods(tree, class_name):
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == class_name:
            return {n.name: n for n in node.body if isinstance(n, ast.FunctionDef)}
    raise AssertionError(f"class {class_name} not found")


def _call_sites(names):
    """[(file, lineno, method, n_positional, keyword names)] of calls <expr>.<method>(...) in the callers"""
    sites = []
    for rel in CALLERS:
        for node in ast.walk(_parse(rel)):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr in names:
                if any(isinstance(a, ast.Starred) for a in node.args):
                    continue
                recv = node.func.value
                if isinstance(recv, ast.Name) and recv.id in ("logger", "logging", "np", "os", "json", "parser", "self"):
                    continue  # logger.log(...), the solver's own self.log(...): not a model state
                sites.append((rel, node.lineno, node.func.attr, len(node.args), [k.arg for k in node.keywords if k.arg]))
    return sites


def _mirror_classes():
    from nk_ooc_b200.model_state_base import ModelStateBase, get_model_state_class

    return ModelStateBase, [get_model_state_class("py_driver_2d"), get_model_state_class("test_problem")]


def test_every_model_state_method_the_solvers_call_exists_and_binds():
    ref_methods = _class_methods(_parse("model_state_base.py"), "ModelStateBase")
    public = {n for n in ref_methods if not n.startswith("_")}
    sites = _call_sites(public)
    used = sorted({s[2] for s in sites})
    # the survey's list of what the solvers call must be a subset of what this finds
    for must in ("comp_fcn", "apply_precond_jacobian", "gen_precond_jacobian", "comp_jacobian_fcn_state_prod", "norm",
                 "mod_gram_schmidt", "apply_limiter", "dump", "log", "log_vals", "def_stats_vars",
                 "put_stats_vars_iteration_invariant", "put_stats_vars", "copy_shadow_tracers_to_real_tracers",
                 "copy_real_tracers_to_shadow_tracers"):
        assert must in used, f"{must} not found among the reference's call sites"
    base, classes = _mirror_classes()
    for cls in classes:
        for rel, lineno, name, npos, kws in sites:
            fn = getattr(cls, name, None)
            # comp_fcn & co are "must be implemented in derived class" in the reference's base as well
            assert fn is not None, f"{cls.__module__}.{cls.__name__} lacks {name} (called at nk_ooc/{rel}:{lineno})"
            sig = inspect.signature(fn)
            try:
                sig.bind(None, *([None] * npos), **{k: None for k in kws})
            except TypeError as err:
                raise AssertionError(f"{cls.__name__}.{name}{sig} does not accept the call at nk_ooc/{rel}:{lineno} "
                                     f"({npos} positional, keywords {kws}): {err}") from None


def test_every_operator_of_the_reference_exists():
    ref_methods = _class_methods(_parse("model_state_base.py"), "ModelStateBase")
    dunders = sorted(n for n in ref_methods if n.startswith("__") and n.endswith("__") and n != "__init__")
    assert "__radd__" in dunders and "__rtruediv__" in dunders and "__itruediv__" in dunders
    base, classes = _mirror_classes()
    for cls in classes:
        for name in dunders:
            assert callable(getattr(cls, name, None)), f"{cls.__name__} lacks {name}"
        assert cls.__array_priority__ == 100  # numpy defers to the reversed operators (model_state_base.py:27-29)


def test_module_level_entry_points():
    from nk_ooc_b200 import model_state_base as ours

    tree = _parse("model_state_base.py")
    ref_fns = {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef)}
    for name in ("lin_comb", "get_model_state_class"):
        assert name in ref_fns
        ref_args = [a.arg for a in ref_fns[name].args.args]
        our_params = list(inspect.signature(getattr(ours, name)).parameters)
        assert our_params[: len(ref_args)] == ref_args, f"{name}: reference {ref_args}, mirror {our_params}"


@pytest.mark.parametrize("model,modules", [("py_driver_2d", ["iage", "forced", "phosphorus"]),
                                           ("test_problem", ["iage", "dye_decay", "phosphorus"])])
def test_tracer_module_hooks_exist_with_the_reference_parameters(model, modules):
    from nk_ooc_b200.model_state_base import get_tracer_module_state_class

    hooks = ("comp_tend", "comp_jacobian", "comp_jacobian_sparsity", "apply_precond_jacobian")
    base_methods = _class_methods(_parse(f"{model}/tracer_module_state.py"), "TracerModuleState")
    for mod in modules:
        ref = dict(base_methods)
        ref.update(_class_methods(_parse(f"{model}/{mod}.py"), mod))
        cls = get_tracer_module_state_class(model, mod, {"py_mod_name": mod})
        assert cls.__name__ == mod and cls.__module__ == f"nk_ooc_b200.{model}.{mod}"
        for hook in hooks:
            if hook not in ref:
                continue
            if model == "test_problem" and hook == "apply_precond_jacobian":
                continue  # served by ModelState.apply_precond_jacobian from the precond file's variables (a-10)
            ref_args = [a.arg for a in ref[hook].args.args][1:]
            fn = getattr(cls, hook, None)
            assert fn is not None, f"{cls.__name__} lacks {hook}"
            ours = list(inspect.signature(fn).parameters)[1:]
            assert len(ours) >= len(ref_args), f"{mod}.{hook}: reference {ref_args}, mirror {ours}"
            inspect.signature(fn).bind(None, *([None] * len(ref_args)))
